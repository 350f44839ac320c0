# InvertibleNetworksB200.jl - the reference-side binding of libinb200.so (include/inb200.h).
#
# `using InvertibleNetworks, InvertibleNetworksB200` adds methods that are strictly more specific than
# the reference's `AbstractArray` methods (CuArray{Float32,N} inputs), so existing user code
#     Z, lgdet = G.forward(X);  ΔX, X = G.backward(ΔZ, Z);  get_params(G);  clear_grad!(G)
# runs unchanged and lands in the B200 library - the same mechanism the reference already uses to
# specialise on CuArray (src/utils/compute_utils.jl:6-18, src/layers/invertible_layer_conv1x1.jl:89).
# Every overload first asks `on_b200(...)` whether the object is one the library implements (checkerboard
# squeezer, ResidualBlock with fan / ReLU / unit strides / "same" padding, SigmoidLayer activation, all ActNorms
# in the same initialisation state); everything else goes back to the reference method through `invoke`.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia toolchain.  What IS checked here, on every
# test run (tests/test_julia_shim.py): every `obj.field` access below against the reference's struct definitions,
# every overloaded signature against the reference's method of the same name / arity / object type / keywords,
# every `ccall` (symbol, argument count, argument types) against include/inb200.h.  The identical C ABI is
# exercised from Python ctypes (invertiblenetworks.jl_b200/lib.py, tests/).
module InvertibleNetworksB200

using CUDA
using InvertibleNetworks
import InvertibleNetworks: forward, inverse, backward, squeeze, unsqueeze, wavelet_squeeze, wavelet_unsqueeze,
                           Haar_squeeze, invHaar_unsqueeze, get_params,
                           NetworkGlow, NetworkConditionalGlow, NetworkMultiScaleHINT, CouplingLayerHINT,
                           CouplingLayerBasic, ActNorm, Conv1x1, CouplingLayerGlow, ConditionalLayerGlow,
                           ResidualBlock, Parameter, ActivationFunction, Squeezer, ReLU

const LIB = get(ENV, "INB200_LIB", joinpath(@__DIR__, "..", "invertiblenetworks.jl_b200", "libinb200.so"))

# mirrors `inb_glow_desc`
struct GlowDesc
    ndims::Cint; nx::Cint; ny::Cint; nz::Cint
    n_in::Cint; n_cond::Cint; n_hidden::Cint; L::Cint; K::Cint; batch::Cint
    split_scales::Cint; logdet::Cint
    k1::Cint; k2::Cint; p1::Cint; p2::Cint
    sig_low::Cfloat; sig_high::Cfloat
    freeze_conv::Cint; precision::Cint
end

# 0 fp32, 1 bf16x3, 2 bf16, 3 fp16x3 (INB_PREC_*); fp16x3 = float32-level products on the tensor cores
const PRECISION = Ref{Cint}(parse(Cint, get(ENV, "INB200_PRECISION", "3")))

check(rc) = rc == 0 || error(unsafe_string(ccall((:inb_last_error, LIB), Cstring, ())))
stream() = CUDA.stream().handle
dptr(x::CuArray{Float32}) = reinterpret(Ptr{Cfloat}, pointer(x))
dptr(::Nothing) = Ptr{Cfloat}(C_NULL)

# ---------------------------------------------------------------- what the library implements
# SigmoidLayer(low, high) returns an ActivationFunction whose fields are forward / inverse / backward ONLY
# (src/utils/activation_functions.jl:16-20); low and high live in the closures it builds (:30-35).  A closure is a
# struct whose fields are the captured variables, so they are read from `act.forward`, and the result is verified
# on two probe values (σ(0) = (low + high) / 2, σ(+big) = high).  Anything else (ExpClampLayer, Sigmoid2Layer,
# a user activation) returns `nothing` and the call stays on the reference.
function sigmoid_bounds(act::ActivationFunction)
    f = act.forward
    (hasproperty(f, :low) && hasproperty(f, :high)) || return nothing
    lo, hi = Float32(getfield(f, :low)), Float32(getfield(f, :high))
    y = f(Float32[0f0, 60f0])
    (isapprox(y[1], (lo + hi) / 2; atol=1f-6) && isapprox(y[2], hi; atol=1f-6) && hi > lo) || return nothing
    return (lo, hi)
end
# channel counts the per-pixel kernels are instantiated for (csrc/elementwise.cu INB_FOR_C); a layer or network with any
# other count (Conv1x1.k, invertible_layer_conv1x1.jl:42-49; ActNorm.k, invertible_layer_actnorm.jl:42-48) stays on the
# reference
const SUPPORTED_C = (1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64)
is_relu(act::ActivationFunction) = act.forward === ReLU                        # ReLUlayer(), :22-24
# ShuffleLayer(; pattern) is a FUNCTION that returns Squeezer(x -> squeeze(x; pattern=pattern), ...)
# (src/utils/dimensionality_operations.jl:12-19): the pattern is the closure's captured variable
function is_checkerboard(sq::Squeezer)
    f = sq.forward
    return hasproperty(f, :pattern) && getfield(f, :pattern) == "checkerboard"
end
# ResidualBlock(W1, W2, W3, b1, b2, fan, strides, pad, activation) (src/layers/layer_residual_block.jl:67-77);
# dense=true builds a FluxBlock instead (invertible_layer_glow.jl:65)
function rb_on_b200(RB)
    RB isa ResidualBlock || return false
    (RB.fan && is_relu(RB.activation) && all(==(1), RB.strides)) || return false
    k1, k2 = size(RB.W1.data, 1), size(RB.W2.data, 1)
    return (k1 == 1 || k1 == 3) && (k2 == 1 || k2 == 3) && RB.pad[1] == (k1 - 1) ÷ 2 && RB.pad[2] == (k2 - 1) ÷ 2 &&
           RB.W1.data isa CuArray{Float32}
end
layer_on_b200(L::Union{CouplingLayerGlow,ConditionalLayerGlow}) =
    rb_on_b200(L.RB) && sigmoid_bounds(L.activation) !== nothing && L.C.k in SUPPORTED_C
# :uninit - every ActNorm still has s.data === nothing (the library runs the data-dependent initialisation),
# :ready - all set, :mixed - a partially initialised network stays on the reference
function actnorm_state(ANs)
    n_unset = count(AN -> AN.s.data === nothing, ANs)
    return n_unset == 0 ? :ready : (n_unset == length(ANs) ? :uninit : :mixed)
end
all_actnorms(G::NetworkGlow) = vec(G.AN)
all_actnorms(G::NetworkConditionalGlow) = vcat(vec(G.AN), [G.AN_C])
all_actnorms(H::NetworkMultiScaleHINT) = vec(H.AN)
function on_b200(G::Union{NetworkGlow,NetworkConditionalGlow})
    is_checkerboard(G.squeezer) || return false
    all(layer_on_b200, G.CL) || return false
    all(L -> L.C.k in SUPPORTED_C, G.CL) || return false
    all(AN -> AN.k in SUPPORTED_C, all_actnorms(G)) || return false
    b = sigmoid_bounds(G.CL[1, 1].activation)
    all(L -> sigmoid_bounds(L.activation) == b, G.CL) || return false
    all(AN -> !AN.is_reversed, all_actnorms(G)) || return false
    return actnorm_state(all_actnorms(G)) != :mixed
end

# ---------------------------------------------------------------- plans and their persistent buffers
# One plan per network object and input geometry.  The plan's CUDA graphs replay only when the pointers of a call
# repeat (api.cu run_graphed), so the shim keeps what it can stable: the parameter / gradient pointer tables, ONE
# flat gradient buffer in the library's canonical layout (inb_glow_flat_layout; p.grad becomes a view into it, and a
# data-parallel all-reduce covers it with one collective per scale), and a ring of two output sets.
mutable struct PlanState
    key::Any
    plan::Ptr{Cvoid}
    θ::Vector{Ptr{Cfloat}}
    ∇flat::CuVector{Float32}
    ∇::Vector{Ptr{Cfloat}}
    ∇views::Vector{Any}
    out::Vector{Dict{Symbol,Any}}      # ring of output buffers
    turn::Int
    comm::Ptr{Cvoid}
end
const PLANS = IdDict{Any,PlanState}()
const COMM = Ref{Ptr{Cvoid}}(C_NULL)   # communicator attached to every plan created after `attach_comm!`

# get_params order == the library's pointer-table order (src/utils/neuralnet.jl:72-88)
ptr_table(ps::Vector{Parameter}) = Ptr{Cfloat}[dptr(p.data) for p in ps]

function destroy_plan!(st::PlanState)
    ccall((:inb_glow_plan_destroy, LIB), Cint, (Ptr{Cvoid},), st.plan)
end

# dims = size of the network input (nx, ny[, nz], C, B)
function plan_for(G::Union{NetworkGlow,NetworkConditionalGlow}, dims::NTuple{N,Int}, n_cond::Int) where N
    key = (dims[1:N-2]..., dims[N], n_cond, PRECISION[])
    if haskey(PLANS, G)
        PLANS[G].key == key && return PLANS[G]
        destroy_plan!(PLANS[G])
    end
    nd = N - 2
    L1 = G.CL[1, 1]
    rb = L1.RB
    k1, k2 = size(rb.W1.data, 1), size(rb.W2.data, 1)
    lo, hi = sigmoid_bounds(L1.activation)
    desc = GlowDesc(nd, dims[1], dims[2], nd == 3 ? dims[3] : 1,
                    dims[N-1], n_cond, size(rb.W2.data, N), G.L, G.K, dims[N],
                    G.split_scales, G isa NetworkGlow ? G.logdet : true,
                    k1, k2, (k1 - 1) ÷ 2, (k2 - 1) ÷ 2, lo, hi, L1.C.freeze, PRECISION[])
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:inb_glow_plan_create, LIB), Cint, (Ref{GlowDesc}, Ref{Ptr{Cvoid}}), desc, p))
    n = ccall((:inb_glow_num_params, LIB), Cint, (Ptr{Cvoid},), p[])
    offs = zeros(Clonglong, n)
    total = Ref{Clonglong}(0)
    check(ccall((:inb_glow_flat_layout, LIB), Cint, (Ptr{Cvoid}, Ptr{Clonglong}, Ref{Clonglong}), p[], offs, total))
    ∇flat = CUDA.zeros(Float32, total[])
    st = PlanState(key, p[], Ptr{Cfloat}[], ∇flat, [dptr(∇flat) + 4 * o for o in offs], Any[], Dict{Symbol,Any}[Dict{Symbol,Any}(), Dict{Symbol,Any}()], 0, C_NULL)
    st.∇views = Any[offs[i] for i in 1:n]
    COMM[] != C_NULL && (check(ccall((:inb_glow_plan_set_comm, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), p[], COMM[])); st.comm = COMM[])
    PLANS[G] = st
    return st
end

# the parameter table is rebuilt only when a parameter array was replaced (set_params! rebinds p.data)
function param_table!(st::PlanState, ps::Vector{Parameter})
    if length(st.θ) != length(ps) || any(i -> st.θ[i] != dptr(ps[i].data), 1:length(ps))
        st.θ = ptr_table(ps)
    end
    return st.θ
end
# p.grad of parameter i = a view of the flat gradient buffer with the parameter's shape
grad_view(st::PlanState, i::Int, p::Parameter) =
    unsafe_wrap(CuArray, pointer(st.∇flat) + 4 * st.∇views[i], size(p.data))
# output ring: buffer `name` of the current turn with the requested size
function outbuf(st::PlanState, name::Symbol, dims)
    d = st.out[st.turn % 2 + 1]
    (haskey(d, name) && size(d[name]) == Tuple(dims)) || (d[name] = CUDA.zeros(Float32, dims...))
    return d[name]
end

# ActNorm parameters start as `nothing` (invertible_layer_actnorm.jl:53-57): allocate them; the library runs the
# data-dependent initialisation layer by layer inside its forward (init_actnorm = 1)
function alloc_actnorms!(ANs)
    for AN in ANs
        AN.s.data = CUDA.zeros(Float32, AN.k)
        AN.b.data = CUDA.zeros(Float32, AN.k)
    end
end

function fill_zdims!(G, plan, B)
    (G.split_scales && G.Z_dims !== nothing) || return
    dims = zeros(Cint, 5)
    for i in 1:length(G.Z_dims)
        n = ccall((:inb_glow_zdims, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cint}), plan, B, i - 1, dims)
        G.Z_dims[i] = collect(Int, reverse(dims[1:n]))        # (nx, ny[, nz], C, B), invertible_network_glow.jl:123
    end
end

# The library WRITES gradients into the flat buffer; the reference's rules are applied here:
#   ActNorm / ResidualBlock grads are overwritten (invertible_layer_actnorm.jl:113-114,
#   layer_residual_block.jl:168-172), Conv1x1 grads accumulate unless cleared (conv1x1.jl:237-239).
hh_params(G) = (hh = Set{Parameter}(); for L in G.CL; push!(hh, L.C.v1, L.C.v2, L.C.v3); end; hh)
function saved_hh_grads(G, ps)
    hh = hh_params(G)
    return Dict{Int,Any}(i => copy(p.grad) for (i, p) in enumerate(ps) if p in hh && p.grad !== nothing)
end
function assign_grads!(st::PlanState, ps, old::Dict{Int,Any})
    for (i, p) in enumerate(ps)
        p.grad = grad_view(st, i, p)
        haskey(old, i) && (p.grad .+= old[i])
    end
end

# ---------------------------------------------------------------- NetworkGlow
# replaces src/networks/invertible_network_glow.jl:109-129
function forward(X::CuArray{Float32,N}, G::NetworkGlow) where N
    (on_b200(G) && N in (4, 5)) || return invoke(forward, Tuple{AbstractArray{Float32,N},NetworkGlow}, X, G)
    init = actnorm_state(all_actnorms(G)) == :uninit
    init && alloc_actnorms!(all_actnorms(G))
    st = plan_for(G, size(X), 0)
    st.turn += 1
    θ = param_table!(st, get_params(G))
    Z = G.split_scales ? outbuf(st, :Z, (length(X),)) : outbuf(st, :Z, size(X))
    ld = outbuf(st, :ld, (1,))
    check(ccall((:inb_glow_forward, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Cfloat}, Ptr{Cfloat}, Cint, Ptr{Cvoid}),
                st.plan, size(X, N), dptr(X), θ, dptr(Z), G.logdet ? dptr(ld) : Ptr{Cfloat}(C_NULL), init, stream()))
    fill_zdims!(G, st.plan, size(X, N))
    G.logdet ? (return Z, Array(ld)[1]) : (return Z)
end

# Z_dims is filled by forward (:123); the input shape follows from it
function input_shape(G::NetworkGlow, Z)
    G.split_scales || return size(Z)
    zd = G.Z_dims[1]
    nd = length(zd) - 2
    c_in = zd[end-1] * 2 ÷ (2^nd)
    return (2 .* zd[1:nd]..., c_in, zd[end])
end

# replaces :132-147
function inverse(Z::CuArray{Float32,N}, G::NetworkGlow) where N
    (on_b200(G) && actnorm_state(all_actnorms(G)) == :ready && length(input_shape(G, Z)) in (4, 5)) ||
        return invoke(inverse, Tuple{AbstractArray{Float32,N},NetworkGlow}, Z, G)
    Xshape = input_shape(G, Z)
    st = plan_for(G, Tuple(Xshape), 0)
    st.turn += 1
    X = outbuf(st, :X, Xshape)
    check(ccall((:inb_glow_inverse, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Cfloat}, Ptr{Cvoid}),
                st.plan, Xshape[end], dptr(Z), param_table!(st, get_params(G)), dptr(X), stream()))
    return X
end

# replaces :150-191 (set_grad = true)
function backward(ΔZ::CuArray{Float32,N}, Z::CuArray{Float32,N}, G::NetworkGlow; set_grad::Bool=true) where N
    (set_grad && on_b200(G) && actnorm_state(all_actnorms(G)) == :ready && length(input_shape(G, Z)) in (4, 5)) ||
        return invoke(backward, Tuple{AbstractArray{Float32,N},AbstractArray{Float32,N},NetworkGlow},
                      ΔZ, Z, G; set_grad=set_grad)          # Jacobian paths stay on the reference
    Xshape = input_shape(G, Z)
    st = plan_for(G, Tuple(Xshape), 0)
    st.turn += 1
    X, ΔX = outbuf(st, :X, Xshape), outbuf(st, :ΔX, Xshape)
    ps = get_params(G)
    old = saved_hh_grads(G, ps)
    check(ccall((:inb_glow_backward, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Ptr{Cfloat}}, Ptr{Cfloat},
                 Ptr{Cfloat}, Ptr{Cvoid}),
                st.plan, Xshape[end], dptr(ΔZ), dptr(Z), param_table!(st, ps), st.∇, dptr(ΔX), dptr(X), stream()))
    assign_grads!(st, ps, old)
    return ΔX, X
end

# ---------------------------------------------------------------- NetworkConditionalGlow
# replaces src/networks/invertible_network_conditional_glow.jl:107-130
function forward(X::CuArray{Float32,N}, C::CuArray{Float32,N}, G::NetworkConditionalGlow) where N
    (on_b200(G) && N in (4, 5)) || return invoke(forward, Tuple{AbstractArray{Float32,N},AbstractArray{Float32,N},NetworkConditionalGlow}, X, C, G)
    init = actnorm_state(all_actnorms(G)) == :uninit
    init && alloc_actnorms!(all_actnorms(G))
    st = plan_for(G, size(X), size(C, N - 1))
    st.turn += 1
    ZX = outbuf(st, :ZX, size(X))
    f = G.split_scales ? 2^G.L : 1
    nd = N - 2
    ZC = outbuf(st, :ZC, ((size(C)[1:nd] .÷ f)..., size(C, N - 1) * (G.split_scales ? (2^nd)^G.L : 1), size(C, N)))
    ld = outbuf(st, :ld, (1,))
    check(ccall((:inb_cglow_forward, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat},
                 Cint, Ptr{Cvoid}),
                st.plan, size(X, N), dptr(X), dptr(C), param_table!(st, get_params(G)), dptr(ZX), dptr(ZC), dptr(ld),
                init, stream()))
    fill_zdims!(G, st.plan, size(X, N))
    return ZX, ZC, Array(ld)[1]
end

cond_channels(ZC, G::NetworkConditionalGlow, N) = size(ZC, N - 1) ÷ (G.split_scales ? (2^(N - 2))^G.L : 1)

# replaces :133-148
function inverse(ZX::CuArray{Float32,N}, ZC::CuArray{Float32,N}, G::NetworkConditionalGlow) where N
    (on_b200(G) && N in (4, 5) && actnorm_state(all_actnorms(G)) == :ready) ||
        return invoke(inverse, Tuple{AbstractArray{Float32,N},AbstractArray{Float32,N},NetworkConditionalGlow}, ZX, ZC, G)
    st = plan_for(G, size(ZX), cond_channels(ZC, G, N))
    st.turn += 1
    X = outbuf(st, :X, size(ZX))
    check(ccall((:inb_cglow_inverse, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Cfloat}, Ptr{Cvoid}),
                st.plan, size(X, N), dptr(ZX), dptr(ZC), param_table!(st, get_params(G)), dptr(X), stream()))
    return X
end

# replaces :151-181
function backward(ΔZX::CuArray{Float32,N}, ZX::CuArray{Float32,N}, ZC::CuArray{Float32,N},
                  G::NetworkConditionalGlow) where N
    (on_b200(G) && N in (4, 5) && actnorm_state(all_actnorms(G)) == :ready) ||
        return invoke(backward, Tuple{AbstractArray{Float32,N},AbstractArray{Float32,N},AbstractArray{Float32,N},NetworkConditionalGlow},
                      ΔZX, ZX, ZC, G)
    n_cond = cond_channels(ZC, G, N)
    st = plan_for(G, size(ZX), n_cond)
    st.turn += 1
    X, ΔX = outbuf(st, :X, size(ZX)), outbuf(st, :ΔX, size(ZX))
    ΔC = outbuf(st, :ΔC, (size(ZX)[1:N-2]..., n_cond, size(ZX, N)))
    ps = get_params(G)
    old = saved_hh_grads(G, ps)
    check(ccall((:inb_cglow_backward, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Ptr{Cfloat}},
                 Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                st.plan, size(X, N), dptr(ΔZX), dptr(ZX), dptr(ZC), param_table!(st, ps), st.∇, dptr(ΔX), dptr(X),
                dptr(ΔC), stream()))
    assign_grads!(st, ps, old)
    return ΔX, X, ΔC
end

# ---------------------------------------------------------------- layers (one ccall each; outputs are fresh arrays)
geom(X::CuArray{Float32,N}) where N = (Cint(N - 2), Cint(size(X, 1)), Cint(size(X, 2)), Cint(N == 5 ? size(X, 3) : 1))
bcs(X::CuArray{Float32,N}) where N = (Cint(size(X, N)), Cint(size(X, N - 1)), Clonglong(prod(size(X)[1:N-2])))

# replaces src/layers/invertible_layer_actnorm.jl:60-77
function forward(X::CuArray{Float32,N}, AN::ActNorm; logdet=nothing) where N
    (AN.is_reversed || !(size(X, N - 1) in SUPPORTED_C)) &&
        return invoke(forward, Tuple{AbstractArray{Float32,N},ActNorm}, X, AN; logdet=logdet)
    isnothing(logdet) ? logdet = (AN.logdet && ~AN.is_reversed) : logdet = logdet
    B, C, sp = bcs(X)
    if AN.s.data === nothing
        AN.s.data = CUDA.zeros(Float32, C); AN.b.data = CUDA.zeros(Float32, C)
        check(ccall((:inb_actnorm_init, LIB), Cint, (Cint, Cint, Clonglong, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                    B, C, sp, dptr(X), dptr(AN.s.data), dptr(AN.b.data), stream()))
    end
    Y = similar(X); ld = CUDA.zeros(Float32, 1)
    check(ccall((:inb_actnorm_forward, LIB), Cint,
                (Cint, Cint, Clonglong, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                B, C, sp, dptr(X), dptr(AN.s.data), dptr(AN.b.data), dptr(Y), logdet ? dptr(ld) : Ptr{Cfloat}(C_NULL), stream()))
    logdet ? (return Y, Array(ld)[1]) : (return Y)
end

# replaces :80-97 (the logdet = true variant returns -logdet of the forward: left to the reference)
function inverse(Y::CuArray{Float32,N}, AN::ActNorm; logdet=nothing) where N
    isnothing(logdet) ? logdet = (AN.logdet && AN.is_reversed) : logdet = logdet
    (logdet || AN.is_reversed || AN.s.data === nothing || !(size(Y, N - 1) in SUPPORTED_C)) &&
        return invoke(inverse, Tuple{AbstractArray{Float32,N},ActNorm}, Y, AN; logdet=logdet)
    B, C, sp = bcs(Y)
    X = similar(Y)
    check(ccall((:inb_actnorm_inverse, LIB), Cint,
                (Cint, Cint, Clonglong, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                B, C, sp, dptr(Y), dptr(AN.s.data), dptr(AN.b.data), dptr(X), stream()))
    return X
end

# replaces :100-123 (set_grad = true: grads overwritten, :113-114)
function backward(ΔY::CuArray{Float32,N}, Y::CuArray{Float32,N}, AN::ActNorm; set_grad::Bool=true) where N
    (set_grad && !AN.is_reversed && size(Y, N - 1) in SUPPORTED_C) ||
        return invoke(backward, Tuple{AbstractArray{Float32,N},AbstractArray{Float32,N},ActNorm}, ΔY, Y, AN; set_grad=set_grad)
    B, C, sp = bcs(Y)
    ΔX, X = similar(Y), similar(Y)
    Δs, Δb = CUDA.zeros(Float32, C), CUDA.zeros(Float32, C)
    check(ccall((:inb_actnorm_backward, LIB), Cint,
                (Cint, Cint, Clonglong, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Cint, Ptr{Cfloat}, Ptr{Cfloat},
                 Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                B, C, sp, dptr(ΔY), dptr(Y), dptr(AN.s.data), dptr(AN.b.data), Cint(AN.logdet), dptr(ΔX), dptr(X), dptr(Δs),
                dptr(Δb), stream()))
    AN.s.grad = Δs
    AN.b.grad = Δb
    return ΔX, X
end

# replaces src/layers/invertible_layer_conv1x1.jl:174-189 / 209-224 (logdet of an orthogonal map is 0)
for (fn, sym) in ((:forward, :inb_conv1x1_forward), (:inverse, :inb_conv1x1_inverse))
    @eval function $fn(X::CuArray{Float32,N}, C::Conv1x1; logdet=nothing) where N
        C.k in SUPPORTED_C || return invoke($fn, Tuple{AbstractArray{Float32,N},Conv1x1}, X, C; logdet=logdet)
        isnothing(logdet) ? logdet = C.logdet : logdet = logdet
        Y = similar(X)
        B, k, sp = bcs(X)
        check(ccall(($(QuoteNode(sym)), LIB), Cint,
                    (Cint, Cint, Clonglong, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                    B, k, sp, dptr(X), dptr(C.v1.data), dptr(C.v2.data), dptr(C.v3.data), dptr(Y), stream()))
        logdet ? (return Y, 0) : (return Y)
    end
end

# replaces :227-245: ΔX, X = C.inverse((ΔY, Y)) with the gradients w.r.t. v1, v2, v3 (accumulated unless cleared, :237-239)
function inverse(Y_tuple::Tuple{CuArray{Float32,N},CuArray{Float32,N}}, C::Conv1x1; set_grad::Bool=true) where N
    (set_grad && C.k in SUPPORTED_C) || return invoke(inverse, Tuple{Tuple,Conv1x1}, Y_tuple, C; set_grad=set_grad)
    ΔY, Y = Y_tuple
    B, k, sp = bcs(Y)
    ΔX, X = similar(Y), similar(Y)
    Δv = [CUDA.zeros(Float32, k) for _ in 1:3]
    check(ccall((:inb_conv1x1_backward, LIB), Cint,
                (Cint, Cint, Clonglong, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Cint, Ptr{Cfloat},
                 Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                B, k, sp, dptr(ΔY), dptr(Y), dptr(C.v1.data), dptr(C.v2.data), dptr(C.v3.data), Cint(C.freeze), dptr(ΔX),
                dptr(X), dptr(Δv[1]), dptr(Δv[2]), dptr(Δv[3]), stream()))
    for (p, g) in zip((C.v1, C.v2, C.v3), Δv)
        p.grad = p.grad === nothing ? g : p.grad .+ g
    end
    return ΔX, X
end

# ResidualBlock (src/layers/layer_residual_block.jl:119-178); weights are passed as the reference's arrays, unmodified
rb_ints(X::CuArray{Float32,N}, RB::ResidualBlock) where N =
    (geom(X)..., Cint(size(X, N)), Cint(size(RB.W1.data, N - 1)), Cint(size(RB.W2.data, N)), Cint(size(RB.W3.data, N - 1)),
     Cint(size(RB.W1.data, 1)), Cint(size(RB.W2.data, 1)), PRECISION[])
const RB_ARGT = (Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint)

# replaces :119-134
function forward(X1::CuArray{Float32,N}, RB::ResidualBlock; save=false) where N
    (rb_on_b200(RB) && N in (4, 5) && !save) || return invoke(forward, Tuple{AbstractArray{Float32,N},ResidualBlock}, X1, RB; save=save)
    Y = CUDA.zeros(Float32, size(X1)[1:N-2]..., size(RB.W3.data, N - 1), size(X1, N))
    check(ccall((:inb_resblock_forward, LIB), Cint,
                (RB_ARGT..., Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                rb_ints(X1, RB)..., dptr(X1), dptr(RB.W1.data), dptr(RB.W2.data), dptr(RB.W3.data), dptr(RB.b1.data),
                dptr(RB.b2.data), dptr(Y), stream()))
    return Y
end

# replaces :137-178 (set_grad = true: grads overwritten, :168-172)
function backward(ΔX4::CuArray{Float32,N}, X1::CuArray{Float32,N}, RB::ResidualBlock; set_grad::Bool=true) where N
    (rb_on_b200(RB) && N in (4, 5) && set_grad) ||
        return invoke(backward, Tuple{AbstractArray{Float32,N},AbstractArray{Float32,N},ResidualBlock}, ΔX4, X1, RB; set_grad=set_grad)
    ΔX1 = similar(X1)
    ps = (RB.W1, RB.W2, RB.W3, RB.b1, RB.b2)
    g = [CUDA.zeros(Float32, size(p.data)) for p in ps]
    check(ccall((:inb_resblock_backward, LIB), Cint,
                (RB_ARGT..., Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat},
                 Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                rb_ints(X1, RB)..., dptr(ΔX4), dptr(X1), dptr(RB.W1.data), dptr(RB.W2.data), dptr(RB.W3.data),
                dptr(RB.b1.data), dptr(RB.b2.data), dptr(ΔX1), dptr(g[1]), dptr(g[2]), dptr(g[3]), dptr(g[4]), dptr(g[5]),
                stream()))
    for (p, gi) in zip(ps, g)
        p.grad = gi
    end
    return ΔX1
end

# CouplingLayerGlow (src/layers/invertible_layer_glow.jl:104-170) and ConditionalLayerGlow
# (src/conditional_layers/conditional_layer_glow.jl:94-158): one library call per direction
const GlowLayer = Union{CouplingLayerGlow,ConditionalLayerGlow}
cl_params(L::GlowLayer) = Parameter[L.C.v1, L.C.v2, L.C.v3, L.RB.W1, L.RB.W2, L.RB.W3, L.RB.b1, L.RB.b2]
function cl_ints(X::CuArray{Float32,N}, n_cond::Int, L::GlowLayer) where N
    lo, hi = sigmoid_bounds(L.activation)
    return (geom(X)..., Cint(size(X, N)), Cint(size(X, N - 1)), Cint(n_cond), Cint(size(L.RB.W2.data, N)),
            Cint(size(L.RB.W1.data, 1)), Cint(size(L.RB.W2.data, 1)), Cfloat(lo), Cfloat(hi))
end
const CL_ARGT = (Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cfloat, Cfloat)

function cl_forward(X::CuArray{Float32,N}, C, L::GlowLayer) where N
    Y = similar(X); ld = CUDA.zeros(Float32, 1)
    check(ccall((:inb_coupling_forward, LIB), Cint,
                (CL_ARGT..., Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                cl_ints(X, C === nothing ? 0 : size(C, N - 1), L)..., PRECISION[], dptr(X), dptr(C), ptr_table(cl_params(L)),
                dptr(Y), L.logdet ? dptr(ld) : Ptr{Cfloat}(C_NULL), stream()))
    L.logdet ? (return Y, Array(ld)[1]) : (return Y)
end
function cl_inverse(Y::CuArray{Float32,N}, C, L::GlowLayer) where N
    X = similar(Y)
    check(ccall((:inb_coupling_inverse, LIB), Cint,
                (CL_ARGT..., Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Cfloat}, Ptr{Cvoid}),
                cl_ints(Y, C === nothing ? 0 : size(C, N - 1), L)..., PRECISION[], dptr(Y), dptr(C), ptr_table(cl_params(L)),
                dptr(X), stream()))
    return X
end
function cl_backward(ΔY::CuArray{Float32,N}, Y::CuArray{Float32,N}, C, L::GlowLayer) where N
    ΔX, X = similar(Y), similar(Y)
    ΔC = C === nothing ? nothing : similar(C)
    ps = cl_params(L)
    g = [CUDA.zeros(Float32, size(p.data)) for p in ps]
    check(ccall((:inb_coupling_backward, LIB), Cint,
                (CL_ARGT..., Cint, Cint, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Ptr{Cfloat}},
                 Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                cl_ints(Y, C === nothing ? 0 : size(C, N - 1), L)..., Cint(L.logdet), Cint(L.C.freeze), PRECISION[], dptr(ΔY),
                dptr(Y), dptr(C), ptr_table(ps), Ptr{Cfloat}[dptr(gi) for gi in g], dptr(ΔX), dptr(X), dptr(ΔC), stream()))
    for (i, (p, gi)) in enumerate(zip(ps, g))   # conv1x1.jl:237-239 for v1..v3, overwrite for the block
        p.grad = (i <= 3 && p.grad !== nothing) ? p.grad .+ gi : gi
    end
    return ΔX, X, ΔC
end

# replaces invertible_layer_glow.jl:104-117
function forward(X::CuArray{Float32,N}, L::CouplingLayerGlow) where N
    (layer_on_b200(L) && N in (4, 5)) || return invoke(forward, Tuple{AbstractArray{Float32,N},CouplingLayerGlow}, X, L)
    return cl_forward(X, nothing, L)
end
# replaces :120-133 (save = true hands the intermediates to the reference's own backward: left there)
function inverse(Y::CuArray{Float32,N}, L::CouplingLayerGlow; save=false) where N
    (layer_on_b200(L) && N in (4, 5) && !save) || return invoke(inverse, Tuple{AbstractArray{Float32,N},CouplingLayerGlow}, Y, L; save=save)
    return cl_inverse(Y, nothing, L)
end
# replaces :136-170 (set_grad = true)
function backward(ΔY::CuArray{Float32,N}, Y::CuArray{Float32,N}, L::CouplingLayerGlow; set_grad::Bool=true) where N
    (layer_on_b200(L) && N in (4, 5) && set_grad) ||
        return invoke(backward, Tuple{AbstractArray{Float32,N},AbstractArray{Float32,N},CouplingLayerGlow}, ΔY, Y, L; set_grad=set_grad)
    ΔX, X, _ = cl_backward(ΔY, Y, nothing, L)
    return ΔX, X
end
# replaces conditional_layer_glow.jl:94-112
function forward(X::CuArray{Float32,N}, C::CuArray{Float32,N}, L::ConditionalLayerGlow) where N
    (layer_on_b200(L) && N in (4, 5)) || return invoke(forward, Tuple{AbstractArray{Float32,N},AbstractArray{Float32,N},ConditionalLayerGlow}, X, C, L)
    return cl_forward(X, C, L)
end
# replaces :115-131
function inverse(Y::CuArray{Float32,N}, C::CuArray{Float32,N}, L::ConditionalLayerGlow; save=false) where N
    (layer_on_b200(L) && N in (4, 5) && !save) ||
        return invoke(inverse, Tuple{AbstractArray{Float32,N},AbstractArray{Float32,N},ConditionalLayerGlow}, Y, C, L; save=save)
    return cl_inverse(Y, C, L)
end
# replaces :134-158
function backward(ΔY::CuArray{Float32,N}, Y::CuArray{Float32,N}, C::CuArray{Float32,N}, L::ConditionalLayerGlow) where N
    (layer_on_b200(L) && N in (4, 5)) ||
        return invoke(backward, Tuple{AbstractArray{Float32,N},AbstractArray{Float32,N},AbstractArray{Float32,N},ConditionalLayerGlow}, ΔY, Y, C, L)
    return cl_backward(ΔY, Y, C, L)
end

# replaces src/utils/dimensionality_operations.jl:79-107 / 137-166 (checkerboard pattern; the others stay on the reference)
function squeeze(X::CuArray{Float32,N}; pattern="column") where N
    pattern == "checkerboard" || return invoke(squeeze, Tuple{AbstractArray{Float32,N}}, X; pattern=pattern)
    nd = N - 2
    Y = CUDA.zeros(Float32, (size(X)[1:nd] .÷ 2)..., size(X, N - 1) * 2^nd, size(X, N))
    check(ccall((:inb_squeeze, LIB), Cint, (Cint, Cint, Cint, Cint, Cint, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                geom(X)..., size(X, N), size(X, N - 1), dptr(X), dptr(Y), stream()))
    return Y
end
function unsqueeze(Y::CuArray{Float32,N}; pattern="column") where N
    pattern == "checkerboard" || return invoke(unsqueeze, Tuple{AbstractArray{Float32,N}}, Y; pattern=pattern)
    nd = N - 2
    X = CUDA.zeros(Float32, (size(Y)[1:nd] .* 2)..., size(Y, N - 1) ÷ 2^nd, size(Y, N))
    check(ccall((:inb_unsqueeze, LIB), Cint, (Cint, Cint, Cint, Cint, Cint, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                geom(Y)..., size(Y, N), size(Y, N - 1), dptr(Y), dptr(X), stream()))
    return X
end

# ---------------------------------------------------------------- data-parallel training (SURVEY 8e)
# One Julia process per GPU (MPI.jl or Distributed), the batch sharded along its last dimension.  Rank 0 draws the
# NCCL id, the caller ships its 128 bytes (`MPI.Bcast!(id, 0, comm)`), every rank joins:
#     id = rank == 0 ? InvertibleNetworksB200.unique_id() : zeros(UInt8, 128);  MPI.Bcast!(id, 0, MPI.COMM_WORLD)
#     InvertibleNetworksB200.attach_comm!(InvertibleNetworksB200.create_comm(nranks, rank, id))
# From then on the first G.forward(X) initialises ActNorm from the GLOBAL batch and G.backward averages the gradients
# over the ranks (per scale, overlapped with the rest of the backward pass).
function unique_id()
    id = zeros(UInt8, 128)
    check(ccall((:inb_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id))
    return id
end
function create_comm(nranks::Integer, rank::Integer, id::Vector{UInt8})
    c = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:inb_comm_create, LIB), Cint, (Cint, Cint, Ptr{UInt8}, Ref{Ptr{Cvoid}}), nranks, rank, id, c))
    return c[]
end
# an existing NCCL.jl communicator: wrap_comm(comm.handle)
function wrap_comm(nccl_comm::Ptr{Cvoid})
    c = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:inb_comm_wrap, LIB), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), nccl_comm, c))
    return c[]
end
function attach_comm!(comm::Ptr{Cvoid})
    COMM[] = comm
    for st in values(PLANS)
        check(ccall((:inb_glow_plan_set_comm, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), st.plan, comm))
        st.comm = comm
    end
end
# explicit average of the gradients of G's last backward (callers that do not attach)
function allreduce_grads!(G::Union{NetworkGlow,NetworkConditionalGlow}, comm::Ptr{Cvoid})
    st = PLANS[G]
    check(ccall((:inb_allreduce_grads, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cfloat}}, Ptr{Cvoid}, Ptr{Cvoid}), st.plan, st.∇, comm, stream()))
end
function broadcast_params!(G::Union{NetworkGlow,NetworkConditionalGlow}, comm::Ptr{Cvoid}; root::Integer=0)
    st = PLANS[G]
    check(ccall((:inb_broadcast_params, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cfloat}}, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
                st.plan, param_table!(st, get_params(G)), comm, root, stream()))
end

# ---------------------------------------------------------------- HINT family (SURVEY 8f rank 3)
# mirrors `inb_hint_desc`
struct HintDesc
    nx::Cint; ny::Cint; n_in::Cint; n_hidden::Cint; L::Cint; K::Cint; batch::Cint; split_scales::Cint
    k1::Cint; k2::Cint; p1::Cint; p2::Cint
    sig_low::Cfloat; sig_high::Cfloat
    squeeze_type::Cint; shared_grads::Cint; precision::Cint
end
# 0 = sum the gradients of coupling layers the recursion visits several times (the true gradient, what the
# reference's set_grad=false path returns, invertible_layer_hint.jl:222); 1 = keep the last visit only (what its
# set_grad=true path leaves in .grad, layer_residual_block.jl:168-172)
const HINT_SHARED_GRADS = Ref{Cint}(parse(Cint, get(ENV, "INB200_HINT_SHARED_GRADS", "0")))
const HINT_PLANS = IdDict{Any,Tuple{Any,Ptr{Cvoid}}}()
# the fused chain (and with it fp16x3) needs k2 = 1; the reference's HINT default is k2 = 3: those run in bf16x3
hint_precision(k2) = (PRECISION[] == 3 && k2 != 1) ? Cint(1) : PRECISION[]

basic_on_b200(L::CouplingLayerBasic) = rb_on_b200(L.RB) && !L.is_reversed && sigmoid_bounds(L.activation) !== nothing
hint_on_b200(H::CouplingLayerHINT) = !H.is_reversed && haskey(PERMUTE, H.permute) && all(basic_on_b200, H.CL) &&
                                     (H.C === nothing || H.C.k in SUPPORTED_C)
hint_net_on_b200(H::NetworkMultiScaleHINT) = all(hint_on_b200, H.CL) && all(AN -> !AN.is_reversed, H.AN) &&
                                             actnorm_state(all_actnorms(H)) != :mixed

function hint_plan_for(H::NetworkMultiScaleHINT, X::CuArray{Float32,4})
    key = (size(X, 1), size(X, 2), size(X, 4), PRECISION[])
    haskey(HINT_PLANS, H) && HINT_PLANS[H][1] == key && return HINT_PLANS[H][2]
    haskey(HINT_PLANS, H) && ccall((:inb_hint_plan_destroy, LIB), Cint, (Ptr{Cvoid},), HINT_PLANS[H][2])
    L1 = H.CL[1, 1].CL[1]
    rb = L1.RB
    k1, k2 = size(rb.W1.data, 1), size(rb.W2.data, 1)
    lo, hi = sigmoid_bounds(L1.activation)
    desc = HintDesc(size(X, 1), size(X, 2), size(X, 3), size(rb.W2.data, 4), H.L, H.K, size(X, 4), H.split_scales,
                    k1, k2, (k1 - 1) ÷ 2, (k2 - 1) ÷ 2, lo, hi, 0, HINT_SHARED_GRADS[], hint_precision(k2))
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:inb_hint_plan_create, LIB), Cint, (Ref{HintDesc}, Ref{Ptr{Cvoid}}), desc, p))
    HINT_PLANS[H] = (key, p[])
    return p[]
end

# replaces src/networks/invertible_network_hint_multiscale.jl:98-117
function forward(X::CuArray{Float32,4}, H::NetworkMultiScaleHINT)
    hint_net_on_b200(H) || return invoke(forward, Tuple{AbstractArray{Float32,4},NetworkMultiScaleHINT}, X, H)
    init = actnorm_state(all_actnorms(H)) == :uninit
    init && alloc_actnorms!(all_actnorms(H))
    plan = hint_plan_for(H, X)
    f = 2^H.L
    Z = H.split_scales ? CUDA.zeros(Float32, length(X)) :
        CUDA.zeros(Float32, size(X, 1) ÷ f, size(X, 2) ÷ f, size(X, 3) * 4^H.L, size(X, 4))
    ld = CUDA.zeros(Float32, 1)
    check(ccall((:inb_hint_forward, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Cfloat}, Ptr{Cfloat}, Cint, Ptr{Cvoid}),
                plan, size(X, 4), dptr(X), ptr_table(get_params(H)), dptr(Z), dptr(ld), init, stream()))
    HINT_INPUT[H] = size(X)   # stands in for the input size the reference keeps in its Z_dims bookkeeping (:111)
    return Z, Array(ld)[1]
end
const HINT_INPUT = IdDict{Any,Any}()
hint_input_shape(H, Z) = haskey(HINT_INPUT, H) && prod(HINT_INPUT[H]) == length(Z) ? HINT_INPUT[H] :
    (size(Z, 1) * 2^H.L, size(Z, 2) * 2^H.L, size(Z, 3) ÷ 4^H.L, size(Z, 4))

# replaces :120-133
function inverse(Z::CuArray{Float32,N}, H::NetworkMultiScaleHINT) where N
    (hint_net_on_b200(H) && actnorm_state(all_actnorms(H)) == :ready) ||
        return invoke(inverse, Tuple{AbstractArray{Float32,N},NetworkMultiScaleHINT}, Z, H)
    X = CUDA.zeros(Float32, hint_input_shape(H, Z)...)
    plan = hint_plan_for(H, X)
    check(ccall((:inb_hint_inverse, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Cfloat}, Ptr{Cvoid}),
                plan, size(X, 4), dptr(Z), ptr_table(get_params(H)), dptr(X), stream()))
    return X
end

# replaces :136-174 (set_grad = true; the Jacobian paths stay on the reference)
function backward(ΔZ::CuArray{Float32,N}, Z::CuArray{Float32,N}, H::NetworkMultiScaleHINT; set_grad::Bool=true) where N
    (set_grad && hint_net_on_b200(H) && actnorm_state(all_actnorms(H)) == :ready) ||
        return invoke(backward, Tuple{AbstractArray{Float32,N},AbstractArray{Float32,N},NetworkMultiScaleHINT},
                      ΔZ, Z, H; set_grad=set_grad)
    shape = hint_input_shape(H, Z)
    X, ΔX = CUDA.zeros(Float32, shape...), CUDA.zeros(Float32, shape...)
    plan = hint_plan_for(H, X)
    ps = get_params(H)
    fresh = [CUDA.zeros(Float32, size(p.data)) for p in ps]
    check(ccall((:inb_hint_backward, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Ptr{Cfloat}}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                plan, shape[end], dptr(ΔZ), dptr(Z), ptr_table(ps), Ptr{Cfloat}[dptr(g) for g in fresh],
                dptr(ΔX), dptr(X), stream()))
    hh = Set{Parameter}()
    for HL in H.CL   # the Conv1x1 of every CouplingLayerHINT accumulates (conv1x1.jl:237-239)
        HL.C === nothing || push!(hh, HL.C.v1, HL.C.v2, HL.C.v3)
    end
    for (p, g) in zip(ps, fresh)
        p.grad = (p in hh && p.grad !== nothing) ? p.grad .+ g : g
    end
    return ΔX, X
end

# ---- layer level: CouplingLayerHINT (src/layers/invertible_layer_hint.jl:105-297) and CouplingLayerBasic
# (src/layers/invertible_layer_basic.jl:90-149).  Reversed layers and set_grad = false fall through.
const PERMUTE = Dict("none" => Cint(0), "full" => Cint(1), "lower" => Cint(2), "both" => Cint(3))
function hint_ints(X::CuArray{Float32,N}, H::CouplingLayerHINT) where N
    rb = H.CL[1].RB
    lo, hi = sigmoid_bounds(H.CL[1].activation)
    return (geom(X)..., Cint(size(X, N)), Cint(size(X, N - 1)), Cint(size(rb.W2.data, N)), Cint(size(rb.W1.data, 1)),
            Cint(size(rb.W2.data, 1)), Cfloat(lo), Cfloat(hi), PERMUTE[H.permute])
end
const HINT_ARGT = (Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cfloat, Cfloat, Cint)
hint_layer_precision(H::CouplingLayerHINT) = hint_precision(size(H.CL[1].RB.W2.data, 1))

# replaces :105-156 (scale = 1 entry; the recursion runs inside the library)
function forward(X::CuArray{Float32,N}, H::CouplingLayerHINT; scale=1, permute=nothing, logdet=nothing) where N
    (hint_on_b200(H) && scale == 1 && permute === nothing) ||
        return invoke(forward, Tuple{AbstractArray{Float32,N},CouplingLayerHINT}, X, H; scale=scale, permute=permute, logdet=logdet)
    logdet = logdet === nothing ? H.logdet : logdet
    Y = similar(X)
    ld = CUDA.zeros(Float32, 1)
    check(ccall((:inb_hint_coupling_forward, LIB), Cint,
                (HINT_ARGT..., Cint, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                hint_ints(X, H)..., hint_layer_precision(H), dptr(X), ptr_table(get_params(H)), dptr(Y),
                logdet ? dptr(ld) : Ptr{Cfloat}(C_NULL), stream()))
    logdet ? (return Y, Array(ld)[1]) : (return Y)
end

# replaces :159-204
function inverse(Y::CuArray{Float32,N}, H::CouplingLayerHINT; scale=1, permute=nothing, logdet=nothing) where N
    (hint_on_b200(H) && scale == 1 && permute === nothing && logdet !== true) ||
        return invoke(inverse, Tuple{AbstractArray{Float32,N},CouplingLayerHINT}, Y, H; scale=scale, permute=permute, logdet=logdet)
    X = similar(Y)
    check(ccall((:inb_hint_coupling_inverse, LIB), Cint,
                (HINT_ARGT..., Cint, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Cfloat}, Ptr{Cvoid}),
                hint_ints(Y, H)..., hint_layer_precision(H), dptr(Y), ptr_table(get_params(H)), dptr(X), stream()))
    return X
end

# replaces :207-297 (set_grad = true)
function backward(ΔY::CuArray{Float32,N}, Y::CuArray{Float32,N}, H::CouplingLayerHINT; scale=1, permute=nothing,
                  set_grad::Bool=true) where N
    (hint_on_b200(H) && scale == 1 && permute === nothing && set_grad) ||
        return invoke(backward, Tuple{AbstractArray{Float32,N},AbstractArray{Float32,N},CouplingLayerHINT}, ΔY, Y, H;
                      scale=scale, permute=permute, set_grad=set_grad)
    ΔX, X = similar(Y), similar(Y)
    ps = get_params(H)
    fresh = [CUDA.zeros(Float32, size(p.data)) for p in ps]
    check(ccall((:inb_hint_coupling_backward, LIB), Cint,
                (HINT_ARGT..., Cint, Cint, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Ptr{Cfloat}}, Ptr{Cfloat},
                 Ptr{Cfloat}, Ptr{Cvoid}),
                hint_ints(Y, H)..., Cint(H.logdet), HINT_SHARED_GRADS[], hint_layer_precision(H), dptr(ΔY), dptr(Y),
                ptr_table(ps), Ptr{Cfloat}[dptr(g) for g in fresh], dptr(ΔX), dptr(X), stream()))
    hh = H.C === nothing ? Set{Parameter}() : Set{Parameter}((H.C.v1, H.C.v2, H.C.v3))
    for (p, g) in zip(ps, fresh)   # Conv1x1 gradients accumulate unless cleared (conv1x1.jl:237-239)
        p.grad = (p in hh && p.grad !== nothing) ? p.grad .+ g : g
    end
    return ΔX, X
end

# CouplingLayerBasic: forward :90-105, inverse :108-121, backward :124-149 (non-reversed, set_grad = true, save = false)
function basic_ints(X1::CuArray{Float32,N}, L::CouplingLayerBasic) where N
    rb = L.RB
    lo, hi = sigmoid_bounds(L.activation)
    return (geom(X1)..., Cint(size(X1, N)), Cint(size(X1, N - 1)), Cint(size(rb.W2.data, N)), Cint(size(rb.W1.data, 1)),
            Cint(size(rb.W2.data, 1)), Cfloat(lo), Cfloat(hi))
end
const BASIC_ARGT = (Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cfloat, Cfloat)
basic_precision(L::CouplingLayerBasic) = hint_precision(size(L.RB.W2.data, 1))
function forward(X1::CuArray{Float32,N}, X2::CuArray{Float32,N}, L::CouplingLayerBasic; save::Bool=false, logdet=nothing) where N
    (basic_on_b200(L) && !save) ||
        return invoke(forward, Tuple{AbstractArray{Float32,N},AbstractArray{Float32,N},CouplingLayerBasic}, X1, X2, L; save=save, logdet=logdet)
    logdet = logdet === nothing ? L.logdet : logdet
    Y2 = similar(X2)
    ld = CUDA.zeros(Float32, 1)
    check(ccall((:inb_basic_coupling_forward, LIB), Cint,
                (BASIC_ARGT..., Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                basic_ints(X1, L)..., basic_precision(L), dptr(X1), dptr(X2), ptr_table(get_params(L)), dptr(Y2),
                logdet ? dptr(ld) : Ptr{Cfloat}(C_NULL), stream()))
    logdet ? (return X1, Y2, Array(ld)[1]) : (return X1, Y2)
end
function inverse(Y1::CuArray{Float32,N}, Y2::CuArray{Float32,N}, L::CouplingLayerBasic; save::Bool=false, logdet=nothing) where N
    (basic_on_b200(L) && !save && logdet !== true) ||
        return invoke(inverse, Tuple{AbstractArray{Float32,N},AbstractArray{Float32,N},CouplingLayerBasic}, Y1, Y2, L; save=save, logdet=logdet)
    X2 = similar(Y2)
    check(ccall((:inb_basic_coupling_inverse, LIB), Cint,
                (BASIC_ARGT..., Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Cfloat}, Ptr{Cvoid}),
                basic_ints(Y1, L)..., basic_precision(L), dptr(Y1), dptr(Y2), ptr_table(get_params(L)), dptr(X2), stream()))
    return Y1, X2
end
function backward(ΔY1::CuArray{Float32,N}, ΔY2::CuArray{Float32,N}, Y1::CuArray{Float32,N}, Y2::CuArray{Float32,N},
                  L::CouplingLayerBasic; set_grad::Bool=true) where N
    (basic_on_b200(L) && set_grad) ||
        return invoke(backward, Tuple{AbstractArray{Float32,N},AbstractArray{Float32,N},AbstractArray{Float32,N},
                                      AbstractArray{Float32,N},CouplingLayerBasic}, ΔY1, ΔY2, Y1, Y2, L; set_grad=set_grad)
    ΔX1, ΔX2, X2 = similar(Y1), similar(Y2), similar(Y2)
    ps = get_params(L)
    fresh = [CUDA.zeros(Float32, size(p.data)) for p in ps]
    check(ccall((:inb_basic_coupling_backward, LIB), Cint,
                (BASIC_ARGT..., Cint, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Ptr{Cfloat}},
                 Ptr{Ptr{Cfloat}}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                basic_ints(Y1, L)..., Cint(L.logdet), basic_precision(L), dptr(ΔY1), dptr(ΔY2), dptr(Y1), dptr(Y2),
                ptr_table(ps), Ptr{Cfloat}[dptr(g) for g in fresh], dptr(ΔX1), dptr(ΔX2), dptr(X2), stream()))
    for (p, g) in zip(ps, fresh)   # layer_residual_block.jl:168-172: overwritten
        p.grad = g
    end
    return ΔX1, ΔX2, Y1, X2
end

# replaces wavelet_squeeze / wavelet_unsqueeze with the default type = WT.db1 and Haar_squeeze / invHaar_unsqueeze
# (src/utils/dimensionality_operations.jl:199-258, 318-371), 4-D tensors
function haar_call(sym::Symbol, X::CuArray{Float32,4}, ty::Integer, up::Bool)
    Y = up ? CUDA.zeros(Float32, 2size(X, 1), 2size(X, 2), size(X, 3) ÷ 4, size(X, 4)) :
             CUDA.zeros(Float32, size(X, 1) ÷ 2, size(X, 2) ÷ 2, 4size(X, 3), size(X, 4))
    # (nx, ny, C) describe the INPUT of either direction (include/inb200.h)
    if sym === :inb_haar_squeeze
        check(ccall((:inb_haar_squeeze, LIB), Cint, (Cint, Cint, Cint, Cint, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                    size(X, 1), size(X, 2), size(X, 4), size(X, 3), ty, dptr(X), dptr(Y), stream()))
    else
        check(ccall((:inb_haar_unsqueeze, LIB), Cint, (Cint, Cint, Cint, Cint, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                    size(X, 1), size(X, 2), size(X, 4), size(X, 3), ty, dptr(X), dptr(Y), stream()))
    end
    return Y
end
Haar_squeeze(X::CuArray{Float32,4}) = haar_call(:inb_haar_squeeze, X, 1, false)
invHaar_unsqueeze(X::CuArray{Float32,4}) = haar_call(:inb_haar_unsqueeze, X, 1, true)
# the reference runs the wavelet transform on the CPU channel by channel (:199-216); other wavelet types stay there
function wavelet_squeeze(X::CuArray{Float32,4}; type=InvertibleNetworks.WT.db1)
    type == InvertibleNetworks.WT.db1 || return invoke(wavelet_squeeze, Tuple{AbstractArray{Float32,4}}, X; type=type)
    return haar_call(:inb_haar_squeeze, X, 0, false)
end
function wavelet_unsqueeze(X::CuArray{Float32,4}; type=InvertibleNetworks.WT.db1)
    type == InvertibleNetworks.WT.db1 || return invoke(wavelet_unsqueeze, Tuple{AbstractArray{Float32,4}}, X; type=type)
    return haar_call(:inb_haar_unsqueeze, X, 0, true)
end

# Optional replacement of the `for p in get_params(G); update!(opt, p.data, p.grad); end` loop with Flux.ADAM
# (examples/networks/network_glow.jl:38-42) for callers that keep parameters and gradients in flat CuArrays:
# one launch instead of 10*L*K.  m, v start at zero; t is the 1-based step count.
function adam_update!(θ::CuArray{Float32}, ∇θ::CuArray{Float32}, m::CuArray{Float32}, v::CuArray{Float32};
                      η=1f-3, β=(0.9f0, 0.999f0), ϵ=1f-8, t::Int=1)
    check(ccall((:inb_adam_update, LIB), Cint,
                (Clonglong, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Cfloat, Cfloat, Cfloat, Cfloat, Cfloat, Cfloat, Ptr{Cvoid}),
                length(θ), dptr(θ), dptr(∇θ), dptr(m), dptr(v), η, β[1], β[2], ϵ, β[1]^t, β[2]^t, stream()))
    return θ
end

end # module
