# InvertibleNetworksB200.jl - the reference-side binding of libinb200.so (include/inb200.h).
#
# `using InvertibleNetworks, InvertibleNetworksB200` adds methods that are strictly more specific than
# the reference's `AbstractArray` methods (CuArray{Float32,N} inputs), so existing user code
#     Z, lgdet = G.forward(X);  ΔX, X = G.backward(ΔZ, Z);  get_params(G);  clear_grad!(G)
# runs unchanged and lands in the B200 library - the same mechanism the reference already uses to
# specialise on CuArray (src/utils/compute_utils.jl:6-18, src/layers/invertible_layer_conv1x1.jl:89).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia toolchain.  The identical C ABI is
# exercised from Python ctypes (invertiblenetworks.jl_b200/lib.py, tests/); this file is the stub a
# maintainer drops into the reference (see INTEGRATION.md).
module InvertibleNetworksB200

using CUDA
using InvertibleNetworks
import InvertibleNetworks: forward, inverse, backward, NetworkGlow, NetworkConditionalGlow, NetworkMultiScaleHINT,
                           CouplingLayerHINT, CouplingLayerBasic, ActNorm,
                           Conv1x1, CouplingLayerGlow, ResidualBlock, get_params, Parameter

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

const PRECISION = Ref{Cint}(parse(Cint, get(ENV, "INB200_PRECISION", "1")))  # 0 fp32, 1 bf16x3, 2 bf16

check(rc) = rc == 0 || error(unsafe_string(ccall((:inb_last_error, LIB), Cstring, ())))
stream() = CUDA.stream().handle
dptr(x::CuArray{Float32}) = reinterpret(Ptr{Cfloat}, pointer(x))
dptr(::Nothing) = Ptr{Cfloat}(C_NULL)

# ---------------------------------------------------------------- plans, cached per (network, input size)
const PLANS = IdDict{Any,Tuple{Any,Ptr{Cvoid}}}()

function plan_for(G, X::CuArray{Float32,N}, n_cond::Int) where N
    key = (size(X)[1:N-2]..., size(X, N))
    haskey(PLANS, G) && PLANS[G][1] == key && return PLANS[G][2]
    haskey(PLANS, G) && ccall((:inb_glow_plan_destroy, LIB), Cint, (Ptr{Cvoid},), PLANS[G][2])
    nd = N - 2
    rb = G.CL[1, 1].RB
    k1, k2 = size(rb.W1.data, 1), size(rb.W2.data, 1)
    desc = GlowDesc(nd, size(X, 1), size(X, 2), nd == 3 ? size(X, 3) : 1,
                    size(X, N - 1), n_cond, size(rb.W2.data, N), G.L, G.K, size(X, N),
                    G.split_scales, hasproperty(G, :logdet) ? G.logdet : true,
                    k1, k2, (k1 - 1) ÷ 2, (k2 - 1) ÷ 2,
                    G.CL[1, 1].activation.low, G.CL[1, 1].activation.high,   # SigmoidLayer(low, high)
                    G.CL[1, 1].C.freeze, PRECISION[])
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:inb_glow_plan_create, LIB), Cint, (Ref{GlowDesc}, Ref{Ptr{Cvoid}}), desc, p))
    PLANS[G] = (key, p[])
    return p[]
end

# get_params order == the library's pointer-table order (src/utils/neuralnet.jl:72-88)
ptr_table(ps::Vector{Parameter}, f) = Ptr{Cfloat}[dptr(getfield(p, f)) for p in ps]

# ActNorm parameters start as `nothing` (invertible_layer_actnorm.jl:53-57): allocate them and ask the
# library to run the data-dependent initialisation inside forward.
function ensure_actnorm!(G, T, ::Type{A}) where A
    init = false
    for AN in G.AN
        if AN.s.data === nothing
            AN.s.data = CUDA.zeros(T, AN.k); AN.b.data = CUDA.zeros(T, AN.k); init = true
        end
    end
    return init
end

function fill_zdims!(G, plan, B)
    G.split_scales || return
    dims = zeros(Cint, 5)
    for i in 1:length(G.Z_dims)
        n = ccall((:inb_glow_zdims, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cint}), plan, B, i - 1, dims)
        G.Z_dims[i] = collect(Int, reverse(dims[1:n]))        # (nx, ny[, nz], C, B)
    end
end

# ---------------------------------------------------------------- NetworkGlow
# replaces src/networks/invertible_network_glow.jl:109-129
function forward(X::CuArray{Float32,N}, G::NetworkGlow) where N
    plan = plan_for(G, X, 0)
    init = ensure_actnorm!(G, Float32, CuArray)
    θ = ptr_table(get_params(G), :data)
    Z = G.split_scales ? CUDA.zeros(Float32, length(X)) : similar(X)
    ld = CUDA.zeros(Float32, 1)
    check(ccall((:inb_glow_forward, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Cfloat}, Ptr{Cfloat}, Cint, Ptr{Cvoid}),
                plan, size(X, N), dptr(X), θ, dptr(Z), G.logdet ? dptr(ld) : C_NULL, init, stream()))
    fill_zdims!(G, plan, size(X, N))
    G.logdet ? (return Z, Array(ld)[1]) : (return Z)
end

# replaces :132-147
function inverse(Z::CuArray{Float32,N}, G::NetworkGlow) where N
    Xshape = input_shape(G, Z)
    X = CUDA.zeros(Float32, Xshape...)
    plan = plan_for(G, X, 0)
    check(ccall((:inb_glow_inverse, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Cfloat}, Ptr{Cvoid}),
                plan, Xshape[end], dptr(Z), ptr_table(get_params(G), :data), dptr(X), stream()))
    return X
end

# replaces :150-191 (set_grad = true)
function backward(ΔZ::CuArray{Float32,N}, Z::CuArray{Float32,N}, G::NetworkGlow; set_grad::Bool=true) where N
    set_grad || return invoke(backward, Tuple{AbstractArray{Float32,N},AbstractArray{Float32,N},NetworkGlow},
                              ΔZ, Z, G; set_grad=false)          # Jacobian paths stay on the reference
    Xshape = input_shape(G, Z)
    X, ΔX = CUDA.zeros(Float32, Xshape...), CUDA.zeros(Float32, Xshape...)
    plan = plan_for(G, X, 0)
    ps = get_params(G)
    fresh = [CUDA.zeros(Float32, size(p.data)) for p in ps]
    check(ccall((:inb_glow_backward, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Ptr{Cfloat}}, Ptr{Cfloat},
                 Ptr{Cfloat}, Ptr{Cvoid}),
                plan, Xshape[end], dptr(ΔZ), dptr(Z), ptr_table(ps, :data), Ptr{Cfloat}[dptr(g) for g in fresh],
                dptr(ΔX), dptr(X), stream()))
    assign_grads!(G, ps, fresh)
    return ΔX, X
end

# the library WRITES gradients; the reference's rules are applied here:
#   ActNorm / ResidualBlock grads are overwritten (invertible_layer_actnorm.jl:113-114,
#   layer_residual_block.jl:168-172), Conv1x1 grads accumulate unless cleared (conv1x1.jl:237-239).
function assign_grads!(G, ps, fresh)
    hh = Set{Parameter}()
    for CL in G.CL
        push!(hh, CL.C.v1, CL.C.v2, CL.C.v3)
    end
    for (p, g) in zip(ps, fresh)
        p.grad = (p in hh && p.grad !== nothing) ? p.grad .+ g : g
    end
end

# Z_dims is filled by forward (:123); the input shape follows from it
function input_shape(G::NetworkGlow, Z)
    G.split_scales || return size(Z)
    zd = G.Z_dims[1]
    nd = length(zd) - 2
    c_in = zd[end-1] * 2 ÷ (2^nd)
    return (2 .* zd[1:nd]..., c_in, zd[end])
end

# ---------------------------------------------------------------- NetworkConditionalGlow
# replaces src/networks/invertible_network_conditional_glow.jl:107-130
function forward(X::CuArray{Float32,N}, C::CuArray{Float32,N}, G::NetworkConditionalGlow) where N
    plan = plan_for(G, X, size(C, N - 1))
    init = ensure_actnorm!(G, Float32, CuArray)
    if G.AN_C.s.data === nothing
        G.AN_C.s.data = CUDA.zeros(Float32, G.AN_C.k); G.AN_C.b.data = CUDA.zeros(Float32, G.AN_C.k); init = true
    end
    ZX = similar(X)
    f = G.split_scales ? 2^G.L : 1
    nd = N - 2
    ZC = CUDA.zeros(Float32, (size(C)[1:nd] .÷ f)..., size(C, N - 1) * (G.split_scales ? (2^nd)^G.L : 1), size(C, N))
    ld = CUDA.zeros(Float32, 1)
    check(ccall((:inb_cglow_forward, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat},
                 Cint, Ptr{Cvoid}),
                plan, size(X, N), dptr(X), dptr(C), ptr_table(get_params(G), :data), dptr(ZX), dptr(ZC), dptr(ld),
                init, stream()))
    fill_zdims!(G, plan, size(X, N))
    return ZX, ZC, Array(ld)[1]
end

# replaces :133-148
function inverse(ZX::CuArray{Float32,N}, ZC::CuArray{Float32,N}, G::NetworkConditionalGlow) where N
    X = similar(ZX)
    plan = plan_for(G, X, size(ZC, N - 1) ÷ (G.split_scales ? (2^(N - 2))^G.L : 1))
    check(ccall((:inb_cglow_inverse, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Cfloat}, Ptr{Cvoid}),
                plan, size(X, N), dptr(ZX), dptr(ZC), ptr_table(get_params(G), :data), dptr(X), stream()))
    return X
end

# replaces :151-181
function backward(ΔZX::CuArray{Float32,N}, ZX::CuArray{Float32,N}, ZC::CuArray{Float32,N},
                  G::NetworkConditionalGlow) where N
    X, ΔX = similar(ZX), similar(ZX)
    n_cond = size(ZC, N - 1) ÷ (G.split_scales ? (2^(N - 2))^G.L : 1)
    ΔC = CUDA.zeros(Float32, size(ZX)[1:N-2]..., n_cond, size(ZX, N))
    plan = plan_for(G, X, n_cond)
    ps = get_params(G)
    fresh = [CUDA.zeros(Float32, size(p.data)) for p in ps]
    check(ccall((:inb_cglow_backward, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Ptr{Cfloat}},
                 Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                plan, size(X, N), dptr(ΔZX), dptr(ZX), dptr(ZC), ptr_table(ps, :data),
                Ptr{Cfloat}[dptr(g) for g in fresh], dptr(ΔX), dptr(X), dptr(ΔC), stream()))
    assign_grads!(G, ps, fresh)
    return ΔX, X, ΔC
end

# ---------------------------------------------------------------- layers (same pattern, one ccall each)
# replaces src/layers/invertible_layer_actnorm.jl:60-77
function forward(X::CuArray{Float32,N}, AN::ActNorm; logdet=nothing) where N
    isnothing(logdet) ? logdet = (AN.logdet && ~AN.is_reversed) : logdet = logdet
    B, C, sp = size(X, N), size(X, N - 1), prod(size(X)[1:N-2])
    if AN.s.data === nothing && !AN.is_reversed
        AN.s.data = CUDA.zeros(Float32, C); AN.b.data = CUDA.zeros(Float32, C)
        check(ccall((:inb_actnorm_init, LIB), Cint, (Cint, Cint, Clonglong, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                    B, C, sp, dptr(X), dptr(AN.s.data), dptr(AN.b.data), stream()))
    end
    Y = similar(X); ld = CUDA.zeros(Float32, 1)
    check(ccall((:inb_actnorm_forward, LIB), Cint,
                (Cint, Cint, Clonglong, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                B, C, sp, dptr(X), dptr(AN.s.data), dptr(AN.b.data), dptr(Y), logdet ? dptr(ld) : C_NULL, stream()))
    logdet ? (return Y, Array(ld)[1]) : (return Y)
end

# replaces src/layers/invertible_layer_conv1x1.jl:174-189 / 209-224
for (fn, sym) in ((:forward, :inb_conv1x1_forward), (:inverse, :inb_conv1x1_inverse))
    @eval function $fn(X::CuArray{Float32,N}, C::Conv1x1; logdet=nothing) where N
        Y = similar(X)
        check(ccall(($(QuoteNode(sym)), LIB), Cint,
                    (Cint, Cint, Clonglong, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                    size(X, N), size(X, N - 1), prod(size(X)[1:N-2]), dptr(X), dptr(C.v1.data), dptr(C.v2.data),
                    dptr(C.v3.data), dptr(Y), stream()))
        return Y
    end
end

# replaces src/utils/dimensionality_operations.jl:79-107 (checkerboard)
function InvertibleNetworks.squeeze(X::CuArray{Float32,N}; pattern="column") where N
    pattern == "checkerboard" || return invoke(InvertibleNetworks.squeeze, Tuple{AbstractArray{Float32,N}}, X; pattern=pattern)
    nd = N - 2
    Y = CUDA.zeros(Float32, (size(X)[1:nd] .÷ 2)..., size(X, N - 1) * 2^nd, size(X, N))
    check(ccall((:inb_squeeze, LIB), Cint, (Cint, Cint, Cint, Cint, Cint, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                nd, size(X, 1), size(X, 2), nd == 3 ? size(X, 3) : 1, size(X, N), size(X, N - 1), dptr(X), dptr(Y), stream()))
    return Y
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

function hint_plan_for(H::NetworkMultiScaleHINT, X::CuArray{Float32,4})
    key = (size(X, 1), size(X, 2), size(X, 4))
    haskey(HINT_PLANS, H) && HINT_PLANS[H][1] == key && return HINT_PLANS[H][2]
    haskey(HINT_PLANS, H) && ccall((:inb_hint_plan_destroy, LIB), Cint, (Ptr{Cvoid},), HINT_PLANS[H][2])
    rb = H.CL[1, 1].CL[1].RB
    k1, k2 = size(rb.W1.data, 1), size(rb.W2.data, 1)
    act = H.CL[1, 1].CL[1].activation
    desc = HintDesc(size(X, 1), size(X, 2), size(X, 3), size(rb.W2.data, 4), H.L, H.K, size(X, 4), H.split_scales,
                    k1, k2, (k1 - 1) ÷ 2, (k2 - 1) ÷ 2, act.low, act.high, 0, HINT_SHARED_GRADS[], PRECISION[])
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:inb_hint_plan_create, LIB), Cint, (Ref{HintDesc}, Ref{Ptr{Cvoid}}), desc, p))
    HINT_PLANS[H] = (key, p[])
    return p[]
end

# replaces src/networks/invertible_network_hint_multiscale.jl:98-117
function forward(X::CuArray{Float32,4}, H::NetworkMultiScaleHINT)
    plan = hint_plan_for(H, X)
    init = ensure_actnorm!(H, Float32, CuArray)
    f = 2^H.L
    Z = H.split_scales ? CUDA.zeros(Float32, length(X)) :
        CUDA.zeros(Float32, size(X, 1) ÷ f, size(X, 2) ÷ f, size(X, 3) * 4^H.L, size(X, 4))
    ld = CUDA.zeros(Float32, 1)
    check(ccall((:inb_hint_forward, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Cfloat}, Ptr{Cfloat}, Cint, Ptr{Cvoid}),
                plan, size(X, 4), dptr(X), ptr_table(get_params(H), :data), dptr(Z), dptr(ld), init, stream()))
    HINT_INPUT[H] = size(X)   # stands in for H.X_dims (:111)
    return Z, Array(ld)[1]
end
const HINT_INPUT = IdDict{Any,Any}()
hint_input_shape(H, Z) = haskey(HINT_INPUT, H) && prod(HINT_INPUT[H]) == length(Z) ? HINT_INPUT[H] :
    (size(Z, 1) * 2^H.L, size(Z, 2) * 2^H.L, size(Z, 3) ÷ 4^H.L, size(Z, 4))

# replaces :120-133
function inverse(Z::CuArray{Float32}, H::NetworkMultiScaleHINT)
    X = CUDA.zeros(Float32, hint_input_shape(H, Z)...)
    plan = hint_plan_for(H, X)
    check(ccall((:inb_hint_inverse, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Cfloat}, Ptr{Cvoid}),
                plan, size(X, 4), dptr(Z), ptr_table(get_params(H), :data), dptr(X), stream()))
    return X
end

# replaces :136-174 (set_grad = true; the Jacobian paths stay on the reference)
function backward(ΔZ::CuArray{Float32}, Z::CuArray{Float32}, H::NetworkMultiScaleHINT; set_grad::Bool=true)
    set_grad || return invoke(backward, Tuple{AbstractArray{Float32},AbstractArray{Float32},NetworkMultiScaleHINT},
                              ΔZ, Z, H; set_grad=false)
    shape = hint_input_shape(H, Z)
    X, ΔX = CUDA.zeros(Float32, shape...), CUDA.zeros(Float32, shape...)
    plan = hint_plan_for(H, X)
    ps = get_params(H)
    fresh = [CUDA.zeros(Float32, size(p.data)) for p in ps]
    check(ccall((:inb_hint_backward, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Ptr{Cfloat}}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                plan, shape[end], dptr(ΔZ), dptr(Z), ptr_table(ps, :data), Ptr{Cfloat}[dptr(g) for g in fresh],
                dptr(ΔX), dptr(X), stream()))
    assign_grads!(H, ps, fresh)   # H.CL[i,j].C is the Conv1x1 whose gradients accumulate (conv1x1.jl:237-239)
    return ΔX, X
end

# ---- layer level: CouplingLayerHINT (src/layers/invertible_layer_hint.jl:105-297) and CouplingLayerBasic
# (src/layers/invertible_layer_basic.jl:90-149).  Reversed layers and set_grad = false fall through.
const PERMUTE = Dict("none" => Cint(0), "full" => Cint(1), "lower" => Cint(2), "both" => Cint(3))
hint_on_b200(H::CouplingLayerHINT) = !H.is_reversed && haskey(PERMUTE, H.permute)
function hint_ints(X::CuArray{Float32,N}, H::CouplingLayerHINT) where N
    rb = H.CL[1].RB
    act = H.CL[1].activation
    nd = N - 2
    return (Cint(nd), Cint(size(X, 1)), Cint(size(X, 2)), Cint(nd == 3 ? size(X, 3) : 1), Cint(size(X, N)),
            Cint(size(X, N - 1)), Cint(size(rb.W2.data, N)), Cint(size(rb.W1.data, 1)), Cint(size(rb.W2.data, 1)),
            Cfloat(act.low), Cfloat(act.high), PERMUTE[H.permute])
end
const HINT_ARGT = (Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cfloat, Cfloat, Cint)

# replaces :105-156 (scale = 1 entry; the recursion runs inside the library)
function forward(X::CuArray{Float32,N}, H::CouplingLayerHINT; scale=1, permute=nothing, logdet=nothing) where N
    (hint_on_b200(H) && scale == 1 && permute === nothing) ||
        return invoke(forward, Tuple{AbstractArray{Float32,N},CouplingLayerHINT}, X, H; scale=scale, permute=permute, logdet=logdet)
    logdet = logdet === nothing ? H.logdet : logdet
    Y = similar(X)
    ld = CUDA.zeros(Float32, 1)
    check(ccall((:inb_hint_coupling_forward, LIB), Cint,
                (HINT_ARGT..., Cint, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                hint_ints(X, H)..., PRECISION[], dptr(X), ptr_table(get_params(H), :data), dptr(Y),
                logdet ? dptr(ld) : C_NULL, stream()))
    logdet ? (return Y, Array(ld)[1]) : (return Y)
end

# replaces :159-204
function inverse(Y::CuArray{Float32,N}, H::CouplingLayerHINT; scale=1, permute=nothing, logdet=nothing) where N
    (hint_on_b200(H) && scale == 1 && permute === nothing && logdet !== true) ||
        return invoke(inverse, Tuple{AbstractArray{Float32,N},CouplingLayerHINT}, Y, H; scale=scale, permute=permute, logdet=logdet)
    X = similar(Y)
    check(ccall((:inb_hint_coupling_inverse, LIB), Cint,
                (HINT_ARGT..., Cint, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Cfloat}, Ptr{Cvoid}),
                hint_ints(Y, H)..., PRECISION[], dptr(Y), ptr_table(get_params(H), :data), dptr(X), stream()))
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
                hint_ints(Y, H)..., Cint(H.logdet), HINT_SHARED_GRADS[], PRECISION[], dptr(ΔY), dptr(Y),
                ptr_table(ps, :data), Ptr{Cfloat}[dptr(g) for g in fresh], dptr(ΔX), dptr(X), stream()))
    hh = H.C === nothing ? Set{Parameter}() : Set{Parameter}((H.C.v1, H.C.v2, H.C.v3))
    for (p, g) in zip(ps, fresh)   # Conv1x1 gradients accumulate unless cleared (conv1x1.jl:237-239)
        p.grad = (p in hh && p.grad !== nothing) ? p.grad .+ g : g
    end
    return ΔX, X
end

# CouplingLayerBasic: forward :90-105, inverse :108-121, backward :124-149 (non-reversed, set_grad = true, save = false)
function basic_ints(X1::CuArray{Float32,N}, L::CouplingLayerBasic) where N
    rb = L.RB
    nd = N - 2
    return (Cint(nd), Cint(size(X1, 1)), Cint(size(X1, 2)), Cint(nd == 3 ? size(X1, 3) : 1), Cint(size(X1, N)),
            Cint(size(X1, N - 1)), Cint(size(rb.W2.data, N)), Cint(size(rb.W1.data, 1)), Cint(size(rb.W2.data, 1)),
            Cfloat(L.activation.low), Cfloat(L.activation.high))
end
const BASIC_ARGT = (Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cfloat, Cfloat)
function forward(X1::CuArray{Float32,N}, X2::CuArray{Float32,N}, L::CouplingLayerBasic; save::Bool=false, logdet=nothing) where N
    (L.RB isa ResidualBlock && !save && !L.is_reversed) ||
        return invoke(forward, Tuple{AbstractArray{Float32,N},AbstractArray{Float32,N},CouplingLayerBasic}, X1, X2, L; save=save, logdet=logdet)
    logdet = logdet === nothing ? L.logdet : logdet
    Y2 = similar(X2)
    ld = CUDA.zeros(Float32, 1)
    check(ccall((:inb_basic_coupling_forward, LIB), Cint,
                (BASIC_ARGT..., Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                basic_ints(X1, L)..., PRECISION[], dptr(X1), dptr(X2), ptr_table(get_params(L), :data), dptr(Y2),
                logdet ? dptr(ld) : C_NULL, stream()))
    logdet ? (return X1, Y2, Array(ld)[1]) : (return X1, Y2)
end
function inverse(Y1::CuArray{Float32,N}, Y2::CuArray{Float32,N}, L::CouplingLayerBasic; save::Bool=false, logdet=nothing) where N
    (L.RB isa ResidualBlock && !save && !L.is_reversed && logdet !== true) ||
        return invoke(inverse, Tuple{AbstractArray{Float32,N},AbstractArray{Float32,N},CouplingLayerBasic}, Y1, Y2, L; save=save, logdet=logdet)
    X2 = similar(Y2)
    check(ccall((:inb_basic_coupling_inverse, LIB), Cint,
                (BASIC_ARGT..., Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Ptr{Cfloat}}, Ptr{Cfloat}, Ptr{Cvoid}),
                basic_ints(Y1, L)..., PRECISION[], dptr(Y1), dptr(Y2), ptr_table(get_params(L), :data), dptr(X2), stream()))
    return Y1, X2
end
function backward(ΔY1::CuArray{Float32,N}, ΔY2::CuArray{Float32,N}, Y1::CuArray{Float32,N}, Y2::CuArray{Float32,N},
                  L::CouplingLayerBasic; set_grad::Bool=true) where N
    (L.RB isa ResidualBlock && set_grad && !L.is_reversed) ||
        return invoke(backward, Tuple{AbstractArray{Float32,N},AbstractArray{Float32,N},AbstractArray{Float32,N},
                                      AbstractArray{Float32,N},CouplingLayerBasic}, ΔY1, ΔY2, Y1, Y2, L; set_grad=set_grad)
    ΔX1, ΔX2, X2 = similar(Y1), similar(Y2), similar(Y2)
    ps = get_params(L)
    fresh = [CUDA.zeros(Float32, size(p.data)) for p in ps]
    check(ccall((:inb_basic_coupling_backward, LIB), Cint,
                (BASIC_ARGT..., Cint, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Ptr{Cfloat}},
                 Ptr{Ptr{Cfloat}}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                basic_ints(Y1, L)..., Cint(L.logdet), PRECISION[], dptr(ΔY1), dptr(ΔY2), dptr(Y1), dptr(Y2),
                ptr_table(ps, :data), Ptr{Cfloat}[dptr(g) for g in fresh], dptr(ΔX1), dptr(ΔX2), dptr(X2), stream()))
    for (p, g) in zip(ps, fresh)   # layer_residual_block.jl:168-172: overwritten
        p.grad = g
    end
    return ΔX1, ΔX2, Y1, X2
end

# replaces wavelet_squeeze / wavelet_unsqueeze with type = WT.db1 and Haar_squeeze / invHaar_unsqueeze
# (src/utils/dimensionality_operations.jl:199-258, 318-371), 4-D tensors
for (fn, sym, ty, up) in ((:wavelet_squeeze, :inb_haar_squeeze, 0, false), (:wavelet_unsqueeze, :inb_haar_unsqueeze, 0, true),
                          (:Haar_squeeze, :inb_haar_squeeze, 1, false), (:invHaar_unsqueeze, :inb_haar_unsqueeze, 1, true))
    @eval function InvertibleNetworks.$fn(X::CuArray{Float32,4})
        Y = $up ? CUDA.zeros(Float32, 2size(X, 1), 2size(X, 2), size(X, 3) ÷ 4, size(X, 4)) :
                  CUDA.zeros(Float32, size(X, 1) ÷ 2, size(X, 2) ÷ 2, 4size(X, 3), size(X, 4))
        check(ccall(($(QuoteNode(sym)), LIB), Cint, (Cint, Cint, Cint, Cint, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                    size(X, 1), size(X, 2), size(X, 4), size(X, 3), $ty, dptr(X), dptr(Y), stream()))
        return Y
    end
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
