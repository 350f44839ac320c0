"""GPU parity at the FULL size of the BASELINE configurations, on the path bench.py runs, against the float64
oracle (SURVEY.md 8c parity metric): per tensor ||a - b|| / ||b|| for Z, logdet, the recomputed X, dX and every
parameter gradient, in two columns - CUDA vs truth | float32 oracle vs truth - so that the CUDA path is judged
against the error the reference's own float32 arithmetic shows on the same inputs (ReLU-mask flips included,
DESIGN.md section 8), not against a loose constant.

  cfg2  NetworkGlow(3, 256, L=3, K=16; split_scales) on 256x256x3, B = 2   (the benched tcgen05 path + fp32 path)
  cfg3  NetworkConditionalGlow(1, 1, 32, L=2, K=10; split_scales) on 64x64, B = 8
  cfg5  NetworkGlow3D(1, 32, L=2, K=2) on 64^3, B = 2
  Sigmoid(low=0.5, high=1) of examples/applications/application_glow_seismic/glow_seismic.jl:86 at layer and network level.

Every report is also written to gpurun_out/parity_<name>.json (copied to profiles/ for the record)."""
import json
import os

import pytest
import torch

from _util import O, TOL_GRAD, TOL_LOGDET, TOL_OUT, FragileUnits, clone_oracle, rel

import inb200

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# Gradient parity at depth.  _relugrad (activation_functions.jl:84) is discontinuous: two arithmetics that differ by eps
# in a pre-activation disagree on the mask of the units within eps of zero, every such flip is an O(1) change of the
# gradients behind it, and WHERE the flips fall is random.  So the float32 oracle (= the reference's own arithmetic) is
# itself 1e-3 away from the float64 truth after 48 flow steps, tensor by tensor at different places than the CUDA path.
# The comparison is therefore made on the distribution over the 10*L*K gradient tensors:
#   (i)  the flow step backward visits FIRST sees identical inputs in every arithmetic: its ten gradients must meet
#        STRICT[precision] (the float32 bar TOL_GRAD for fp32 / fp16x3; bf16x3's own 2^-17 operand rounding already flips
#        ~10 units per million there, stated as such);
#   (ii) median and maximum over all tensors of the CUDA path's error <= DRIFT[precision] x the float32 oracle's.
import statistics

STRICT = {"fp32": TOL_GRAD, "fp16x3": TOL_GRAD, "bf16x3": 1e-2}
DRIFT = {"fp32": 2.5, "fp16x3": 5.0, "bf16x3": 25.0}
# ... or below FLOOR[precision] (median; 10x for the worst tensor) where the float32 oracle happens to see no flip at all:
# the tensor-core modes carry ~3e-6 of arithmetic noise per block (both x3 splits; the fp32 kernels 1e-7), i.e. ~30x
# as many units within reach of a flip as the reference's own float32 arithmetic
FLOOR = {"fp32": TOL_GRAD, "fp16x3": 5e-4, "bf16x3": 2e-3}
# invertibility ||X - inverse(forward(X))|| / ||X||: reference bound 1f-5 on its small test nets (test_glow.jl:46);
# at full depth the float32 oracle itself is measured next to the CUDA path and printed.
INV_TOL = {"fp32": 1e-5, "bf16x3": 5e-5, "fp16x3": 1e-5}


def g(t):
    return t.to(DEV)


def _dump(name, rep):
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, f"parity_{name}.json"), "w") as fh:
            json.dump(rep, fh, indent=1)
    except OSError:
        pass


def _table(name, rows):
    print(f"\n== parity {name}: tensor | CUDA vs float64 oracle | float32 oracle vs float64 oracle")
    for r in rows:
        print(f"  {r[0]:<28s} {r[1]:10.3e} {r[2]:10.3e}")


def _grad_rows(ps, q64, q32, names):
    rows = []
    for i, (p, a, b) in enumerate(zip(ps, q64, q32)):
        rows.append((names(i), rel(p.grad, a.grad), rel(b.grad, a.grad)))
    return rows


def _check_grads(rows, precision, fragile, strict_idx=()):
    e_cuda = [r[1] for r in rows]
    e_o32 = [r[2] for r in rows]
    for i in strict_idx:
        nm, a, b = rows[i]
        assert a < STRICT[precision], f"{nm} (first flow step of backward): CUDA {a:.3e} vs {STRICT[precision]:.1e} " \
                                      f"(float32 oracle {b:.3e})"
    med_c, med_o, max_c, max_o = statistics.median(e_cuda), statistics.median(e_o32), max(e_cuda), max(e_o32)
    stats = {"median_cuda": med_c, "median_oracle_f32": med_o, "max_cuda": max_c, "max_oracle_f32": max_o}
    d, floor = DRIFT[precision], FLOOR[precision]
    assert med_c < max(floor, d * med_o), f"median gradient error {med_c:.3e} vs max({floor:.0e}, {d} x {med_o:.3e}) " \
                                          f"(float32 oracle)"
    assert max_c < max(10 * floor, d * max_o), f"worst gradient error {max_c:.3e} vs max({10 * floor:.0e}, {d} x {max_o:.3e}) " \
                                               f"(float32 oracle); fragile ReLU units {fragile.count}/{fragile.total}"
    return stats


def glow_names(L, K):
    an = 2 * L * K

    def f(i):
        if i < an:
            l, w = divmod(i, 2)
            return f"AN[{l // K + 1},{l % K + 1}].{'sb'[w]}"
        l, w = divmod(i - an, 8)
        return f"CL[{l // K + 1},{l % K + 1}].{'v1 v2 v3 W1 W2 W3 b1 b2'.split()[w]}"
    return f


_ORACLE_CACHE = {}


def _oracle_glow(n_in, nh, L, K, shape, ndims, low, high, seed):
    """Both oracle arithmetics on one seeded problem, computed once per configuration (the float64 backward of full
    cfg2 takes about a minute on the host cores) and shared by the precisions under test."""
    key = (n_in, nh, L, K, tuple(shape), ndims, low, high, seed)
    if key in _ORACLE_CACHE:
        return _ORACLE_CACHE[key]
    torch.manual_seed(seed)
    mk = lambda dt: O.NetworkGlow(n_in, nh, L, K, split_scales=True, ndims=ndims, low=low, high=high, seed=3,
                                  dtype=dt, faithful=False)
    G32, G64 = mk(torch.float32), clone_oracle(mk, torch.float64)
    X = torch.rand(*shape)
    B = shape[0]
    Z32, ld32 = G32.forward(X)  # the float32 oracle's data-dependent ActNorm init defines everyone's parameters
    for p, q in zip(G64.get_params(), G32.get_params()):
        p.data = q.data.double()
    Z64, ld64 = G64.forward(X.double())
    r = {"X": X, "params": [p.data.clone() for p in G32.get_params()], "Z32": Z32, "ld32": ld32, "Z64": Z64, "ld64": ld64,
         "inv_o32": rel(G32.inverse(Z32), X)}
    # backward from the same latent (the truth's Z rounded to float32) in both arithmetics
    Zin = Z64.float()
    dZ = Zin / B
    with FragileUnits(1e-6) as fr:
        dX64, X64 = G64.backward(dZ.double(), Zin.double())
    dX32, X32 = G32.backward(dZ, Zin)
    r.update(Zin=Zin, dZ=dZ, dX64=dX64, X64=X64, dX32=dX32, X32=X32, fragile=fr,
             g64=[p.grad for p in G64.get_params()], g32=[p.grad for p in G32.get_params()])
    _ORACLE_CACHE[key] = r
    return r


def run_glow_full(name, n_in, nh, L, K, shape, precision, *, ndims=2, low=0.0, high=1.0, seed=11, tol_out=TOL_OUT):
    o = _oracle_glow(n_in, nh, L, K, shape, ndims, low, high, seed)
    X, Z64, ld64, Z32, ld32, fr = o["X"], o["Z64"], o["ld64"], o["Z32"], o["ld32"], o["fragile"]
    G = inb200.NetworkGlow(n_in, nh, L, K, split_scales=True, ndims=ndims, low=low, high=high, precision=precision,
                           device=DEV)
    inb200.set_params(G, o["params"])
    Z, ld = G.forward(g(X))
    rows = [("Z", rel(Z, Z64), rel(Z32, Z64)),
            ("logdet", abs(ld.item() - ld64.item()) / abs(ld64.item()), abs(ld32.item() - ld64.item()) / abs(ld64.item()))]
    assert rows[0][1] < max(tol_out, 2 * rows[0][2])
    assert rows[1][1] < TOL_LOGDET * (tol_out / TOL_OUT)
    inv_cuda, inv_o32 = rel(G.inverse(Z), X), o["inv_o32"]
    rows.append(("X - inverse(forward(X))", inv_cuda, inv_o32))
    assert inv_cuda < max(INV_TOL[precision], 2 * inv_o32), (inv_cuda, inv_o32)
    dX, Xr = G.backward(g(o["dZ"]), g(o["Zin"]))
    rows.append(("X recomputed by backward", rel(Xr, o["X64"]), rel(o["X32"], o["X64"])))
    assert rows[-1][1] < tol_out
    rows.append(("dX", rel(dX, o["dX64"]), rel(o["dX32"], o["dX64"])))
    assert rows[-1][1] < max(tol_out, DRIFT[precision] * rows[-1][2])
    ps = G.get_params()
    assert sum(p.grad is not None for p in ps) == 10 * L * K  # test_glow.jl:50-62
    names = glow_names(L, K)
    grows = [(names(i), rel(p.grad, a), rel(b, a)) for i, (p, a, b) in enumerate(zip(ps, o["g64"], o["g32"]))]
    # the flow step backward visits first (scale L, step K) sees the same inputs in every arithmetic: strict
    last = (L - 1) * K + (K - 1)
    strict = [2 * last, 2 * last + 1] + [2 * L * K + 8 * last + k for k in range(8)]
    _table(f"{name} [{precision}]", rows + [grows[i] for i in strict])
    _dump(f"{name}_{precision}", {"config": name, "precision": precision, "shape": list(shape),
                                  "columns": ["tensor", "cuda_vs_f64", "oracle_f32_vs_f64"],
                                  "rows": [list(r) for r in rows + grows],
                                  "fragile_relu_units": [fr.count, fr.total]})
    st = _check_grads(grows, precision, fr, strict)
    print(f"  all {len(grows)} gradients: median CUDA {st['median_cuda']:.3e} | float32 oracle {st['median_oracle_f32']:.3e}; "
          f"worst CUDA {st['max_cuda']:.3e} | float32 oracle {st['max_oracle_f32']:.3e}; fragile ReLU units "
          f"{fr.count}/{fr.total}")
    return rows, grows


@pytest.mark.parametrize("precision", ["fp16x3", "bf16x3", "fp32"])
def test_cfg2_full_size_against_the_float64_oracle(precision):
    """BASELINE configs[1] at full depth and resolution (L=3, K=16, n_hidden=256, 256x256x3), B=2 - the configuration
    bench.py times, on the kernels it times (bf16x3 = the tcgen05 CTA-pair chain + TMEM weight gradients)."""
    run_glow_full("cfg2", 3, 256, 3, 16, (2, 3, 256, 256), precision)


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
def test_cfg5_full_size_3d_against_the_float64_oracle(precision):
    """BASELINE configs[4]: NetworkGlow3D(1, 32, L=2, K=2) on 64^3 x 1 volumes, B=2.  fp16x3: scale 1 on the
    fused tcgen05 chain (n_hidden padded 32 -> 128), scale 2 (27 taps x 32 channels = 864 expanded columns, more than the
    chain's TMEM region) on the fp32 kernels."""
    run_glow_full("cfg5", 1, 32, 2, 2, (2, 1, 64, 64, 64), precision, ndims=3)


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
def test_cfg1_sigmoid_half_one_network(precision):
    """SigmoidLayer(low=0.5, high=1) (activation_functions.jl:30-35) as glow_seismic.jl:86 passes it to NetworkGlow:
    cfg1's network with that activation, full parity report."""
    run_glow_full("cfg1_sigmoid_0.5_1", 1, 32, 2, 2, (8, 1, 64, 64), precision, low=0.5, high=1.0)


@pytest.mark.parametrize("precision", ["fp16x3", "bf16x3"])
def test_cfg2_channel_plan_sigmoid_half_one_tensor_core(precision):
    """The same activation on the tcgen05 path: cfg2's channel plan (n_hidden=256, L=3) at 64x64, K=2."""
    run_glow_full("cfg2small_sigmoid_0.5_1", 3, 256, 3, 2, (2, 3, 64, 64), precision, low=0.5, high=1.0)


@pytest.mark.parametrize("low,high", [(0.5, 1.0), (0.1, 2.0)])
@pytest.mark.parametrize("n_cond", [0, 3])
def test_coupling_layer_sigmoid_low_high(low, high, n_cond):
    """CouplingLayerGlow / ConditionalLayerGlow with SigmoidLayer(low, high): forward, logdet, inverse, backward and
    all eight gradients (invertible_layer_glow.jl:104-170 with activation_functions.jl:160-217)."""
    torch.manual_seed(6)
    B, Cc, nh, sp = 2, 12, 32, (16, 16)
    CL = inb200.CouplingLayerGlow(Cc, nh, n_cond=n_cond, logdet=True, low=low, high=high,
                                  gen=torch.Generator().manual_seed(9), device=DEV)
    CL.RB.b1.data.copy_(torch.randn(nh) * 0.1)
    CL.RB.b2.data.copy_(torch.randn(nh) * 0.1)
    ws = [p.data.cpu() for p in CL.get_params()]
    mk = lambda dt: O.CouplingLayerGlow(O.Conv1x1(*[w.to(dt) for w in ws[:3]]),
                                        O.ResidualBlock(*[w.to(dt) for w in ws[3:]]), logdet=True, low=low, high=high)
    C32, C64 = mk(torch.float32), mk(torch.float64)
    X, dY = torch.randn(B, Cc, *sp), torch.randn(B, Cc, *sp)
    cond = torch.randn(B, n_cond, *sp) if n_cond else None
    cg = g(cond) if n_cond else None
    c64 = cond.double() if n_cond else None
    Y, ld = CL.forward(g(X), cg)
    Y64, ld64 = C64.forward(X.double(), c64)
    Y32, _ = C32.forward(X, cond)
    assert rel(Y, Y64) < max(TOL_OUT, 2 * rel(Y32, Y64))
    assert abs(ld.item() - ld64.item()) / abs(ld64.item()) < TOL_LOGDET
    # S in [low, high): the log-determinant has the sign of log(S)
    assert rel(CL.inverse(Y, cg), X) < 1e-5
    res = CL.backward(g(dY), Y, cg)
    r64 = C64.backward(dY.double(), Y.cpu().double(), c64)
    for a, b in zip(res, r64):
        assert rel(a, b) < TOL_OUT
    for nm, p, q in zip("v1 v2 v3 W1 W2 W3 b1 b2".split(), CL.get_params(), C64.params()):
        assert rel(p.grad, q.grad) < TOL_GRAD, nm


def cglow_names(L, K):
    an = 2 * L * K

    def f(i):
        if i < an:
            l, w = divmod(i, 2)
            return f"AN[{l // K + 1},{l % K + 1}].{'sb'[w]}"
        if i < an + 2:
            return f"AN_C.{'sb'[i - an]}"
        l, w = divmod(i - an - 2, 8)
        return f"CL[{l // K + 1},{l % K + 1}].{'v1 v2 v3 W1 W2 W3 b1 b2'.split()[w]}"
    return f


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
def test_cfg3_full_size_conditional_against_the_float64_oracle(precision):
    """BASELINE configs[2]: NetworkConditionalGlow(1, 1, 32, L=2, K=10; split_scales) on 64x64 with a 64x64
    condition (amortized_glow_mnist_inpainting.jl:82-89 at the SURVEY 8d size), B=8."""
    torch.manual_seed(21)
    n_in, n_cond, nh, L, K, B = 1, 1, 32, 2, 10, 8
    mk = lambda dt: O.NetworkConditionalGlow(n_in, n_cond, nh, L, K, split_scales=True, seed=5, dtype=dt,
                                             faithful=False)
    G32, G64 = mk(torch.float32), clone_oracle(mk, torch.float64)
    X, Cn = torch.rand(B, n_in, 64, 64), torch.rand(B, n_cond, 64, 64)
    ZX32, ZC32, ld32 = G32.forward(X, Cn)
    for p, q in zip(G64.get_params(), G32.get_params()):
        p.data = q.data.double()
    ZX64, ZC64, ld64 = G64.forward(X.double(), Cn.double())
    G = inb200.NetworkConditionalGlow(n_in, n_cond, nh, L, K, split_scales=True, precision=precision, device=DEV)
    inb200.set_params(G, [p.data for p in G32.get_params()])
    ZX, ZC, ld = G.forward(g(X), g(Cn))
    rows = [("ZX", rel(ZX, ZX64), rel(ZX32, ZX64)), ("ZC", rel(ZC, ZC64), rel(ZC32, ZC64)),
            ("logdet", abs(ld.item() - ld64.item()) / abs(ld64.item()), abs(ld32.item() - ld64.item()) / abs(ld64.item()))]
    assert rows[0][1] < max(TOL_OUT, 2 * rows[0][2]) and rows[1][1] < TOL_OUT and rows[2][1] < TOL_LOGDET
    inv_cuda, inv_o32 = rel(G.inverse(ZX, ZC), X), rel(G32.inverse(ZX32, ZC32), X)
    rows.append(("X - inverse(forward(X))", inv_cuda, inv_o32))
    assert inv_cuda < max(INV_TOL[precision], 2 * inv_o32)  # test_conditional_glow_network.jl:35
    Zin, ZCin = ZX64.float(), ZC64.float()
    dX, Xr, dC = G.backward(g(Zin / B), g(Zin), g(ZCin))
    with FragileUnits(1e-6) as fr:
        dX64, X64, dC64 = G64.backward((Zin / B).double(), Zin.double(), ZCin.double())
    dX32, X32, dC32 = G32.backward(Zin / B, Zin, ZCin)
    rows += [("X recomputed by backward", rel(Xr, X64), rel(X32, X64)), ("dX", rel(dX, dX64), rel(dX32, dX64)),
             ("dC", rel(dC, dC64), rel(dC32, dC64))]
    assert rows[-3][1] < TOL_OUT
    assert rows[-2][1] < max(TOL_OUT, DRIFT[precision] * rows[-2][2])
    assert rows[-1][1] < max(TOL_OUT, DRIFT[precision] * rows[-1][2])
    ps = G.get_params()
    assert sum(p.grad is not None for p in ps) == 10 * L * K + 2  # test_conditional_glow_network.jl:46
    grows = _grad_rows(ps, G64.get_params(), G32.get_params(), cglow_names(L, K))
    last = (L - 1) * K + (K - 1)
    strict = [2 * last, 2 * last + 1] + [2 * L * K + 2 + 8 * last + k for k in range(8)]
    _table(f"cfg3 [{precision}]", rows + [grows[i] for i in strict])
    st = _check_grads(grows, precision, fr, strict)
    print(f"  all {len(grows)} gradients: median CUDA {st['median_cuda']:.3e} | float32 oracle {st['median_oracle_f32']:.3e}; "
          f"worst CUDA {st['max_cuda']:.3e} | float32 oracle {st['max_oracle_f32']:.3e}")
    _dump(f"cfg3_{precision}", {"config": "cfg3", "precision": precision, "shape": [B, n_in, 64, 64],
                                "columns": ["tensor", "cuda_vs_f64", "oracle_f32_vs_f64"],
                                "rows": [list(r) for r in rows + grows], "fragile_relu_units": [fr.count, fr.total]})
