"""GPU parity, network level: NetworkGlow / NetworkConditionalGlow through the C ABI against the
oracle (outputs, logdet, recomputed X, dX and every one of the 10*L*K gradients), plus the
reference's own network properties (test_glow.jl:20-66, test_conditional_glow_network.jl:31-46)
and size-independent properties at larger sizes."""
import pytest
import torch

from _util import O, TOL_GRAD, TOL_LOGDET, TOL_OUT, FragileUnits, assert_grad_close, clone_oracle, rel

import inb200

pytestmark = pytest.mark.gpu
DEV = "cuda"


def g(t):
    return t.to(DEV)


def run_glow_parity(n_in, nh, L, K, shape, *, logdet=True, split=True, ndims=2, tol_out=TOL_OUT,
                    tol_grad=TOL_GRAD, precision="fp32", seed=11, fragile_thr=1e-6, inv_tol=1e-5):
    torch.manual_seed(seed)
    mk = lambda dt: O.NetworkGlow(n_in, nh, L, K, logdet=logdet, split_scales=split, ndims=ndims, seed=3, dtype=dt,
                                  faithful=False)
    G32, G64 = mk(torch.float32), clone_oracle(mk, torch.float64)
    X = torch.rand(*shape)
    # float32 oracle first: its data-dependent ActNorm init defines the parameters both sides use
    o32 = G32.forward(X)
    for p, q in zip(G64.get_params(), G32.get_params()):
        p.data = q.data.double()
    G = inb200.NetworkGlow(n_in, nh, L, K, logdet=logdet, split_scales=split, ndims=ndims, precision=precision,
                           device=DEV)
    inb200.set_params(G, [p.data for p in G32.get_params()])
    o64 = G64.forward(X.double())
    out = G.forward(g(X))
    Z, Z64, Z32 = (out[0], o64[0], o32[0]) if logdet else (out, o64, o32)
    assert Z.shape == Z64.shape
    e32 = rel(Z32, Z64)
    assert rel(Z, Z64) < max(tol_out, 2 * e32)
    if logdet:
        assert abs(out[1].item() - o64[1].item()) / abs(o64[1].item()) < TOL_LOGDET * (tol_out / TOL_OUT)
    if split:
        assert [tuple(z) for z in G.Z_dims] == [tuple(z) for z in G64.Z_dims]
    # invertibility, reference bound 1f-5 (test_glow.jl:46)
    Xi = G.inverse(Z)
    inv_cuda, inv_oracle = rel(Xi, X), rel(G32.inverse(Z32), X)
    assert inv_cuda < max(inv_tol, 2 * inv_oracle)
    # backward from the same (float64-rounded-to-float32) latent
    Zin = Z64.float()
    dZ = Zin / shape[0]
    dX, Xr = G.backward(g(dZ), g(Zin))
    with FragileUnits(fragile_thr) as fr:
        dX64, X64 = G64.backward(dZ.double(), Zin.double())
    assert rel(Xr, X64) < tol_out
    assert_grad_close(dX, dX64, tol_out, fr, "dX")
    ps, qs = G.get_params(), G64.get_params()
    assert sum(p.grad is not None for p in ps) == 10 * L * K  # test_glow.jl:50-62
    worst = 0.0
    for i, (p, q) in enumerate(zip(ps, qs)):
        worst = max(worst, rel(p.grad, q.grad))
        assert_grad_close(p.grad, q.grad, tol_grad, fr, f"gradient {i}")
    inb200.clear_grad(G)
    assert all(p.grad is None for p in G.get_params())  # test_glow.jl:64-66
    return worst


@pytest.mark.parametrize("logdet", [True, False])
@pytest.mark.parametrize("split", [True, False])
def test_glow_2d_reference_test_shape(logdet, split):
    # test_glow.jl:20-35: 32x32, n_in=2, n_hidden=4, B=2, L=K=2
    run_glow_parity(2, 4, 2, 2, (2, 2, 32, 32), logdet=logdet, split=split)


@pytest.mark.parametrize("split", [True, False])
def test_glow_3d_reference_test_shape(split):
    run_glow_parity(2, 4, 2, 2, (2, 2, 16, 16, 16), split=split, ndims=3)


def test_glow_cfg1_full_size():
    # BASELINE configs[0]: NetworkGlow(1, 32, 2, 2) on 64x64x1, batch 8 (examples/networks/network_glow.jl)
    run_glow_parity(1, 32, 2, 2, (8, 1, 64, 64))


def test_glow_cfg2_channel_plan_small():
    # cfg2's channel plan (3 -> 12/24/48, L=3) at a size the oracle finishes in seconds
    run_glow_parity(3, 16, 3, 2, (2, 3, 32, 32))


def test_glow_L1_splits_once():
    run_glow_parity(2, 4, 1, 2, (2, 2, 8, 8))  # invertible_network_glow.jl:120 (i == 1)


def test_glow_cfg5_3d_small():
    run_glow_parity(1, 8, 2, 2, (2, 1, 16, 16, 16), ndims=3)


def test_householder_grads_accumulate_unless_cleared():
    G = inb200.NetworkGlow(2, 4, 1, 1, split_scales=True, device=DEV)
    X = g(torch.rand(2, 2, 8, 8))
    Z, _ = G.forward(X)
    G.backward(Z / 2, Z)
    g1 = [p.grad.clone() for p in G.get_params()]
    G.backward(Z / 2, Z)  # no clear_grad!: conv1x1.jl:237-239 accumulates v grads, the rest overwrite
    for i, p in enumerate(G.get_params()):
        want = 2 * g1[i] if 2 <= i < 5 else g1[i]
        assert rel(p.grad, want) < 1e-5


def test_frozen_conv_gives_zero_householder_grads():
    G = inb200.NetworkGlow(2, 4, 1, 2, split_scales=True, freeze_conv=True, device=DEV)
    X = g(torch.rand(2, 2, 8, 8))
    Z, _ = G.forward(X)
    G.backward(Z / 2, Z)
    for i in G._hh_idx:
        assert G.get_params()[i].grad.abs().max().item() == 0.0


@pytest.mark.parametrize("split", [True, False])
def test_conditional_glow_parity(split):
    torch.manual_seed(21)
    n_in, n_cond, nh, L, K, B = (1, 1, 8, 2, 3, 4) if split else (4, 2, 8, 2, 2, 3)
    mk = lambda dt: O.NetworkConditionalGlow(n_in, n_cond, nh, L, K, split_scales=split, seed=5, dtype=dt,
                                             faithful=False)
    G32, G64 = mk(torch.float32), clone_oracle(mk, torch.float64)
    X, Cn = torch.rand(B, n_in, 16, 16), torch.rand(B, n_cond, 16, 16)
    ZX32, ZC32, ld32 = G32.forward(X, Cn)
    for p, q in zip(G64.get_params(), G32.get_params()):
        p.data = q.data.double()
    ZX64, ZC64, ld64 = G64.forward(X.double(), Cn.double())
    G = inb200.NetworkConditionalGlow(n_in, n_cond, nh, L, K, split_scales=split, device=DEV)
    inb200.set_params(G, [p.data for p in G32.get_params()])
    ZX, ZC, ld = G.forward(g(X), g(Cn))
    assert ZX.shape == X.shape and ZC.shape == ZC64.shape
    assert rel(ZX, ZX64) < max(TOL_OUT, 2 * rel(ZX32, ZX64)) and rel(ZC, ZC64) < TOL_OUT
    assert abs(ld.item() - ld64.item()) / abs(ld64.item()) < TOL_LOGDET
    assert rel(G.inverse(ZX, ZC), X) < 1e-5  # test_conditional_glow_network.jl:35
    Zin, ZCin = ZX64.float(), ZC64.float()
    dX, Xr, dC = G.backward(g(Zin / B), g(Zin), g(ZCin))
    dX64, X64, dC64 = G64.backward((Zin / B).double(), Zin.double(), ZCin.double())
    assert rel(dX, dX64) < TOL_OUT and rel(Xr, X64) < TOL_OUT and rel(dC, dC64) < TOL_OUT
    ps = G.get_params()
    assert sum(p.grad is not None for p in ps) == 10 * L * K + 2  # test_conditional_glow_network.jl:46
    for i, (p, q) in enumerate(zip(ps, G64.get_params())):
        assert rel(p.grad, q.grad) < TOL_GRAD, f"gradient {i}"


def test_actnorm_init_inside_network_forward():
    # first forward initialises every ActNorm from the data (actnorm.jl:67-72): compare to the oracle
    torch.manual_seed(5)
    G32 = O.NetworkGlow(2, 4, 2, 2, split_scales=True, seed=3, faithful=False)
    X = torch.rand(4, 2, 16, 16)
    Z32, ld32 = G32.forward(X)
    G = inb200.NetworkGlow(2, 4, 2, 2, split_scales=True, device=DEV)
    # copy only the non-ActNorm parameters; s, b stay unset
    for p, q in zip(G.get_params()[8:], G32.get_params()[8:]):
        p.data.copy_(q.data)
    Z, ld = G.forward(g(X))
    for p, q in zip(G.get_params()[:8], G32.get_params()[:8]):
        assert rel(p.data, q.data) < 1e-4
    assert rel(Z, Z32) < 1e-4 and abs(ld.item() - ld32.item()) / abs(ld32.item()) < 1e-4


def test_full_size_properties_cfg2_one_sample_pair():
    """BASELINE configs[1] shape (256x256x3, L=3, n_hidden=256) at B=2, K=2: too big for the oracle, so
    size-independent properties: invertibility (test_glow.jl:46), backward's recomputed X equals the
    input, dX finite, gradient count."""
    torch.manual_seed(0)
    G = inb200.NetworkGlow(3, 256, 3, 2, split_scales=True, device=DEV)
    X = g(torch.rand(2, 3, 256, 256))
    Z, ld = G.forward(X)
    assert torch.isfinite(Z).all() and torch.isfinite(ld)
    assert rel(G.inverse(Z), X) < 1e-5
    dX, Xr = G.backward(Z / 2, Z)
    assert rel(Xr, X) < 1e-5 and torch.isfinite(dX).all()
    assert sum(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in G.get_params()) == 60


def test_adam_matches_flux_update_rule():
    """inb_adam_update over the flat parameter buffer == Flux.Optimise.ADAM applied parameter by parameter
    (examples/networks/network_glow.jl:38-42); restated in float64 for the check, three steps."""
    torch.manual_seed(0)
    G = inb200.NetworkGlow(2, 8, 2, 2, split_scales=True, device=DEV)
    X = torch.rand(4, 2, 16, 16)
    opt = inb200.ADAM(G, lr=1e-2)
    x = None
    m = v = None
    b1, b2, eps, lr = 0.9, 0.999, 1e-8, 1e-2
    for t in range(1, 4):
        Z, ld = G.forward(g(X))
        nll, dZ = inb200.nll_grad(Z, X.shape[0])
        G.backward(dZ, Z)
        gr = G.flat_grads.double().cpu()
        if x is None:
            x = G.flat_params.double().cpu()
            m, v = torch.zeros_like(x), torch.zeros_like(x)
        m = b1 * m + (1 - b1) * gr
        v = b2 * v + (1 - b2) * gr * gr
        x = x - lr * (m / (1 - b1 ** t)) / (torch.sqrt(v / (1 - b2 ** t)) + eps)
        opt.step()
        inb200.clear_grad(G)
        assert rel(G.flat_params, x) < 1e-6, t


def test_cuda_matches_golden():
    """The CUDA path against the committed oracle vectors (tests/golden/glow_small.npz): forward, logdet, dX and every
    gradient of a small NetworkGlow, float32 tolerance."""
    import os
    import numpy as np
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "glow_small.npz"))
    n_in, nh, L, K = [int(v) for v in gold["glow_cfg"]]
    G = inb200.NetworkGlow(n_in, nh, L, K, split_scales=True, device=DEV)
    nparam = sum(1 for k in gold.files if k.startswith("glow_p"))
    inb200.set_params(G, [torch.from_numpy(gold[f"glow_p{i:03d}"]) for i in range(nparam)])
    X = torch.from_numpy(gold["glow_X"])
    Z, ld = G.forward(g(X))
    assert rel(Z, torch.from_numpy(gold["glow_Z"])) < TOL_OUT
    assert abs(ld.item() - float(gold["glow_logdet"])) < TOL_LOGDET * abs(float(gold["glow_logdet"])) + 1e-4
    dZ = Z / X.shape[0]
    dX, Xr = G.backward(dZ, Z)
    assert rel(Xr, X) < 1e-5
    assert rel(dX, torch.from_numpy(gold["glow_dX"])) < 10 * TOL_OUT
    for i, p in enumerate(G.get_params()):
        assert rel(p.grad, torch.from_numpy(gold[f"glow_g{i:03d}"])) < 10 * TOL_GRAD, i


# Full-size parity of cfg2 / cfg3 / cfg5 against the float64 oracle: tests/test_gpu_fullsize.py.


def test_checkpoint_roundtrip_in_get_params_order(tmp_path):
    """save_params / load_params: a second network of the same architecture loaded from the file
    gives the same outputs; arrays are stored in the reference's axis order."""
    import numpy as np
    torch.manual_seed(4)
    X = torch.rand(2, 2, 16, 16, device=DEV)
    G = inb200.NetworkGlow(2, 8, 2, 2, split_scales=True, device=DEV, seed=1)
    Z, ld = G.forward(X)  # initialises ActNorm
    path = str(tmp_path / "net_params.npz")
    inb200.save_params(G, path)
    z = np.load(path)
    assert len(z.files) == len(G.get_params())
    W1 = G.get_params()[2 * 2 * 2 + 3].data  # CL[1,1].RB.W1, torch (nh, Cin, ky, kx)
    assert z[f"p{2 * 2 * 2 + 3:03d}"].shape == tuple(reversed(W1.shape))  # Julia (kx, ky, Cin, nh)
    G2 = inb200.NetworkGlow(2, 8, 2, 2, split_scales=True, device=DEV, seed=99)
    inb200.load_params(G2, path)
    Z2, ld2 = G2.forward(X)
    assert rel(Z2, Z) < 1e-6 and abs(ld2.item() - ld.item()) < 1e-6 * abs(ld.item())
