"""GPU parity of the HINT family (BASELINE configs[3]) through the C ABI against
oracle/hint_oracle.py: Haar / wavelet squeezes, CouplingLayerHINT (recursive, permute none / full / lower),
NetworkMultiScaleHINT - outputs, logdet, recomputed X, dX and every gradient - plus the reference's own
properties (test_squeeze.jl, test_coupling_layer_hint.jl:17-31, test_multiscale_hint_network.jl:25-35)."""
import pytest
import torch

from _util import O, TOL_GRAD, TOL_LOGDET, TOL_OUT, FragileUnits, assert_grad_close, rel
from oracle import hint_oracle as H

import inb200

pytestmark = pytest.mark.gpu
DEV = "cuda"


def g(t):
    return t.to(DEV)


@pytest.mark.parametrize("shape", [(4, 2, 28, 28), (2, 3, 6, 10), (1, 8, 64, 32)])
def test_haar_squeezes_match_oracle_and_roundtrip(shape):
    torch.manual_seed(11)
    X = torch.randn(*shape)
    for sq, unsq, osq, ounsq in ((inb200.wavelet_squeeze, inb200.wavelet_unsqueeze, H.wavelet_squeeze, H.wavelet_unsqueeze),
                                 (inb200.Haar_squeeze, inb200.invHaar_unsqueeze, H.haar_squeeze, H.inv_haar_unsqueeze)):
        Y = sq(g(X))
        assert rel(Y, osq(X.double())) < 1e-6
        assert abs(Y.norm().item() - X.norm().item()) < 1e-5 * X.norm().item()   # test_squeeze.jl:16
        assert rel(unsq(Y), X) < 1e-6                                            # :19-22 (reference: 1f-5)
        Yr = torch.randn_like(Y)
        assert rel(unsq(Yr), ounsq(Yr.cpu().double())) < 1e-6
        a, b = torch.sum(Yr * Y).item(), torch.sum(g(X) * unsq(Yr)).item()       # :25-31 adjoint
        assert abs(a / b - 1) < 1e-4


def test_haar_squeeze_errors():
    with pytest.raises(inb200.InbError, match="multiple of 2"):
        inb200.wavelet_squeeze(g(torch.randn(1, 1, 5, 4)))
    with pytest.raises(inb200.InbError, match="divisible by 4"):
        inb200.wavelet_unsqueeze(g(torch.randn(1, 3, 4, 4)))


def _load(layer, oracle_params):
    inb200.set_params(layer, [p.data.float() for p in oracle_params])


def run_hint_layer(C, nh, shape, *, permute, logdet, k2, shared="sum", precision="fp32", tol_out=TOL_OUT,
                   tol_grad=TOL_GRAD, fragile_thr=1e-6):
    torch.manual_seed(C + k2)
    p2 = (k2 - 1) // 2
    mk = lambda dt: H.make_hint_coupling(torch.Generator().manual_seed(5), C, nh, logdet=logdet, permute=permute, k2=k2,
                                         p2=p2, dtype=dt, shared_grads=shared)
    H32 = mk(torch.float32)
    H64 = mk(torch.float64)
    for p, q in zip(H64.params(), H32.params()):
        p.data = q.data.double()
    HL = inb200.CouplingLayerHINT(C, nh, logdet=logdet, permute=permute, k2=k2, p2=p2, shared_grads=shared,
                                  precision=precision, device=DEV)
    _load(HL, H32.params())
    X = torch.randn(*shape)
    Y64, ld64 = H64.forward(X.double())
    out = HL.forward(g(X))
    Y = out[0] if logdet else out
    assert rel(Y, Y64) < tol_out
    if logdet:
        assert abs(out[1].item() - ld64.item()) / abs(ld64.item()) < TOL_LOGDET * (tol_out / TOL_OUT)
    assert rel(HL.inverse(Y), X) < max(1e-5, tol_out)        # test_coupling_layer_hint.jl:26-27
    Yin = Y64.float()
    dY = torch.randn_like(Yin)
    dX, Xr = HL.backward(g(dY), g(Yin))
    with FragileUnits(fragile_thr) as fr:
        dX64, X64 = H64.backward(dY.double(), Yin.double())
    assert rel(Xr, X64) < tol_out
    assert_grad_close(dX, dX64, tol_out, fr, "dX")
    for i, (p, q) in enumerate(zip(HL.get_params(), H64.params())):
        assert_grad_close(p.grad, q.grad, tol_grad, fr, f"gradient {i}")


@pytest.mark.parametrize("C", [2, 4, 8, 16])
@pytest.mark.parametrize("permute", ["none", "full", "lower", "both"])
def test_hint_coupling_parity(C, permute):
    if permute == "lower" and C == 2:
        pytest.skip("Conv1x1 on one channel")
    run_hint_layer(C, 8, (2, C, 16, 16), permute=permute, logdet=True, k2=1)


@pytest.mark.parametrize("logdet", [False, True])
def test_hint_coupling_parity_default_kernels_and_last_visit_gradients(logdet):
    # the constructor's defaults k1 = k2 = 3 (hint.jl:78); "last" = what set_grad=true leaves in .grad
    run_hint_layer(16, 8, (2, 16, 12, 8), permute="full", logdet=logdet, k2=3, shared="last")
    run_hint_layer(32, 6, (1, 32, 8, 8), permute="none", logdet=logdet, k2=3)


def test_hint_coupling_errors():
    with pytest.raises(inb200.InbError, match="cannot be halved"):
        HL = inb200.CouplingLayerHINT(8, 4, device=DEV)
        HL.n_in = 12
        HL.forward(g(torch.randn(1, 12, 4, 4)))
    with pytest.raises(inb200.InbError, match="permute"):
        inb200.CouplingLayerHINT(8, 4, permute="upper", device=DEV)


def run_hint_network(n_in, nh, L, K, shape, *, split, k2=1, squeezer="wavelet", precision="fp32", tol_out=TOL_OUT,
                     tol_grad=TOL_GRAD, fragile_thr=1e-6, inv_tol=1e-5):
    torch.manual_seed(11)
    p2 = (k2 - 1) // 2
    mk = lambda dt: H.NetworkMultiScaleHINT(n_in, nh, L, K, split_scales=split, k2=k2, p2=p2, seed=3, dtype=dt,
                                            squeezer=squeezer)
    N32, N64 = mk(torch.float32), mk(torch.float64)
    X = torch.randn(*shape)
    Z32, _ = N32.forward(X)  # data-dependent ActNorm init in float32 defines the parameters of both sides
    for p, q in zip(N64.get_params(), N32.get_params()):
        p.data = q.data.double()
    net = inb200.NetworkMultiScaleHINT(n_in, nh, L, K, split_scales=split, k2=k2, p2=p2, squeezer=squeezer,
                                       precision=precision, device=DEV)
    inb200.set_params(net, [p.data for p in N32.get_params()])
    Z64, ld64 = N64.forward(X.double())
    Z, ld = net.forward(g(X))
    assert Z.shape == Z64.shape
    assert rel(Z, Z64) < max(tol_out, 2 * rel(Z32, Z64))
    assert abs(ld.item() - ld64.item()) / abs(ld64.item()) < TOL_LOGDET * (tol_out / TOL_OUT)
    inv_cuda, inv_oracle = rel(net.inverse(Z), X), rel(N32.inverse(Z32), X)
    assert inv_cuda < max(inv_tol, 2 * inv_oracle)  # reference bound rtol 1f-3 (test_multiscale_hint_network.jl:30,35)
    Zin = Z64.float()
    dZ = Zin / shape[0]
    dX, Xr = net.backward(g(dZ), g(Zin))
    with FragileUnits(fragile_thr) as fr:
        dX64, X64 = N64.backward(dZ.double(), Zin.double())
    # two columns: the CUDA path against truth beside the float32 oracle against truth.  The sigmoid scale
    # of CouplingLayerBasic reaches down to 0 (SigmoidLayer(), low = 0), so inverting 2^depth couplings per layer
    # amplifies float32 rounding - the reference's own bound on the recomputed X is rtol 1f-3 (:30); the CUDA path
    # must stay within 3x of what the reference's float32 arithmetic shows on the same numbers.
    dX32, X32 = N32.backward(dZ, Zin)
    assert rel(Xr, X64) < max(tol_out, 3 * rel(X32, X64))
    assert_grad_close(dX, dX64, max(tol_out, 3 * rel(dX32, dX64)), fr, "dX")
    ps, qs, q32 = net.get_params(), N64.get_params(), N32.get_params()
    assert len(ps) == len(qs) and all(p.grad is not None for p in ps)
    for i, (p, q, r) in enumerate(zip(ps, qs, q32)):
        assert_grad_close(p.grad, q.grad, max(tol_grad, 3 * rel(r.grad, q.grad)), fr, f"gradient {i}")
    inb200.clear_grad(net)
    assert all(p.grad is None for p in net.get_params())


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("squeezer", ["wavelet", "haar"])
def test_multiscale_hint_reference_test_shape(split, squeezer):
    # test_multiscale_hint_network.jl:10-18: 64x64, n_in=2, n_hidden=4, L=K=2, k1=3, k2=1
    run_hint_network(2, 4, 2, 2, (2, 2, 64, 64), split=split, squeezer=squeezer)


def test_multiscale_hint_default_kernels_three_scales():
    run_hint_network(1, 8, 3, 1, (3, 1, 32, 32), split=True, k2=3)


def test_multiscale_hint_actnorm_init_inside_forward():
    torch.manual_seed(2)
    X = torch.randn(4, 2, 32, 32) * 2 + 0.5
    net = inb200.NetworkMultiScaleHINT(2, 4, 2, 2, k2=1, p2=0, device=DEV, seed=7)
    N32 = H.NetworkMultiScaleHINT(2, 4, 2, 2, k2=1, p2=0)
    for p, q in zip(N32.get_params(), net.get_params()):
        if tuple(q.data.shape) and p.data is not None:
            p.data = q.data.cpu().clone()
    Z, ld = net.forward(g(X))           # first forward: initialises s, b layer by layer (actnorm.jl:67-72)
    Z32, ld32 = N32.forward(X)
    assert rel(Z, Z32) < 1e-4 and abs(ld.item() - ld32.item()) < 1e-4 * abs(ld32.item())
    for p, q in zip(net.get_params(), N32.get_params()):
        assert rel(p.data, q.data) < 1e-4


def test_multiscale_hint_cfg4_channel_plan_tensor_cores():
    # BASELINE configs[3] channel plan (n_in = 2 -> 8 and 32 channels) with n_hidden = 128 on the tcgen05 path
    # (k2 = 1: fused chain; bf16x3 = float32-equivalent arithmetic), reduced spatial size and batch
    run_hint_network(2, 128, 2, 1, (2, 2, 64, 64), split=False, k2=1, precision="bf16x3", fragile_thr=1e-4,
                     inv_tol=1e-4)


def test_multiscale_hint_cfg4_full_size_properties():
    # BASELINE configs[3]: 128x128x2, batch 32 - size-independent properties at full size
    torch.manual_seed(0)
    B = 32
    net = inb200.NetworkMultiScaleHINT(2, 64, 2, 4, device=DEV, seed=1)   # default 3x3 kernels, fp32 path
    X = torch.randn(B, 2, 128, 128, device=DEV)
    Z, ld = net.forward(X)
    assert Z.shape == (B, 32, 32, 32) and torch.isfinite(Z).all() and torch.isfinite(ld)
    assert rel(net.inverse(Z), X) < 1e-3                       # test_multiscale_hint_network.jl:33-35
    dX, Xr = net.backward(Z / B, Z)
    assert rel(Xr, X) < 1e-3                                   # :29-30
    assert all(torch.isfinite(p.grad).all() for p in net.get_params())
    # ActNorm init statistics of the first layer (test_actnorm.jl:57-58) seen through the first squeeze
    Y = inb200.wavelet_squeeze(X)
    s, b = net.get_params()[0].data, net.get_params()[1].data
    Yn = Y * s.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)
    assert Yn.mean(dim=(0, 2, 3)).abs().max() < 1e-4 and (Yn.var(dim=(0, 2, 3)) - 1).abs().max() < 1e-3
    # adjoint test of the backward pass at full size: backward is affine in dZ, dX(dZ) = J' dZ + c (c = the logdet
    # gradient), so <dX(dZ) - dX(0), d> = <dZ, J d>; J d from a central difference of the forward pass and dZ := J d,
    # which makes both sides ||J d||^2 (no cancellation between random vectors).  (A Taylor test of the scalar loss,
    # :46-66, is limited here by the float32 logdet: 1e-7 * |logdet| / h.)
    d = torch.randn_like(X)
    h = 1e-3
    Jd = (net.forward(X + h * d)[0].double() - net.forward(X - h * d)[0].double()) / (2 * h)
    dZr = Jd.float()
    dX0 = net.backward(0 * dZr, Z)[0]
    dX1 = net.backward(dZr, Z)[0]
    lhs = torch.sum((dX1 - dX0).double() * d.double()).item()
    rhs = torch.sum(Jd * Jd).item()
    assert abs(lhs - rhs) < 1e-2 * rhs, (lhs, rhs)
    # and it is linear in dZ beyond the constant: dX(2 dZ) - dX(dZ) = dX(dZ) - dX(0)
    assert rel(net.backward(2 * dZr, Z)[0] - dX1, dX1 - dX0) < 1e-4


def test_cuda_hint_matches_golden():
    """The CUDA path against the committed oracle vectors (tests/golden/hint_small.npz)."""
    import os
    import numpy as np
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hint_small.npz"))
    n_in, nh, Ls, K = [int(v) for v in gold["hint_cfg"]]
    net = inb200.NetworkMultiScaleHINT(n_in, nh, Ls, K, split_scales=True, k2=1, p2=0, device=DEV)
    nparam = sum(1 for k in gold.files if k.startswith("hint_p"))
    inb200.set_params(net, [torch.from_numpy(gold[f"hint_p{i:03d}"]) for i in range(nparam)])
    X = torch.from_numpy(gold["hint_X"])
    Z, ld = net.forward(g(X))
    assert rel(Z, torch.from_numpy(gold["hint_Z"])) < TOL_OUT
    assert abs(ld.item() - float(gold["hint_logdet"])) < TOL_LOGDET * abs(float(gold["hint_logdet"])) + 1e-4
    dX, Xr = net.backward(Z / X.shape[0], Z)
    assert rel(Xr, X) < 1e-3
    assert rel(dX, torch.from_numpy(gold["hint_dX"])) < 10 * TOL_OUT
    for i, p in enumerate(net.get_params()):
        assert rel(p.grad, torch.from_numpy(gold[f"hint_g{i:03d}"])) < 10 * TOL_GRAD, i


@pytest.mark.parametrize("C1,k2,logdet", [(2, 3, True), (4, 1, True), (6, 3, False)])
def test_basic_coupling_parity(C1, k2, logdet):
    """CouplingLayerBasic (invertible_layer_basic.jl:90-149) against the oracle; invertibility of
    test_coupling_layer_basic.jl (forward -> inverse, forward -> backward recompute)."""
    torch.manual_seed(C1)
    nh, p2 = 8, (k2 - 1) // 2
    gen = torch.Generator().manual_seed(9)
    ws = [O.glorot_uniform(gen, nh, C1, 3, 3), O.glorot_uniform(gen, nh, nh, k2, k2), O.glorot_uniform(gen, nh, 2 * C1, 3, 3),
          0.1 * torch.randn(nh), 0.1 * torch.randn(nh)]
    L64 = H.CouplingLayerBasic(O.ResidualBlock(*[w.double() for w in ws], p1=1, p2=p2), logdet=logdet)
    L = inb200.CouplingLayerBasic(C1, nh, k2=k2, p2=p2, logdet=logdet, device=DEV)
    inb200.set_params(L, ws)
    X1, X2 = torch.randn(3, C1, 12, 10), torch.randn(3, C1, 12, 10)
    _, Y2_64, ld64 = L64.forward(X1.double(), X2.double())
    out = L.forward(g(X1), g(X2))
    assert rel(out[1], Y2_64) < TOL_OUT
    if logdet:
        assert abs(out[2].item() - ld64.item()) < TOL_LOGDET * abs(ld64.item()) + 1e-6
    assert rel(L.inverse(g(X1), out[1])[1], X2) < 1e-5
    Y2 = Y2_64.float()
    dY1, dY2 = torch.randn_like(X1), torch.randn_like(X2)
    dX1, dX2, _, X2r = L.backward(g(dY1), g(dY2), g(X1), g(Y2))
    with FragileUnits(1e-6) as fr:
        dX1_64, dX2_64, _, X2_64 = L64.backward(dY1.double(), dY2.double(), X1.double(), Y2.double())
    assert rel(X2r, X2_64) < TOL_OUT and rel(dX2, dX2_64) < TOL_OUT
    assert_grad_close(dX1, dX1_64, TOL_OUT, fr, "dX1")
    for i, (p, q) in enumerate(zip(L.get_params(), L64.params())):
        assert_grad_close(p.grad, q.grad, TOL_GRAD, fr, f"gradient {i}")
