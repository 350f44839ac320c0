"""CPU tests of the HINT oracle (oracle/hint_oracle.py): every property the reference's own tests assert for this family
(test_squeeze.jl:15-31, test_coupling_layer_hint.jl:17-31, test_multiscale_hint_network.jl:25-66), float64 autograd for the
hand-derived backward, and the C-ABI plan bookkeeping (get_params order / sizes) without a GPU."""
import ctypes

import pytest
import torch

from _util import O, rel
from oracle import hint_oracle as H

import inb200
from inb200 import lib as L

DT = torch.float64


def test_haar_and_wavelet_squeeze_properties():
    torch.manual_seed(11)
    X = torch.randn(4, 2, 28, 28, dtype=DT)  # test_squeeze.jl:7
    for sq, unsq in ((H.wavelet_squeeze, H.wavelet_unsqueeze), (H.haar_squeeze, H.inv_haar_unsqueeze)):
        Y = sq(X)
        assert Y.shape == (4, 8, 14, 14)
        assert abs(Y.norm() - X.norm()) < 1e-12 * X.norm()          # :16 orthogonal transform
        assert rel(unsq(Y), X) < 1e-14                               # :19-22 invertibility
        Yr = torch.randn_like(Y)
        a, b = torch.sum(Yr * sq(X)), torch.sum(X * unsq(Yr))        # :25-31 adjoint
        assert abs(a / b - 1) < 1e-12
    # the two squeezes are the same four butterflies in another channel order:
    # Haar_squeeze channel q*C + c with q = (a, v, h, d) <-> wavelet channel 4c + (0, 2, 1, 3)
    Yw = H.wavelet_squeeze(X).reshape(4, 2, 4, 14, 14)
    Yh = torch.cat([Yw[:, :, q] for q in (0, 2, 1, 3)], dim=1)
    assert rel(H.haar_squeeze(X), Yh) < 1e-14
    # a constant image has only the approximation band, a left-to-right step only the x detail, positive for
    # a falling edge (detail = first - second, dimensionality_operations.jl:272-281)
    Xc = torch.zeros(1, 1, 4, 4, dtype=DT)
    Xc[..., 0::2] = 1.0  # x even = 1, x odd = 0
    Yc = H.wavelet_squeeze(Xc)
    assert torch.allclose(Yc[0, 0], torch.full((2, 2), 1.0, dtype=DT)) and torch.allclose(Yc[0, 1], torch.full((2, 2), 1.0, dtype=DT))
    assert Yc[0, 2].abs().max() == 0 and Yc[0, 3].abs().max() == 0
    with pytest.raises(ValueError, match="multiple of 2"):
        H.wavelet_squeeze(torch.zeros(1, 1, 5, 4))


def test_get_depth():
    assert [H.get_depth(c) for c in (2, 4, 8, 16, 32, 64)] == [1, 1, 2, 3, 4, 5]  # hint.jl:63-71
    assert [inb200.get_depth(c) for c in (2, 4, 8, 16, 32, 64)] == [1, 1, 2, 3, 4, 5]


def _autograd_check(layer_params, run_forward, X0, hand):
    """hand = (f, dX, grads) from the hand-derived backward; compare with autograd of the forward."""
    leaves = [p.data.clone().requires_grad_(True) for p in layer_params]
    for p, l in zip(layer_params, leaves):
        p.data = l
    Xl = X0.clone().requires_grad_(True)
    f = run_forward(Xl)
    ga = torch.autograd.grad(f, [Xl] + leaves)
    for p, l in zip(layer_params, leaves):
        p.data = l.detach()
    assert abs(f.item() - hand[0].item()) < 1e-10 * abs(f.item())
    assert rel(hand[1], ga[0]) < 2e-5  # the eps(Float32) of basic.jl:114 is part of the reference's inverse
    for g_hand, g_auto in zip(hand[2], ga[1:]):
        assert rel(g_hand, g_auto) < 1e-4


@pytest.mark.parametrize("C", [4, 8, 16])
@pytest.mark.parametrize("permute", ["none", "full", "lower", "both"])
@pytest.mark.parametrize("logdet", [False, True])
def test_hint_coupling_invertibility_and_autograd(C, permute, logdet):
    torch.manual_seed(C)
    gen = torch.Generator().manual_seed(5)
    HL = H.make_hint_coupling(gen, C, 6, logdet=logdet, permute=permute, k2=1, p2=0, dtype=DT)
    X = torch.randn(2, C, 8, 8, dtype=DT)
    Y, ld = HL.forward(X)
    assert rel(HL.inverse(Y), X) < 1e-5                     # test_coupling_layer_hint.jl:26-27
    assert rel(HL.backward(0 * Y, Y)[1], X) < 1e-5          # :29-30

    def fwd(Xl):
        Yl, ldl = HL.forward(Xl)
        return 0.5 * torch.sum(Yl * Yl) / 2 - ldl if logdet else 0.5 * torch.sum(Yl * Yl)

    f = fwd(X)
    O.clear_grad(HL.params())
    dY = Y / 2 if logdet else Y
    dX, _ = HL.backward(dY, Y)
    _autograd_check(HL.params(), fwd, X, (f, dX, [p.grad for p in HL.params()]))


def test_shared_gradient_modes_differ_only_on_revisited_layers():
    gen = torch.Generator().manual_seed(5)
    mk = lambda mode: H.make_hint_coupling(torch.Generator().manual_seed(5), 16, 6, k2=1, p2=0, dtype=DT, shared_grads=mode)
    A, B = mk("sum"), mk("last")
    X = torch.randn(2, 16, 8, 8, dtype=DT, generator=gen)
    Y, _ = A.forward(X)
    dXa, _ = A.backward(Y, Y)
    dXb, _ = B.backward(Y, Y)
    assert torch.equal(dXa, dXb)
    for i, (p, q) in enumerate(zip(A.params(), B.params())):
        if i < 5:  # CL[1] is visited once
            assert torch.equal(p.grad, q.grad)
        else:      # CL[2], CL[3] are visited 2 and 4 times: "last" keeps one visit's contribution
            assert rel(q.grad, p.grad) > 1e-2


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("squeezer", ["wavelet", "haar"])
def test_multiscale_hint_network(split, squeezer):
    torch.manual_seed(11)
    net = H.NetworkMultiScaleHINT(2, 4, 2, 2, split_scales=split, k1=3, k2=1, p1=1, p2=0, dtype=DT, seed=1,
                                  squeezer=squeezer)  # test_multiscale_hint_network.jl:10-18
    X = torch.randn(3, 2, 16, 16, dtype=DT)
    Z, ld = net.forward(X)
    assert Z.numel() == X.numel()
    assert rel(net.backward(0 * Z, Z)[1], X) < 1e-3   # :29-30
    assert rel(net.inverse(Z), X) < 1e-3              # :33-35

    def fwd(Xl):
        Zl, ldl = net.forward(Xl)
        return 0.5 * torch.sum(Zl * Zl) / 3 - ldl

    O.clear_grad(net.get_params())  # the Householder gradients accumulate otherwise (conv1x1.jl:237-239)
    f, dX = H.hint_train_step(net, X)
    _autograd_check(net.get_params(), fwd, X, (f, dX, [p.grad for p in net.get_params()]))


@pytest.mark.parametrize("cfg", [
    dict(n_in=2, n_hidden=4, L=2, K=2, split_scales=0, k2=1, p2=0),    # test_multiscale_hint_network.jl:10-18
    dict(n_in=2, n_hidden=64, L=2, K=4, split_scales=0, nx=128, ny=128, batch=32),   # BASELINE configs[3]
    dict(n_in=2, n_hidden=8, L=3, K=2, split_scales=1),
    dict(n_in=1, n_hidden=8, L=2, K=1, split_scales=0),
])
def test_hint_plan_parameter_order_and_sizes(cfg):
    base = dict(nx=16, ny=16, n_in=2, n_hidden=4, L=2, K=2, batch=2, split_scales=0, k1=3, k2=3, p1=1, p2=1,
                sig_low=0.0, sig_high=1.0, squeeze_type=0, shared_grads=0, precision=0)
    base.update(cfg)
    d = L.HintDesc(*[base[f[0]] for f in L.HintDesc._fields_])
    plan = L.P()
    L.call("inb_hint_plan_create", ctypes.byref(d), ctypes.byref(plan))
    net = H.NetworkMultiScaleHINT(base["n_in"], base["n_hidden"], base["L"], base["K"], split_scales=bool(base["split_scales"]),
                                  k1=base["k1"], k2=base["k2"], p1=base["p1"], p2=base["p2"])
    want = []
    for row in net.AN:
        for an in row:
            want += [an.k, an.k]
    for row in net.CL:
        for cl in row:
            want += [p.data.numel() for p in cl.params()]
    lib = L.load()
    n = ctypes.c_longlong()
    got = []
    for i in range(lib.inb_hint_num_params(plan)):
        assert lib.inb_hint_param_numel(plan, i, ctypes.byref(n)) == 0
        got.append(n.value)
    assert got == want
    assert lib.inb_hint_workspace_bytes(plan) > 0
    L.call("inb_hint_plan_destroy", plan)


def test_hint_plan_errors():
    def mk(**kw):
        base = dict(nx=16, ny=16, n_in=2, n_hidden=4, L=2, K=2, batch=2, split_scales=0, k1=3, k2=3, p1=1, p2=1,
                    sig_low=0.0, sig_high=1.0, squeeze_type=0, shared_grads=0, precision=0)
        base.update(kw)
        d = L.HintDesc(*[base[f[0]] for f in L.HintDesc._fields_])
        plan = L.P()
        L.call("inb_hint_plan_create", ctypes.byref(d), ctypes.byref(plan))
    with pytest.raises(L.InbError, match="multiple of 2"):
        mk(nx=18, L=2)   # 18 -> 9: odd at the second squeeze (dimensionality_operations.jl:82-84)
    with pytest.raises(L.InbError, match="cannot be halved"):
        mk(n_in=3)       # 12 channels: Int(12/8) is inexact (hint.jl:86)
    with pytest.raises(L.InbError, match="padding"):
        mk(p2=0)


def test_hint_oracle_matches_golden():
    """tests/golden/hint_small.npz (make_golden.py: float64 oracle on seeded inputs) pins the HINT oracle against drift;
    oracle-generated, not reference-generated ('parity unpinned', oracle/hint_oracle.py)."""
    import os
    import numpy as np
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hint_small.npz"))
    n_in, nh, Ls, K = [int(v) for v in gold["hint_cfg"]]
    N = H.NetworkMultiScaleHINT(n_in, nh, Ls, K, split_scales=True, k2=1, p2=0, seed=7, dtype=DT)
    ps = N.get_params()
    for i, p in enumerate(ps):
        p.data = torch.from_numpy(gold[f"hint_p{i:03d}"]).double()
    X = torch.from_numpy(gold["hint_X"]).double()
    Z, ld = N.forward(X)
    dX, Xr = N.backward(Z / X.shape[0], Z)
    assert rel(Z, torch.from_numpy(gold["hint_Z"])) < 1e-12
    assert abs(float(ld) - float(gold["hint_logdet"])) < 1e-10 * abs(float(gold["hint_logdet"]))
    assert rel(dX, torch.from_numpy(gold["hint_dX"])) < 1e-10
    assert rel(Xr, X) < 1e-4
    for i, p in enumerate(ps):
        assert rel(p.grad, torch.from_numpy(gold[f"hint_g{i:03d}"])) < 1e-10, i


@pytest.mark.parametrize("logdet", [True, False])
def test_basic_coupling_properties(logdet):
    """test_coupling_layer_basic.jl:12-66: forward -> inverse, forward -> backward (recompute), inverse -> forward, all to
    atol 1e-2 there; plus float64 autograd of the hand-derived backward (:124-149)."""
    torch.manual_seed(11)
    gen = torch.Generator().manual_seed(11)
    n_in, nh = 2, 4
    RB = O.ResidualBlock(O.glorot_uniform(gen, nh, n_in, 3, 3, dtype=DT), O.glorot_uniform(gen, nh, nh, 3, 3, dtype=DT),
                         O.glorot_uniform(gen, nh, 2 * n_in, 3, 3, dtype=DT), torch.zeros(nh, dtype=DT),
                         torch.zeros(nh, dtype=DT), p1=1, p2=1)
    L = H.CouplingLayerBasic(RB, logdet=logdet)
    Xa, Xb = torch.randn(1, n_in, 24, 24, dtype=DT), torch.randn(1, n_in, 24, 24, dtype=DT)
    Ya, Yb, ld = L.forward(Xa, Xb)
    assert torch.equal(Ya, Xa)
    assert rel(L.inverse(Ya, Yb)[1], Xb) < 1e-5
    assert rel(L.backward(0 * Ya, 0 * Yb, Ya, Yb)[3], Xb) < 1e-5
    Ya2, Yb2, _ = L.inverse(Xa, Xb)
    assert rel(L.forward(Ya2, Yb2)[1], Xb) < 1e-5

    def fwd(leaves):
        Xa_l, Xb_l = leaves
        Y1, Y2, ldl = L.forward(Xa_l, Xb_l)
        f = 0.5 * (torch.sum(Y1 * Y1) + torch.sum(Y2 * Y2))
        return f - ldl if logdet else f

    Xal, Xbl = Xa.clone().requires_grad_(True), Xb.clone().requires_grad_(True)
    ps = L.params()
    leaves = [p.data.clone().requires_grad_(True) for p in ps]
    for p, l in zip(ps, leaves):
        p.data = l
    ga = torch.autograd.grad(fwd((Xal, Xbl)), [Xal, Xbl] + leaves)
    for p, l in zip(ps, leaves):
        p.data = l.detach()
    dXa, dXb, _, _ = L.backward(Ya.clone(), Yb.clone(), Ya, Yb)
    assert rel(dXa, ga[0]) < 2e-5 and rel(dXb, ga[1]) < 2e-5
    for p, g_auto in zip(ps, ga[2:]):
        assert rel(p.grad, g_auto) < 1e-4


@pytest.mark.parametrize("split", [False, True])
def test_hint_host_mirror_param_shapes_match_oracle(split):
    """The Python mirror's flat parameter buffer (get_params order, torch layout) against the oracle's containers."""
    net = inb200.NetworkMultiScaleHINT(2, 6, 3, 2, split_scales=split, device="cpu")
    ora = H.NetworkMultiScaleHINT(2, 6, 3, 2, split_scales=split)
    ps, qs = net.get_params(), ora.get_params()
    assert len(ps) == len(qs)
    n_an = 2 * 3 * 2
    for i, (p, q) in enumerate(zip(ps, qs)):
        if i < n_an:
            assert q.data is None and p.data.numel() == ora.AN[i // 4][(i // 2) % 2].k
        else:
            assert tuple(p.data.shape) == tuple(q.data.shape), i
    for permute in ("none", "full", "lower", "both"):
        HL = inb200.CouplingLayerHINT(16, 5, permute=permute, device="cpu")
        OL = H.make_hint_coupling(torch.Generator().manual_seed(0), 16, 5, permute=permute)
        assert [tuple(p.data.shape) for p in HL.get_params()] == [tuple(p.data.shape) for p in OL.params()]
