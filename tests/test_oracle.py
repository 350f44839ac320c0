"""Pins the CPU oracle (oracle/glow_oracle.py) against every property the reference's own tests
assert for the Glow path and against torch.autograd in float64.

The reference holds no golden vectors for this path, so these properties ARE its pins:
  test/test_layers/test_actnorm.jl:22-39,57-62   test/test_layers/test_layer_conv1x1.jl:49-61,80-91
  test/test_layers/test_coupling_layer_glow.jl:28-31   test/test_utils/test_squeeze.jl:11-13
  test/test_networks/test_glow.jl:46,50-66   test/test_networks/test_conditional_glow_network.jl:35,46
"""
import math

import pytest
import torch

from oracle import glow_oracle as O

torch.manual_seed(0)
D = torch.float64


def rel(a, b):
    return (torch.linalg.norm((a - b).reshape(-1)) / torch.linalg.norm(b.reshape(-1))).item()


# ---------------------------------------------------------------- squeeze / split / states
@pytest.mark.parametrize("shape", [(2, 3, 8, 6), (2, 2, 4, 6, 8)])
def test_squeeze_roundtrip_and_indexmap(shape):
    X = torch.randn(*shape)
    Y = O.squeeze(X)
    assert rel(O.unsqueeze(Y), X) < 1e-6  # test_squeeze.jl:11-13
    C = shape[1]
    # closed form of SURVEY §9.2: Y[b, p*C+c, y', x'] = X[b, c, 2y'+iy, 2x'+ix], p = ix+2iy(+4iz)
    if len(shape) == 4:
        for p in range(4):
            ix, iy = p % 2, p // 2
            assert torch.equal(Y[:, p * C:(p + 1) * C], X[:, :, iy::2, ix::2])
    else:
        for p in range(8):
            ix, iy, iz = p % 2, (p // 2) % 2, p // 4
            assert torch.equal(Y[:, p * C:(p + 1) * C], X[:, :, iz::2, iy::2, ix::2])


def test_squeeze_odd_throws():
    with pytest.raises(ValueError):
        O.squeeze(torch.randn(1, 1, 5, 4))


def test_split_ties_to_even():
    a, b = O.tensor_split(torch.randn(1, 3, 2, 2))
    assert a.shape[1] == 2 and b.shape[1] == 1  # round(1.5) = 2
    a, b = O.tensor_split(torch.randn(1, 5, 2, 2))
    assert a.shape[1] == 2 and b.shape[1] == 3  # round(2.5) = 2


# ---------------------------------------------------------------- ActNorm
def test_actnorm_init_invert_logdet():
    X = torch.randn(4, 3, 6, 5, dtype=D) * 3 + 1
    AN = O.ActNorm(3, logdet=True)
    Y, ld = AN.forward(X)
    assert abs(Y.mean(dim=(0, 2, 3))).max() < 1e-6  # test_actnorm.jl:57
    assert abs(Y.var(dim=(0, 2, 3), unbiased=True) - 1).max() < 1e-3  # :58
    assert rel(AN.inverse(Y), X) < 1e-6  # :61-62
    # explicit log|det J| per sample (test_actnorm.jl:22-39): J is diagonal with s_c per pixel
    assert abs(ld.item() - 30 * torch.log(AN.s.data.abs()).sum().item()) < 1e-9


def test_actnorm_backward_matches_autograd():
    X = torch.randn(3, 4, 5, 6, dtype=D, requires_grad=True)
    AN = O.ActNorm(4, logdet=True)
    with torch.no_grad():
        AN.forward(X)
    s = (AN.s.data * 1.3).requires_grad_(True)  # move off the init point (where Δb == 0)
    b = (AN.b.data + 0.2).requires_grad_(True)
    AN.s.data, AN.b.data = s, b
    Y, ld = AN.forward(X)
    f = 0.5 * (Y ** 2).sum() / 3 - ld
    gX, gs, gb = torch.autograd.grad(f, (X, s, b))
    with torch.no_grad():
        dX, X_ = AN.backward(Y / 3, Y)
    assert rel(dX, gX) < 1e-12 and rel(X_, X.detach()) < 1e-12
    assert rel(AN.s.grad, gs) < 1e-12 and rel(AN.b.grad, gb) < 1e-12


# ---------------------------------------------------------------- Conv1x1
def _conv1x1(k, freeze=False):
    return O.Conv1x1(*(torch.randn(k, dtype=D) for _ in range(3)), freeze=freeze)


def test_conv1x1_invertible_and_orthogonal():
    C = _conv1x1(6)
    X = torch.randn(3, 6, 4, 5, dtype=D)
    Y = C.forward(X)
    assert rel(C.inverse(Y), X) < 1e-6  # test_layer_conv1x1.jl:49
    assert rel(C.forward(C.inverse(X)), X) < 1e-6  # :54
    assert abs(torch.linalg.norm(Y) - torch.linalg.norm(X)) < 1e-9  # orthogonal => logdet 0


@pytest.mark.parametrize("faithful", [True, False])
def test_conv1x1_grad_matches_autograd(faithful):
    k = 5
    vs = [torch.randn(k, dtype=D, requires_grad=True) for _ in range(3)]
    C = O.Conv1x1(*vs)
    X = torch.randn(2, k, 3, 4, dtype=D, requires_grad=True)
    Y = C.forward(X)
    W = torch.randn_like(Y)
    f = (Y * W).sum()
    g = torch.autograd.grad(f, [X] + vs)
    with torch.no_grad():
        dX, X_ = C.inverse_tuple(W, Y, faithful_batch_loop=faithful)
    assert rel(dX, g[0]) < 1e-10 and rel(X_, X.detach()) < 1e-10
    for p, gv in zip(C.params(), g[1:]):
        assert rel(p.grad, gv) < 1e-9
    # gradients accumulate (conv1x1.jl:237-239)
    with torch.no_grad():
        C.inverse_tuple(W, Y, faithful_batch_loop=faithful)
    assert rel(C.v1.grad, 2 * g[1]) < 1e-9


def test_conv1x1_frozen_zero_grad():
    C = _conv1x1(4, freeze=True)
    X = torch.randn(2, 4, 3, 3, dtype=D)
    C.inverse_tuple(torch.randn_like(X), X)
    assert all(float(p.grad.abs().max()) == 0 for p in C.params())  # test_layer_conv1x1.jl:80-91


# ---------------------------------------------------------------- NNlib restatement: adjointness
@pytest.mark.parametrize("k,p", [(3, 1), (1, 0)])
def test_conv_adjoint_pair(k, p):
    """test_nnlib_convolution.jl:19-27 : <y, conv(x)> == <x, ∇conv_data(y)>."""
    x = torch.randn(2, 3, 6, 7, dtype=D)
    w = torch.randn(5, 3, k, k, dtype=D)
    y = torch.randn(2, 5, 6, 7, dtype=D)
    a = (y * O.nn_conv(x, w, p)).sum()
    b = (x * O.nn_conv_data(y, w, p)).sum()
    assert abs(a - b) / abs(a) < 1e-12
    # ∇conv_filter is d/dw of conv
    wr = w.clone().requires_grad_(True)
    g, = torch.autograd.grad((y * O.nn_conv(x, wr, p)).sum(), wr)
    assert rel(O.nn_conv_filter(x, y, w.shape, p), g) < 1e-12


def test_conv_is_true_convolution():
    """A delta kernel at Julia index w[1,1] (0-based (0,0)) shifts the image by +1 (true
    convolution), the convention NNlib documents for flipped=false."""
    x = torch.zeros(1, 1, 5, 5, dtype=D)
    x[0, 0, 2, 2] = 1
    w = torch.zeros(1, 1, 3, 3, dtype=D)
    w[0, 0, 0, 0] = 1  # (ky=0, kx=0)
    y = O.nn_conv(x, w, 1)
    # y[Y,X] = sum x[Y+p-b, X+p-a] w[a,b]  ->  peak where Y+1-0 = 2, X+1-0 = 2
    assert y[0, 0, 1, 1] == 1 and y.sum() == 1


# ---------------------------------------------------------------- ResidualBlock
def test_residual_block_backward_matches_autograd():
    gen = torch.Generator().manual_seed(1)
    cl = O.make_coupling(gen, 6, 8, dtype=D)
    RB = cl.RB
    RB.b1.data = torch.randn(8, dtype=D) * 0.1
    RB.b2.data = torch.randn(8, dtype=D) * 0.1
    ps = [p.data.clone().requires_grad_(True) for p in RB.params()]
    for p, q in zip(RB.params(), ps):
        p.data = q
    X = torch.randn(2, 3, 6, 5, dtype=D, requires_grad=True)
    out = RB.forward(X)
    W = torch.randn_like(out)
    g = torch.autograd.grad((out * W).sum(), [X] + ps)
    with torch.no_grad():
        dX = RB.backward(W, X)
    assert rel(dX, g[0]) < 1e-10
    for p, gv in zip(RB.params(), g[1:]):
        assert rel(p.grad, gv) < 1e-10


# ---------------------------------------------------------------- coupling layer
@pytest.mark.parametrize("n_cond", [0, 3])
def test_coupling_invertible_and_grads(n_cond):
    gen = torch.Generator().manual_seed(2)
    L = O.make_coupling(gen, 4, 8, n_cond=n_cond, dtype=D, logdet=True)
    X = torch.randn(2, 4, 6, 6, dtype=D)
    Cn = torch.randn(2, n_cond, 6, 6, dtype=D) if n_cond else None
    Y, _ = L.forward(X, Cn)
    assert rel(L.inverse(Y, Cn), X) < 1e-6  # test_coupling_layer_glow.jl:28 (1e-2 there)
    ps = [p.data.clone().requires_grad_(True) for p in L.params()]
    for p, q in zip(L.params(), ps):
        p.data = q
    Xr = X.clone().requires_grad_(True)
    ins = [Xr] + ps
    if n_cond:
        Cr = Cn.clone().requires_grad_(True)
        ins.append(Cr)
    Y, ld = L.forward(Xr, Cr if n_cond else None)
    f = 0.5 * (Y ** 2).sum() / 2 - ld
    g = torch.autograd.grad(f, ins)
    with torch.no_grad():
        res = L.backward(Y / 2, Y, Cn)
    assert rel(res[0], g[0]) < 1e-6 and rel(res[1], X) < 1e-6
    for p, gv in zip(L.params(), g[1:1 + len(ps)]):
        assert rel(p.grad, gv) < 1e-6, "param grad"
    if n_cond:
        assert rel(res[2], g[-1]) < 1e-6


# ---------------------------------------------------------------- networks
@pytest.mark.parametrize("logdet", [True, False])
@pytest.mark.parametrize("split_scales", [True, False])
@pytest.mark.parametrize("nd", [2, 3])
def test_glow_network_properties(logdet, split_scales, nd):
    """test_glow.jl:20-66 shapes (shrunk): invertibility 1e-5, grad bookkeeping L*K*10."""
    L, K, n_in, nh, B = 2, 2, 2, 4, 2
    sp = (8, 8) if nd == 2 else (4, 4, 4)
    G = O.NetworkGlow(n_in, nh, L, K, logdet=logdet, split_scales=split_scales, ndims=nd,
                      dtype=torch.float32, faithful=False)
    X = torch.rand(B, n_in, *sp)
    out = G.forward(X)
    Z = out[0] if logdet else out
    assert rel(G.inverse(Z), X) < 1e-5  # test_glow.jl:46
    G.backward(Z, Z)
    P = G.get_params()
    assert len(P) == L * K * 10
    assert sum(p.grad is not None for p in P) == L * K * 10  # :50-59
    O.clear_grad(P)
    assert sum(p.grad is not None for p in P) == 0  # :61-66


def _autograd_check(G, X, cond=None):
    """loss f = ||Z||^2/(2B) - logdet; hand backward must equal autograd in float64."""
    B = X.shape[0]
    with torch.no_grad():  # data-dependent ActNorm init
        G.forward(X) if cond is None else G.forward(X, cond)
    P = G.get_params()
    leaves = [p.data.clone().requires_grad_(True) for p in P]
    for p, q in zip(P, leaves):
        p.data = q
    Xr = X.clone().requires_grad_(True)
    if cond is None:
        out = G.forward(Xr)
        Z, ld = out if isinstance(out, tuple) else (out, 0.0)
        ins = [Xr] + leaves
    else:
        Cr = cond.clone().requires_grad_(True)
        Z, ZC, ld = G.forward(Xr, Cr)
        ins = [Xr, Cr] + leaves
    f = 0.5 * (Z ** 2).sum() / B - ld
    g = torch.autograd.grad(f, ins)
    O.clear_grad(P)
    with torch.no_grad():
        if cond is None:
            dX, X_ = G.backward(Z.detach() / B, Z.detach())
            assert rel(dX, g[0]) < 1e-6 and rel(X_, X) < 1e-6
            gp = g[1:]
        else:
            dX, X_, dC = G.backward(Z.detach() / B, Z.detach(), ZC.detach())
            assert rel(dX, g[0]) < 1e-6 and rel(X_, X) < 1e-6
            assert rel(dC, g[1]) < 1e-6
            gp = g[2:]
    for i, (p, gv) in enumerate(zip(P, gp)):
        assert rel(p.grad, gv) < 1e-6, f"param {i}"


@pytest.mark.parametrize("split_scales", [True, False])
def test_glow_backward_matches_autograd(split_scales):
    G = O.NetworkGlow(2, 6, 2, 2, logdet=True, split_scales=split_scales, dtype=D, faithful=False)
    _autograd_check(G, torch.rand(3, 2, 8, 8, dtype=D))


def test_glow3d_backward_matches_autograd():
    G = O.NetworkGlow(1, 4, 2, 1, logdet=True, ndims=3, dtype=D, faithful=False)
    _autograd_check(G, torch.rand(2, 1, 4, 4, 4, dtype=D))


def test_glow_L1_split_quirk():
    """L == 1 still splits at i == 1 (invertible_network_glow.jl:120,136)."""
    G = O.NetworkGlow(1, 4, 1, 2, logdet=True, dtype=D, faithful=False)
    X = torch.rand(2, 1, 4, 4, dtype=D)
    Z, _ = G.forward(X)
    assert Z.dim() == 1 and Z.numel() == X.numel()
    assert rel(G.inverse(Z), X) < 1e-6


@pytest.mark.parametrize("split_scales", [True, False])
def test_conditional_glow(split_scales):
    """test_conditional_glow_network.jl:35,46 : invertibility 1e-5, L*K*10+2 grads."""
    L, K = 2, 2
    G = O.NetworkConditionalGlow(2, 3, 6, L, K, split_scales=split_scales, dtype=D, faithful=False)
    X = torch.rand(3, 2, 8, 8, dtype=D)
    Cn = torch.rand(3, 3, 8, 8, dtype=D)
    Z, ZC, ld = G.forward(X, Cn)
    assert Z.shape == X.shape
    assert rel(G.inverse(Z, ZC), X) < 1e-5
    G.backward(Z, Z, ZC)
    assert sum(p.grad is not None for p in G.get_params()) == L * K * 10 + 2
    G2 = O.NetworkConditionalGlow(2, 3, 6, L, K, split_scales=split_scales, dtype=D, faithful=False)
    _autograd_check(G2, X, Cn)


def test_param_order_and_counts():
    """SURVEY §3.4/§8 : cfg1 14 840 params, per-step 2 836 / 4 584; AN first then CL."""
    G = O.NetworkGlow(1, 32, 2, 2, faithful=False)
    G.forward(torch.rand(2, 1, 16, 16))
    P = G.get_params()
    assert sum(p.data.numel() for p in P) == 14840
    assert [tuple(p.data.shape) for p in P[:2]] == [(4,), (4,)]
    step0 = P[8:16]
    assert [tuple(p.data.shape) for p in step0] == [(4,), (4,), (4,), (32, 2, 3, 3), (32, 32, 1, 1),
                                                    (32, 4, 3, 3), (32,), (32,)]


def test_oracle_matches_golden():
    """The committed vectors (tests/golden/make_golden.py: float64 oracle on seeded inputs) pin the oracle against
    drift: any edit of the restatement that changes its numbers shows up here.  They are oracle-generated, not
    reference-generated (the oracle itself stays 'parity unpinned', see oracle/glow_oracle.py)."""
    import os
    import numpy as np
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "glow_small.npz"))
    n_in, nh, L, K = [int(v) for v in gold["glow_cfg"]]
    G = O.NetworkGlow(n_in, nh, L, K, split_scales=True, seed=5, dtype=torch.float64, faithful=False)
    ps = G.get_params()
    for i, p in enumerate(ps):
        p.data = torch.from_numpy(gold[f"glow_p{i:03d}"]).double()
    X = torch.from_numpy(gold["glow_X"]).double()
    Z, ld = G.forward(X)
    dX, Xr = G.backward(Z / X.shape[0], Z)
    r = lambda a, b: float(torch.linalg.norm(a.reshape(-1) - torch.from_numpy(b).double().reshape(-1)) /
                           max(torch.linalg.norm(torch.from_numpy(b).double().reshape(-1)), 1e-300))
    assert r(Z, gold["glow_Z"]) < 1e-12
    assert abs(float(ld) - float(gold["glow_logdet"])) < 1e-10 * abs(float(gold["glow_logdet"]))
    assert r(dX, gold["glow_dX"]) < 1e-11
    assert r(Xr, gold["glow_X"]) < 1e-6   # float32 input recovered by inversion
    for i, p in enumerate(ps):
        assert r(p.grad, gold[f"glow_g{i:03d}"]) < 1e-10, i
