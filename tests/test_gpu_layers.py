"""GPU parity, layer level: every layer-level entry point of the C ABI against the oracle on the same
seeded inputs and weights (float32 oracle = what the reference's arithmetic gives; float64 oracle =
truth).  Tolerances are in tests/_util.py.  Properties asserted by the reference's own layer tests
(invertibility, init statistics, frozen grads, grad accumulation) are re-asserted on the CUDA path."""
import pytest
import torch

from _util import O, TOL_GRAD, TOL_LOGDET, TOL_OUT, rel

import inb200

pytestmark = pytest.mark.gpu
DEV = "cuda"


def g(t):
    return t.to(DEV)


@pytest.mark.parametrize("shape", [(2, 3, 8, 6), (3, 2, 4, 6, 8), (1, 1, 64, 64), (2, 12, 32, 32)])
def test_squeeze_matches_index_map_and_roundtrips(shape):
    X = torch.randn(*shape)
    Y = inb200.squeeze(g(X))
    assert torch.equal(Y.cpu(), O.squeeze(X))  # pure index map: bit exact
    assert torch.equal(inb200.unsqueeze(Y).cpu(), X)  # test_squeeze.jl:11-13


def test_squeeze_odd_throws():
    with pytest.raises(inb200.InbError, match="multiple of 2"):
        inb200.squeeze(g(torch.randn(1, 1, 5, 4)))
    with pytest.raises(inb200.InbError, match="divisible"):
        inb200.unsqueeze(g(torch.randn(1, 3, 4, 4)))


@pytest.mark.parametrize("shape", [(4, 3, 6, 5), (2, 12, 32, 32), (2, 8, 4, 6, 8), (8, 48, 16, 16)])
def test_actnorm_parity(shape):
    torch.manual_seed(1)
    X = torch.randn(*shape) * 3 + 1
    dY = torch.randn(*shape)
    AN = inb200.ActNorm(shape[1], logdet=True)
    Y, ld = AN.forward(g(X))
    A32, A64 = O.ActNorm(shape[1], logdet=True), O.ActNorm(shape[1], logdet=True)
    Y32, ld32 = A32.forward(X)
    Y64, ld64 = A64.forward(X.double())
    # data-dependent init (actnorm.jl:67-72)
    assert rel(AN.s.data, A64.s.data) < 1e-5 and rel(AN.b.data, A64.b.data) < 1e-5
    red = [0] + list(range(2, X.dim()))
    assert Y.mean(dim=red).abs().max().item() < 1e-5           # test_actnorm.jl:57
    assert (Y.var(dim=red, unbiased=True) - 1).abs().max().item() < 1e-3  # :58
    assert rel(Y, Y64) < max(TOL_OUT, 2 * rel(Y32, Y64))
    assert abs(ld.item() - ld64.item()) / abs(ld64.item()) < TOL_LOGDET
    assert rel(AN.inverse(Y), X) < 1e-6                          # test_actnorm.jl:61-62
    # backward with the oracle's float32 parameters loaded so both sides see identical s, b
    AN.s.data.copy_(A32.s.data)
    AN.b.data.copy_(A32.b.data)
    A64.s.data, A64.b.data = A32.s.data.double(), A32.b.data.double()
    Yo = A64.forward(X.double())[0]
    dX, Xr = AN.backward(g(dY), g(Yo.float()))
    dX64, X64 = A64.backward(dY.double(), Yo.float().double())
    assert rel(dX, dX64) < TOL_OUT and rel(Xr, X64) < TOL_OUT
    assert rel(AN.s.grad, A64.s.grad) < TOL_GRAD and rel(AN.b.grad, A64.b.grad) < TOL_GRAD


@pytest.mark.parametrize("shape", [(2, 4, 8, 8), (3, 12, 16, 16), (2, 24, 8, 8), (2, 48, 8, 8), (2, 8, 4, 4, 4),
                                   (2, 3, 6, 6), (5, 2, 10, 10)])
def test_conv1x1_parity(shape):
    torch.manual_seed(2)
    k = shape[1]
    X, dY = torch.randn(*shape), torch.randn(*shape)
    Cn = inb200.Conv1x1(k, gen=torch.Generator().manual_seed(5), device=DEV)
    vs = [p.data.cpu() for p in Cn.get_params()]
    C64 = O.Conv1x1(*[v.double() for v in vs])
    C32 = O.Conv1x1(*vs)
    Y = Cn.forward(g(X))
    Y64 = C64.forward(X.double())
    assert rel(Y, Y64) < max(TOL_OUT, 2 * rel(C32.forward(X), Y64))
    assert rel(Cn.inverse(Y), X) < 1e-6                          # test_layer_conv1x1.jl:49-61
    dX, Xr = Cn.inverse((g(dY), Y))
    dX64, X64 = C64.inverse_tuple(dY.double(), Y.cpu().double())
    assert rel(dX, dX64) < TOL_OUT and rel(Xr, X64) < TOL_OUT
    for p, q in zip(Cn.get_params(), C64.params()):
        assert rel(p.grad, q.grad) < TOL_GRAD
    # conv1x1.jl:237-239: a second call accumulates
    Cn.inverse((g(dY), Y))
    for p, q in zip(Cn.get_params(), C64.params()):
        assert rel(p.grad, 2 * q.grad) < TOL_GRAD
    # frozen layer gives zero gradients (test_layer_conv1x1.jl:80-91)
    Cf = inb200.Conv1x1(k, freeze=True, device=DEV)
    Cf.inverse((g(dY), Y))
    assert all(p.grad.abs().max().item() == 0 for p in Cf.get_params())


RB_CASES = [
    # B, Cin, nh, Cout, spatial, k1, k2
    (2, 2, 8, 4, (8, 8), 3, 1),
    (2, 6, 32, 12, (16, 16), 3, 1),
    (1, 3, 16, 6, (12, 10), 3, 3),
    (2, 2, 8, 4, (4, 6, 8), 3, 1),
    (3, 5, 70, 20, (9, 7), 1, 1),
]


def make_rb(case, seed=3):
    B, Cin, nh, Cout, sp, k1, k2 = case
    nd = len(sp)
    RB = inb200.ResidualBlock(Cin, nh, n_out=Cout, k1=k1, k2=k2, p1=(k1 - 1) // 2, p2=(k2 - 1) // 2, ndims=nd,
                              gen=torch.Generator().manual_seed(seed), device=DEV)
    RB.b1.data.copy_(torch.randn(nh) * 0.1)
    RB.b2.data.copy_(torch.randn(nh) * 0.1)
    ws = [p.data.cpu() for p in RB.get_params()]
    mk = lambda dt: O.ResidualBlock(*[w.to(dt) for w in ws], p1=(k1 - 1) // 2, p2=(k2 - 1) // 2)
    return RB, mk(torch.float32), mk(torch.float64)


@pytest.mark.parametrize("case", RB_CASES)
def test_resblock_parity(case):
    torch.manual_seed(4)
    B, Cin, nh, Cout, sp, k1, k2 = case
    RB, R32, R64 = make_rb(case)
    X = torch.randn(B, Cin, *sp)
    dY = torch.randn(B, Cout, *sp)
    Y = RB.forward(g(X))
    Y64 = R64.forward(X.double())
    assert rel(Y, Y64) < max(TOL_OUT, 2 * rel(R32.forward(X), Y64))
    dX = RB.backward(g(dY), g(X))
    dX64 = R64.backward(dY.double(), X.double())
    assert rel(dX, dX64) < TOL_OUT
    for name, p, q in zip("W1 W2 W3 b1 b2".split(), RB.get_params(), R64.params()):
        assert rel(p.grad, q.grad) < TOL_GRAD, name


CL_CASES = [
    # B, C, n_cond, nh, spatial
    (2, 4, 0, 8, (8, 8)),
    (2, 12, 0, 32, (16, 16)),
    (2, 3, 0, 8, (6, 6)),       # odd channel count: split 2|1 (ties to even)
    (2, 8, 0, 8, (4, 4, 4)),
    (2, 4, 4, 8, (8, 8)),       # ConditionalLayerGlow
]


@pytest.mark.parametrize("case", CL_CASES)
@pytest.mark.parametrize("logdet", [True, False])
def test_coupling_layer_parity(case, logdet):
    torch.manual_seed(6)
    B, Cc, n_cond, nh, sp = case
    nd = len(sp)
    CL = inb200.CouplingLayerGlow(Cc, nh, n_cond=n_cond, logdet=logdet, ndims=nd,
                                  gen=torch.Generator().manual_seed(9), device=DEV)
    CL.RB.b1.data.copy_(torch.randn(nh) * 0.1)
    CL.RB.b2.data.copy_(torch.randn(nh) * 0.1)
    ws = [p.data.cpu() for p in CL.get_params()]

    def mk(dt):
        return O.CouplingLayerGlow(O.Conv1x1(*[w.to(dt) for w in ws[:3]]),
                                   O.ResidualBlock(*[w.to(dt) for w in ws[3:]]), logdet=logdet)
    C32, C64 = mk(torch.float32), mk(torch.float64)
    X = torch.randn(B, Cc, *sp)
    cond = torch.randn(B, n_cond, *sp) if n_cond else None
    dY = torch.randn(B, Cc, *sp)
    out = CL.forward(g(X), g(cond) if n_cond else None)
    o64 = C64.forward(X.double(), cond.double() if n_cond else None)
    o32 = C32.forward(X, cond)
    Y, Y64, Y32 = (out[0], o64[0], o32[0]) if logdet else (out, o64, o32)
    assert rel(Y, Y64) < max(TOL_OUT, 2 * rel(Y32, Y64))
    if logdet:
        assert abs(out[1].item() - o64[1].item()) / abs(o64[1].item()) < TOL_LOGDET
    Xi = CL.inverse(Y, g(cond) if n_cond else None)
    assert rel(Xi, X) < 1e-2  # test_coupling_layer_glow.jl:28-31 (reference bound)
    assert rel(Xi, X) < 1e-5
    res = CL.backward(g(dY), Y, g(cond) if n_cond else None)
    r64 = C64.backward(dY.double(), Y.cpu().double(), cond.double() if n_cond else None)
    assert rel(res[0], r64[0]) < TOL_OUT and rel(res[1], r64[1]) < TOL_OUT
    if n_cond:
        assert rel(res[2], r64[2]) < TOL_OUT
    names = "v1 v2 v3 W1 W2 W3 b1 b2".split()
    for name, p, q in zip(names, CL.get_params(), C64.params()):
        assert rel(p.grad, q.grad) < TOL_GRAD, name


def test_nll_grad():
    Z = torch.randn(3 * 1000)
    f, dZ = inb200.nll_grad(g(Z), 3)
    assert abs(f.item() - (-O.log_likelihood(Z.double(), 3)).item()) / f.item() < 1e-6
    assert rel(dZ, Z / 3) < 1e-7
