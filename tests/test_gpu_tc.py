"""GPU parity of the tcgen05 tensor-core path (precision "bf16x3" = 3-term split bf16, fp32-equivalent;
precision "bf16" = single pass) against the float64 oracle, at ResidualBlock, coupling-layer and network
level.  Tolerances: bf16x3 must meet the float32 bar of tests/_util.py; bf16 states its own (TOL_BF16)."""
import pytest
import torch

from _util import O, TOL_GRAD, TOL_LOGDET, TOL_OUT, FragileUnits, assert_grad_close, clone_oracle, rel

import inb200

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL_BF16 = 1.5e-1   # single-pass bf16 operands (8-bit mantissa), fp32 accumulate; ReLU-mask flips dominate the backward error
TOLS = {"bf16x3": (TOL_OUT, TOL_GRAD), "fp16x3": (TOL_OUT, TOL_GRAD), "bf16": (TOL_BF16, 2 * TOL_BF16)}
# a ReLU unit is "fragile" when its pre-activation is closer to zero than the arithmetic's own error
FRAGILE = {"bf16x3": 1e-4, "fp16x3": 1e-5, "bf16": 3e-2}
INV_TOL = {"bf16x3": 1e-5, "fp16x3": 1e-5, "bf16": 1e-2}  # bf16 rounding of the block input is discontinuous: ~1e-4 per layer


def g(t):
    return t.to(DEV)


RB_CASES = [
    # B, Cin, nh, Cout, spatial, k1, k2
    (2, 6, 128, 12, (16, 16), 3, 1),
    (3, 2, 256, 4, (8, 8), 3, 1),          # two samples per 128-pixel tile, ragged last tile
    (2, 24, 128, 48, (32, 32), 3, 1),      # cfg2 scale-3 channel plan
    (1, 3, 128, 6, (2, 256), 3, 1),        # W > tile
    (2, 4, 128, 8, (8, 8, 8), 3, 1),       # 3-D, 27 taps
    (2, 6, 128, 12, (16, 16), 3, 3),       # k2 = 3 (HINT-style block)
    (2, 5, 128, 10, (16, 16), 1, 1),
]


@pytest.mark.parametrize("prec", ["bf16x3", "fp16x3", "bf16"])
@pytest.mark.parametrize("case", RB_CASES)
def test_resblock_tc(case, prec):
    torch.manual_seed(4)
    tol_out, tol_grad = TOLS[prec]
    B, Cin, nh, Cout, sp, k1, k2 = case
    nd = len(sp)
    RB = inb200.ResidualBlock(Cin, nh, n_out=Cout, k1=k1, k2=k2, p1=(k1 - 1) // 2, p2=(k2 - 1) // 2, ndims=nd,
                              precision=prec, gen=torch.Generator().manual_seed(3), device=DEV)
    RB.b1.data.copy_(torch.randn(nh) * 0.1)
    RB.b2.data.copy_(torch.randn(nh) * 0.1)
    ws = [p.data.cpu().double() for p in RB.get_params()]
    R64 = O.ResidualBlock(*ws, p1=(k1 - 1) // 2, p2=(k2 - 1) // 2)
    X = torch.randn(B, Cin, *sp)
    dY = torch.randn(B, Cout, *sp)
    Y = RB.forward(g(X))
    Y64 = R64.forward(X.double())
    assert rel(Y, Y64) < tol_out, f"forward {rel(Y, Y64)}"
    dX = RB.backward(g(dY), g(X))
    with FragileUnits(FRAGILE[prec]) as fr:
        dX64 = R64.backward(dY.double(), X.double())
    assert_grad_close(dX, dX64, tol_out, fr, "dX")
    for name, p, q in zip("W1 W2 W3 b1 b2".split(), RB.get_params(), R64.params()):
        assert_grad_close(p.grad, q.grad, tol_grad, fr, name)


@pytest.mark.parametrize("prec", ["bf16x3", "fp16x3", "bf16"])
@pytest.mark.parametrize("case", [(2, 12, 0, 128, (16, 16)), (2, 4, 4, 128, (16, 16)), (2, 8, 0, 128, (8, 8, 8))])
def test_coupling_layer_tc(case, prec):
    torch.manual_seed(6)
    tol_out, tol_grad = TOLS[prec]
    B, Cc, n_cond, nh, sp = case
    CL = inb200.CouplingLayerGlow(Cc, nh, n_cond=n_cond, logdet=True, ndims=len(sp), precision=prec,
                                  gen=torch.Generator().manual_seed(9), device=DEV)
    ws = [p.data.cpu().double() for p in CL.get_params()]
    C64 = O.CouplingLayerGlow(O.Conv1x1(*ws[:3]), O.ResidualBlock(*ws[3:]), logdet=True)
    X = torch.randn(B, Cc, *sp)
    cond = torch.randn(B, n_cond, *sp) if n_cond else None
    dY = torch.randn(B, Cc, *sp)
    Y, ld = CL.forward(g(X), g(cond) if n_cond else None)
    Y64, ld64 = C64.forward(X.double(), cond.double() if n_cond else None)
    assert rel(Y, Y64) < tol_out
    assert abs(ld.item() - ld64.item()) / abs(ld64.item()) < TOL_LOGDET * (tol_out / TOL_OUT)
    assert rel(CL.inverse(Y, g(cond) if n_cond else None), X) < 1e-5  # same arithmetic both ways
    res = CL.backward(g(dY), Y, g(cond) if n_cond else None)
    with FragileUnits(FRAGILE[prec]) as fr:
        r64 = C64.backward(dY.double(), Y.cpu().double(), cond.double() if n_cond else None)
    assert rel(res[1], r64[1]) < tol_out
    assert_grad_close(res[0], r64[0], tol_out, fr, "dX")
    if n_cond:
        assert_grad_close(res[2], r64[2], tol_out, fr, "dC")
    for name, p, q in zip("v1 v2 v3 W1 W2 W3 b1 b2".split(), CL.get_params(), C64.params()):
        assert_grad_close(p.grad, q.grad, tol_grad, fr, name)


@pytest.mark.parametrize("case", [
    (4, 6, 256, 12, (32, 32)),     # cfg2 scale-1 channel plan, 4 image rows per tile
    (3, 12, 256, 24, (64, 64)),    # scale-2 plan: GEMM3 wider than 128 columns (full P path), 2 rows per tile
    (5, 6, 128, 12, (16, 16)),     # odd number of tiles: the last pair has an empty tile
    (2, 24, 256, 48, (32, 32)),    # scale-3 plan: forward on the single-CTA kernel, backward on pairs
])
@pytest.mark.parametrize("prec", ["bf16x3", "fp16x3"])
def test_resblock_tc_many_tiles_per_pair(case, prec, monkeypatch):
    """The persistent loop of the CTA-pair kernel (ring wrap, barrier parities, TMEM region swap, staging reuse)
    over many tiles per pair: the number of resident pairs is capped at 2, so a small input already walks 4-16
    tile pairs per cluster."""
    monkeypatch.setenv("INB_CHAIN_MAXPAIRS", "2")
    test_resblock_tc(case + (3, 1), prec)


@pytest.mark.parametrize("prec", ["bf16x3", "fp16x3", "bf16"])
def test_glow_network_tc(prec):
    """cfg2's channel plan (3 -> 12/24/48) with n_hidden = 128 at 64x64, L = 3."""
    from test_gpu_networks import run_glow_parity
    tol_out, tol_grad = TOLS[prec]
    run_glow_parity(3, 128, 3, 2, (2, 3, 64, 64), precision=prec, tol_out=tol_out, tol_grad=tol_grad,
                    fragile_thr=FRAGILE[prec], inv_tol=INV_TOL[prec])


def test_fp16x3_gradient_scale_covers_tiny_and_huge_gradients():
    """fp16x3 derives one power-of-two scale per ResidualBlock backward from max|dY| on the device: gradients of
    magnitude 1e-9 or 1e+6 (far outside the range of an IEEE half) must come out as accurately as O(1) ones."""
    torch.manual_seed(8)
    B, Cin, nh, Cout, sp = 2, 6, 128, 12, (16, 16)
    RB = inb200.ResidualBlock(Cin, nh, n_out=Cout, k1=3, k2=1, p1=1, p2=0, precision="fp16x3",
                              gen=torch.Generator().manual_seed(3), device=DEV)
    ws = [p.data.cpu().double() for p in RB.get_params()]
    R64 = O.ResidualBlock(*ws, p1=1, p2=0)
    X, dY = torch.randn(B, Cin, *sp), torch.randn(B, Cout, *sp)
    for mag in (1e-9, 1.0, 1e6):
        dX = RB.backward(g(dY * mag), g(X))
        with FragileUnits(1e-5) as fr:
            dX64 = R64.backward(dY.double() * mag, X.double())
        assert_grad_close(dX, dX64, TOL_OUT, fr, f"dX at |dY| ~ {mag}")
        for name, p, q in zip("W1 W2 W3 b1 b2".split(), RB.get_params(), R64.params()):
            assert_grad_close(p.grad, q.grad, TOL_GRAD, fr, f"{name} at |dY| ~ {mag}")
    assert torch.equal(RB.backward(g(dY * 0), g(X)).cpu(), torch.zeros(B, Cin, *sp))  # max|dY| = 0: scale 1


SMALL_NH_CASES = [
    # B, Cin, nh, Cout, spatial, k1, k2 - blocks narrower than the 128 hidden channels the chain kernels run at
    (2, 2, 32, 4, (32, 32), 3, 1),         # cfg1 / cfg3 scale-1 plan (n_hidden = 32)
    (2, 4, 64, 8, (16, 16), 3, 1),
    (2, 5, 70, 6, (16, 16), 3, 1),         # not a multiple of anything
    (2, 4, 32, 8, (8, 8, 8), 3, 1),        # cfg5 scale-1 plan (3-D, 27 taps)
    (2, 6, 160, 12, (16, 16), 3, 1),       # between 128 and 256
    (2, 16, 32, 32, (8, 8, 8), 3, 1),      # cfg5 scale-2 plan: 27 taps x 32 channels = 864 expanded columns, four GEMM3 passes
]


@pytest.mark.parametrize("prec", ["bf16x3", "fp16x3"])
@pytest.mark.parametrize("case", SMALL_NH_CASES)
def test_resblock_tc_small_n_hidden(case, prec):
    """n_hidden of 32 / 64 (BASELINE configs[0], [2], [4]) on the fused tcgen05 chain: the hidden width is zero padded
    to 128 (or 256) inside the packed operands; outputs and all five gradients have the block's own shapes."""
    test_resblock_tc(case, prec)


@pytest.mark.parametrize("prec", ["bf16x3", "fp16x3"])
def test_tc_modes_fall_back_for_shapes_the_tensor_core_kernels_cannot_tile(prec):
    """A tensor-core precision never refuses a block: k2 = 3 runs on the unfused tcgen05 convolutions (bf16x3 arithmetic
    under either mode), a spatial size that cannot be cut into 64 / 128-pixel boxes on the fp32 CUDA-core kernels."""
    test_resblock_tc((2, 6, 128, 12, (16, 16), 3, 3), "bf16x3" if prec == "bf16x3" else "fp16x3")
    torch.manual_seed(4)
    RB = inb200.ResidualBlock(2, 32, n_out=4, k1=3, k2=1, p1=1, p2=0, precision=prec,
                              gen=torch.Generator().manual_seed(3), device=DEV)
    ws = [p.data.cpu().double() for p in RB.get_params()]
    R64 = O.ResidualBlock(*ws, p1=1, p2=0)
    X, dY = torch.randn(1, 2, 12, 12), torch.randn(1, 4, 12, 12)
    assert rel(RB.forward(g(X)), R64.forward(X.double())) < 1e-6       # fp32 kernels
    dX = RB.backward(g(dY), g(X))
    assert rel(dX, R64.backward(dY.double(), X.double())) < 1e-5


@pytest.mark.parametrize("prec", ["fp32", "bf16x3", "fp16x3"])
@pytest.mark.parametrize("sp", [(16, 16), (8, 8, 8)])
def test_resblock_exact_lattice(prec, sp):
    """Flip-free gradient parity: inputs and weights on an integer lattice with half-integer biases, so
    every pre-activation is exactly representable, never zero-ambiguous, and every sum is exact in fp32
    accumulation (and in the hi/lo bf16 split, whose dropped lo*lo term vanishes because one operand of
    every product has an empty lo part).  Forward, dX and all five parameter gradients must then equal
    the float64 oracle to rounding of the final sums only - no tolerance for ReLU-mask flips is needed."""
    gen = torch.Generator().manual_seed(17)
    B, Cin, nh, Cout = 2, 6, 128, 12
    nd = len(sp)

    def tern(shape, density):
        v = torch.randint(-1, 2, shape, generator=gen).float()
        return v * (torch.rand(shape, generator=gen) < density).float()
    RB = inb200.ResidualBlock(Cin, nh, n_out=Cout, k1=3, k2=1, p1=1, p2=0, ndims=nd, precision=prec, device=DEV)
    W = [tern((nh, Cin) + (3,) * nd, 0.3), tern((nh, nh) + (1,) * nd, 0.05), tern((nh, Cout) + (3,) * nd, 0.05 if nd == 2 else 0.02),
         torch.full((nh,), 0.5), torch.full((nh,), 0.25)]
    for p, w in zip(RB.get_params(), W):
        p.data.copy_(w)
    R64 = O.ResidualBlock(*[w.double() for w in W], p1=1, p2=0)
    X = torch.randint(-1, 2, (B, Cin) + sp, generator=gen).float()
    dY = torch.randint(-1, 2, (B, Cout) + sp, generator=gen).float()
    Y1, Y2, _ = R64.forward(X.double(), save=True)
    assert (Y1.abs() >= 0.25).all() and (Y2.abs() >= 0.25).all()  # no unit anywhere near the ReLU kink
    assert rel(RB.forward(g(X)), R64.forward(X.double())) < 1e-7
    dX = RB.backward(g(dY), g(X))
    dX64 = R64.backward(dY.double(), X.double())
    assert rel(dX, dX64) < 1e-7
    for name, p, q in zip("W1 W2 W3 b1 b2".split(), RB.get_params(), R64.params()):
        assert rel(p.grad, q.grad) < 1e-6, f"{name} {rel(p.grad, q.grad)}"
