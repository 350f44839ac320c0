import torch, sys
sys.path.insert(0, "tests")
from _util import O, rel
import inb200
torch.manual_seed(4)
B, Cin, nh, Cout, sp, k1, k2 = (2, 24, 128, 48, (32, 32), 3, 1)
outs = {}
X = torch.randn(B, Cin, *sp); dY = torch.randn(B, Cout, *sp)
for prec in ["fp32", "bf16x3", "bf16"]:
    RB = inb200.ResidualBlock(Cin, nh, n_out=Cout, k1=k1, k2=k2, p1=1, p2=0, precision=prec, gen=torch.Generator().manual_seed(3), device="cuda")
    RB.b1.data.copy_(torch.randn(nh, generator=torch.Generator().manual_seed(1)) * 0.1)
    RB.b2.data.copy_(torch.randn(nh, generator=torch.Generator().manual_seed(2)) * 0.1)
    if prec == "fp32":
        ws = [p.data.cpu().double() for p in RB.get_params()]
        R64 = O.ResidualBlock(*ws, p1=1, p2=0)
        dX64 = R64.backward(dY.double(), X.double())
        R32 = O.ResidualBlock(*[w.float() for w in ws], p1=1, p2=0)
        dX32 = R32.backward(dY, X)
        print("oracle fp32 vs fp64 dX", rel(dX32, dX64))
    dX = RB.backward(dY.cuda(), X.cuda()).cpu().double()
    e = (dX - dX64).abs().reshape(-1)
    scale = dX64.abs().max()
    srt = e.sort().values
    n = e.numel()
    print(prec, "relL2", rel(dX, dX64), "median", (srt[n//2]/scale).item(), "p99", (srt[int(n*0.99)]/scale).item(), "p999", (srt[int(n*0.999)]/scale).item(), "max", (srt[-1]/scale).item(),
          "frac>1e-4*scale", (e > 1e-4*scale).double().mean().item())
    for name, p, q in zip("W1 W2 W3 b1 b2".split(), RB.get_params(), R64.params()):
        print("   ", name, rel(p.grad, q.grad))
