"""Full-size cfg2: gradients of the tcgen05 path against the fp32 CUDA-core path, flow step by flow step."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import inb200
K = int(sys.argv[1]) if len(sys.argv) > 1 else 16
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
def rel(a, b):
    return (torch.linalg.norm((a - b).double().reshape(-1)) / torch.linalg.norm(b.double().reshape(-1))).item()
torch.manual_seed(0)
G32 = inb200.NetworkGlow(3, 256, 3, K, split_scales=True, precision="fp32", seed=1, device="cuda")
Gtc = inb200.NetworkGlow(3, 256, 3, K, split_scales=True, precision="bf16x3", seed=1, device="cuda")
X = torch.rand(B, 3, 256, 256, device="cuda")
Z32, ld32 = G32.forward(X)
inb200.set_params(Gtc, [p.data for p in G32.get_params()])
Ztc, ldtc = Gtc.forward(X)
print("Z", rel(Ztc, Z32), "logdet", abs(ldtc.item() - ld32.item()) / abs(ld32.item()))
dZ = Z32 / B
dX32, X32 = G32.backward(dZ, Z32)
dXtc, Xtc = Gtc.backward(dZ, Z32)
print("X recomputed", rel(Xtc, X32), "dX", rel(dXtc, dX32))
ps32, pstc = G32.get_params(), Gtc.get_params()
nAN = 2 * 3 * K
names = ["v1", "v2", "v3", "W1", "W2", "W3", "b1", "b2"]
for i in range(3):
    for j in (0, K // 2, K - 1):
        base = nAN + 8 * (i * K + j)
        errs = [rel(pstc[base + k].grad, ps32[base + k].grad) for k in range(8)]
        an = [rel(pstc[2 * (i * K + j) + k].grad, ps32[2 * (i * K + j) + k].grad) for k in range(2)]
        print(f"scale {i + 1} step {j:2d}: " + " ".join(f"{n} {e:.1e}" for n, e in zip(names, errs)) + f" | s {an[0]:.1e} b {an[1]:.1e}")
