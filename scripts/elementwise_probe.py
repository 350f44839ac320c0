"""One forward + backward of cfg2 with K = 1 (three flow steps, full-size tensors, B = 64): ncu probe for the per-pixel kernels."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import inb200
os.environ.setdefault("INB_GRAPHS", "0")
G = inb200.NetworkGlow(3, 256, 3, 1, split_scales=True, precision="bf16x3", seed=0, device="cuda")
X = torch.rand(64, 3, 256, 256, device="cuda")
for _ in range(2):
    Z, ld = G.forward(X)
    nll, dZ = inb200.nll_grad(Z, 64)
    G.backward(dZ, Z)
    Xr = G.inverse(Z)
torch.cuda.synchronize()
print("ok")
