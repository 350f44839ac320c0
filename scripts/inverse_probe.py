"""Sampling throughput: X = G.inverse(Z) for cfg2, samples/s on one GPU."""
import json, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import inb200
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
G = inb200.NetworkGlow(3, 256, 3, 16, split_scales=True, precision=prec, seed=0, device="cuda")
X = torch.rand(B, 3, 256, 256, device="cuda")
Z, ld = G.forward(X)
for _ in range(3):
    Xr = G.inverse(Z)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    Xr = G.inverse(Z)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
err = (torch.linalg.norm((Xr - X).reshape(-1)) / torch.linalg.norm(X.reshape(-1))).item()
print(f"inverse: {ms:.2f} ms per batch of {B} = {B / ms * 1e3:.1f} samples/s, ||X - inverse(forward(X))|| / ||X|| = {err:.2e}")
print(json.dumps({"metric": "glow_inverse_samples_per_sec", "value": B / ms * 1e3, "unit": "samples/s", "ms_per_batch": ms,
                  "config": {"workload": "cfg2: NetworkGlow(3,256,L=3,K=16) on 256x256x3, X = G.inverse(Z)", "batch": B,
                             "precision": prec}, "invertibility_rel_l2": err}))
