"""BASELINE configs[3]: NetworkMultiScaleHINT on 128x128x2, batch 32 - forward + loss gradient + backward samples/s on
one GPU, per-family time table, and the torch-CPU oracle on a bounded sample beside it.
usage: python scripts/hint_probe.py [n_hidden] [precision] [k2] [L] [K] [out.json]"""
import json, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import inb200
nh = int(sys.argv[1]) if len(sys.argv) > 1 else 64
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
k2 = int(sys.argv[3]) if len(sys.argv) > 3 else 3
L = int(sys.argv[4]) if len(sys.argv) > 4 else 2
K = int(sys.argv[5]) if len(sys.argv) > 5 else 4
out = sys.argv[6] if len(sys.argv) > 6 else None
B = 32
net = inb200.NetworkMultiScaleHINT(2, nh, L, K, k2=k2, p2=(k2 - 1) // 2, precision=prec, seed=0, device="cuda")
X = torch.randn(B, 2, 128, 128, device="cuda")


def step():
    Z, ld = net.forward(X)
    nll, dZ = inb200.nll_grad(Z, B)
    dX, Xr = net.backward(dZ, Z)
    return nll - ld, Xr


for _ in range(3):
    f, Xr = step()
torch.cuda.synchronize()
n0 = inb200.lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    f, Xr = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
launches = (inb200.lib.launch_count() - n0) // 5
inv = (torch.linalg.norm((Xr - X).reshape(-1)) / torch.linalg.norm(X.reshape(-1))).item()
inb200.lib.load().inb_prof_enable(1); inb200.lib.load().inb_prof_reset()
step(); torch.cuda.synchronize()
fam = {r["name"]: round(r["ms"], 3) for r in inb200.lib.prof_table()}
inb200.lib.load().inb_prof_enable(0)
# CPU baseline: the oracle on 2 samples of the same network
from oracle import hint_oracle as H  # checker / baseline only
torch.set_num_threads(os.cpu_count())
N = H.NetworkMultiScaleHINT(2, nh, L, K, k2=k2, p2=(k2 - 1) // 2, seed=0)
Xc = torch.randn(2, 2, 128, 128)
H.hint_train_step(N, Xc)
t0 = time.time(); reps = 0
while time.time() - t0 < 8:
    H.hint_train_step(N, Xc); reps += 1
cpu = 2 * reps / (time.time() - t0)
rec = {"workload": f"cfg4: NetworkMultiScaleHINT(2,{nh},L={L},K={K};k1=3,k2={k2}) on 128x128x2", "batch": B,
       "precision": prec, "ms_per_step": ms, "samples_per_s": B / ms * 1e3, "gpu_launches_per_step": launches,
       "loss": f.item(), "invertibility": inv, "family_ms": fam,
       "cpu_baseline": {"value": cpu, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                        "sample": f"{reps} steps of batch 2 of the same network (torch-CPU oracle)"}}
print(json.dumps(rec))
if out:
    os.makedirs(os.path.dirname(out), exist_ok=True)
    open(out, "w").write(json.dumps(rec) + "\n")
