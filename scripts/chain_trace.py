"""Phase timeline of the fused ResidualBlock kernel (CTA 0): cycles per phase for the first tiles of
(1) the plain forward pass, (2) the recompute pass that stores the hidden tensors, (3) the backward pass."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import inb200
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
Cin, Cout = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (6, 12)
S = int(sys.argv[5]) if len(sys.argv) > 5 else 128
nh, sp = 256, (S, S)
RB = inb200.ResidualBlock(Cin, nh, n_out=Cout, k1=3, k2=1, p1=1, p2=0, precision=prec, device="cuda")
X = torch.randn(B, Cin, *sp, device="cuda"); dY = torch.randn(B, Cout, *sp, device="cuda")
for _ in range(2):
    Y = RB.forward(X); RB.backward(dY, X)
torch.cuda.synchronize()
buf = torch.zeros(4, 32, 16, dtype=torch.int64, device="cuda")
L = inb200.lib.load()
L.inb_debug_chain_trace(buf.data_ptr())
Y = RB.forward(X)
RB.backward(dY, X)
torch.cuda.synchronize()
L.inb_debug_chain_trace(None)
names = ["tile", "G1iss", "hrdy1", "G2iss", "hrdy2", "G3iss", "|", "d1full", "E1done", "d2full", "E2done", "d3full", "E3done"]
for li, title in enumerate(["forward (no store)", "recompute (stores H1, H2)", "backward (masks, stores dY2, dY1)"]):
    t = buf[li].cpu()
    t0 = t[0, 0].item()
    print(title)
    print(" ".join(n.rjust(8) for n in names))
    for i in range(10):
        r = t[i]
        if r[0].item() == 0:
            break
        v = [(x.item() - t0) for x in r[:12]]
        print(" ".join(str(x).rjust(8) for x in v[:6]) + "        | " + " ".join(str(x).rjust(8) for x in v[6:12]),
              "| TMA waits G1/G2/G3", r[12].item(), r[13].item(), r[14].item())
    f = t[16, :8]
    if f[0].item():
        names3 = ["start", "pq_read", "staged", "bar1", "summed", "fenced", "bar2", "issued"]
        print("  E3 of the third tile (thread 0), cycles after d3full: " +
              ", ".join(f"{n} {f[k].item() - t[2, 10].item()}" for k, n in enumerate(names3)))
