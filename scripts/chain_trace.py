"""Phase timeline of the fused ResidualBlock kernel (CTA 0): cycles per phase for the first tiles."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import inb200
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
Cin, nh, Cout, sp = 6, 256, 12, (128, 128)
RB = inb200.ResidualBlock(Cin, nh, n_out=Cout, k1=3, k2=1, p1=1, p2=0, precision=prec, device="cuda")
X = torch.randn(B, Cin, *sp, device="cuda"); dY = torch.randn(B, Cout, *sp, device="cuda")
for _ in range(2):
    Y = RB.forward(X)
torch.cuda.synchronize()
buf = torch.zeros(32, 16, dtype=torch.int64, device="cuda")
L = inb200.lib.load()
L.inb_debug_chain_trace(buf.data_ptr())
Y = RB.forward(X)
torch.cuda.synchronize()
L.inb_debug_chain_trace(None)
t = buf.cpu()
t0 = t[0, 0].item()
names = ["tile", "G1iss", "hrdy1", "G2iss", "hrdy2", "G3iss", "|", "d1full", "E1done", "d2full", "E2done", "d3full", "E3done"]
print(" ".join(n.rjust(8) for n in names))
for i in range(12):
    r = t[i]
    if r[0].item() == 0:
        break
    v = [(x.item() - t0) for x in r[:12]]
    e3 = [(x.item() - t0) for x in t[16 + i][:3]]
    print(" ".join(str(x).rjust(8) for x in v[:6]) + "        | " + " ".join(str(x).rjust(8) for x in v[6:12]),
          "| waits G1/G2/G3", r[12].item(), r[13].item(), r[14].item(), "| E3: sts", e3[0], "bar", e3[1], "out", e3[2])
