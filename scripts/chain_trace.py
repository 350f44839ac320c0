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
    fine = t[16:24, :9]
    if fine[0, 0].item():
        print("  epilogue warp 0, third tile: per chunk (E1 c0..3, E2 c0..3), cycles relative to the tile's d1full:")
        print("  " + " ".join(n.rjust(8) for n in ["ldwait0", "ldwait1", "emit", "packed", "st_iss", "sfree", "staged", "st_done", "arrived"]))
        base = t[2, 6].item()
        for rr in fine:
            v = [rr[7].item(), rr[8].item(), rr[0].item(), rr[1].item(), rr[2].item(), rr[3].item(), rr[4].item(), rr[5].item(), rr[6].item()]
            print("  " + " ".join((str(x - base) if x else "-").rjust(8) for x in v))
