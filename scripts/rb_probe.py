"""One ResidualBlock forward+backward at cfg2's scale-1 shape (profiling probe for ncu)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import inb200
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
Cin, nh, Cout, sp = 6, 256, 12, (128, 128)
RB = inb200.ResidualBlock(Cin, nh, n_out=Cout, k1=3, k2=1, p1=1, p2=0, precision=prec, device="cuda")
X = torch.randn(B, Cin, *sp, device="cuda"); dY = torch.randn(B, Cout, *sp, device="cuda")
for _ in range(3):
    Y = RB.forward(X); dX = RB.backward(dY, X)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    Y = RB.forward(X); dX = RB.backward(dY, X)
e1.record(); torch.cuda.synchronize()
print("rb fwd+bwd ms", e0.elapsed_time(e1) / 5)
L = inb200.lib.load(); L.inb_prof_reset(); L.inb_prof_enable(1)
for _ in range(3):
    Y = RB.forward(X); dX = RB.backward(dY, X)
torch.cuda.synchronize(); L.inb_prof_enable(0)
for r in sorted(inb200.lib.prof_table(), key=lambda r: -r["ms"]):
    print(f"  {r['name']:16s} {r['ms']/3:8.3f} ms/iter  launches/iter {r['launches']/3:5.1f}  "
          f"{(r['flops']/r['ms']/1e9 if r['ms'] else 0):8.1f} TF/s  {(r['bytes']/r['ms']/1e6 if r['ms'] else 0):8.1f} GB/s")
