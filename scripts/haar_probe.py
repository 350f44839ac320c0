"""HBM throughput of the Haar squeeze kernels (k_haar<fwd/inv>) on a tensor far larger than L2 (64 x 8 x 512 x 512 floats =
537 MB in, 537 MB out): CUDA events around 10 launches each, algorithmic bytes = 8 B / element."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import inb200
X = torch.randn(64, 8, 512, 512, device="cuda")
res = {}
for name, fn, arg in (("wavelet_squeeze", inb200.wavelet_squeeze, X), ("Haar_squeeze", inb200.Haar_squeeze, X)):
    Y = fn(arg)
    inv = inb200.wavelet_unsqueeze if name == "wavelet_squeeze" else inb200.invHaar_unsqueeze
    for label, f, a in ((name, fn, arg), (inv.__name__, inv, Y)):
        for _ in range(3):
            f(a)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            f(a)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res[label] = {"ms": ms, "GB/s": 8.0 * X.numel() / ms / 1e6}
    err = (torch.linalg.norm((inv(Y) - X).reshape(-1)) / torch.linalg.norm(X.reshape(-1))).item()
    res[name]["roundtrip"] = err
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
print(json.dumps({"kernel": "k_haar", "tensor": list(X.shape), "results": res, "peaks": peaks}))
