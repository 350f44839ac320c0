"""Device time of one training step of a BASELINE configuration, with the per-family table of the library's profiler:

  python scripts/step_time.py cfg2 64 fp16x3 [label]

Environment switches of the library (INB_L2_HINTS, INB_FUSE_COUPLING, ...) are read at first use, so an A/B is two
processes."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
import inb200  # noqa: E402


def main():
    cfg_name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    cfg = bench.CONFIGS[cfg_name]
    B = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["gb"]
    prec = sys.argv[3] if len(sys.argv) > 3 else cfg["precision"]
    label = sys.argv[4] if len(sys.argv) > 4 else ""
    steps = 5
    W = bench.Workload(cfg, prec, B, 0, torch.device("cuda", 0))
    for _ in range(3):
        W.step()
    torch.cuda.synchronize()
    n0 = inb200.lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        W.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    launches = (inb200.lib.launch_count() - n0) // steps
    L = inb200.lib.load()
    L.inb_prof_reset()
    L.inb_prof_enable(1)
    for _ in range(steps):
        W.step()
    torch.cuda.synchronize()
    L.inb_prof_enable(0)
    fam = {r["name"]: round(r["ms"] / steps, 3) for r in sorted(inb200.lib.prof_table(), key=lambda r: -r["ms"])}
    print(json.dumps({"label": label, "config": cfg_name, "batch": B, "precision": prec, "samples_per_s": B / ms * 1e3,
                      "ms_per_step": ms, "launches_per_step": launches, "families_ms": fam}))


if __name__ == "__main__":
    main()
