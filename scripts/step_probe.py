"""One training step of a BASELINE configuration between cudaProfilerStart / cudaProfilerStop, for ncu:

  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/step_probe.py cfg2 64 fp16x3

CUDA graphs are switched off (INB_GRAPHS=0) so that every kernel is a launch of its own with its name; warm-up steps
run before the profiled one."""
import os
import sys

os.environ.setdefault("INB_GRAPHS", "0")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    cfg_name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    cfg = bench.CONFIGS[cfg_name]
    B = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["gb"]
    prec = sys.argv[3] if len(sys.argv) > 3 else cfg["precision"]
    nsteps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    if len(sys.argv) > 5:  # flow steps per scale (a K = 1 network has every kernel of the step in ~40 launches)
        cfg = dict(cfg, K=int(sys.argv[5]))
    dev = torch.device("cuda", 0)
    W = bench.Workload(cfg, prec, B, 0, dev)
    for _ in range(3):
        W.step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(nsteps):
        W.step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled", nsteps, "step(s) of", cfg_name, "batch", B, prec)


if __name__ == "__main__":
    main()
