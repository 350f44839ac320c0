#!/bin/bash
# usage: gpu_session_multi.sh N   (run under gpurun --gpus N)
N=$1; O=gpurun_out/r02_final; mkdir -p $O
if [ "$N" = "2" ]; then python -m pytest tests/test_gpu_dp.py -q > $O/pytest_gpu_dp.log 2>&1; echo "dp tests rc=$?" | tee -a $O/pytest_gpu_dp.log; tail -3 $O/pytest_gpu_dp.log; fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N > $O/bench_cfg2_${N}gpu.json 2> $O/bench_cfg2_${N}gpu.err; echo "bench N=$N rc=$?"
python - <<PY
import json
d=json.loads(open("$O/bench_cfg2_${N}gpu.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("dp"))
PY
