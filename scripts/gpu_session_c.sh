#!/bin/bash
O=gpurun_out/r02d; mkdir -p $O
python -m pytest tests/test_gpu_tc.py tests/test_gpu_layers.py tests/test_gpu_networks.py -x -q > $O/pytest_tc.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_tc.log; tail -25 $O/pytest_tc.log
for v in 0 1; do INB_PLANE_LO8=$v python scripts/step_time.py cfg2 64 fp16x3 lo8_$v | tee -a $O/ab_lo8.jsonl; done
for v in 0 1; do INB_PLANE_LO8=$v python scripts/step_time.py cfg2 8 fp16x3 lo8_${v}_b8 | tee -a $O/ab_lo8.jsonl; done
