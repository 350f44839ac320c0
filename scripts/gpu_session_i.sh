#!/bin/bash
O=gpurun_out/r02j; mkdir -p $O
python -m pytest tests/test_gpu_tc.py tests/test_gpu_networks.py tests/test_gpu_layers.py -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log; tail -3 $O/pytest.log
python scripts/step_time.py cfg2 8 fp16x3 b8 | tee -a $O/step.jsonl
python scripts/step_time.py cfg2 64 fp16x3 b64 | tee -a $O/step.jsonl
