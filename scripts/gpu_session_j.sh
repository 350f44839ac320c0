#!/bin/bash
O=gpurun_out/r02k; mkdir -p $O
python -m pytest tests/test_gpu_hint.py -x -q > $O/pytest_hint.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_hint.log; tail -3 $O/pytest_hint.log
python scripts/step_time.py cfg4 32 bf16x3 cfg4 | tee -a $O/step.jsonl
