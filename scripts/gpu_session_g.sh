#!/bin/bash
O=gpurun_out/r02h; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log; tail -4 $O/pytest.log
python scripts/step_time.py cfg2 64 fp16x3 ones | tee -a $O/step.jsonl
python scripts/step_time.py cfg2 8 fp16x3 ones_b8 | tee -a $O/step.jsonl
