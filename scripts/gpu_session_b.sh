#!/bin/bash
O=gpurun_out/r02c; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log; tail -3 $O/pytest.log
for h in 0 1 2 3; do INB_L2_HINTS=$h python scripts/step_time.py cfg2 64 fp16x3 hints$h | tee -a $O/ab_hints.jsonl; done
INB_L2_HINTS=0 python scripts/step_time.py cfg2 8 fp16x3 hints0_b8 | tee -a $O/ab_hints.jsonl
INB_L2_HINTS=3 python scripts/step_time.py cfg2 8 fp16x3 hints3_b8 | tee -a $O/ab_hints.jsonl
