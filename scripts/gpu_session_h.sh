#!/bin/bash
O=gpurun_out/r02i; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log; tail -4 $O/pytest.log
for v in 0 1; do INB_CHAIN_QWIDE=$v python scripts/step_time.py cfg2 64 fp16x3 qwide$v | tee -a $O/step.jsonl; done
for v in 0 1; do INB_CHAIN_QWIDE=$v python scripts/step_time.py cfg2 8 fp16x3 qwide${v}_b8 | tee -a $O/step.jsonl; done
