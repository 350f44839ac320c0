#!/bin/bash
O=gpurun_out/r02m; mkdir -p $O
for v in 4 5 6 12; do INB_CHAIN_MAXSTAGES=$v python scripts/step_time.py cfg2 64 fp16x3 maxstages$v | tee -a $O/ab_stages.jsonl; done
