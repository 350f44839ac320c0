#!/bin/bash
O=gpurun_out/r02l; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_rb_chain2' --launch-skip 7 --launch-count 2 -o $O/chain -f \
  python scripts/step_probe.py cfg2 64 fp16x3 1 1 > $O/ncu.log 2>&1
ls -la $O/chain.ncu-rep
