#!/bin/bash
O=gpurun_out/r02e; mkdir -p $O
for v in 1 0; do
INB_PLANE_LO8=$v python -m pytest tests/test_gpu_fullsize.py -x -q -s -k "cfg2_full_size and fp16x3" > $O/full_lo8_$v.log 2>&1; echo "lo8=$v rc=$?"
grep -A14 "== parity" $O/full_lo8_$v.log | head -20; grep "all 480" $O/full_lo8_$v.log
cp gpurun_out/parity_cfg2_fp16x3.json $O/parity_cfg2_fp16x3_lo8_$v.json
done
