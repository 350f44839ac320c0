#!/bin/bash
O=gpurun_out/r02f; mkdir -p $O
for v in 1 0; do
INB_PLANE_LO8=$v timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:'k_wgrad2_tc|k_rb_chain2' -o /tmp/wg_$v -f \
  python scripts/step_probe.py cfg2 64 fp16x3 1 1 > $O/ncu_$v.log 2>&1
ncu -i /tmp/wg_$v.ncu-rep --page raw --csv > $O/wg_lo8_${v}_raw.csv 2>/dev/null
python scripts/ncu_summary.py $O/wg_lo8_${v}_raw.csv > $O/wg_lo8_${v}_summary.tsv; echo "== lo8=$v"; cat $O/wg_lo8_${v}_summary.tsv
done
ls -la /tmp/*.ncu-rep
