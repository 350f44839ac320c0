#!/bin/bash
O=gpurun_out/r02g; mkdir -p $O
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:'k_wgrad2_tc' -o /tmp/wg -f \
  python scripts/step_probe.py cfg2 64 fp16x3 1 1 > $O/ncu.log 2>&1
ncu -i /tmp/wg.ncu-rep --page raw --csv > $O/wg_raw.csv 2>/dev/null
python scripts/ncu_summary.py $O/wg_raw.csv > $O/wg_summary.tsv; cat $O/wg_summary.tsv
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02g/wg_raw.csv')))
hdr=rows[0]; data=rows[2:]
want=[h for h in hdr if any(k in h for k in ('stall','issue_active','warp_cycles_per_issued','smsp__pcsamp','l1tex__data_bank_conflicts','smsp__average_warp'))]
idx={h:i for i,h in enumerate(hdr)}
for r in data[-3:]:
    print(r[idx['Kernel Name']][:20], r[idx['gpu__time_duration.sum']])
    for h in want[:40]:
        print('   ',h,r[idx[h]])
PY
