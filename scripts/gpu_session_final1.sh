#!/bin/bash
# single-GPU evidence of the round's final state (everything lands under gpurun_out/r02_final/, no .ncu-rep kept)
O=gpurun_out/r02_final; mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log; tail -3 $O/smoke.log
python bench.py > $O/bench_cfg2_1gpu.json 2> $O/bench_cfg2_1gpu.err; echo "bench rc=$?"
python bench.py --impl reference > $O/bench_cfg2_reference.json 2> $O/bench_cfg2_reference.err; echo "ref rc=$?"
for c in cfg1 cfg3 cfg4 cfg5; do python bench.py --config $c > $O/bench_${c}_1gpu.json 2> $O/bench_${c}_1gpu.err; echo "$c rc=$?"; done
python scripts/inverse_probe.py 64 fp16x3 > $O/inverse_cfg2_fp16x3.log 2>&1; tail -1 $O/inverse_cfg2_fp16x3.log > $O/inverse_cfg2_fp16x3.json
for b in 64 8; do
  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file $O/launches_b$b.csv python scripts/step_probe.py cfg2 $b fp16x3 > $O/launches_b$b.out 2>&1
  python scripts/launch_summary.py $O/launches_b$b.csv > $O/launches_b$b.tsv; tail -1 $O/launches_b$b.tsv
done
timeout 900 ncu --set full --clock-control none --profile-from-start off -o /tmp/step_k1 -f python scripts/step_probe.py cfg2 64 fp16x3 1 1 > $O/ncu_k1.log 2>&1
ncu -i /tmp/step_k1.ncu-rep --page raw --csv > /tmp/step_k1_raw.csv 2>/dev/null
python scripts/ncu_summary.py /tmp/step_k1_raw.csv > $O/step_k1_summary.tsv; wc -l $O/step_k1_summary.tsv
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitizer_memcheck.log 2>&1; echo "sanitizer rc=$?" | tee -a $O/sanitizer_memcheck.log
du -sh $O
