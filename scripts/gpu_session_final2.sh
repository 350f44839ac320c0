#!/bin/bash
# refresh of the lines that the last code changes touch (full evidence: gpu_session_final1.sh)
O=gpurun_out/r02_final; mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
python bench.py > $O/bench_cfg2_1gpu.json 2> $O/bench_cfg2_1gpu.err; echo "bench rc=$?"
python bench.py --config cfg4 > $O/bench_cfg4_1gpu.json 2> $O/bench_cfg4_1gpu.err; echo "cfg4 rc=$?"
for b in 64 8; do
  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file $O/launches_b$b.csv python scripts/step_probe.py cfg2 $b fp16x3 > $O/launches_b$b.out 2>&1
  python scripts/launch_summary.py $O/launches_b$b.csv > $O/launches_b$b.tsv; tail -1 $O/launches_b$b.tsv
done
python scripts/step_time.py cfg2 8 fp16x3 final_b8 | tee $O/step_b8.json
