#!/bin/bash
# one box: GPU tests, the default bench line, ncu --set full of every kernel of a K = 1 cfg2 step, sanitizer, sampling
O=gpurun_out/r02b; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log; tail -3 $O/pytest.log
python bench.py > $O/bench_cfg2.json 2> $O/bench_cfg2.err; tail -c 600 $O/bench_cfg2.err
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -o $O/step_k1 -f \
  python scripts/step_probe.py cfg2 64 fp16x3 1 1 > $O/ncu_k1.log 2>&1; tail -2 $O/ncu_k1.log
ncu -i $O/step_k1.ncu-rep --page raw --csv > $O/step_k1_raw.csv 2>/dev/null
python scripts/ncu_summary.py $O/step_k1_raw.csv > $O/step_k1_summary.tsv; cat $O/step_k1_summary.tsv
ls -la $O/step_k1.ncu-rep
python scripts/inverse_probe.py 64 fp16x3 > $O/inverse_fp16x3.json 2>&1; tail -1 $O/inverse_fp16x3.json
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitizer_memcheck.log 2>&1; echo "sanitizer rc=$?" | tee -a $O/sanitizer_memcheck.log; tail -5 $O/sanitizer_memcheck.log
