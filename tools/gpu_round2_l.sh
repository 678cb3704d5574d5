#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
L=$O/r2_tc_sweep_l.log
: > $L
for v in tools/ubench_tc_i8 tools/exp/ubench_g2_r6_a3 tools/exp/ubench_g2_r4_a3 tools/exp/ubench_g2_r6_a4 tools/exp/ubench_g2_r8_a3 tools/ubench_tc_i8 tools/exp/ubench_g2_r6_a3; do
  for F in 0 1; do
    echo -n "$v fmt=$F " >> $L
    timeout 200 $v $F 64 3072000 1 2>&1 | grep -o '"mismatches".*' | cut -c1-20,60-200 >> $L
  done
done
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/r2_tests_l.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke_l.log 2>&1
timeout 900 python bench.py > $O/r2_bench_l.json 2> $O/r2_bench_l.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decimate_tc -s 3 -c 1 -f -o $O/tc_prof_l python bench.py --no-e2e --no-spot-check --sustained-s 0 --no-alt --no-alone --steps 2 > $O/r2_tc_ncu_l.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r02_l.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-spot-check --sustained-s 0 --no-alt --no-alone > $O/r2_ncu_launches_l.log 2>&1
cat $L; tail -4 $O/r2_tests_l.log; tail -9 $O/r2_smoke_l.log; cut -c1-400 $O/r2_bench_l.json; tail -3 $O/r2_bench_l.err; tail -2 $O/r2_tc_ncu_l.log
