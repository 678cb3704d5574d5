#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
L=$O/r2_tc_ubench_e.log
: > $L
for F in 1 2; do timeout 180 tools/ubench_tc_i8 $F 512 3072000 1 >> $L 2>&1; echo "# fmt=$F full rc=$?" >> $L; done
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/r2_tests_e.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke_e.log 2>&1
timeout 600 python bench.py > $O/r2_bench_e.json 2> $O/r2_bench_e.err
timeout 600 python bench.py --impl reference > $O/r2_bench_e_reference.json 2> $O/r2_bench_e_reference.err
timeout 300 python bench.py --noise-only --no-e2e > $O/r2_bench_e_noise.json 2> $O/r2_bench_e_noise.err
timeout 300 python bench.py --noise-only --no-e2e --pipeline serial > $O/r2_bench_e_noise_serial.json 2> $O/r2_bench_e_noise_serial.err
for W in c1 c2 c3; do timeout 600 python bench.py --workload $W > $O/r2_bench_e_$W.json 2> $O/r2_bench_e_$W.err; done
timeout 600 python bench.py --format sc16 --frontend tc > $O/r2_bench_e_sc16_tc.json 2> $O/r2_bench_e_sc16_tc.err
timeout 600 python bench.py --format sc8 --frontend tc --no-e2e-formats > $O/r2_bench_e_sc8_tc.json 2> $O/r2_bench_e_sc8_tc.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r02_e.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-spot-check --sustained-s 0 > $O/r2_ncu_launches_e.log 2>&1
cat $L | cut -c1-300; tail -4 $O/r2_tests_e.log; tail -8 $O/r2_smoke_e.log
for f in e e_reference e_noise e_noise_serial e_c1 e_c2 e_c3 e_sc16_tc e_sc8_tc; do echo "== $f"; cut -c1-260 $O/r2_bench_$f.json; tail -2 $O/r2_bench_$f.err; done
