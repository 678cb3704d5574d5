#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/r2_tests_w.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke_w.log 2>&1
timeout 900 python bench.py > $O/r2_bench_w.json 2> $O/r2_bench_w.err
timeout 600 python bench.py --impl reference > $O/r2_bench_w_reference.json 2> $O/r2_bench_w_reference.err
timeout 600 python bench.py --format sc16 --no-e2e-formats > $O/r2_bench_w_sc16.json 2> $O/r2_bench_w_sc16.err
timeout 600 python bench.py --format sc8 --no-e2e-formats > $O/r2_bench_w_sc8.json 2> $O/r2_bench_w_sc8.err
timeout 600 python bench.py --noise-only --no-e2e > $O/r2_bench_w_noise.json 2> $O/r2_bench_w_noise.err
for W in c1 c2 c3; do timeout 600 python bench.py --workload $W > $O/r2_bench_w_$W.json 2> $O/r2_bench_w_$W.err; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:ltb|decimate|pss_|sss_|tail_kernel|chain_order|ingest" -c 60 --csv --log-file $O/launches_r02_w.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-spot-check --sustained-s 0 --no-alt --no-alone > $O/r2_ncu_launches_w.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:decimate_tc_kernel|pss_corr_fft|pss_track" -s 9 -c 3 -f -o $O/prof_r02_w python bench.py --no-e2e --no-spot-check --sustained-s 0 --no-alt --no-alone --steps 2 --pipeline serial > $O/r2_ncu_full_w.log 2>&1
tail -4 $O/r2_tests_w.log; tail -10 $O/r2_smoke_w.log
for f in w w_reference w_sc16 w_sc8 w_noise w_c1 w_c2 w_c3; do echo "== $f"; cut -c1-200 $O/r2_bench_$f.json; tail -2 $O/r2_bench_$f.err; done
tail -2 $O/r2_ncu_full_w.log | cut -c1-200
