#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/r2_tests_p.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke_p.log 2>&1
timeout 900 python bench.py > $O/r2_bench_p.json 2> $O/r2_bench_p.err
timeout 600 python bench.py --impl reference > $O/r2_bench_p_reference.json 2> $O/r2_bench_p_reference.err
timeout 600 python bench.py --format sc16 --no-e2e-formats > $O/r2_bench_p_sc16.json 2> $O/r2_bench_p_sc16.err
timeout 600 python bench.py --format sc8 --no-e2e-formats > $O/r2_bench_p_sc8.json 2> $O/r2_bench_p_sc8.err
timeout 600 python bench.py --decim 12 --format sc16 --no-e2e --sustained-s 0 > $O/r2_bench_p_d12_sc16.json 2> $O/r2_bench_p_d12_sc16.err
timeout 600 python bench.py --decim 8 --no-e2e --sustained-s 0 > $O/r2_bench_p_d8.json 2> $O/r2_bench_p_d8.err
timeout 600 python bench.py --noise-only --no-e2e > $O/r2_bench_p_noise.json 2> $O/r2_bench_p_noise.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:ltb|decimate|pss_|sss_|tail_kernel|chain_order|ingest" -c 60 --csv --log-file $O/launches_r02_p.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-spot-check --sustained-s 0 --no-alt --no-alone > $O/r2_ncu_launches_p.log 2>&1
timeout 400 python tests/fuzz_parity.py --seconds 240 --seed 11 > $O/r2_fuzz_p.log 2>&1
tail -4 $O/r2_tests_p.log; tail -10 $O/r2_smoke_p.log
for f in p p_reference p_sc16 p_sc8 p_d12_sc16 p_d8 p_noise; do echo "== $f"; cut -c1-230 $O/r2_bench_$f.json; tail -2 $O/r2_bench_$f.err; done
tail -3 $O/r2_fuzz_p.log; grep -c '^ok' $O/r2_fuzz_p.log; grep -c 'fe=tc' $O/r2_fuzz_p.log
