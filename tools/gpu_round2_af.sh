#!/bin/bash
# final tree: the whole GPU suite, then the default bench line
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > $O/r2_gpu_tests_af.log; cat $O/r2_gpu_tests_af.log
timeout 300 python bench.py > $O/r2_bench_af_default.json 2> $O/r2_bench_af_default.err; echo "bench rc=$?"; cut -c1-400 $O/r2_bench_af_default.json
