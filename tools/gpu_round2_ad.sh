#!/bin/bash
# time-segment sharding on the engine: the three new GPU tests, the sequential-vs-segmented bench, smoke
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_segments.py -m gpu -q -x 2>&1 | tail -15 | tee $O/r2_segments_tests_ad.log
timeout 300 python tools/bench_segments.py 8 sc16 fc32 > $O/r2_bench_segments_ad.json 2> $O/r2_bench_segments_ad.err; echo "bench_segments rc=$?"; tail -3 $O/r2_bench_segments_ad.err; cat $O/r2_bench_segments_ad.json
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -12 | tee $O/r2_smoke_ad.log
