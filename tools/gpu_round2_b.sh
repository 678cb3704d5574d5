#!/bin/bash
# TC front end session: prototype variants (each its own process, under timeout), TC tests, sc16 benches
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
: > $O/r2_tc_ubench.log
GOOD=""
for G in 1 2 4 8; do
  timeout 120 tools/ubench_tc_i8 $G 8 768000 1 >> $O/r2_tc_ubench.log 2>&1; rc=$?
  echo "# G=$G single-call rc=$rc" >> $O/r2_tc_ubench.log
  if [ $rc -eq 0 ] && [ -z "$GOOD" ]; then GOOD=$G; fi
done
echo "# first good G: '$GOOD'" >> $O/r2_tc_ubench.log
if [ -z "$GOOD" ]; then
  timeout 120 tools/ubench_tc_i8_notma 1 8 768000 1 >> $O/r2_tc_ubench.log 2>&1; echo "# notma G=1 rc=$?" >> $O/r2_tc_ubench.log
  timeout 120 tools/ubench_tc_i8_notma 8 8 768000 1 >> $O/r2_tc_ubench.log 2>&1; echo "# notma G=8 rc=$?" >> $O/r2_tc_ubench.log
fi
if [ -n "$GOOD" ]; then
  timeout 120 tools/ubench_tc_i8 $GOOD 8 768000 5 >> $O/r2_tc_ubench.log 2>&1; echo "# G=$GOOD chunked rc=$?" >> $O/r2_tc_ubench.log
  timeout 120 tools/ubench_tc_i8 $GOOD 512 3072000 1 >> $O/r2_tc_ubench.log 2>&1; echo "# G=$GOOD full size rc=$?" >> $O/r2_tc_ubench.log
  LTB_TC_G=$GOOD timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -x 2>&1 | tail -25 > $O/r2_tests_tc.log
  LTB_TC_G=$GOOD timeout 300 python bench.py --format sc16 --frontend tc --no-e2e > $O/r2_bench_sc16_tc.json 2> $O/r2_bench_sc16_tc.err
fi
echo "=== ubench"; cat $O/r2_tc_ubench.log | cut -c1-400
echo "=== tc tests"; tail -12 $O/r2_tests_tc.log 2>/dev/null
echo "== sc16_tc"; cut -c1-300 $O/r2_bench_sc16_tc.json 2>/dev/null; tail -3 $O/r2_bench_sc16_tc.err 2>/dev/null
