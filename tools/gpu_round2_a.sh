#!/bin/bash
# one GPU-box session: TC prototype variants, full GPU suite, smoke, bench variants (everything under timeout)
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
if [ -n "$GOOD" ]; then
  timeout 120 tools/ubench_tc_i8 $GOOD 8 768000 5 >> $O/r2_tc_ubench.log 2>&1; echo "# G=$GOOD chunked rc=$?" >> $O/r2_tc_ubench.log
  timeout 120 tools/ubench_tc_i8 $GOOD 512 3072000 1 >> $O/r2_tc_ubench.log 2>&1; echo "# G=$GOOD full size rc=$?" >> $O/r2_tc_ubench.log
fi
timeout 900 python -m pytest tests -m gpu -x -q --ignore=tests/test_gpu_tc.py 2>&1 | tail -15 > $O/r2_tests_a.log
if [ -n "$GOOD" ]; then
  LTB_TC_G=$GOOD timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q 2>&1 | tail -25 > $O/r2_tests_tc.log
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke_a.log 2>&1
timeout 600 python bench.py > $O/r2_bench_overlap.json 2> $O/r2_bench_overlap.err
timeout 300 python bench.py --pipeline serial --no-e2e > $O/r2_bench_serial.json 2> $O/r2_bench_serial.err
L=$PWD/gr-ltetrigger_b200/lib/libltetrigger_b200_r128.so
if [ -f $L ]; then
  LTB200_LIB=$L timeout 300 python bench.py --pipeline serial --no-e2e --no-spot-check > $O/r2_bench_serial_r128.json 2>&1
  LTB200_LIB=$L timeout 300 python bench.py --pipeline overlap --no-e2e --no-spot-check > $O/r2_bench_overlap_r128.json 2>&1
fi
timeout 300 python bench.py --format sc16 --no-e2e > $O/r2_bench_sc16_fp32.json 2> $O/r2_bench_sc16_fp32.err
if [ -n "$GOOD" ]; then
  LTB_TC_G=$GOOD timeout 300 python bench.py --format sc16 --frontend tc --no-e2e > $O/r2_bench_sc16_tc.json 2> $O/r2_bench_sc16_tc.err
fi
echo "=== ubench"; cat $O/r2_tc_ubench.log | cut -c1-300
echo "=== tests"; tail -4 $O/r2_tests_a.log; echo "=== tc tests"; tail -8 $O/r2_tests_tc.log 2>/dev/null
echo "=== smoke"; tail -6 $O/r2_smoke_a.log
for f in overlap serial serial_r128 overlap_r128 sc16_fp32 sc16_tc; do echo "== $f"; cut -c1-230 $O/r2_bench_$f.json 2>/dev/null; done
tail -3 $O/r2_bench_overlap.err $O/r2_bench_sc16_tc.err 2>/dev/null
