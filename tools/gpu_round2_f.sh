#!/bin/bash
# fc32 through the integer tensor-core front end (23-bit fixed point): exactness at kernel level, timing, engine tests, bench, ncu
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
L=$O/r2_tc_ubench_f.log
: > $L
ok=1
timeout 120 tools/ubench_tc_i8 0 8 768000 1 >> $L 2>&1; rc=$?; echo "# fmt=0 small rc=$rc" >> $L; [ $rc -ne 0 ] && ok=0
timeout 120 tools/ubench_tc_i8 0 8 768000 5 >> $L 2>&1; rc=$?; echo "# fmt=0 chunked rc=$rc" >> $L; [ $rc -ne 0 ] && ok=0
if [ $ok -eq 1 ]; then
  timeout 400 tools/ubench_tc_i8 0 512 3072000 1 >> $L 2>&1; rc=$?; echo "# fmt=0 full rc=$rc" >> $L; [ $rc -ne 0 ] && ok=0
fi
timeout 120 tools/ubench_tc_i8 1 64 3072000 1 >> $L 2>&1; echo "# fmt=1 rc=$?" >> $L
if [ $ok -eq 1 ]; then
  timeout 400 python -m pytest tests/test_gpu_tc.py -m gpu -q -x 2>&1 | tail -25 > $O/r2_tests_tc_f.log
  timeout 300 python bench.py --frontend tc --no-e2e > $O/r2_bench_fc32_tc_f.json 2> $O/r2_bench_fc32_tc_f.err
  timeout 300 python bench.py --frontend tc --no-e2e --pipeline serial > $O/r2_bench_fc32_tc_serial_f.json 2> $O/r2_bench_fc32_tc_serial_f.err
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:decimate_tc -s 1 -c 1 -f -o $O/tc_prof_f tools/ubench_tc_i8 0 256 3072000 1 > $O/r2_tc_ncu_f.log 2>&1
fi
cat $L | cut -c1-420; tail -8 $O/r2_tests_tc_f.log 2>/dev/null
for f in fc32_tc fc32_tc_serial; do cut -c1-300 $O/r2_bench_${f}_f.json 2>/dev/null; tail -2 $O/r2_bench_${f}_f.err 2>/dev/null; done
tail -2 $O/r2_tc_ncu_f.log 2>/dev/null
