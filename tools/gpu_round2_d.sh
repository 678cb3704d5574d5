#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
L=$O/r2_tc_ubench_d.log
: > $L
ok=1
for F in 1 2; do
  timeout 120 tools/ubench_tc_i8 $F 8 768000 1 >> $L 2>&1; rc=$?; echo "# fmt=$F small rc=$rc" >> $L; [ $rc -ne 0 ] && ok=0
  timeout 120 tools/ubench_tc_i8 $F 8 768000 5 >> $L 2>&1; rc=$?; echo "# fmt=$F chunked rc=$rc" >> $L; [ $rc -ne 0 ] && ok=0
  timeout 180 tools/ubench_tc_i8 $F 512 3072000 1 >> $L 2>&1; rc=$?; echo "# fmt=$F full rc=$rc" >> $L; [ $rc -ne 0 ] && ok=0
done
if [ $ok -eq 1 ]; then
  timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -x 2>&1 | tail -25 > $O/r2_tests_tc_d.log
  timeout 300 python bench.py --format sc16 --frontend tc --no-e2e > $O/r2_bench_sc16_tc_d.json 2> $O/r2_bench_sc16_tc_d.err
  timeout 300 python bench.py --format sc8 --frontend tc --no-e2e > $O/r2_bench_sc8_tc_d.json 2> $O/r2_bench_sc8_tc_d.err
  timeout 300 python bench.py --format sc8 --no-e2e > $O/r2_bench_sc8_fp32_d.json 2> $O/r2_bench_sc8_fp32_d.err
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:decimate_tc -s 1 -c 1 -f -o $O/tc_prof_d tools/ubench_tc_i8 1 512 3072000 1 > $O/r2_tc_ncu_d.log 2>&1
fi
cat $L | cut -c1-420; tail -6 $O/r2_tests_tc_d.log 2>/dev/null
for f in sc16_tc sc8_tc sc8_fp32; do cut -c1-220 $O/r2_bench_${f}_d.json 2>/dev/null; tail -2 $O/r2_bench_${f}_d.err 2>/dev/null; done
tail -2 $O/r2_tc_ncu_d.log 2>/dev/null
