#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
: > $O/r2_tc_ubench_c.log
timeout 120 tools/ubench_tc_i8 1 8 768000 1 >> $O/r2_tc_ubench_c.log 2>&1; echo "# small rc=$?" >> $O/r2_tc_ubench_c.log
timeout 120 tools/ubench_tc_i8 1 8 768000 5 >> $O/r2_tc_ubench_c.log 2>&1; echo "# chunked rc=$?" >> $O/r2_tc_ubench_c.log
timeout 120 tools/ubench_tc_i8 1 512 3072000 1 >> $O/r2_tc_ubench_c.log 2>&1; rc=$?; echo "# full rc=$rc" >> $O/r2_tc_ubench_c.log
if [ $rc -eq 0 ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:decimate_tc -s 1 -c 1 -f -o $O/tc_prof_c tools/ubench_tc_i8 1 512 3072000 1 > $O/r2_tc_ncu_c.log 2>&1
fi
cat $O/r2_tc_ubench_c.log | cut -c1-420; tail -3 $O/r2_tc_ncu_c.log
