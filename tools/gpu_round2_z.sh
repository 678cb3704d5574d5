#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
L=$O/r2_tc_diet_z.log
: > $L
export UBENCH_FORCE_TIME=1
for v in tools/ubench_tc_i8_old tools/ubench_tc_i8 tools/ubench_tc_i8_prod tools/ubench_tc_i8_old tools/ubench_tc_i8_prod; do
  for F in 0 1 2; do
    echo -n "$v fmt=$F " >> $L
    timeout 200 $v $F 64 3072000 1 2>&1 | grep -o '"mismatches".*' | cut -c1-20,60-200 >> $L
  done
done
cat $L
