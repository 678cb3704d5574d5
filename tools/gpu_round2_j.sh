#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
L=$O/r2_tc_bisect_j.log
: > $L
export UBENCH_FORCE_TIME=1
for b in "" 1 2 4 8 3 5 6 7 15; do
  v=tools/exp/ubench_b$b; [ -z "$b" ] && v=tools/ubench_tc_i8
  for F in 0 1; do
    echo -n "bisect=$b fmt=$F " >> $L
    timeout 200 $v $F 128 3072000 1 0 2>&1 | grep -o '"ms".*' >> $L
  done
done
cat $L
