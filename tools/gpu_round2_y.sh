#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
L=$O/r2_tc_hint_y.log
: > $L
for v in tools/ubench_tc_i8 tools/exp/ubench_h100 tools/exp/ubench_h400 tools/exp/ubench_h2000 tools/ubench_tc_i8; do
  for F in 0 1; do
    echo -n "$v fmt=$F " >> $L
    timeout 200 $v $F 64 3072000 1 2>&1 | grep -o '"mismatches".*' | cut -c1-20,60-200 >> $L
  done
done
cat $L
