#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
L=$O/r2_tc_sweep_h.log
: > $L
for v in tools/exp/ubench_g1 tools/ubench_tc_i8 tools/exp/ubench_g2_r6_a3; do
  for F in 0 1 2; do
    echo "# $v fmt=$F" >> $L
    timeout 200 $v $F 128 3072000 1 2>&1 | cut -c1-40,200-400 >> $L
  done
done
timeout 100 tools/ubench_tc_i8 0 8 768000 5 2>&1 | cut -c1-40,200-400 >> $L
timeout 100 tools/ubench_tc_i8 1 8 768000 5 2>&1 | cut -c1-40,200-400 >> $L
cat $L
