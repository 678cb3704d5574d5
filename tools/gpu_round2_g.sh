#!/bin/bash
# tuning sweep: pipeline depths of the tensor-core front end; register budget vs co-residency with the tracker
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
L=$O/r2_tc_sweep_g.log
: > $L
for v in tools/ubench_tc_i8 tools/exp/ubench_r4_a2 tools/exp/ubench_r6_a2 tools/exp/ubench_r6_a3 tools/exp/ubench_r8_a2 tools/exp/ubench_r8_a3; do
  for F in 0 1; do
    echo "# $v fmt=$F" >> $L
    timeout 200 $v $F 128 3072000 1 2>&1 | cut -c1-40,200-400 >> $L
  done
done
for lib in "" gr-ltetrigger_b200/lib/exp/lib_r80.so gr-ltetrigger_b200/lib/exp/lib_r64.so; do
  for fmt in fc32 sc16; do
    echo "# lib=$lib fmt=$fmt" >> $L
    LTB200_LIB=$lib timeout 300 python bench.py --format $fmt --frontend tc --no-e2e --no-spot-check --sustained-s 0 2>>$L | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['roofline']['stage_ms'])" >> $L 2>&1
  done
done
cat $L
