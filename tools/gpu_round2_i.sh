#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
L=$O/r2_tc_ldg_i.log
: > $L
ok=1
for F in 0 1 2; do
  timeout 100 tools/ubench_tc_i8 $F 8 768000 1 1 2>&1 | cut -c1-100,210-420 >> $L; rc=${PIPESTATUS[0]}; echo "# fmt=$F ldg small rc=$rc" >> $L; [ $rc -ne 0 ] && ok=0
  timeout 100 tools/ubench_tc_i8 $F 8 768000 5 1 2>&1 | cut -c1-100,210-420 >> $L; rc=${PIPESTATUS[0]}; echo "# fmt=$F ldg chunked rc=$rc" >> $L; [ $rc -ne 0 ] && ok=0
done
if [ $ok -eq 1 ]; then
  for F in 0 1 2; do
    for K in 1 0; do
      timeout 200 tools/ubench_tc_i8 $F 128 3072000 1 $K 2>&1 | cut -c1-100,210-420 >> $L
    done
  done
  timeout 400 python -m pytest tests/test_gpu_tc.py -m gpu -q -x 2>&1 | tail -5 >> $L
  for fmt in fc32 sc16; do
    echo "# bench fmt=$fmt" >> $L
    timeout 300 python bench.py --format $fmt --frontend tc --no-e2e --no-spot-check --sustained-s 0 2>>$L | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['roofline']['stage_ms'])" >> $L 2>&1
    timeout 300 python bench.py --format $fmt --frontend tc --no-e2e --no-spot-check --sustained-s 0 --pipeline serial 2>>$L | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['roofline']['stage_ms'])" >> $L 2>&1
  done
fi
cat $L
