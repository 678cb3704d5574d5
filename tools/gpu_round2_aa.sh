#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
L=$O/r2_tc_epi_aa.log
: > $L
ok=1
for v in "0 16" "1 16" "2 16" "1 4" "0 12" "2 32"; do set -- $v
  timeout 100 tools/ubench_tc_i8 $1 8 1536000 5 0 $2 2>&1 | cut -c1-110,230-420 >> $L; rc=${PIPESTATUS[0]}; echo "# fmt=$1 decim=$2 chunked rc=$rc" >> $L; [ $rc -ne 0 ] && ok=0
done
export UBENCH_FORCE_TIME=1
for v in tools/ubench_tc_i8_old tools/ubench_tc_i8_prod tools/ubench_tc_i8_old tools/ubench_tc_i8_prod; do
  for F in 0 1 2; do
    echo -n "$v fmt=$F " >> $L
    timeout 200 $v $F 64 3072000 1 2>&1 | grep -o '"mismatches".*' | cut -c1-20,60-200 >> $L
  done
done
if [ $ok -eq 1 ]; then
  timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -x 2>&1 | tail -3 >> $L
  for fmt in fc32 sc16; do
    timeout 300 python bench.py --format $fmt --no-e2e --sustained-s 0 --no-alt 2>>$L | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(round(j['value']/1e3,1), round(j['ms_per_step'],3), {k[:5]:round(v,3) for k,v in j['roofline']['kernel_alone']['stage_ms'].items()}, j['parity_spot_check']['bit_identical_to_oracle'])" >> $L 2>&1
  done
fi
cat $L
