#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
L=$O/r2_tc_rates_o.log
: > $L
ok=1
for v in "0 2" "0 4" "0 8" "0 12" "0 16" "1 4" "1 8" "1 12" "1 16" "2 8" "2 16"; do set -- $v
  timeout 100 tools/ubench_tc_i8 $1 8 768000 1 0 $2 2>&1 | cut -c1-110,230-420 >> $L; rc=${PIPESTATUS[0]}; echo "# fmt=$1 decim=$2 small rc=$rc" >> $L; [ $rc -ne 0 ] && ok=0
  timeout 100 tools/ubench_tc_i8 $1 8 768000 5 0 $2 2>&1 | cut -c1-110,230-420 >> $L; rc=${PIPESTATUS[0]}; echo "# fmt=$1 decim=$2 chunked rc=$rc" >> $L; [ $rc -ne 0 ] && ok=0
done
if [ $ok -eq 1 ]; then
  for v in "0 8" "0 4" "1 8" "1 4" "0 12" "1 12" "0 16" "1 16"; do set -- $v
    timeout 200 tools/ubench_tc_i8 $1 64 3072000 1 0 $2 2>&1 | cut -c1-110,230-420 >> $L
  done
  timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -x 2>&1 | tail -5 >> $L
  for d in 8 4; do for fmt in fc32 sc16; do
    echo "# bench decim=$d fmt=$fmt" >> $L
    timeout 300 python bench.py --decim $d --format $fmt --no-e2e --sustained-s 0 2>>$L | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(j['config']['frontend'][:3], j['value'], j['ms_per_step'], j['roofline']['kernel_alone']['stage_ms'], 'other', j['other_frontend']['value'], j['other_frontend']['roofline']['kernel_alone']['stage_ms'], j['parity_spot_check']['bit_identical_to_oracle'], j['tc_vs_fp32']['decisions_identical'], j['tc_vs_fp32']['max_rel_diff_psr_peak'])" >> $L 2>&1
  done; done
fi
cat $L
