#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
L=$O/r2_tc_d32_x.log
: > $L
ok=1
for v in "0 24" "0 32" "1 24" "1 32" "2 24" "2 32"; do set -- $v
  timeout 100 tools/ubench_tc_i8 $1 8 1536000 1 0 $2 2>&1 | cut -c1-110,230-420 >> $L; rc=${PIPESTATUS[0]}; echo "# fmt=$1 decim=$2 small rc=$rc" >> $L; [ $rc -ne 0 ] && ok=0
  timeout 100 tools/ubench_tc_i8 $1 8 1536000 5 0 $2 2>&1 | cut -c1-110,230-420 >> $L; rc=${PIPESTATUS[0]}; echo "# fmt=$1 decim=$2 chunked rc=$rc" >> $L; [ $rc -ne 0 ] && ok=0
done
if [ $ok -eq 1 ]; then
  timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -x 2>&1 | tail -4 >> $L
  for d in 32 24; do for fmt in sc16 fc32; do
    echo "# bench decim=$d fmt=$fmt (256 streams)" >> $L
    timeout 400 python bench.py --decim $d --format $fmt --streams 256 --no-e2e --sustained-s 0 2>>$L | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(j['config']['frontend'][:3], round(j['value']/1e3,1), round(j['ms_per_step'],3), {k[:5]:round(v,3) for k,v in j['roofline']['kernel_alone']['stage_ms'].items()}, 'other', round(j['other_frontend']['value']/1e3,1), {k[:5]:round(v,3) for k,v in j['other_frontend']['roofline']['kernel_alone']['stage_ms'].items()}, j['parity_spot_check']['bit_identical_to_oracle'], j['tc_vs_fp32']['decisions_identical'], j['tc_vs_fp32']['max_rel_diff_psr_peak'])" >> $L 2>&1
  done; done
fi
cat $L
