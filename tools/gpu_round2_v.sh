#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
L=$O/r2_cfo_v.log
: > $L
for cfo in 0 500 3000 7000; do
  echo -n "cfo=$cfo " >> $L
  timeout 300 python bench.py --cfo-hz $cfo --no-e2e --sustained-s 0 --no-alt 2>>$L | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(round(j['value']/1e3,1), round(j['ms_per_step'],3), {k[:5]:round(v,3) for k,v in j['roofline']['kernel_alone']['stage_ms'].items()}, j['config']['cells_tagged_per_step'], j['parity_spot_check']['bit_identical_to_oracle'], j['tc_vs_fp32']['decisions_identical'])" >> $L 2>&1
done
cat $L
