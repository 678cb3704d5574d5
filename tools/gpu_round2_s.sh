#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
L=$O/r2_corrstream_s.log
: > $L
for fmt in fc32 sc16; do
for cfg in ":0" ":1" "gr-ltetrigger_b200/lib/exp/lib_r96s4.so:1" "gr-ltetrigger_b200/lib/exp/lib_r80s4.so:0" "gr-ltetrigger_b200/lib/exp/lib_r80s4.so:1" "gr-ltetrigger_b200/lib/exp/lib_r72s4.so:1"; do
  lib=${cfg%%:*}; cs=${cfg#*:}
  echo -n "fmt=$fmt lib=$lib corr_stream=$cs " >> $L
  LTB200_LIB=$lib LTB_CORR_STREAM=$cs timeout 300 python bench.py --format $fmt --no-e2e --sustained-s 0 --no-alt --no-alone 2>>$L | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(round(j['value']/1e3,1), round(j['ms_per_step'],3), {k[:5]:round(v,2) for k,v in j['roofline']['stage_ms'].items()}, j['parity_spot_check']['bit_identical_to_oracle'])" >> $L 2>&1
done; done
cat $L
