#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
L=$O/r2_track_q.log
: > $L
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_blocks.py -m gpu -q -x 2>&1 | tail -5 >> $L
for extra in "" "--noise-only" "--format sc16" "--decim 1 --streams 256 --frontend fp32"; do
  echo "# bench $extra" >> $L
  timeout 300 python bench.py $extra --no-e2e --sustained-s 0 --no-alt 2>>$L | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['roofline'].get('kernel_alone',{}).get('stage_ms'), j['parity_spot_check']['bit_identical_to_oracle'])" >> $L 2>&1
done
cat $L
