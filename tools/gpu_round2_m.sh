#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
L=$O/r2_pipe_m.log
: > $L
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pipeline_modes or in_flight" 2>&1 | tail -5 >> $L
for fe in tc fp32; do for pipe in overlap overlap_corr serial; do
  echo -n "frontend=$fe pipeline=$pipe " >> $L
  timeout 300 python bench.py --frontend $fe --pipeline $pipe --no-e2e --no-spot-check --sustained-s 0 --no-alt --no-alone 2>>$L | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['roofline']['stage_ms'])" >> $L 2>&1
done; done
echo -n "sc16 tc overlap_corr " >> $L
timeout 300 python bench.py --format sc16 --pipeline overlap_corr --no-e2e --no-spot-check --sustained-s 0 --no-alt --no-alone 2>>$L | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['roofline']['stage_ms'])" >> $L 2>&1
echo -n "sc16 tc overlap " >> $L
timeout 300 python bench.py --format sc16 --pipeline overlap --no-e2e --no-spot-check --sustained-s 0 --no-alt --no-alone 2>>$L | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['roofline']['stage_ms'])" >> $L 2>&1
cat $L
