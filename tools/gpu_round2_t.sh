#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
L=$O/r2_walk_t.log
: > $L
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_blocks.py -m gpu -q -x 2>&1 | tail -4 >> $L
timeout 300 python bench.py --no-e2e --sustained-s 0 --no-alt 2>>$L | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['roofline'].get('kernel_alone',{}).get('stage_ms'), j['parity_spot_check']['bit_identical_to_oracle'])" >> $L 2>&1
timeout 600 python bench.py --workload c2 2>>$L | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('c2', j['value'], j['single_stream']['c_abi'], j['single_stream']['parity'])" >> $L 2>&1
cat $L
