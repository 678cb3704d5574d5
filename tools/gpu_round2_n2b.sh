#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N > $O/r2_bench_n${N}b.json 2> $O/r2_bench_n${N}b.err
echo "rc=$?"; python -c "
import json,sys
j=json.loads(open('$O/r2_bench_n${N}b.json').read().strip().splitlines()[-1])
print(j['value'], j['ms_per_step'], j['e2e']['value'], j['e2e'].get('balance'), j['e2e_by_format'], j['parity_spot_check'], j['tc_vs_fp32']['decisions_identical'], j['other_frontend']['value'])
"; tail -3 $O/r2_bench_n${N}b.err
timeout 300 python -m pytest tests/test_gpu_blocks.py -m gpu -q -k "tensor_core_front_end" 2>&1 | tail -3
