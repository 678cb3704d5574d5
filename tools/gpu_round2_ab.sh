#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
for W in c1 c2 c3; do timeout 600 python bench.py --workload $W > $O/r2_bench_ab_$W.json 2> $O/r2_bench_ab_$W.err; python -c "
import json
j=json.load(open('$O/r2_bench_ab_$W.json')); s=j['single_stream']
print('$W', round(j['value'],1), s['c_abi']['process_host_msamples_per_s'], s['c_abi'].get('best_submit_collect_msamples_per_s'), s['c_abi']['last_call_stage_ms'], s['first_track']['wall_ms'], s['parity']['bit_identical_to_oracle'], j['cpu_baseline']['value'])
"; tail -2 $O/r2_bench_ab_$W.err; done
