#!/bin/bash
# time-segment sharding across GPUs: run with gpurun --gpus N
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
N=${1:-2}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/bench_segments.py 8 sc16 fc32 > $O/r2_bench_segments_n$N.json 2> $O/r2_bench_segments_n$N.err; echo "rc=$?"; tail -5 $O/r2_bench_segments_n$N.err; cat $O/r2_bench_segments_n$N.json
