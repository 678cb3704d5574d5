#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > $O/r2_bench_n2.json 2> $O/r2_bench_n2.err
echo "rc=$?"; cut -c1-300 $O/r2_bench_n2.json; tail -3 $O/r2_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > $O/r2_bench_n2_ref.json 2> $O/r2_bench_n2_ref.err
echo "rc=$?"; cut -c1-300 $O/r2_bench_n2_ref.json; tail -3 $O/r2_bench_n2_ref.err
timeout 300 python -m pytest tests/test_sigmf.py -m gpu -q 2>&1 | tail -5
