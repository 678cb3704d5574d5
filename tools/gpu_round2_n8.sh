#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
N=${1:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/e2e_scaling_probe.py --gib 1 --reps 6 > $O/r2_e2e_probe_n$N.json 2> $O/r2_e2e_probe_n$N.err
echo "probe rc=$?"; cut -c1-600 $O/r2_e2e_probe_n$N.json; tail -3 $O/r2_e2e_probe_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N > $O/r2_bench_n$N.json 2> $O/r2_bench_n$N.err
echo "bench rc=$?"; cut -c1-200 $O/r2_bench_n$N.json; tail -3 $O/r2_bench_n$N.err
nvidia-smi topo -m > $O/r2_topo_n$N.txt 2>&1; lscpu | head -30 > $O/r2_lscpu.txt; numactl -H >> $O/r2_lscpu.txt 2>&1
