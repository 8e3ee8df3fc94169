#!/usr/bin/env bash
# round 2, GPU call D (8 GPUs): bench at N=8 with the parity gate and per-rank diagnostics
set -u
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench n$N exit $?"; cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
