#!/usr/bin/env bash
# round 2, GPU call H (N GPUs): two-process peer-exchange test (N>=2) + bench at N with the parity gate
set -u
mkdir -p gpurun_out
N=${1:-2}
if [ "$N" = "2" ]; then
timeout 600 python -m pytest tests/test_peer_exchange_gpu.py -q -m gpu -x > gpurun_out/pytest_peer.log 2>&1
echo "pytest peer exit $?"; tail -4 gpurun_out/pytest_peer.log
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench n$N exit $?"; cut -c1-300 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
