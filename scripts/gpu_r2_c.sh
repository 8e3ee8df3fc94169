#!/usr/bin/env bash
# round 2, GPU call C (2 GPUs): exchange v2 (in-kernel barriers, compacted expansion) -- tests incl. the two-process one,
# emulated 8-rank kernel costs, 2-GPU bench with the parity gate
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_peer_exchange_gpu.py -q -m gpu -x > gpurun_out/pytest_peer.log 2>&1
echo "pytest peer exit $?"; tail -8 gpurun_out/pytest_peer.log
timeout 300 python scripts/peer_emulate.py 1000000 3 8 sparse > gpurun_out/peer_emulate8s.json 2> gpurun_out/peer_emulate8s.err; echo "emulate8 sparse exit $?"; cat gpurun_out/peer_emulate8s.json; tail -3 gpurun_out/peer_emulate8s.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 exit $?"; cat gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench n1 exit $?"; cut -c1-400 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
