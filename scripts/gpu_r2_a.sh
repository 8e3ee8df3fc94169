#!/usr/bin/env bash
# round 2, GPU call A: full GPU test suite (Chamfer un-gated), peer-exchange variants under emulation, Chamfer timing
# against the reference kernels, a quick single-GPU bench line, host enqueue cost.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/pytest_gpu.log
GSL_EXPAND_COMPACT=1 timeout 300 python -m pytest tests/test_peer_exchange_gpu.py -q -m gpu > gpurun_out/pytest_compact.log 2>&1
echo "compact exit $?"; tail -4 gpurun_out/pytest_compact.log
GSL_PEER_EARLY_FACTORS=1 timeout 300 python -m pytest tests/test_peer_exchange_gpu.py -q -m gpu > gpurun_out/pytest_early.log 2>&1
echo "early exit $?"; tail -4 gpurun_out/pytest_early.log
timeout 300 python scripts/bench_chamfer.py > gpurun_out/bench_chamfer.json 2> gpurun_out/bench_chamfer.err
echo "chamfer bench exit $?"; cat gpurun_out/bench_chamfer.json; tail -3 gpurun_out/bench_chamfer.err
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
echo "bench exit $?"; cat gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
timeout 120 python scripts/host_overhead.py > gpurun_out/host_overhead.txt 2>&1; cat gpurun_out/host_overhead.txt
