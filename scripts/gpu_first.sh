#!/usr/bin/env bash
# first on-GPU shake-out: parity vs the reference CUDA rasterizer, then timing
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python scripts/gpu_check.py --sizes 2000 --shapes 66x1030x180,66x515x90 --json gpurun_out/check_small.json > gpurun_out/check_small.log 2>&1
echo "small exit $?" | tee -a gpurun_out/status.txt
tail -5 gpurun_out/check_small.log
timeout 1200 python scripts/gpu_check.py --sizes 100000,1000000 --shapes 66x1030x180 --time --json gpurun_out/check_big.json > gpurun_out/check_big.log 2>&1
echo "big exit $?" | tee -a gpurun_out/status.txt
tail -8 gpurun_out/check_big.log
