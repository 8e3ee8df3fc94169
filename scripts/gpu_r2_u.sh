#!/usr/bin/env bash
# round 2, call u: GPU suite, then ncu evidence of the final kernels (launch list + full capture of one step)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
bash scripts/gpu_r2_m.sh
