#!/usr/bin/env bash
# round 2, call p: GPU tests after the grouped binning (no library sort left), then the library variants in $VARIANTS
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"
tail -8 gpurun_out/pytest_gpu.log
bash scripts/gpu_variants.sh
