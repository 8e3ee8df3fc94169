#!/usr/bin/env bash
# round 2, call q: the glue fold of the peer exchange (new tests first, then the whole GPU suite)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_peer_glue_gpu.py -x -q -m gpu > gpurun_out/pytest_glue.log 2>&1
echo "glue pytest exit $?"
tail -30 gpurun_out/pytest_glue.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"
tail -8 gpurun_out/pytest_gpu.log
