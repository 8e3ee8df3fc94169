#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/one_step.py --impl ours --iters 2 > gpurun_out/plain_ours2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render -s 2 -c 2 -f -o gpurun_out/prof_render \
   python scripts/one_step.py --impl ours --iters 2 > gpurun_out/ncu_full.log 2>&1
echo "exit $?"
tail -3 gpurun_out/ncu_full.log
