#!/usr/bin/env bash
# ncu evidence -- launch list of one bench-shaped run + full capture of every kernel of the step
set -u
mkdir -p gpurun_out
timeout 300 python scripts/one_step.py --impl ours --iters 3 > gpurun_out/plain_ours.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_ours.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_render|k_preprocess|k_bin|k_tile|k_sort|k_depth" -s 28 -c 14 -f -o gpurun_out/prof_step \
   python scripts/one_step.py --impl ours --iters 3 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -2 gpurun_out/ncu_full.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ours.csv \
   python scripts/one_step.py --impl ours --iters 3 > gpurun_out/ncu_ours.log 2>&1
echo "ncu launches exit $?"
