#!/usr/bin/env bash
# ncu captures: launch list of one bench run + full capture of the two compositors (+ optional others)
set -u
mkdir -p gpurun_out
timeout 300 python scripts/one_step.py --impl ours --iters 3 > gpurun_out/plain_ours.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_ours.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${NCU_K:-k_render}" -s "${NCU_SKIP:-4}" -c "${NCU_COUNT:-2}" -f -o gpurun_out/prof_render \
   python scripts/one_step.py --impl ours --iters 3 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ours.csv \
   python scripts/one_step.py --impl ours --iters 3 > gpurun_out/ncu_ours.log 2>&1
echo "ncu launches exit $?"
