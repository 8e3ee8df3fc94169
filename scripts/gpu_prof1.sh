#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for impl in ours ref; do
  timeout 300 python scripts/one_step.py --impl $impl --iters 3 > gpurun_out/plain_$impl.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$impl.csv \
     python scripts/one_step.py --impl $impl --iters 3 > gpurun_out/ncu_$impl.log 2>&1
  echo "$impl exit $?"
done
