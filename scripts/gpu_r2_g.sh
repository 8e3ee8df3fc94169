#!/usr/bin/env bash
# round 2, GPU call G (re-entry): state of the tree -- full GPU suite, bench (graph on/off in one line), launch list
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench n1 exit $?"; cut -c1-600 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ours.csv \
   python scripts/one_step.py --impl ours --iters 3 > gpurun_out/ncu_ours.log 2>&1
echo "ncu launches exit $?"
