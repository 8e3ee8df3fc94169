#!/usr/bin/env bash
# GPU iteration: parity tests, compositor work counters, bench (ours + reference).
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"
tail -15 gpurun_out/pytest_gpu.log
if [ -f gs_lidar_b200/libgsl_b200_stats.so ]; then
  GSL_B200_LIB=$PWD/gs_lidar_b200/libgsl_b200_stats.so timeout 300 python scripts/stats_step.py > gpurun_out/stats.json 2> gpurun_out/stats.err
  echo "stats exit $?"; cat gpurun_out/stats.json; tail -3 gpurun_out/stats.err
fi
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
echo "bench exit $?"; cat gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
if [ "${SKIP_REF:-0}" != "1" ]; then
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "bench ref exit $?"; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
fi
