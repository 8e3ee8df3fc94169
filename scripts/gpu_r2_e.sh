#!/usr/bin/env bash
# round 2, GPU call E: graph-mode tests + full GPU suite + bench (graph on / off)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_graph_gpu.py -q -m gpu -x > gpurun_out/pytest_graph.log 2>&1
echo "pytest graph exit $?"; tail -25 gpurun_out/pytest_graph.log
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench n1 exit $?"; cut -c1-300 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
