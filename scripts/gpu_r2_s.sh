#!/usr/bin/env bash
# round 2, call s: binning changes (prefix bytes in the scatter, warp scans in the block lists, k_bin_bases folded into
# k_bin_scan) through the whole GPU suite; the two-entries-per-turn forward compositor through the parity tests; benches
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
GSL_B200_LIB=$PWD/gs_lidar_b200/libgsl_b200_fwdilp2.so timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_wrap_gpu.py -x -q -m gpu > gpurun_out/pytest_ilp2.log 2>&1
echo "pytest ilp2 exit $?"; tail -3 gpurun_out/pytest_ilp2.log
VARIANTS="base fwdilp2" bash scripts/gpu_variants.sh
timeout 300 python bench.py --config c4 --steps 30 --warmup 5 > gpurun_out/bench_c4_n1.json 2> gpurun_out/bench_c4_n1.err
echo "c4 exit $?"; tail -3 gpurun_out/bench_c4_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c4_n1.json')); print('c4', d['value'], d['ms_per_step'])"
