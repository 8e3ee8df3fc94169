#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/graph_probe.py > gpurun_out/graph_probe.json 2> gpurun_out/graph_probe.err; echo "probe exit $?"; cat gpurun_out/graph_probe.json; tail -3 gpurun_out/graph_probe.err
GSL_B200_LIB=$PWD/gs_lidar_b200/libgsl_b200_noprio.so timeout 300 python scripts/graph_probe.py > gpurun_out/graph_probe_noprio.json 2> gpurun_out/graph_probe_noprio.err; echo "probe noprio exit $?"; cat gpurun_out/graph_probe_noprio.json; tail -3 gpurun_out/graph_probe_noprio.err
