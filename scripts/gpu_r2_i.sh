#!/usr/bin/env bash
# round 2, GPU call I: does the nvidia-smi clock sampler disturb the measurement?  N=1 bench at three sampling periods
set -u
mkdir -p gpurun_out
for ms in nvml smi off; do
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --clock-sampler $ms > gpurun_out/bench_n1_s$ms.json 2> gpurun_out/bench_n1_s$ms.err
echo "sampler $ms exit $?"; python - <<PY
import json
d=json.load(open('gpurun_out/bench_n1_s$ms.json'))
print(d['ms_per_step'], d['cuda_graph']['ms_per_step_graph_off'], d['step_spread'], d['e2e']['value'], d['clocks'])
PY
done
