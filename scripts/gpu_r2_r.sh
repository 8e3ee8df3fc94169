#!/usr/bin/env bash
# round 2, call r: the dynamic training step (--config c4) and the headline on one GPU
set -u
mkdir -p gpurun_out
timeout 300 python bench.py --config c4 --steps 30 --warmup 5 > gpurun_out/bench_c4_n1.json 2> gpurun_out/bench_c4_n1.err
echo "c4 exit $?"; tail -3 gpurun_out/bench_c4_n1.err; cut -c1-1500 gpurun_out/bench_c4_n1.json
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "c3 exit $?"; tail -3 gpurun_out/bench_n1.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print(round(d['value'],1), d['ms_per_step'], 'e2e', d['e2e']['value'], d['cuda_graph']['ms_per_step_graph_off'])
print({k: round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})
PY
