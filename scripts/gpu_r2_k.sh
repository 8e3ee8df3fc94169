#!/usr/bin/env bash
# round 2, GPU call K: full GPU suite + bench + (optional) ncu full capture of the kernels named in NCU_K
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench n1 exit $?"; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print(d['ms_per_step'], d['cuda_graph']['ms_per_step_graph_off'], d['step_spread'], 'e2e', d['e2e']['value'])
print({k: round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})
PY
if [ -n "${NCU_K:-}" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$NCU_K" -s "${NCU_SKIP:-2}" -c "${NCU_COUNT:-2}" -f -o gpurun_out/prof_${NCU_NAME:-k} \
   python scripts/one_step.py --impl ours --iters 3 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -2 gpurun_out/ncu_full.log
fi
