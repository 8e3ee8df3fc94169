#!/usr/bin/env bash
# round 2, GPU call L: new tests (post-ops, batch frames), full suite, bench c3 + c2 + c5 (bounded)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --tb=short > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench n1 exit $?"; tail -3 gpurun_out/bench_n1.err
timeout 300 python bench.py --config c2 --steps 200 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
echo "bench c2 exit $?"; tail -3 gpurun_out/bench_c2.err
timeout 600 python bench.py --config c5 --frames ${C5_FRAMES:-64} > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
echo "bench c5 exit $?"; tail -3 gpurun_out/bench_c5.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print(d['ms_per_step'], d['cuda_graph']['ms_per_step_graph_off'], 'e2e', d['e2e']['value'])
print({k: round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})
for c in ('c2','c5'):
    try:
        d=json.load(open('gpurun_out/bench_%s.json'%c))
        print(c, d['value'], d['ms_per_step'], d['run'], d['roofline']['frac'], 'e2e', d['e2e']['value'], d['config']['visible_surfels'], d['config']['tile_instances'])
    except Exception as ex:
        print(c, 'no line', ex)
PY
