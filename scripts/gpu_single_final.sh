#!/usr/bin/env bash
# one GPU: GPU suite, the headline line with both CPU baselines, c4, c5 (64 frames), c2
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "c3 exit $?"; tail -3 gpurun_out/bench_n1.err
timeout 300 python bench.py --config c4 --steps 30 --warmup 5 > gpurun_out/bench_c4_n1.json 2> gpurun_out/bench_c4_n1.err; echo "c4 exit $?"
timeout 300 python bench.py --config c5 --frames 64 > gpurun_out/bench_c5_n1.json 2> gpurun_out/bench_c5_n1.err; echo "c5 exit $?"
timeout 300 python bench.py --config c2 --steps 50 > gpurun_out/bench_c2_n1.json 2> gpurun_out/bench_c2_n1.err; echo "c2 exit $?"
python - <<PY
import json
for f in ('bench_n1','bench_c4_n1','bench_c5_n1','bench_c2_n1'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f, round(d['value'],1), round(d['ms_per_step'],4), 'e2e', d.get('e2e',{}).get('value'))
        print('   ', {k: round(v['ms_per_launch'],4) for k,v in d.get('kernels',{}).items()})
    except Exception as ex:
        print(f, 'no line', ex)
PY
