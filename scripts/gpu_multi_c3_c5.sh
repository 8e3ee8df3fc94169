#!/usr/bin/env bash
# N GPUs (arg 1): the headline step with the default exchange schedule (+ early-high at 8) and c5 frame-sharded
set -u
mkdir -p gpurun_out
N=${1:-2}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
if [ "$N" = "2" ]; then
timeout 600 python -m pytest tests/test_peer_exchange_gpu.py -q -m gpu > gpurun_out/pytest_peer.log 2>&1
echo "pytest peer exit $?"; tail -2 gpurun_out/pytest_peer.log
fi
run --steps 50 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "c3 n$N exit $?"; tail -2 gpurun_out/bench_n$N.err | cut -c1-200
run --config c5 --frames 512 > gpurun_out/bench_c5_n$N.json 2> gpurun_out/bench_c5_n$N.err; echo "c5 n$N exit $?"; tail -2 gpurun_out/bench_c5_n$N.err | cut -c1-200
if [ "$N" = "8" ]; then
run --steps 50 --warmup 5 --exchange-schedule early-high > gpurun_out/bench_n${N}_early.json 2> gpurun_out/bench_n${N}_early.err; echo "c3 early n$N exit $?"
fi
python - <<PY
import json
for f in ('bench_n$N', 'bench_n${N}_early', 'bench_c5_n$N'):
    try:
        d=json.load(open('gpurun_out/%s.json' % f))
        print(f, round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d.get('exchange_parity') and d['exchange_parity']['ok'], d.get('per_rank_ms_without_exchange'))
        print('   ', {k: round(v['ms_per_launch'],4) for k,v in d.get('kernels',{}).items() if 'peer' in k or 'bwd' in k})
    except Exception as ex:
        print(f, 'no line', ex)
PY
