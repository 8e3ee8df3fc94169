#!/usr/bin/env bash
# N GPUs (arg 1): schedules of the fused exchange step side by side (bench at N, one line per schedule)
set -u
mkdir -p gpurun_out
N=${1:-2}
if [ "$N" = "2" ]; then
timeout 600 python -m pytest tests/test_peer_exchange_gpu.py tests/test_graph_gpu.py -q -m gpu -x > gpurun_out/pytest_peer.log 2>&1
echo "pytest peer exit $?"; tail -4 gpurun_out/pytest_peer.log
fi
for sch in ${SCHEDULES:-late early-high early-low}; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 --exchange-schedule $sch > gpurun_out/bench_n${N}_$sch.json 2> gpurun_out/bench_n${N}_$sch.err
echo "bench n$N $sch exit $?"; tail -2 gpurun_out/bench_n${N}_$sch.err | cut -c1-300
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_n${N}_$sch.json'))
    print('$sch', d['ms_per_step'], d['cuda_graph'], d['exchange_parity']['ok'], d['per_rank_ms_without_exchange'])
    print({k: round(v['ms_per_launch'],4) for k,v in d['kernels'].items() if 'peer' in k or 'preprocess_bwd' in k})
except Exception as ex:
    print('no line', ex)
PY
done
