#!/usr/bin/env bash
# ncu capture of the peer-exchange kernels (two ranks emulated on one GPU)
set -u
mkdir -p gpurun_out
timeout 300 python scripts/peer_emulate.py 1000000 3 > gpurun_out/peer_emulate.json 2> gpurun_out/peer_emulate.err || { echo "plain run failed"; tail -5 gpurun_out/peer_emulate.err; exit 1; }
cat gpurun_out/peer_emulate.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_peer_sh_expand|k_peer_reduce_rows|k_peer_unpack|k_preprocess_bwd" -s 12 -c 6 -f -o gpurun_out/prof_peer \
   python scripts/peer_emulate.py 1000000 3 > gpurun_out/ncu_peer.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_peer.log
