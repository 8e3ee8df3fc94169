#!/usr/bin/env bash
# round 2, GPU call B: the exchange kernels with 8 ranks emulated on one GPU (kernel-only cost), plain and under ncu
set -u
mkdir -p gpurun_out
timeout 300 python scripts/peer_emulate.py 1000000 3 8 > gpurun_out/peer_emulate8.json 2> gpurun_out/peer_emulate8.err; echo "emulate8 exit $?"; cat gpurun_out/peer_emulate8.json; tail -3 gpurun_out/peer_emulate8.err
GSL_EXPAND_COMPACT=1 timeout 300 python scripts/peer_emulate.py 1000000 3 8 > gpurun_out/peer_emulate8c.json 2> gpurun_out/peer_emulate8c.err; echo "emulate8 compact exit $?"; cat gpurun_out/peer_emulate8c.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_peer_sh_expand|k_peer_reduce_rows|k_peer_unpack" -s 24 -c 6 -f -o gpurun_out/prof_peer8 \
   python scripts/peer_emulate.py 1000000 2 8 > gpurun_out/ncu_peer8.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_peer8.log
GSL_EXPAND_COMPACT=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_peer_sh_expand" -s 8 -c 2 -f -o gpurun_out/prof_peer8c \
   python scripts/peer_emulate.py 1000000 2 8 > gpurun_out/ncu_peer8c.log 2>&1
echo "ncu compact exit $?"; tail -3 gpurun_out/ncu_peer8c.log
