#!/usr/bin/env bash
# bench every built library variant (gs_lidar_b200/libgsl_b200_<v>.so) listed in $VARIANTS
set -u
mkdir -p gpurun_out
for v in ${VARIANTS}; do
  lib=$PWD/gs_lidar_b200/libgsl_b200.so
  [ "$v" != "base" ] && lib=$PWD/gs_lidar_b200/libgsl_b200_$v.so
  GSL_B200_LIB=$lib timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  echo "$v exit $?"
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_$v.json'))
print('$v', {k:round(x['ms_per_launch'],4) for k,x in d['kernels'].items()}, round(d['ms_per_step'],4), round(d['value'],1))
PY
done
