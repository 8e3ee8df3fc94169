"""Work counters of the compositors (needs the -DGSL_STATS build: GSL_B200_LIB=.../libgsl_b200_stats.so)."""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
from gs_lidar_b200 import synth  # noqa: E402
import gs_lidar_b200.diff_gaussian_rasterization_2d as G  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
scene = synth.make_scene(P).to("cuda")
cot = {k: v.cuda() for k, v in synth.make_cotangents(scene.H, scene.W, 4).items()}
buf = (C.c_ulonglong * 16)()
G._lib.gsl_stats_read(buf, 1)
G._lib.gsl_stats_read_bwd(buf, 1)
out, state, grads = common.run_ours(scene, cot, export=False)
G._lib.gsl_stats_read(buf, 1)
f = list(buf)[:5]
G._lib.gsl_stats_read_bwd(buf, 1)
b = list(buf)[:5]
print(json.dumps(dict(P=P, fwd=dict(staged_entries=f[0], warp_iterations=f[1], contributing_entries=f[2],
                                    valid_pairs=f[3], evaluated_pairs=f[4]),
                      bwd=dict(staged_entries=b[0], warp_iterations=b[1], evaluated_pairs=b[3], uniform_iterations=b[4]))))
