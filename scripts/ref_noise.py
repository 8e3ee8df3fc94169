"""How reproducible are the reference's own gradients (atomicAdd order)?  grad_err(ref run A, ref run B) and
grad_err(ours, ref) for the parity cases."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
from gs_lidar_b200 import synth
import test_parity_gpu as T

for i, kw in enumerate(T.CASES):
    kw = dict(kw); P = kw.pop("P")
    scene = synth.make_scene(P, **kw).to("cuda")
    S = scene.features.shape[1]
    cot = {k: v.cuda() for k, v in synth.make_cotangents(scene.H, scene.W, S, seed=99).items()}
    _, _, g = common.run_ours(scene, cot, export=False)
    _, _, ra, _ = common.run_ref(scene, cot)
    ra = {k: v.clone() for k, v in ra.items()}
    _, _, rb, _ = common.run_ref(scene, cot)
    row = {}
    for k, rk in T.GRAD_KEYS.items():
        if g.get(k) is None or rk not in ra or g[k].numel() == 0:
            continue
        a = ra[rk].reshape(-1)[:g[k].numel()] if rk != "dL_dfeatures" else ra[rk][:, :S].reshape(-1)
        b = rb[rk].reshape(-1)[:g[k].numel()] if rk != "dL_dfeatures" else rb[rk][:, :S].reshape(-1)
        row[k] = dict(ref_vs_ref="%.2e" % common.grad_err(a, b)[0], ours_vs_ref="%.2e" % common.grad_err(g[k].reshape(-1), a)[0])
    print(i, json.dumps(row))
