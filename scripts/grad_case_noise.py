"""Repeats one parity case several times and prints the element-wise / norm-wise gradient errors of the product against the
reference CUDA kernels, and of each implementation against its own second run (run-to-run noise of the atomics)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import common
from gs_lidar_b200 import synth
import test_parity_gpu as T

case = int(sys.argv[1]) if len(sys.argv) > 1 else 3
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
kw = dict(T.CASES[case]); P = kw.pop("P")
scene = synth.make_scene(P, **kw).to("cuda")
S = scene.features.shape[1]
cot = {k: v.cuda() for k, v in synth.make_cotangents(scene.H, scene.W, S, seed=99).items()}
rows = []
for it in range(reps):
    g1 = common.run_ours(scene, cot, export=False)[2]
    g2 = common.run_ours(scene, cot, export=False)[2]
    r1 = {k: v.clone() for k, v in common.run_ref(scene, cot)[2].items()}
    r2 = {k: v.clone() for k, v in common.run_ref(scene, cot)[2].items()}
    row = {}
    for k, rk in T.GRAD_KEYS.items():
        if g1.get(k) is None or rk not in r1:
            continue
        a, b = r1[rk], r2[rk]
        if rk == "dL_dfeatures":
            a, b = a[:, :S], b[:, :S]
        a = a.reshape(g1[k].shape); b = b.reshape(g1[k].shape)
        row[k] = dict(ours_vs_ref=common.grad_err(g1[k], a), ref_vs_ref=common.grad_err(b, a)[0], ours_vs_ours=common.grad_err(g2[k], g1[k])[0])
    rows.append(row)
for k in rows[0]:
    print(k, "elem ours-ref", ["%.2e" % r[k]["ours_vs_ref"][0] for r in rows], "norm", "%.1e" % max(r[k]["ours_vs_ref"][1] for r in rows),
          "| ref-ref", ["%.2e" % r[k]["ref_vs_ref"] for r in rows], "| ours-ours", ["%.2e" % r[k]["ours_vs_ours"] for r in rows])
