import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
from gs_lidar_b200 import synth
scene = synth.make_scene(6000, seed=21, footprint_px=6.0).to("cuda")
m = scene.means3D.clone(); m[:50] *= 0.05; m[50:100] *= 40.0
scene = scene._replace(means3D=m)
cp = torch.rand(6000, 4, generator=torch.Generator().manual_seed(5)).cuda()
cot = {k: v.cuda() for k, v in synth.make_cotangents(scene.H, scene.W, 4, seed=98).items()}
g1 = common.run_ours(scene, cot, colors_precomp=cp, export=False)[2]
g2 = common.run_ours(scene, cot, colors_precomp=cp, export=False)[2]
_, _, ra, ref = common.run_ref(scene, cot, colors_precomp=cp)
ra = {k: v.clone() for k, v in ra.items()}
rb = {k: v.clone() for k, v in common.run_ref(scene, cot, colors_precomp=cp, ref=ref)[2].items()}
for k, rk in dict(colors_precomp="dL_dcolors", opacities="dL_dopacity", means3D="dL_dmeans3D", features="dL_dfeatures").items():
    a = ra[rk].reshape(g1[k].shape); b = rb[rk].reshape(g1[k].shape)
    print(k, "ours/ours %.2e  ref/ref %.2e  ours/ref %.2e" % (common.grad_err(g1[k], g2[k])[0], common.grad_err(a, b)[0], common.grad_err(g1[k], a)[0]))
    d = (g1[k] - a).abs().flatten(); i = int(d.argmax()); row = i // g1[k].shape[1]
    print("   worst row", row, "ours", g1[k][row].tolist(), "ref", a[row].tolist(), "max|ref|", float(a.abs().max()))
