"""Small fwd+bwd cases for compute-sanitizer (memcheck / racecheck): fast binning, general binning, S=0, glue."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
import glue_oracle as GO
from gs_lidar_b200 import synth, renderer
for kw in (dict(P=3001, seed=1), dict(P=2000, seed=2, H=50, W=70, vfov=(-60.0, 60.0), hfov=(-100.0, 100.0), S=0, sh_degree=0, footprint_px=3.0),
           dict(P=2500, seed=3, H=272, W=1040, vfov=(-40.0, 20.0), footprint_px=2.0), dict(P=1500, seed=4, footprint_px=12.0, S=10, sh_degree=2)):
    P = kw.pop("P")
    scene = synth.make_scene(P, **kw).to("cuda")
    S = scene.features.shape[1]
    cot = {k: v.cuda() for k, v in synth.make_cotangents(scene.H, scene.W, S, seed=9).items()}
    out, state, grads = common.run_ours(scene, cot)
    torch.cuda.synchronize()
    print("case ok", P, state["R"])
pc = GO.make_model(2001, seed=5, device="cuda")
o = renderer.activate_surfels(pc, 0.1, 0.02, True, None)
torch.autograd.backward(list(o[:4]), [torch.ones_like(x) for x in o[:4]])
torch.cuda.synchronize()
print("glue ok")
