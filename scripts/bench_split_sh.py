"""Split-SH measurement: one fwd+bwd of the rasterizer at the headline workload with the SH coefficients passed (a) as
the reference does, shs = torch.cat((_features_dc, _features_rest), 1) built per call (scene/gaussian_model.py:167-171),
and (b) as the two parameter tensors (shs=dc, shs_rest=rest).  CUDA events, gradients w.r.t. dc/rest in both arms."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gs_lidar_b200 import GaussianRasterizer, synth

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
scene = synth.make_scene(P, seed=0).to("cuda")
cot = {k: v.cuda() for k, v in synth.make_cotangents(scene.H, scene.W, 4, seed=1).items()}
rast = GaussianRasterizer(synth.settings_for(scene))
names = ("means3D", "opacities", "scales", "rotations", "features")
leaves = {k: getattr(scene, k).detach().clone().requires_grad_(True) for k in names}
leaves["dc"] = scene.shs[:, :1].contiguous().requires_grad_(True)
leaves["rest"] = scene.shs[:, 1:].contiguous().requires_grad_(True)
m2 = torch.zeros((P, 4), device="cuda", requires_grad=True)


def step(split):
    sh = dict(shs=leaves["dc"], shs_rest=leaves["rest"]) if split else \
        dict(shs=torch.cat((leaves["dc"], leaves["rest"]), dim=1))
    contrib, color, feature, depth, alpha, radii = rast(
        means3D=leaves["means3D"], means2D=m2, opacities=leaves["opacities"], features=leaves["features"],
        scales=leaves["scales"], rotations=leaves["rotations"], mask=scene.mask, **sh)
    torch.autograd.backward([color, feature, depth, alpha], [cot["color"], cot["feature"], cot["depth"], cot["alpha"]])
    for l in list(leaves.values()) + [m2]:
        l.grad = None


res = {}
for name, split in (("concatenated", False), ("split", True), ("concatenated_again", False), ("split_again", True)):
    for _ in range(5):
        step(split)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        step(split)
    e1.record(); torch.cuda.synchronize()
    res[name] = e0.elapsed_time(e1) / 30
print(json.dumps(dict(P=P, ms_per_step=res, saved_ms=res["concatenated"] - res["split"],
                      note="fwd+bwd incl. autograd; the concatenated arm pays cat (read+write 256 MB) and its backward "
                           "(two slice copies of the 256 MB gradient)")))
