"""Host-side enqueue cost of one forward / backward through the public API (GPU work is asynchronous)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gs_lidar_b200 import synth
import gs_lidar_b200.diff_gaussian_rasterization_2d as G

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
scene = synth.make_scene(P).to("cuda")
cot = {k: v.cuda() for k, v in synth.make_cotangents(scene.H, scene.W, 4).items()}
rast = G.GaussianRasterizer(synth.settings_for(scene))
leaves = dict(means3D=scene.means3D.clone(), means2D=torch.zeros((P, 4), device="cuda"), opacities=scene.opacities.clone(),
              shs=scene.shs.clone(), features=scene.features.clone(), scales=scene.scales.clone(), rotations=scene.rotations.clone())
for v in leaves.values():
    v.requires_grad_(True)
tf, tb, tt = [], [], []
for it in range(30):
    for v in leaves.values():
        v.grad = None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = rast(mask=scene.mask, **leaves)
    t1 = time.perf_counter()
    torch.autograd.backward([out[1], out[2], out[3], out[4]], [cot["color"], cot["feature"], cot["depth"], cot["alpha"]])
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    if it >= 5:
        tf.append(t1 - t0); tb.append(t2 - t1); tt.append(t3 - t0)
med = lambda x: sorted(x)[len(x) // 2] * 1e3
print("host ms: forward call %.3f (includes the wait for the instance count), backward call %.3f, step wall %.3f" % (med(tf), med(tb), med(tt)))
