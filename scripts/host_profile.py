"""Host-side cost of a step through the public API in CUDA-graph mode with the deferred instance count: issue N steps without
any synchronisation and divide (the GPU is far behind, so this is pure host time), then the top of a cProfile of the same loop."""
import cProfile, os, pstats, sys, time
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

def step():
    for v in leaves.values():
        v.grad = None
    out = rast(mask=scene.mask, **leaves)
    torch.autograd.backward([out[1], out[2], out[3], out[4]], [cot["color"], cot["feature"], cot["depth"], cot["alpha"]])

for mode in ("eager", "graph", "graph+deferred"):
    G.set_cuda_graphs(mode != "eager", deferred_count=mode == "graph+deferred")
    for _ in range(6):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(40):
        step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("%-15s host %.3f ms/step issued, %.3f ms/step wall" % (mode, (t1 - t0) / 40 * 1e3, (t2 - t0) / 40 * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(40):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
