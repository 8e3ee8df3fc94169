"""Device time of the forward and of forward+backward, un-graphed vs CUDA-graph replay (CUDA events)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gs_lidar_b200 import synth
import gs_lidar_b200.diff_gaussian_rasterization_2d as G

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
scene = synth.make_scene(P).to("cuda")
cot = {k: v.cuda() for k, v in synth.make_cotangents(scene.H, scene.W, 4).items()}
rast = G.GaussianRasterizer(synth.settings_for(scene))
lv = dict(means3D=scene.means3D.clone(), means2D=torch.zeros((P, 4), device="cuda"), opacities=scene.opacities.clone(),
          shs=scene.shs.clone(), features=scene.features.clone(), scales=scene.scales.clone(), rotations=scene.rotations.clone())
for v in lv.values():
    v.requires_grad_(True)


def fwd():
    with torch.no_grad():
        rast(mask=scene.mask, **lv)


def fb():
    for v in lv.values():
        v.grad = None
    o = rast(mask=scene.mask, **lv)
    torch.autograd.backward([o[1], o[2], o[3], o[4]], [cot["color"], cot["feature"], cot["depth"], cot["alpha"]])


def timed(fn, n=40):
    for _ in range(6):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def split_times(n=40):
    """device time of the forward part and of the backward part of a step (events between the two)"""
    for _ in range(6):
        fb()
    torch.cuda.synchronize()
    evs = []
    for _ in range(n):
        for v in lv.values():
            v.grad = None
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record()
        o = rast(mask=scene.mask, **lv)
        b.record()
        torch.autograd.backward([o[1], o[2], o[3], o[4]], [cot["color"], cot["feature"], cot["depth"], cot["alpha"]])
        c.record()
        evs.append((a, b, c))
    torch.cuda.synchronize()
    f = sorted(a.elapsed_time(b) for a, b, c in evs)[n // 2]
    bw = sorted(b.elapsed_time(c) for a, b, c in evs)[n // 2]
    return f, bw


res = {}
for mode in (False, True):
    G.set_cuda_graphs(mode)
    res["fwd_graph_%s" % mode] = timed(fwd)
    res["fwdbwd_graph_%s" % mode] = timed(fb)
    res["split_fwd_bwd_graph_%s" % mode] = split_times()
G.set_cuda_graphs(False)
print(json.dumps(res))
