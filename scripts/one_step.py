"""Runs a few fwd(+bwd) steps of one implementation on one synthetic scene (for ncu / timing)."""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
from gs_lidar_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--impl", default="ours")
ap.add_argument("--P", type=int, default=1000000)
ap.add_argument("--H", type=int, default=66)
ap.add_argument("--W", type=int, default=1030)
ap.add_argument("--hfov", type=float, default=180.0)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--fwd-only", action="store_true")
args = ap.parse_args()

scene = synth.make_scene(args.P, H=args.H, W=args.W, hfov=(-args.hfov, args.hfov)).to("cuda")
cot = {k: v.cuda() for k, v in synth.make_cotangents(args.H, args.W, 4).items()}
if args.impl == "ours":
    import gs_lidar_b200.diff_gaussian_rasterization_2d as G
    rast = G.GaussianRasterizer(synth.settings_for(scene))
    P = args.P
    leaves = [scene.means3D.clone().requires_grad_(True), torch.zeros((P, 4), device="cuda", requires_grad=True),
              scene.opacities.clone().requires_grad_(True), scene.shs.clone().requires_grad_(True),
              scene.features.clone().requires_grad_(True), scene.scales.clone().requires_grad_(True),
              scene.rotations.clone().requires_grad_(True)]

    def step():
        for l in leaves:
            l.grad = None
        contrib, color, feature, depth, alpha, radii = rast(means3D=leaves[0], means2D=leaves[1], opacities=leaves[2],
                                                            shs=leaves[3], features=leaves[4], scales=leaves[5],
                                                            rotations=leaves[6], mask=scene.mask)
        if not args.fwd_only:
            torch.autograd.backward([color, feature, depth, alpha], [cot["color"], cot["feature"], cot["depth"], cot["alpha"]])
else:
    import oracle
    ref = oracle.RefCuda()
    a = common.ref_args(scene)
    bufs = {}

    def step():
        f = ref.forward(a, zero_fill=True, outs=bufs.get("o"))
        bufs["o"] = {k: v for k, v in f.items() if k != "R"}
        if not args.fwd_only:
            bufs["g"] = ref.backward(a, f, cot, zero_fill=True, grads=bufs.get("g"))

for i in range(args.iters):
    torch.cuda.synchronize()
    t = time.time()
    step()
    torch.cuda.synchronize()
    print("iter", i, "ms", (time.time() - t) * 1e3, flush=True)
