"""GPU parity + timing probe: product vs the reference CUDA rasterizer (and the CPU oracle on small
scenes).  Diagnostic tool for development; the formal checks live in tests/.

    python scripts/gpu_check.py [--sizes 2000,100000,1000000] [--time] [--json out.json]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
from gs_lidar_b200 import synth  # noqa: E402


def compare(P, H, W, hfov, vfov, seed, do_oracle, report):
    scene = synth.make_scene(P, H=H, W=W, hfov=hfov, vfov=vfov, seed=seed).to("cuda")
    S = scene.features.shape[1]
    cot = {k: v.cuda() for k, v in synth.make_cotangents(H, W, S, seed + 1).items()}
    o_out, o_state, o_grads = common.run_ours(scene, cot)
    r_out, r_state, r_grads, ref = common.run_ref(scene, cot)
    res = dict(P=P, H=H, W=W, hfov=hfov, R_ours=o_state["R"], R_ref=int(r_state["point_list"].numel()))
    vis = r_out["radii"] > 0
    res["visible"] = int(vis.sum())
    res["radii_mismatch"] = common.frac_mismatch(o_out["radii"], r_out["radii"])
    res["tiles_mismatch"] = common.frac_mismatch(o_state["tiles_touched"], r_state["tiles_touched"])
    both = vis & (o_out["radii"] > 0)
    for k in ("depths", "means2D", "transMat", "normal_opacity", "rgb"):
        a, b = o_state[k][both], r_state[k][both]
        res[k + "_bitmismatch"] = float((a.contiguous().view(torch.int32) != b.contiguous().view(torch.int32)).double().mean())
        res[k + "_rel"] = common.rel_err(a, b)
    res["clamped_mismatch"] = common.frac_mismatch(o_state["clamped"][both], r_state["clamped"][both])
    if res["R_ours"] == res["R_ref"]:
        res["keys_mismatch"] = common.frac_mismatch(o_state["point_list_keys"], r_state["point_list_keys"])
        res["list_mismatch"] = common.frac_mismatch(o_state["point_list"], r_state["point_list"])
        res["ranges_mismatch"] = common.frac_mismatch(o_state["ranges"], r_state["ranges"])
    for k in ("out_color", "out_feature", "out_depth", "out_alpha"):
        res[k + "_rel"] = common.rel_err(o_out[k], r_out[k])
        res[k + "_bitmismatch"] = float((o_out[k].contiguous().view(torch.int32) != r_out[k].contiguous().view(torch.int32)).double().mean())
    res["contrib_mismatch"] = common.frac_mismatch(o_out["out_contrib"], r_out["out_contrib"])
    res["final_T_rel"] = common.rel_err(o_state["final_T"], r_state["accum_alpha"])
    gmap = dict(means3D="dL_dmeans3D", means2D="dL_dmeans2D", shs="dL_dsh", features="dL_dfeatures",
                opacities="dL_dopacity", scales="dL_dscales", rotations="dL_drotations")
    for k, rk in gmap.items():
        if o_grads.get(k) is None:
            continue
        res["grad_" + k + "_rel"] = common.rel_err(o_grads[k], r_grads[rk].reshape(o_grads[k].shape))
    # pixbox conservativeness is implied by out_* parity; report box stats
    pb = o_state["pixbox"][both].int()
    wrap = pb[:, 0] > pb[:, 2]
    area = ((pb[:, 2] - pb[:, 0] + 1).clamp_min(0) * (pb[:, 3] - pb[:, 1] + 1).clamp_min(0)).double()
    res["pixbox_wrapped"] = int(wrap.sum())
    res["pixbox_area_mean"] = float(area[~wrap].mean()) if (~wrap).any() else 0.0
    res["ncontrib_mean"] = float(r_out["out_contrib"][0].double().mean())
    tl = (r_state["ranges"][:, 1] - r_state["ranges"][:, 0]).double()
    res["tile_list_mean"], res["tile_list_max"] = float(tl.mean()), float(tl.max())
    if do_oracle:
        st, og = common.run_oracle(scene, cot)
        res["oracle_R"] = st["R"]
        res["oracle_radii_mismatch"] = common.frac_mismatch(torch.from_numpy(st["radii"]), r_out["radii"].cpu())
        for k in ("out_color", "out_feature", "out_depth", "out_alpha"):
            res["oracle_" + k + "_rel"] = common.rel_err(torch.from_numpy(st[k]), r_out[k])
        for k, rk in gmap.items():
            ok = dict(means3D="dL_dmeans3D", means2D="dL_dmeans2D", shs="dL_dsh", features="dL_dfeatures",
                      opacities="dL_dopacity", scales="dL_dscales", rotations="dL_drotations")[k]
            res["oracle_grad_" + k + "_rel"] = common.rel_err(torch.from_numpy(og[ok]).reshape(r_grads[rk].shape), r_grads[rk])
    report.append(res)
    print(json.dumps(res), flush=True)
    return scene, cot, ref


def time_both(scene, cot, ref, iters=20, warm=5):
    import gs_lidar_b200.diff_gaussian_rasterization_2d as G
    settings = synth.settings_for(scene)
    rast = G.GaussianRasterizer(settings)
    P = scene.means3D.shape[0]
    leaves = [scene.means3D.clone().requires_grad_(True), torch.zeros((P, 4), device="cuda", requires_grad=True),
              scene.opacities.clone().requires_grad_(True), scene.shs.clone().requires_grad_(True),
              scene.features.clone().requires_grad_(True), scene.scales.clone().requires_grad_(True),
              scene.rotations.clone().requires_grad_(True)]

    def ours_step(bwd=True):
        for l in leaves:
            l.grad = None
        contrib, color, feature, depth, alpha, radii = rast(means3D=leaves[0], means2D=leaves[1], opacities=leaves[2],
                                                            shs=leaves[3], features=leaves[4], scales=leaves[5],
                                                            rotations=leaves[6], mask=scene.mask)
        if bwd:
            torch.autograd.backward([color, feature, depth, alpha], [cot["color"], cot["feature"], cot["depth"], cot["alpha"]])

    a = common.ref_args(scene)
    fwd_bufs, grad_bufs = {}, {}

    def ref_step(bwd=True):
        f = ref.forward(a, zero_fill=True, outs=fwd_bufs.get("o"))
        fwd_bufs["o"] = {k: v for k, v in f.items() if k != "R"}
        if bwd:
            grad_bufs["g"] = ref.backward(a, f, cot, zero_fill=True, grads=grad_bufs.get("g"))

    out = {}
    for name, fn in (("ours", ours_step), ("ref", ref_step)):
        for bwd in (False, True):
            for _ in range(warm):
                fn(bwd)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.time()
            e0.record()
            for _ in range(iters):
                fn(bwd)
            e1.record()
            torch.cuda.synchronize()
            out["%s_%s_ms" % (name, "fwdbwd" if bwd else "fwd")] = e0.elapsed_time(e1) / iters
            out["%s_%s_wall_ms" % (name, "fwdbwd" if bwd else "fwd")] = (time.time() - t0) * 1000 / iters
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="2000,100000,1000000")
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--json", default=None)
    ap.add_argument("--shapes", default="66x1030x180,66x515x90")
    args = ap.parse_args()
    report, timing = [], []
    for shp in args.shapes.split(","):
        H, W, hf = [int(x) for x in shp.split("x")]
        for P in [int(x) for x in args.sizes.split(",")]:
            scene, cot, ref = compare(P, H, W, (-float(hf), float(hf)), synth.KITTI_VFOV, 0, P <= 200000, report)
            if args.time:
                t = time_both(scene, cot, ref)
                t.update(P=P, H=H, W=W, hfov=hf)
                timing.append(t)
                print(json.dumps(t), flush=True)
            del scene, cot, ref
            torch.cuda.empty_cache()
    if args.json:
        with open(args.json, "w") as f:
            json.dump(dict(report=report, timing=timing), f, indent=1)


if __name__ == "__main__":
    main()
