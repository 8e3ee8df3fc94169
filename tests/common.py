"""Shared helpers for the parity tests: run the product (gs_lidar_b200, through its public API and
the C-ABI), the reference CUDA rasterizer (oracle.RefCuda) and the CPU oracle (oracle.CpuOracle) on
the same seeded scene and compare.
"""
import ctypes as C
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from gs_lidar_b200 import synth  # noqa: E402

TANFOV = math.tan(-0.5)


def ref_args(scene, colors_precomp=None):
    """dict of tensors/scalars in the layout oracle.RefCuda expects (scene must be on the GPU)."""
    P = scene.means3D.shape[0]
    S = scene.features.shape[1]
    use_sh = colors_precomp is None
    return dict(P=P, S=S, D=scene.sh_degree, M=scene.shs.shape[1] if use_sh else 0, H=scene.H, W=scene.W, bg=scene.bg,
                means3D=scene.means3D, shs=scene.shs if use_sh else None, colors_precomp=colors_precomp,
                features=scene.features, opacities=scene.opacities, scales=scene.scales, rotations=scene.rotations,
                mask=scene.mask.view(torch.uint8) if scene.mask.dtype == torch.bool else scene.mask,
                viewmatrix=scene.viewmatrix, projmatrix=scene.projmatrix, campos=scene.campos, tanfovx=TANFOV,
                tanfovy=TANFOV, vfov=scene.vfov, hfov=scene.hfov, scale_factor=scene.scale_factor)


def run_ours(scene, cot=None, colors_precomp=None, export=True, debug=False):
    """Forward (+ backward when cot is given) through the public GaussianRasterizer API."""
    import gs_lidar_b200.diff_gaussian_rasterization_2d as G
    from gs_lidar_b200 import _lib as L
    settings = synth.settings_for(scene, debug=debug)
    rast = G.GaussianRasterizer(settings)
    leaves = {}

    def leaf(name, t):
        t = t.detach().clone().requires_grad_(cot is not None or export)
        leaves[name] = t
        return t

    P = scene.means3D.shape[0]
    means3D = leaf("means3D", scene.means3D)
    means2D = leaf("means2D", torch.zeros((P, 4), device=scene.means3D.device))
    opac = leaf("opacities", scene.opacities)
    scales = leaf("scales", scene.scales)
    rots = leaf("rotations", scene.rotations)
    feats = leaf("features", scene.features)
    kw = {}
    if colors_precomp is None:
        kw["shs"] = leaf("shs", scene.shs)
    else:
        kw["colors_precomp"] = leaf("colors_precomp", colors_precomp)
    if export:
        G.set_keep_workspace_after_backward(True)
    try:
        contrib, color, feature, depth, alpha, radii = rast(means3D=means3D, means2D=means2D, opacities=opac,
                                                            features=feats, scales=scales, rotations=rots,
                                                            mask=scene.mask, **kw)
        out = dict(out_contrib=contrib, out_color=color, out_feature=feature, out_depth=depth, out_alpha=alpha,
                   radii=radii)
        node = color.grad_fn
        state = None
        if export and node is not None:
            state = export_state(node, scene, P)
        grads = None
        if cot is not None:
            loss = (color * cot["color"]).sum() + (feature * cot["feature"]).sum() + (depth * cot["depth"]).sum() + \
                   (alpha * cot["alpha"]).sum()
            loss.backward()
            grads = {k: (v.grad.detach() if v.grad is not None else None) for k, v in leaves.items()}
    finally:
        G.set_keep_workspace_after_backward(False)
    return out, state, grads


def export_state(node, scene, P):
    """Dumps the product's internal state in the reference's layouts via gsl_export_state."""
    import gs_lidar_b200.diff_gaussian_rasterization_2d as G
    from gs_lidar_b200 import _lib as L
    ws = node.holder.ws
    params = node.gsl_params
    R = node.num_rendered
    dev = scene.means3D.device
    H, W = scene.H, scene.W
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    t = dict(depths=torch.zeros(P, device=dev), means2D=torch.zeros((P, 2), device=dev),
             transMat=torch.zeros((P, 9), device=dev), normal_opacity=torch.zeros((P, 4), device=dev),
             rgb=torch.zeros((P, 4), device=dev), clamped=torch.zeros((P, 4), dtype=torch.uint8, device=dev),
             tiles_touched=torch.zeros(P, dtype=torch.int32, device=dev),
             point_offsets=torch.zeros(P, dtype=torch.int32, device=dev),
             point_list_keys=torch.zeros(max(R, 1), dtype=torch.int64, device=dev),
             point_list=torch.zeros(max(R, 1), dtype=torch.int32, device=dev),
             ranges=torch.zeros((tiles, 2), dtype=torch.int32, device=dev), final_T=torch.zeros((3, H, W), device=dev),
             pixbox=torch.zeros((P, 4), dtype=torch.int16, device=dev))
    ex = L.gsl_state_export()
    for k, v in t.items():
        setattr(ex, k, v.data_ptr())
    wss = ws.as_struct()
    L.check(G._lib.gsl_export_state(C.byref(params), C.byref(wss), R, C.byref(ex),
                                    C.c_void_p(torch.cuda.current_stream().cuda_stream)), "gsl_export_state")
    torch.cuda.synchronize()
    t["point_list_keys"] = t["point_list_keys"][:R]
    t["point_list"] = t["point_list"][:R]
    t["R"] = R
    return t


def run_ref(scene, cot=None, colors_precomp=None, ref=None):
    import oracle
    ref = ref or oracle.RefCuda()
    a = ref_args(scene, colors_precomp)
    fwd = ref.forward(a)
    P = a["P"]
    state = ref.state(P, fwd["R"], scene.H, scene.W) if P > 0 else None
    grads = None
    if cot is not None and P > 0:
        grads = ref.backward(a, fwd, cot)
        torch.cuda.synchronize()
        if state is not None:
            st2 = ref.state(P, fwd["R"], scene.H, scene.W)
            state["dL_dtransMat"], state["dL_dnormals"] = st2["dL_dtransMat"], st2["dL_dnormals"]
    torch.cuda.synchronize()
    out = dict(out_contrib=fwd["out_contrib"], out_color=fwd["out_color"], out_feature=fwd["out_feature"],
               out_depth=fwd["out_depth"], out_alpha=1 - fwd["out_T"], radii=fwd["radii"][:P])
    return out, state, grads, ref


def run_oracle(scene, cot=None, colors_precomp=None, threads=None, wrap=False):
    """CPU oracle on a CPU copy of the scene; returns (state dict incl. outputs, grads)."""
    import oracle
    o = oracle.CpuOracle(threads)
    s = scene.to("cpu")
    P, S = s.means3D.shape[0], s.features.shape[1]
    M = s.shs.shape[1] if colors_precomp is None else 0
    p = o.params(P, S, s.sh_degree, M, s.W, s.H, s.vfov, s.hfov, s.scale_factor, TANFOV, TANFOV, wrap=wrap)
    shs = s.shs.numpy() if colors_precomp is None else None
    cp = None if colors_precomp is None else colors_precomp.cpu().numpy()
    st = o.forward(p, s.means3D.numpy(), s.scales.numpy(), s.rotations.numpy(), s.opacities.numpy(), shs, cp,
                   s.features.numpy(), s.mask.numpy(), s.viewmatrix.numpy(), s.campos.numpy(), s.bg.numpy())
    grads = None
    if cot is not None:
        c = {k: v.cpu().numpy() for k, v in cot.items()}
        grads = o.backward(p, st, s.means3D.numpy(), s.scales.numpy(), s.rotations.numpy(), shs, s.features.numpy(),
                           s.viewmatrix.numpy(), s.campos.numpy(), s.bg.numpy(), c["color"], c["depth"], c["alpha"],
                           c["feature"])
    return st, grads


def rel_err(a, b, floor=None, floor_frac=1e-3):
    """max |a-b| / max(|b|, floor); floor defaults to floor_frac * max|b| so near-zero entries do not dominate."""
    a = a.detach().double().cpu() if isinstance(a, torch.Tensor) else torch.from_numpy(np.asarray(a)).double()
    b = b.detach().double().cpu() if isinstance(b, torch.Tensor) else torch.from_numpy(np.asarray(b)).double()
    if a.numel() == 0:
        return 0.0
    if floor is None:
        floor = floor_frac * float(b.abs().max()) + 1e-30
    return float(((a - b).abs() / b.abs().clamp_min(floor)).max())


def frac_mismatch(a, b):
    a = a.detach().cpu() if isinstance(a, torch.Tensor) else torch.from_numpy(np.asarray(a))
    b = b.detach().cpu() if isinstance(b, torch.Tensor) else torch.from_numpy(np.asarray(b))
    if a.numel() == 0:
        return 0.0
    return float((a != b).double().mean())


def bits_equal(a, b):
    """bit-exact comparison of float tensors (treats -0 != +0 and compares NaN payloads)."""
    return torch.equal(a.detach().cpu().contiguous().view(torch.int32), b.detach().cpu().contiguous().view(torch.int32))


def grad_err(a, b):
    """Gradient criterion.  A per-surfel gradient is a sum of many signed per-pixel contributions which the
    reference accumulates with atomics in nondeterministic order, so entries far below the tensor's scale are
    cancellation residues that the reference itself does not reproduce run to run.  Elementwise relative error
    is therefore floored at 1% of the tensor's max magnitude; the norm-wise error is returned as well."""
    a = a.detach().double().cpu().flatten()
    b = b.detach().double().cpu().flatten()
    if a.numel() == 0:
        return 0.0, 0.0
    return rel_err(a, b, floor_frac=1e-2), float((a - b).norm() / (b.norm() + 1e-300))


def report(name, measured):
    """Appends the measured error maxima of a parity check to gpurun_out/parity_measured.jsonl (when that directory exists or
    GSL_REPORT names a file) and prints them -- visible with `pytest -s` and in the captured output of a failure."""
    import json
    line = json.dumps({"check": name, "measured": measured}, default=float)
    print(line)
    path = os.environ.get("GSL_REPORT")
    if not path:
        d = os.path.join(ROOT, "gpurun_out")
        path = os.path.join(d, "parity_measured.jsonl") if os.path.isdir(d) else None
    if path:
        try:
            with open(path, "a") as f:
                f.write(line + "\n")
        except OSError:
            pass
