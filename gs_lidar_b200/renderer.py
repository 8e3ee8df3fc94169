"""Drop-in for GS-LiDAR's gaussian_renderer/__init__.py:render() with the per-surfel glue fused.

The reference's render() (gaussian_renderer/__init__.py:16-155) evaluates, in PyTorch and with one kernel + one autograd
node each: the SHM motion model xyz + v sin((t - t0) a) / a (scene/gaussian_model.py:151-157), the marginal
exp(-0.5 (t0 - t)^2 / sigma_t^2) (:185-186), sigmoid / exp / normalize activations (:139-175), opacity * marginal, and
the prefilter mask.  At 1M surfels these ~12 element-wise kernels and their backward passes move more bytes than the
rasterizer's own preprocess.  Here they are ONE streaming kernel per direction (csrc/gsl_glue.cu, C-ABI gsl_glue_forward
/ gsl_glue_backward) feeding the rasterizer of this package; the returned dict has the reference's keys.

`pc` needs the raw-parameter attributes of scene/gaussian_model.py:GaussianModel (_xyz, _velocity, _t, _scaling_t,
_opacity, _scaling, _rotation, get_features, active_sh_degree, T, velocity_decay); `viewpoint_camera` the attributes
render() reads from scene/cameras.py:Camera; `pipe` the flags of configs/base.yaml.  No CPU / PyTorch fallback.
"""
import ctypes as C
import math

import torch

from . import _lib as L
from . import diff_gaussian_rasterization_2d as _G
from .diff_gaussian_rasterization_2d import GaussianRasterizationSettings, GaussianRasterizer, _f32c, _stream_ptr

_lib = L.load()


class _ActivateSurfels(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, velocity, t, scaling_t, opacity, scaling, rotation, mask, timestamp, time_shift, cycle,
                velocity_decay, dynamic):
        if not xyz.is_cuda:
            raise RuntimeError("gs_lidar_b200 runs on CUDA tensors only (no CPU fallback)")
        dev, P = xyz.device, xyz.shape[0]
        raw = [_f32c(x) for x in (xyz, velocity, t, scaling_t, opacity, scaling, rotation)]
        m_in = None
        if mask is not None:
            m_in = mask.contiguous().view(torch.uint8) if mask.dtype == torch.bool else mask.to(torch.uint8).contiguous()
        p = L.gsl_glue_params(P, float(timestamp), float(time_shift or 0.0), float(cycle), float(velocity_decay), int(bool(dynamic)))
        with torch.cuda.device(dev):
            e = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
            means3D, opac, scales, rots, marg = e(P, 3), e(P, 1), e(P, 3), e(P, 4), e(P, 1)
            mask_out = torch.empty((P,), dtype=torch.bool, device=dev)
            gi = L.gsl_glue_inputs(*[x.data_ptr() if x.numel() else None for x in raw],
                                   m_in.data_ptr() if (m_in is not None and m_in.numel()) else None)
            go = L.gsl_glue_outputs(*[x.data_ptr() if x.numel() else None for x in (means3D, opac, scales, rots, marg, mask_out)])
            L.check(_lib.gsl_glue_forward(C.byref(p), C.byref(gi), C.byref(go), _stream_ptr(dev)), "gsl_glue_forward")
        ctx.glue = (p, raw, gi)
        # record of this call for a fused PeerExchange (render() hands it to the rasterizer's autograd node): under frame-
        # parallel training the frame-dependent part of this op's VJP is applied to every rank's rows BEFORE they are summed
        # (include/gsl_b200.h: gsl_peer_glue) and comes back as fold["extras"]
        ctx.fold = dict(p=p, raw=raw, extras=None)
        _ActivateSurfels.last_fold = ctx.fold
        ctx.mark_non_differentiable(marg, mask_out)
        return means3D, opac, scales, rots, marg, mask_out

    @staticmethod
    def backward(ctx, g_means3D, g_opacity, g_scales, g_rotations, _gm, _gk):
        p, raw, gi = ctx.glue
        dev, P = raw[0].device, p.P
        extras, ctx.fold["extras"] = ctx.fold["extras"], None
        if extras is not None:
            # summed over the ranks already, each with its own timestamp: g_means3D is dL/dxyz, g_opacity is
            # dL/d sigmoid(opacity); only the frame-independent Jacobians (sigmoid', exp', normalize') are left
            p = L.gsl_glue_params(p.P, p.timestamp, p.time_shift, p.cycle, p.velocity_decay, 0)
        cots = [None if g is None else _f32c(g) for g in (g_means3D, g_opacity, g_scales, g_rotations)]
        with torch.cuda.device(dev):
            outs = [torch.empty_like(x) for x in raw]
            go = L.gsl_glue_outputs(*[c.data_ptr() if (c is not None and c.numel()) else None for c in cots], None, None)
            gin = L.gsl_glue_inputs_grad(*[o.data_ptr() if o.numel() else None for o in outs])
            L.check(_lib.gsl_glue_backward(C.byref(p), C.byref(gi), C.byref(go), C.byref(gin), _stream_ptr(dev)),
                    "gsl_glue_backward")
        if extras is not None:
            outs[1] = extras["velocity"].reshape(raw[1].shape).contiguous()
            outs[2] = extras["t"].reshape(raw[2].shape).contiguous()
            outs[3] = extras["scaling_t"].reshape(raw[3].shape).contiguous()
        return (*outs, None, None, None, None, None, None)


def activate_surfels(pc, timestamp, time_shift=None, dynamic=False, mask=None):
    """(means3D, opacity, scales, rotations, marginal_t, prefilter mask) of GaussianModel `pc` at `timestamp`: the
    fused equivalent of get_xyz_SHM / get_inst_velocity / get_marginal_t / get_opacity / get_scaling / get_rotation and
    of the prefilter of gaussian_renderer/__init__.py:64-115."""
    # the kernel takes (timestamp, time_shift) and evaluates the motion model at timestamp - time_shift itself
    return _ActivateSurfels.apply(pc._xyz, pc._velocity, pc._t, pc._scaling_t, pc._opacity, pc._scaling, pc._rotation,
                                  mask, timestamp, time_shift or 0.0, pc.T, pc.velocity_decay, dynamic)


def render(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, override_color=None, env_map=None,
           time_shift=None, other=[], mask=None, is_training=False):
    """Same signature and returned dict as gaussian_renderer/__init__.py:render()."""
    # zero tensor whose gradient carries the screen-space (densification) signal, like the reference (:25-29)
    screenspace_points = torch.zeros((pc._xyz.shape[0], 4), dtype=pc._xyz.dtype, requires_grad=True, device=pc._xyz.device)
    if getattr(pipe, "neg_fov", True):
        tanfovx = tanfovy = math.tan(-0.5)
    else:
        tanfovx, tanfovy = math.tan(viewpoint_camera.FoVx * 0.5), math.tan(viewpoint_camera.FoVy * 0.5)
    raster_settings = GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height), image_width=int(viewpoint_camera.image_width),
        tanfovx=tanfovx, tanfovy=tanfovy, bg=bg_color, scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform, projmatrix=viewpoint_camera.full_proj_transform,
        sh_degree=pc.active_sh_degree, campos=viewpoint_camera.camera_center, prefiltered=False,
        debug=getattr(pipe, "debug", False), vfov=viewpoint_camera.vfov, hfov=viewpoint_camera.hfov,
        scale_factor=pipe.scale_factor)
    assert raster_settings.bg.shape[0] == 4, "expected bg color to be RGBA, got {}".format(raster_settings.bg.shape[0])
    if getattr(pipe, "compute_cov3D_python", False) or getattr(pipe, "convert_SHs_python", False):
        raise RuntimeError("compute_cov3D_python / convert_SHs_python are dead paths of the reference rasterizer "
                           "(forward.cu:237 reads scales/rotations unconditionally, NUM_CHANNELS is 4)")
    rasterizer = GaussianRasterizer(raster_settings=raster_settings)

    means3D, opacity, scales, rotations, marginal_t, pre_mask = activate_surfels(
        pc, viewpoint_camera.timestamp, time_shift, bool(getattr(pipe, "dynamic", False)), mask)

    shs, shs_rest, colors_precomp = None, None, override_color
    if override_color is None:
        dc, rest = getattr(pc, "_features_dc", None), getattr(pc, "_features_rest", None)
        if dc is not None and rest is not None and rest.shape[1] > 0:
            shs, shs_rest = dc, rest  # the two parameter tensors as they are: no get_features concatenation (:167-171)
        else:
            shs = pc.get_features
    if len(other) > 0:
        features = torch.cat(other, dim=1)
        S_other = features.shape[1]
    else:
        features = torch.zeros_like(means3D[:, :0])
        S_other = 0

    _G._pending_fold, _ActivateSurfels.last_fold = getattr(_ActivateSurfels, "last_fold", None), None  # consumed by the rasterizer forward
    contrib, rendered_image, rendered_feature, rendered_depth, rendered_opacity, radii = rasterizer(
        means3D=means3D, means2D=screenspace_points, shs=shs, colors_precomp=colors_precomp, features=features,
        opacities=opacity, scales=scales, rotations=rotations, cov3D_precomp=None, mask=pre_mask.view(-1, 1),
        shs_rest=shs_rest)

    _, rendered_intensity_sh, rendered_raydrop = rendered_image.split([2, 1, 1], dim=0)
    rendered_other, rendered_normal = rendered_feature.split([S_other, 3], dim=0)
    rendered_normal = rendered_normal / (rendered_normal.norm(dim=0, keepdim=True) + 1e-8)
    if env_map is not None:
        prior = env_map(viewpoint_camera.towards)
        rendered_raydrop = prior + (1 - prior) * rendered_raydrop
    return {
        "viewspace_points": screenspace_points,
        "visibility_filter": radii > 0,
        "radii": radii,
        "contrib": contrib,
        "depth": rendered_depth[[1]] if getattr(pipe, "median_depth", False) else rendered_depth[[0]],
        "depth_mean": rendered_depth[[0]],
        "depth_median": rendered_depth[[1]],
        "distortion": rendered_depth[[2]],
        "depth_square": rendered_depth[[3]],
        "alpha": rendered_opacity,
        "feature": rendered_other,
        "normal": rendered_normal,
        "intensity_sh": rendered_intensity_sh,
        "raydrop": rendered_raydrop.clamp(0, 1),
    }
