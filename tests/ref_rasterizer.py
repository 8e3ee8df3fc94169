"""TEST INFRASTRUCTURE: the reference's compiled CUDA rasterizer (oracle/_ref/libgslidar_ref.so, unmodified kernels) behind
the reference's own operator API -- GaussianRasterizationSettings / GaussianRasterizer with autograd -- so that the
reference's render() (oracle/ref_python.py) can run on the reference kernels and on this repo's drop-in side by side.
The argument handling follows gaussian_renderer/diff_gaussian_rasterization_2d.py:60-267 (what the torch binding in
rasterize_points.cu does with the tensors is done by oracle.RefCuda)."""
from typing import NamedTuple

import torch
import torch.nn as nn

import oracle


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool
    vfov: tuple
    hfov: tuple
    scale_factor: float


def _args(means3D, sh, colors_precomp, features, opacities, scales, rotations, mask, rs):
    c = lambda t: t.detach().float().contiguous()
    P = means3D.shape[0]
    return dict(P=P, S=features.shape[1] if features.dim() == 2 else 0, D=int(rs.sh_degree),
                M=sh.shape[1] if sh.numel() else 0, W=int(rs.image_width), H=int(rs.image_height), bg=c(rs.bg),
                means3D=c(means3D), shs=c(sh) if sh.numel() else None,
                colors_precomp=c(colors_precomp) if colors_precomp.numel() else None, features=c(features),
                opacities=c(opacities), scales=c(scales), rotations=c(rotations),
                mask=mask.detach().contiguous().view(torch.uint8) if mask.dtype == torch.bool else mask.to(torch.uint8).contiguous(),
                viewmatrix=c(rs.viewmatrix), projmatrix=c(rs.projmatrix), campos=c(rs.campos), tanfovx=float(rs.tanfovx),
                tanfovy=float(rs.tanfovy), vfov=(float(rs.vfov[0]), float(rs.vfov[1])),
                hfov=(float(rs.hfov[0]), float(rs.hfov[1])), scale_factor=float(rs.scale_factor))


class _Ref(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, features, opacities, scales, rotations, cov3Ds_precomp, mask, rs):
        ref = oracle.RefCuda()
        a = _args(means3D, sh, colors_precomp, features, opacities, scales, rotations, mask, rs)
        torch.cuda.synchronize()  # the reference runs on the legacy default stream
        f = ref.forward(a)
        torch.cuda.synchronize()
        ctx.ref, ctx.a, ctx.f = ref, a, f
        ctx.shapes = (tuple(means2D.shape), tuple(cov3Ds_precomp.shape), sh.numel() != 0, colors_precomp.numel() != 0)
        P = a["P"]
        contrib, radii = f["out_contrib"], f["radii"][:P]
        ctx.mark_non_differentiable(contrib, radii)
        return contrib, f["out_color"], f["out_feature"], f["out_depth"], 1 - f["out_T"], radii

    @staticmethod
    def backward(ctx, _gc, g_color, g_feature, g_depth, g_alpha, _gr):
        a, f = ctx.a, ctx.f
        H, W, S = a["H"], a["W"], a["S"]
        z = lambda g, n: torch.zeros((n, H, W), device=a["means3D"].device) if g is None else g.float().contiguous()
        cot = dict(color=z(g_color, 4), feature=z(g_feature, S + 3), depth=z(g_depth, 4), alpha=z(g_alpha, 1))
        torch.cuda.synchronize()
        g = ctx.ref.backward(a, f, cot)
        torch.cuda.synchronize()
        m2d_shape, cov_shape, have_sh, have_cp = ctx.shapes
        P = a["P"]
        return (g["dL_dmeans3D"], g["dL_dmeans2D"].reshape(m2d_shape), g["dL_dsh"] if have_sh else None,
                g["dL_dcolors"] if have_cp else None, g["dL_dfeatures"][:, :S] if S > 0 else g["dL_dfeatures"][:, :0],
                g["dL_dopacity"], g["dL_dscales"], g["dL_drotations"],
                g["dL_dcov3D"] if cov_shape == (P, 6) else None, None, None)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, features=None, scales=None, rotations=None,
                cov3D_precomp=None, mask=None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        e = lambda: torch.empty(0, device=means3D.device)
        shs = e() if shs is None else shs
        colors_precomp = e() if colors_precomp is None else colors_precomp
        features = torch.empty_like(means3D[..., :0]) if features is None else features
        cov3D_precomp = e() if cov3D_precomp is None else cov3D_precomp
        mask = torch.ones_like(means3D[:, :1], dtype=torch.bool) if mask is None else mask
        return _Ref.apply(means3D, means2D, shs, colors_precomp, features, opacities, scales, rotations, cov3D_precomp, mask,
                          self.raster_settings)
