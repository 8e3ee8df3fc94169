"""Callers and post-ops on either side of the rasterizer: the 360-degree range map of
`render_range_map` (gaussian_renderer/__init__.py:158-227 of the reference) and the two panorama post-ops the training
loop runs on every rendered range image, `pano_to_lidar` and `depth_to_normal` (utils/graphics_utils.py:96-149).

Same names, arguments and results as the reference's:

  * `pano_to_lidar` / `depth_to_normal` (and `pano_post_ops`, both from one pass) are CUDA ops of this package's library
    (csrc/gsl_postops.cu, C-ABI gsl_pano_forward / gsl_pano_backward) with autograd: 2 launches instead of the ~15
    element-wise PyTorch kernels each of them is in the reference (meshgrid, sin, cos, stack, normalize, mask, cross ...
    rebuilt on every call, three times per training step, train.py:261-262,306).  CUDA tensors only, like the rest of the
    package;
  * `render_range_map` stitches the two half panoramas with one concatenation per map instead of fifteen slice
    assignments into zero-filled buffers;
  * `render_range_map_360` renders the same maps in ONE rasterizer call with the azimuth wrap-around mode
    (GSL_FLAG_WRAP_AZIMUTH): the stitched panorama's columns are azimuth -180..180 deg of the front
    camera, which is exactly a 360-degree camera with the front camera's pose.

`renderFunc` is `gs_lidar_b200.renderer.render` (or the reference's).
"""
import ctypes as C

import torch

from . import _lib as L
from .diff_gaussian_rasterization_2d import _f32c, _stream_ptr

_lib = L.load()


def _pano_params(h, w, vfov, hfov):
    return L.gsl_pano_params(int(h), int(w), float(vfov[0]), float(vfov[1]), float(hfov[0]), float(hfov[1]))


class _PanoPostOps(torch.autograd.Function):
    """(points (K,3), normals (3,H,W)) of a (1,H,W) range image; either output can be switched off."""

    @staticmethod
    def forward(ctx, range_image, vfov, hfov, want_points, want_normals):
        if not range_image.is_cuda:
            raise RuntimeError("gs_lidar_b200 runs on CUDA tensors only (no CPU fallback)")
        h, w = range_image.shape[-2:]
        if range_image.numel() != h * w:
            raise RuntimeError("range image must be (1, H, W)")
        dev = range_image.device
        rng = _f32c(range_image)
        p = _pano_params(h, w, vfov, hfov)
        n = h * w
        with torch.cuda.device(dev):
            st = _stream_ptr(dev)
            points = index = count = scratch = normals = None
            if want_points:
                points = torch.empty((max(n, 1), 3), dtype=torch.float32, device=dev)
                index = torch.empty((max(n, 1),), dtype=torch.int32, device=dev)
                count = torch.zeros((1,), dtype=torch.int32, device=dev)
                scratch = torch.empty((int(_lib.gsl_pano_scratch_bytes(h, w)),), dtype=torch.uint8, device=dev)
            if want_normals:
                normals = torch.empty((3, h, w), dtype=torch.float32, device=dev)
            ptr = lambda t: None if t is None else t.data_ptr()
            if n > 0:
                L.check(_lib.gsl_pano_forward(C.byref(p), rng.data_ptr(), ptr(points), ptr(index), ptr(count), ptr(normals),
                                              ptr(scratch), st), "gsl_pano_forward")
            k = int(count.item()) if want_points else 0  # the one host read the (K, 3) result shape needs (the reference's
            #                                              boolean-mask indexing synchronises in the same place)
        ctx.pano = (p, rng, index, k, want_points, want_normals)
        out_points = points[:k] if want_points else torch.empty((0, 3), dtype=torch.float32, device=dev)
        out_normals = normals if want_normals else torch.empty((3, 0, 0), dtype=torch.float32, device=dev)
        return out_points.to(range_image.dtype), out_normals.to(range_image.dtype)

    @staticmethod
    def backward(ctx, g_points, g_normals):
        p, rng, index, k, want_points, want_normals = ctx.pano
        dev = rng.device
        gp = _f32c(g_points) if (want_points and g_points is not None and k > 0) else None
        gn = _f32c(g_normals) if (want_normals and g_normals is not None) else None
        with torch.cuda.device(dev):
            g_range = torch.empty_like(rng)
            if rng.numel() > 0:
                L.check(_lib.gsl_pano_backward(C.byref(p), rng.data_ptr(), k if gp is not None else 0,
                                               None if gp is None else gp.data_ptr(),
                                               None if index is None else index.data_ptr(),
                                               None if gn is None else gn.data_ptr(), g_range.data_ptr(), _stream_ptr(dev)),
                        "gsl_pano_backward")
        return g_range, None, None, None, None


def pano_post_ops(range_image, vfov, hfov):
    """(pano_to_lidar(range_image), depth_to_normal(range_image)) from one pass over the range image."""
    return _PanoPostOps.apply(range_image, tuple(vfov), tuple(hfov), True, True)


def pano_to_lidar(range_image, vfov, hfov):
    """(1, H, W) range image -> (K, 3) points of the pixels with range > 0, row-major (graphics_utils.py:96-118)."""
    return _PanoPostOps.apply(range_image, tuple(vfov), tuple(hfov), True, False)[0]


def depth_to_normal(range_image, vfov, hfov):
    """(1, H, W) range image -> (3, H, W) surface normals from central differences of the back-projected points, zero on
    the one-pixel border (graphics_utils.py:121-149)."""
    return _PanoPostOps.apply(range_image, tuple(vfov), tuple(hfov), False, True)[1]


def stitch_half_panoramas(front, back):
    """(C, H, w) front (azimuth -90..90) and back (90..270) half panoramas -> (C, H, 2w) with azimuth -180..180: the
    back half's last w//2 columns, the front half, the back half's first columns (gaussian_renderer/__init__.py:163,
    203-225: breaks = (0, w//2, 3w//2, 2w))."""
    w = front.shape[-1]
    left = w // 2                       # breaks[1] - breaks[0]
    right = 2 * w - (3 * w) // 2        # breaks[3] - breaks[2]
    return torch.cat([back[..., w - left:], front, back[..., :right]], dim=-1)


def _mixed_depth(pkg, args, eps=1e-5):
    """The three depth planes render_range_map keeps per view (:173-197): variance-gated mix of mean and median, mean,
    median; optionally blended with a sky depth."""
    depth, alpha = pkg["depth"], pkg["alpha"]
    var = pkg["depth_square"] - depth ** 2
    median = pkg["depth_median"]
    q = var.median() * 10
    mix = torch.where(var > q, median, torch.where(var <= q, depth, torch.zeros_like(depth)))  # NaN variance -> 0, like :181-183
    out = torch.cat([mix, depth, median])
    if getattr(args, "sky_depth", False):
        sky = 900
        out = out / alpha.clamp_min(eps)
        mode = getattr(args, "depth_blend_mode", 0)
        if mode == 0:    # harmonic mean
            out = 1 / (alpha / out.clamp_min(eps) + (1 - alpha) / sky).clamp_min(eps)
        elif mode == 1:
            out = alpha * out + (1 - alpha) * sky
    return out


def render_range_map(args, cam_front, cam_back, gaussians, renderFunc, renderArgs, env_map, hw):
    """Same signature and 5-tuple as the reference (:158-227): (depth (3,h,2w): mix / mean / median, intensity (1,h,2w),
    raydrop (1,h,2w), ground-truth depth (1,h,2w), ground-truth intensity (1,h,2w))."""
    assert cam_front.towards == "forward" and cam_back.towards == "backward"
    assert cam_front.colmap_id + args.frames == cam_back.colmap_id
    halves = []
    for cam in (cam_front, cam_back):
        pkg = renderFunc(cam, gaussians, *renderArgs, env_map=env_map)
        dev = pkg["depth"].device
        halves.append((_mixed_depth(pkg, args), pkg["intensity_sh"], pkg["raydrop"], cam.pts_depth.to(dev),
                       cam.pts_intensity.to(dev)))
    return tuple(stitch_half_panoramas(f, b) for f, b in zip(*halves))


def render_range_map_360(args, cam_360, gaussians, renderFunc, renderArgs, env_map):
    """The maps of render_range_map from ONE 360-degree call: `cam_360` has the front camera's pose, hfov = (-180, 180)
    and twice the width.  Needs the azimuth wrap-around mode (set_wrap_azimuth(True)) to treat the +-180 degree seam like
    any other column.  Returns (depth (3,h,W), intensity, raydrop); ground truth stays with the caller."""
    from . import diff_gaussian_rasterization_2d as G
    assert tuple(float(x) for x in cam_360.hfov) == (-180.0, 180.0), "render_range_map_360 needs a 360-degree camera"
    was = G._wrap_azimuth
    G.set_wrap_azimuth(True)
    try:
        pkg = renderFunc(cam_360, gaussians, *renderArgs, env_map=env_map)
    finally:
        G.set_wrap_azimuth(was)
    return _mixed_depth(pkg, args), pkg["intensity_sh"], pkg["raydrop"]
