"""Callers and post-ops on either side of the rasterizer: the 360-degree range map of
`render_range_map` (gaussian_renderer/__init__.py:158-227 of the reference) and the two panorama post-ops the training
loop runs on every rendered range image, `pano_to_lidar` and `depth_to_normal` (utils/graphics_utils.py:96-149).

Host-side mirrors in plain PyTorch, same names, arguments and results:

  * the ray directions of a (H, W, vfov, hfov) panorama are built once per shape and device and reused (the reference
    rebuilds meshgrid / sin / cos / normalize -- about ten element-wise kernels -- on every call, three times per
    training step, train.py:261-262,306);
  * `render_range_map` stitches the two half panoramas with one concatenation per map instead of fifteen slice
    assignments into zero-filled buffers;
  * `render_range_map_360` renders the same maps in ONE rasterizer call with the azimuth wrap-around mode
    (GSL_FLAG_WRAP_AZIMUTH): the stitched panorama's columns are azimuth -180..180 deg of the front
    camera, which is exactly a 360-degree camera with the front camera's pose.

Nothing here touches the CUDA library directly; `renderFunc` is `gs_lidar_b200.renderer.render` (or the reference's).
"""

import torch
import torch.nn.functional as F

_dir_cache = {}


def ray_directions(height, width, vfov, hfov, device, dtype=torch.float32):
    """(3, H, W) unit ray directions of the panorama pixels, x right / y down / z forward, as both reference post-ops
    compute them (graphics_utils.py:99-116): theta = (90 - vfov[1] + row / H * (vfov[1] - vfov[0])) deg,
    phi = (hfov[0] + col / W * (hfov[1] - hfov[0])) deg.  Cached per shape, field of view, device and dtype."""
    key = (int(height), int(width), float(vfov[0]), float(vfov[1]), float(hfov[0]), float(hfov[1]), str(device), dtype)
    d = _dir_cache.get(key)
    if d is None:
        rows, cols = torch.meshgrid(torch.arange(height, device=device), torch.arange(width, device=device), indexing="ij")
        theta = (90 - vfov[1] + rows / height * (vfov[1] - vfov[0])) * torch.pi / 180
        phi = (hfov[0] + cols / width * (hfov[1] - hfov[0])) * torch.pi / 180
        d = torch.stack([torch.sin(theta) * torch.sin(phi), -torch.cos(theta), torch.sin(theta) * torch.cos(phi)], dim=0)
        d = F.normalize(d, dim=0).to(dtype)
        if len(_dir_cache) > 32:
            _dir_cache.clear()
        _dir_cache[key] = d
    return d


def pano_to_lidar(range_image, vfov, hfov):
    """(1, H, W) range image -> (K, 3) points of the pixels with range > 0, row-major (graphics_utils.py:96-118)."""
    h, w = range_image.shape[-2:]
    d = ray_directions(h, w, vfov, hfov, range_image.device, range_image.dtype)
    return (d * range_image)[:, range_image[0] > 0].permute(1, 0)


def depth_to_normal(range_image, vfov, hfov):
    """(1, H, W) range image -> (3, H, W) surface normals from central differences of the back-projected points, zero on
    the one-pixel border (graphics_utils.py:121-149)."""
    h, w = range_image.shape[-2:]
    pts = ray_directions(h, w, vfov, hfov, range_image.device, range_image.dtype) * range_image
    out = torch.zeros_like(pts)
    down = pts[:, 2:, 1:-1] - pts[:, :-2, 1:-1]
    right = pts[:, 1:-1, 2:] - pts[:, 1:-1, :-2]
    out[:, 1:-1, 1:-1] = F.normalize(torch.cross(down, right, dim=0), dim=0)
    return out


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
