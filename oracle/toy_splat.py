"""oracle/toy_splat.py -- TEST / BASELINE INFRASTRUCTURE ONLY (never imported by gs_lidar_b200).

CPU restatement of the reference's pure-PyTorch panorama surfel splatting, `surface_splatting` of
scripts/compare_2dgs_3dgs.py (BASELINE.json configs[0], SURVEY.md section 8c/8d "C1").  The reference script hard-codes
`.cuda()` and imports matplotlib, so it cannot run on the GPU box's host cores as it is (and /root/reference does not
exist there); this module states the same arithmetic for CPU tensors and is what bench.py times as the
`cpu_toy_baseline`.  It is a BASELINE, not a parity oracle for the CUDA path: the toy uses the cutoff dist2 < 1, no
low-pass filter, no near/far planes, no alpha clamps or early stop, a global depth sort, forward only.

Parity pinned: tests/golden/toy_*.npz hold outputs of the reference function itself, generated in the build container
by tests/golden/make_toy_golden.py (which imports the unmodified script through import-time shims);
tests/test_toy_splat_cpu.py compares this restatement with them.

Line references are to /root/reference/scripts/compare_2dgs_3dgs.py.
"""
import math

import torch

VFOV_DEG = (-20.0, 20.0)  # hard-coded in surface_splatting (:214-215)
HFOV_DEG = (-90.0, 90.0)


def _fov_radians():
    # :217-220 (torch.pi is a Python float: the arithmetic is float64, the results meet float32 tensors later)
    vmax = math.pi / 2 - VFOV_DEG[0] * math.pi / 180
    vmin = math.pi / 2 - VFOV_DEG[1] * math.pi / 180
    hmax = HFOV_DEG[1] * math.pi / 180
    hmin = HFOV_DEG[0] * math.pi / 180
    return vmin, vmax, hmin, hmax


def rotation_matrices(quats):
    """(P,4) quaternions (r, x, y, z), not necessarily unit -> (P,3,3); build_rotation :32-53."""
    q = quats / torch.sqrt((quats * quats).sum(dim=1, keepdim=True))
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    rows = [1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
            2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
            2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)]
    return torch.stack(rows, dim=1).view(-1, 3, 3)


def splat_setup(means3D, scales, quats, colors, opacities, viewmat, W, H):
    """Per-surfel stage (setup :152-210): view transform, T matrix, centre pixel, 12-sample AABB radii, depth sort.

    Returns T (P,3,3) with rows (u axis, v axis, centre) in view space, colours, opacities, centre pixels (P,3), depth
    (P,), radii (P,4) = extents left / right / up / down of the centre in pixels; all sorted by depth (range)."""
    vmin, vmax, hmin, hmax = _fov_radians()
    rot = viewmat[:3, :3]
    # rows of (R diag(s))^T are the scaled surfel axes (:159, build_scaling_rotation :56-65); then to view space (:163)
    axes = (rotation_matrices(quats) * scales[:, None, :]).transpose(1, 2) @ rot
    p_view = means3D @ rot + viewmat[-1:, :3]                                    # :162
    T = torch.stack([axes[:, 0], axes[:, 1], p_view], dim=1)                     # :164-166 (the 3x3 part of M)
    x, y, z = p_view[:, 0:1], p_view[:, 1:2], p_view[:, 2:3]
    phi = torch.atan2(x, z)                                                       # :171-173
    theta = torch.atan2(torch.sqrt(x ** 2 + z ** 2), -y)
    rng = torch.sqrt(x ** 2 + y ** 2 + z ** 2)
    sx, sy = W / (hmax - hmin), H / (vmax - vmin)
    centre = torch.cat([(phi - hmin) * sx, (theta - vmin) * sy, torch.ones_like(theta)], dim=-1)  # :175-177
    # AABB from 12 points on the 1-sigma ellipse (:182-196)
    ang = 2 * math.pi * torch.arange(0, 1, 1 / 12)
    ring = torch.stack([torch.sin(ang), torch.cos(ang), torch.ones_like(ang)], dim=1).to(T)
    pts = ring @ T                                                                # (P,12,3)
    s_phi = torch.atan2(pts[..., 0], pts[..., 2])
    s_theta = torch.atan2(torch.sqrt(pts[..., 0] ** 2 + pts[..., 2] ** 2), -pts[..., 1])
    radii = torch.cat([(phi - s_phi.min(dim=-1, keepdim=True)[0]) * sx,
                       (s_phi.max(dim=-1, keepdim=True)[0] - phi) * sx,
                       (theta - s_theta.min(dim=-1, keepdim=True)[0]) * sy,
                       (s_theta.max(dim=-1, keepdim=True)[0] - theta) * sy], dim=-1)
    order = rng[:, 0].sort()[1]                                                   # :203
    return T[order], colors[order], opacities[order], centre[order], rng[order, 0], radii[order]


def surface_splatting(means3D, scales, quats, colors, opacities, intrins, viewmat, pixel_chunk=4096):
    """Pure-PyTorch panorama splatting of 2D Gaussian surfels (surface_splatting :213-266 + alpha blending :326-354).

    intrins (3,3): only the principal point is used, W = 2 cx, H = 2 cy (:229-230).  Returns image (H,W,C), depth map
    (H,W,1), centre pixels (P,3), radii (P,4) like the reference (its fifth output, the (pixels, P) distance table, is
    not materialised: pixels are processed `pixel_chunk` at a time, which changes no arithmetic -- compositing is per
    pixel).  Every pixel evaluates every surfel, like the reference."""
    W, H = int((intrins[0, -1] * 2).long()), int((intrins[1, -1] * 2).long())
    vmin, vmax, hmin, hmax = _fov_radians()
    T, colors, opacities, centre, _, radii = splat_setup(means3D, scales, quats, colors, opacities, viewmat, W, H)
    c0, c1, c2 = T[None, :, :, 0], T[None, :, :, 1], T[None, :, :, 2]            # (1,P,3): x / y / z of (u, v, centre)
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")      # row-major pixels like :231-236
    xs, ys = xs.reshape(-1), ys.reshape(-1)
    image = torch.empty((H * W, colors.shape[-1]), dtype=colors.dtype)
    depth = torch.empty((H * W, 1), dtype=T.dtype)
    for lo in range(0, H * W, pixel_chunk):
        x = xs[lo:lo + pixel_chunk].view(-1, 1, 1)
        y = ys[lo:lo + pixel_chunk].view(-1, 1, 1)
        phi = x * (hmax - hmin) / W + hmin                                        # :238-239
        theta = y * (vmax - vmin) / H + vmin
        # the two planes through the pixel's ray, in splat coordinates, and their intersection (:241-244)
        k = torch.cos(phi) * c0 - torch.sin(phi) * c2
        l = torch.cos(theta) * torch.sin(phi) * c0 + torch.sin(theta) * c1 + torch.cos(theta) * torch.cos(phi) * c2
        hit = torch.cross(k, l, dim=-1)
        s = hit[..., :2] / hit[..., -1:]
        dist2 = (s * s).sum(dim=-1)                                               # :250, :255 (no low-pass: dist2 = dist3d)
        hs = torch.cat([s, torch.ones_like(s[..., :1])], dim=-1)
        rng = ((hs * c0).sum(dim=-1) * torch.sin(theta)[..., 0] * torch.sin(phi)[..., 0]     # :256-258
               - (hs * c1).sum(dim=-1) * torch.cos(theta)[..., 0]
               + (hs * c2).sum(dim=-1) * torch.sin(theta)[..., 0] * torch.cos(phi)[..., 0])
        # alpha_blending_with_gaussians :333-354: front-to-back over the depth-sorted surfels, cutoff at 1 sigma
        d2 = dist2.T                                                              # (P, pixels)
        alpha = opacities.unsqueeze(1) * (torch.exp(-0.5 * d2) * (d2 < 1))[..., None]
        trans = torch.cat([torch.ones_like(alpha[-1:]), (1 - alpha).cumprod(dim=0)[:-1]], dim=0)   # :327
        w = trans * alpha
        image[lo:lo + pixel_chunk] = (w * colors.reshape(-1, 1, colors.shape[-1])).sum(dim=0)
        depth[lo:lo + pixel_chunk] = torch.nan_to_num((w * rng.T[..., None]).sum(dim=0), 0, 0)     # :352
    return image.reshape(H, W, -1), depth.reshape(H, W, -1), centre, radii


def make_inputs(num_surfels, W, H, seed=0, dtype=torch.float32):
    """A synthetic scene of the toy's kind for a W x H panorama with the toy's fixed field of view (+-90 x +-20 deg):
    surfels a few pixels wide at 2..8 units, camera at the origin looking along +z (viewmat = identity)."""
    g = torch.Generator().manual_seed(seed)
    u = lambda *s: torch.rand(*s, generator=g, dtype=torch.float64)
    phi = (u(num_surfels) * 2 - 1) * math.radians(88.0)
    theta = math.pi / 2 + (u(num_surfels) * 2 - 1) * math.radians(19.0)
    r = torch.exp(math.log(2.0) + u(num_surfels) * (math.log(8.0) - math.log(2.0)))
    means = torch.stack([r * torch.sin(theta) * torch.sin(phi), -r * torch.cos(theta), r * torch.sin(theta) * torch.cos(phi)], dim=1)
    px = math.pi / W  # radians per pixel
    scales = torch.stack([r * px * (1.0 + 3.0 * u(num_surfels)), r * px * (1.0 + 3.0 * u(num_surfels)), torch.zeros(num_surfels, dtype=torch.float64)], dim=1)
    quats = torch.randn(num_surfels, 4, generator=g, dtype=torch.float64)
    colors = u(num_surfels, 3)
    opac = 0.3 + 0.7 * u(num_surfels, 1)
    intrins = torch.tensor([[700.0, 0.0, W / 2], [0.0, 700.0, H / 2], [0.0, 0.0, 1.0]], dtype=torch.float64)
    viewmat = torch.eye(4, dtype=torch.float64)
    return tuple(t.to(dtype) for t in (means, scales, quats, colors, opac, intrins, viewmat))
