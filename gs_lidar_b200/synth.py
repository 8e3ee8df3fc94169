"""Seeded synthetic surfel scenes at KITTI-360 / OPV2V panorama shapes.

Everything is generated on the CPU with a seeded torch.Generator (so the same scene is reproduced
bit-for-bit anywhere) and moved to the requested device by the caller.  Shapes follow
gaussian_renderer/__init__.py:25-128 (what render() feeds the rasterizer): means3D (P,3), opacity
(P,1), scales (P,3), rotations (P,4) un-normalised, SH (P,16,4), features (P,S), mask (P,1) bool.
"""
import math
from typing import NamedTuple

import torch

KITTI_VFOV = (-24.9, 2.0)       # configs/kitti360_nvs_1908.yaml:16
OPV2V_VFOV = (-25.0, 2.0)       # configs/opv2v_dynamic_2.yaml:18
SCALE_FACTOR = 0.1              # scene/kitti360_loader.py:91-96


class Scene(NamedTuple):
    means3D: torch.Tensor
    opacities: torch.Tensor
    scales: torch.Tensor
    rotations: torch.Tensor
    shs: torch.Tensor
    features: torch.Tensor
    mask: torch.Tensor
    viewmatrix: torch.Tensor   # transposed world->camera (scene/cameras.py:62)
    projmatrix: torch.Tensor
    campos: torch.Tensor
    bg: torch.Tensor
    H: int
    W: int
    vfov: tuple
    hfov: tuple
    scale_factor: float
    sh_degree: int

    def to(self, device):
        return Scene(*[v.to(device) if isinstance(v, torch.Tensor) else v for v in self])


def make_scene(P, H=66, W=1030, vfov=KITTI_VFOV, hfov=(-180.0, 180.0), S=4, sh_degree=3, seed=0,
               scale_factor=SCALE_FACTOR, footprint_px=0.8, view_yaw_deg=0.0, view_shift=(0.0, 0.0, 0.0)):
    g = torch.Generator(device="cpu").manual_seed(seed)
    u = lambda *s: torch.rand(*s, generator=g)
    n = lambda *s: torch.randn(*s, generator=g)
    th_min = math.pi / 2 - math.radians(vfov[1])
    th_max = math.pi / 2 - math.radians(vfov[0])
    d = th_max - th_min
    phi = (u(P) * 2 - 1) * math.pi
    theta = th_min - 0.05 * d + u(P) * 1.1 * d
    r = torch.exp(math.log(0.3) + u(P) * (math.log(8.0) - math.log(0.3)))
    # view-space position: p = r (sin th sin ph, -cos th, sin th cos ph)   (forward.cu:116-125 inverted)
    pv = torch.stack([r * torch.sin(theta) * torch.sin(phi), -r * torch.cos(theta),
                      r * torch.sin(theta) * torch.cos(phi)], dim=1)
    # camera: yaw about the y axis + shift; world = R^T (p_view - t)
    yaw = math.radians(view_yaw_deg)
    Rm = torch.tensor([[math.cos(yaw), 0.0, math.sin(yaw)], [0.0, 1.0, 0.0], [-math.sin(yaw), 0.0, math.cos(yaw)]])
    t = torch.tensor(view_shift, dtype=torch.float32)
    V = torch.eye(4)
    V[:3, :3] = Rm
    V[:3, 3] = t
    means3D = (pv - t) @ Rm  # = R^T (pv - t) for row vectors
    px = 2 * math.pi / 1030.0  # angular size of one KITTI panorama column
    sxy = r[:, None] * px * torch.exp(math.log(footprint_px) + 0.5 * n(P, 2))
    scales = torch.cat([sxy, 0.01 * r[:, None] * torch.exp(n(P, 1))], dim=1)
    rotations = n(P, 4)
    opacities = torch.sigmoid(1.5 * n(P, 1))
    M = 16
    shs = torch.zeros(P, M, 4)
    shs[:, 0, :] = (u(P, 4) * 2 - 1) / 0.28209479177387814
    shs[:, 1:, :] = 0.1 * n(P, M - 1, 4)
    features = n(P, S) if S > 0 else torch.zeros(P, 0)
    mask = opacities > (1.0 / 255.0)
    viewmatrix = V.t().contiguous()  # transposed, as world_view_transform
    campos = torch.linalg.inv(V)[:3, 3].contiguous()
    bg = torch.tensor([0.0, 0.0, 0.0, 1.0])
    return Scene(means3D.float().contiguous(), opacities.float(), scales.float().contiguous(), rotations.float(),
                 shs.float(), features.float().contiguous(), mask, viewmatrix.float(), viewmatrix.float().clone(),
                 campos.float(), bg, H, W, tuple(vfov), tuple(hfov), scale_factor, sh_degree)


def make_cotangents(H, W, S, seed=1):
    g = torch.Generator(device="cpu").manual_seed(seed)
    n = lambda *s: torch.randn(*s, generator=g)
    return dict(color=n(4, H, W), feature=n(S + 3, H, W), depth=n(4, H, W), alpha=n(1, H, W))


def pattern_cotangents(H, W, S):
    """Deterministic cotangents made of exactly representable values (integer arithmetic / 8), so
    golden fixtures need not store them."""
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")

    def chan(c, mul):
        return (((xs * 7 + ys * 13 + c * 5 + mul) % 17) - 8).float() / 8.0

    mk = lambda n, mul: torch.stack([chan(c, mul) for c in range(n)], 0)
    return dict(color=mk(4, 0), feature=mk(S + 3, 3), depth=mk(4, 6), alpha=mk(1, 9))


def settings_for(scene, debug=False):
    from .diff_gaussian_rasterization_2d import GaussianRasterizationSettings
    return GaussianRasterizationSettings(
        image_height=scene.H, image_width=scene.W, tanfovx=math.tan(-0.5), tanfovy=math.tan(-0.5), bg=scene.bg,
        scale_modifier=1.0, viewmatrix=scene.viewmatrix, projmatrix=scene.projmatrix, sh_degree=scene.sh_degree,
        campos=scene.campos, prefiltered=False, debug=debug, vfov=scene.vfov, hfov=scene.hfov,
        scale_factor=scene.scale_factor)
