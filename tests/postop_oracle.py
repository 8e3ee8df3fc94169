"""TEST INFRASTRUCTURE: PyTorch restatement of the reference's panorama post-ops (utils/graphics_utils.py:96-118
pano_to_lidar, :121-149 depth_to_normal), pinned on CPU against outputs of the reference functions themselves
(tests/golden/postop_*.npz, tests/test_range_map_cpu.py); the checker for the CUDA post-ops of gs_lidar_b200.range_map."""
import torch
import torch.nn.functional as F


def ray_directions(height, width, vfov, hfov, device, dtype=torch.float32):
    """(3, H, W) unit ray directions as both reference post-ops compute them (graphics_utils.py:99-116)."""
    rows, cols = torch.meshgrid(torch.arange(height, device=device), torch.arange(width, device=device), indexing="ij")
    theta = (90 - vfov[1] + rows / height * (vfov[1] - vfov[0])) * torch.pi / 180
    phi = (hfov[0] + cols / width * (hfov[1] - hfov[0])) * torch.pi / 180
    d = torch.stack([torch.sin(theta) * torch.sin(phi), -torch.cos(theta), torch.sin(theta) * torch.cos(phi)], dim=0)
    return F.normalize(d, dim=0).to(dtype)


def pano_to_lidar(range_image, vfov, hfov):
    h, w = range_image.shape[-2:]
    d = ray_directions(h, w, vfov, hfov, range_image.device, range_image.dtype)
    return (d * range_image)[:, range_image[0] > 0].permute(1, 0)


def depth_to_normal(range_image, vfov, hfov):
    h, w = range_image.shape[-2:]
    pts = ray_directions(h, w, vfov, hfov, range_image.device, range_image.dtype) * range_image
    out = torch.zeros_like(pts)
    down = pts[:, 2:, 1:-1] - pts[:, :-2, 1:-1]
    right = pts[:, 1:-1, 2:] - pts[:, 1:-1, :-2]
    out[:, 1:-1, 1:-1] = F.normalize(torch.cross(down, right, dim=0), dim=0)
    return out
