"""TEST INFRASTRUCTURE: PyTorch restatement of the per-surfel glue of the reference's render()
(gaussian_renderer/__init__.py:64-115) and of the GaussianModel accessors it calls
(scene/gaussian_model.py:139-186), used as the checker for gs_lidar_b200.renderer."""
import math
from types import SimpleNamespace

import numpy as np
import torch


def make_model(P, seed=0, device="cpu", T=0.2, velocity_decay=1.0, S_sh=16):
    """Raw GaussianModel-like parameters (shapes of scene/gaussian_model.py:266-298) as leaf tensors."""
    g = torch.Generator().manual_seed(seed)
    n = lambda *s: torch.randn(*s, generator=g)
    pc = SimpleNamespace()
    pc._xyz = (n(P, 3) * 2.0).to(device).requires_grad_(True)
    pc._velocity = (n(P, 3) * 0.05).to(device).requires_grad_(True)
    pc._t = (torch.rand(P, 1, generator=g) * 1.2 - 0.6).to(device).requires_grad_(True)
    pc._scaling_t = (math.log(0.1) + 0.3 * n(P, 1)).to(device).requires_grad_(True)
    pc._opacity = (1.5 * n(P, 1)).to(device).requires_grad_(True)
    pc._scaling = (math.log(0.02) + 0.5 * n(P, 3)).to(device).requires_grad_(True)
    pc._rotation = n(P, 4).to(device).requires_grad_(True)
    pc._features_dc = ((torch.rand(P, 1, 4, generator=g) * 2 - 1) / 0.28209479177387814).to(device).requires_grad_(True)
    pc._features_rest = (0.1 * n(P, S_sh - 1, 4)).to(device).requires_grad_(True)
    pc.T, pc.velocity_decay, pc.active_sh_degree = T, velocity_decay, 3
    return pc


RAW = ("_xyz", "_velocity", "_t", "_scaling_t", "_opacity", "_scaling", "_rotation")


def get_features(pc):  # gaussian_model.py:167-171
    return torch.cat((pc._features_dc, pc._features_rest), dim=1)


def reference_glue(pc, timestamp, time_shift=None, dynamic=False, mask=None):
    """means3D, opacity, scales, rotations, marginal_t, mask exactly as render() computes them."""
    scaling_t = torch.exp(pc._scaling_t)                                             # gaussian_model.py:144-145
    def xyz_shm(t):                                                                  # :151-153
        a = 1 / pc.T * np.pi * 2
        return pc._xyz + pc._velocity * torch.sin((t - pc._t) * a) / a
    def marginal(t):                                                                 # :185-186
        return torch.exp(-0.5 * (pc._t - t) ** 2 / scaling_t ** 2)
    if time_shift is not None:                                                       # __init__.py:69-75
        means3D = xyz_shm(timestamp - time_shift)
        inst_v = pc._velocity * torch.exp(-scaling_t / pc.T / 2 * pc.velocity_decay)  # :155-157
        means3D = means3D + inst_v * time_shift
        marginal_t = marginal(timestamp - time_shift)
    else:
        means3D = xyz_shm(timestamp)
        marginal_t = marginal(timestamp)
    opacity = torch.sigmoid(pc._opacity)                                             # :174-175
    if dynamic:                                                                      # __init__.py:77-79
        opacity = opacity * marginal_t
    scales = torch.exp(pc._scaling)                                                  # :140-141
    rotations = torch.nn.functional.normalize(pc._rotation)                          # :148-149
    m = (opacity[:, 0] > 1 / 255) if mask is None else mask & (opacity[:, 0] > 1 / 255)  # __init__.py:112-115
    if dynamic:
        m = m & (marginal_t[:, 0] > 0.05)
    return means3D, opacity, scales, rotations, marginal_t, m
