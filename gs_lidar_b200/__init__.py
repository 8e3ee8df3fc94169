"""gs_lidar_b200 -- B200-native (sm_100a) panoramic 2D-Gaussian-surfel rasterizer.

Drop-in for the ONE hot path of GS-LiDAR: the `GaussianRasterizationSettings` / `GaussianRasterizer`
API of gaussian_renderer/diff_gaussian_rasterization_2d.py, backed by hand-written CUDA behind the
C-ABI in include/gsl_b200.h.  Importing this package loads libgsl_b200.so and fails loudly if it is
missing: there is no CPU or PyTorch fallback.
"""
from .diff_gaussian_rasterization_2d import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    rasterize_gaussians,
    set_keep_workspace_after_backward,
)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians",
           "set_keep_workspace_after_backward"]
