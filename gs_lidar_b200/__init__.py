"""gs_lidar_b200 -- B200-native (sm_100a) panoramic 2D-Gaussian-surfel rasterizer.

Drop-in for the ONE hot path of GS-LiDAR: the `GaussianRasterizationSettings` / `GaussianRasterizer`
API of gaussian_renderer/diff_gaussian_rasterization_2d.py, backed by hand-written CUDA behind the
C-ABI in include/gsl_b200.h.  The first access to any API name loads libgsl_b200.so and fails loudly if
it is missing or stale: there is no CPU or PyTorch fallback.  (The names are resolved lazily only so that
`python -m gs_lidar_b200.build` can rebuild the library when the one on disk is out of date.)
"""
__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians",
           "set_keep_workspace_after_backward", "set_cuda_graphs", "set_wrap_azimuth"]


def __getattr__(name):
    if name in __all__:
        from . import diff_gaussian_rasterization_2d as _m
        return getattr(_m, name)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
