"""gs_lidar_b200/range_map.py (the callers / post-ops either side of the rasterizer) on CPU:
the stitching against a restatement of the reference's slice assignments, and the CHECKER of the CUDA post-ops
(tests/postop_oracle.py, a PyTorch restatement) against outputs of the reference's own functions
(tests/golden/postop_*.npz, made by tests/golden/make_postop_golden.py).  The CUDA post-ops themselves are compared with
that checker, the stored outputs and the live reference functions in tests/test_postops_gpu.py."""
import glob
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from gs_lidar_b200 import range_map
import postop_oracle

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "postop_*.npz")))


def test_fixtures_exist():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_post_op_checker_matches_the_reference_functions(path):
    g = np.load(path)
    rng = torch.from_numpy(g["range_image"])
    vfov, hfov = tuple(g["vfov"].tolist()), tuple(g["hfov"].tolist())
    pts = postop_oracle.pano_to_lidar(rng, vfov, hfov)
    nrm = postop_oracle.depth_to_normal(rng, vfov, hfov)
    assert pts.shape == g["points"].shape
    np.testing.assert_allclose(pts.numpy(), g["points"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(nrm.numpy(), g["normals"], rtol=0, atol=1e-5)
    assert float(nrm[:, 0].abs().sum()) == 0 and float(nrm[:, :, -1].abs().sum()) == 0   # zero border


def test_the_cuda_post_ops_refuse_cpu_tensors():
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        range_map.pano_to_lidar(torch.ones(1, 4, 6), (-24.9, 2.0), (-180.0, 180.0))


def _reference_stitch(front, back):
    """gaussian_renderer/__init__.py:163,203-225 restated: breaks = (0, w//2, 3w//2, 2w), slice assignments."""
    w = front.shape[-1]
    b = (0, w // 2, 3 * w // 2, w * 2)
    pano = torch.zeros(front.shape[:-1] + (2 * w,), dtype=front.dtype)
    pano[..., b[1]:b[2]] = front
    pano[..., b[2]:b[3]] = back[..., 0:(b[3] - b[2])]
    pano[..., b[0]:b[1]] = back[..., (w - b[1] + b[0]):w]
    return pano


@pytest.mark.parametrize("w", [515, 258, 7])
def test_stitching_equals_the_reference_slice_assignments(w):
    g = torch.Generator().manual_seed(w)
    front, back = torch.rand(3, 5, w, generator=g), torch.rand(3, 5, w, generator=g)
    assert torch.equal(range_map.stitch_half_panoramas(front, back), _reference_stitch(front, back))


def _fake_render(cam, gaussians, *args, env_map=None):
    g = torch.Generator().manual_seed(cam.colmap_id)
    h, w = cam.image_height, cam.image_width
    depth = 1 + 5 * torch.rand(1, h, w, generator=g)
    return dict(depth=depth, alpha=0.2 + 0.8 * torch.rand(1, h, w, generator=g), raydrop=torch.rand(1, h, w, generator=g),
                depth_square=depth ** 2 + 0.3 * torch.rand(1, h, w, generator=g), depth_median=depth + 0.1,
                intensity_sh=torch.rand(1, h, w, generator=g))


def _cam(cid, towards, h, w, hfov=(-90.0, 90.0)):
    g = torch.Generator().manual_seed(100 + cid)
    return SimpleNamespace(colmap_id=cid, towards=towards, image_height=h, image_width=w, hfov=hfov,
                           pts_depth=torch.rand(1, h, w, generator=g), pts_intensity=torch.rand(1, h, w, generator=g))


@pytest.mark.parametrize("sky,mode", [(False, 0), (True, 0), (True, 1)])
def test_render_range_map_matches_the_reference_recipe(sky, mode):
    h, w = 6, 11
    args = SimpleNamespace(frames=50, sky_depth=sky, depth_blend_mode=mode)
    front, back = _cam(3, "forward", h, w), _cam(53, "backward", h, w)
    got = range_map.render_range_map(args, front, back, None, _fake_render, (), None, [h, w])
    # the reference recipe (:168-227), restated per view
    halves = []
    for cam in (front, back):
        pkg = _fake_render(cam, None)
        depth, alpha = pkg["depth"], pkg["alpha"]
        var = pkg["depth_square"] - depth ** 2
        q = var.median() * 10
        mix = torch.zeros_like(depth)
        mix[var > q] = pkg["depth_median"][var > q]
        mix[var <= q] = depth[var <= q]
        d = torch.cat([mix, depth, pkg["depth_median"]])
        if sky:
            d = d / alpha.clamp_min(1e-5)
            d = 1 / (alpha / d.clamp_min(1e-5) + (1 - alpha) / 900).clamp_min(1e-5) if mode == 0 else alpha * d + (1 - alpha) * 900
        halves.append((d, pkg["intensity_sh"], pkg["raydrop"], cam.pts_depth, cam.pts_intensity))
    want = [_reference_stitch(f, b) for f, b in zip(*halves)]
    assert len(got) == 5 and got[0].shape == (3, h, 2 * w)
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    with pytest.raises(AssertionError):
        range_map.render_range_map(args, back, front, None, _fake_render, (), None, [h, w])


def test_single_pass_360_sets_and_restores_the_wrap_mode():
    import gs_lidar_b200.diff_gaussian_rasterization_2d as G
    seen = []

    def render(cam, gaussians, *a, env_map=None):
        seen.append(G._wrap_azimuth)
        return _fake_render(cam, gaussians)

    args = SimpleNamespace(frames=50, sky_depth=False)
    cam = _cam(3, "forward", 6, 22, hfov=(-180.0, 180.0))
    assert G._wrap_azimuth is False
    depth, intensity, raydrop = range_map.render_range_map_360(args, cam, None, render, (), None)
    assert seen == [True] and G._wrap_azimuth is False
    assert depth.shape == (3, 6, 22) and intensity.shape == (1, 6, 22) and raydrop.shape == (1, 6, 22)
    with pytest.raises(AssertionError, match="360-degree"):
        range_map.render_range_map_360(args, _cam(3, "forward", 6, 11), None, render, (), None)
