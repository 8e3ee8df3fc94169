"""GPU: the CUDA panorama post-ops (gs_lidar_b200.range_map.pano_to_lidar / depth_to_normal / pano_post_ops,
csrc/gsl_postops.cu) against (1) stored outputs of the reference's own functions (tests/golden/postop_*.npz), (2) the
live reference functions on the GPU (utils/graphics_utils.py:96-149, unmodified, staged under oracle/_ref/py), values and
gradients, and (3) the PyTorch restatement tests/postop_oracle.py for the gradients where the reference is not staged."""
import glob
import os
import types

import numpy as np
import pytest
import torch

import postop_oracle
from gs_lidar_b200 import range_map
from oracle import ref_python

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "postop_*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_cuda_post_ops_match_the_stored_reference_outputs(path):
    g = np.load(path)
    rng = torch.from_numpy(g["range_image"]).cuda()
    vfov, hfov = tuple(g["vfov"].tolist()), tuple(g["hfov"].tolist())
    pts = range_map.pano_to_lidar(rng, vfov, hfov)
    nrm = range_map.depth_to_normal(rng, vfov, hfov)
    both = range_map.pano_post_ops(rng, vfov, hfov)
    assert pts.shape == g["points"].shape and nrm.shape == g["normals"].shape
    np.testing.assert_allclose(pts.cpu().numpy(), g["points"], rtol=0, atol=2e-6)
    # a normal is a normalised cross product of differences of neighbouring points: where two neighbours are almost at the
    # same range the cross product is a cancellation residue, so compare against the scale of the un-normalised vector
    err = (nrm.cpu() - torch.from_numpy(g["normals"])).abs().amax(dim=0)
    assert float(err.median()) < 1e-6 and float((err > 1e-3).float().mean()) < 2e-3, (float(err.median()), float(err.max()))
    assert torch.equal(both[0], pts) and torch.equal(both[1], nrm)
    assert float(nrm[:, 0].abs().sum()) == 0 and float(nrm[:, :, -1].abs().sum()) == 0  # zero border


def _grads(fn_points, fn_normals, rng, vfov, hfov, seed=3):
    r = rng.clone().requires_grad_(True)
    pts, nrm = fn_points(r, vfov, hfov), fn_normals(r, vfov, hfov)
    g = torch.Generator().manual_seed(seed)
    loss = (pts * torch.randn(pts.shape, generator=g).to(r.device)).sum() + (nrm * torch.randn(nrm.shape, generator=g).to(r.device)).sum()
    loss.backward()
    return pts.detach(), nrm.detach(), r.grad.detach()


@pytest.mark.parametrize("h,w,vfov,hfov,holes", [(66, 1030, (-24.9, 2.0), (-180.0, 180.0), 0.3), (128, 2048, (-25.0, 2.0), (-180.0, 180.0), 0.0),
                                                 (66, 515, (-24.9, 2.0), (-90.0, 90.0), 0.7), (5, 7, (-30.0, 10.0), (-40.0, 50.0), 0.5)])
def test_cuda_post_ops_values_and_gradients(h, w, vfov, hfov, holes):
    g = torch.Generator().manual_seed(h * w)
    rng = (0.3 + 8 * torch.rand(1, h, w, generator=g))
    rng = (rng * (torch.rand(1, h, w, generator=g) >= holes)).cuda()  # holes: pixels without a return (range 0)
    checkers = [("restatement", postop_oracle.pano_to_lidar, postop_oracle.depth_to_normal)]
    if ref_python.available():
        dummy = types.ModuleType("no_rasterizer")
        dummy.GaussianRasterizationSettings = dummy.GaussianRasterizer = object
        gu = ref_python.load(dummy).graphics_utils
        checkers.append(("reference functions", gu.pano_to_lidar, gu.depth_to_normal))
    pts, nrm, grad = _grads(range_map.pano_to_lidar, range_map.depth_to_normal, rng, vfov, hfov)
    for name, f_pts, f_nrm in checkers:
        r_pts, r_nrm, r_grad = _grads(f_pts, f_nrm, rng, vfov, hfov)
        assert pts.shape == r_pts.shape, name
        # a few ulp of the range (the azimuth reaches pi: one ulp of it moves sin / cos by 2.4e-7)
        tol_pts = 2e-6 * float(rng.max())
        assert float((pts - r_pts).abs().max()) < tol_pts, (name, float((pts - r_pts).abs().max()))
        # normals are normalised cross products of DIFFERENCES of neighbouring points (cancellation: the finer the grid the
        # closer the neighbours), so the bulk is compared, not the maximum
        err = (nrm - r_nrm).abs().amax(dim=0)
        assert float(err.median()) < 2e-5 and float((err > 1e-2).float().mean()) < 2e-3, (name, float(err.median()), float(err.max()))
        gerr = float((grad.double() - r_grad.double()).norm() / (r_grad.double().norm() + 1e-30))
        assert gerr < 1e-4, (name, gerr)


def test_empty_and_all_zero_images():
    z = torch.zeros(1, 8, 9, device="cuda", requires_grad=True)
    pts = range_map.pano_to_lidar(z, (-24.9, 2.0), (-180.0, 180.0))
    assert pts.shape == (0, 3)
    nrm = range_map.depth_to_normal(z, (-24.9, 2.0), (-180.0, 180.0))
    assert float(nrm.abs().sum()) == 0
    (nrm.sum() + pts.sum()).backward()
    assert torch.isfinite(z.grad).all()
