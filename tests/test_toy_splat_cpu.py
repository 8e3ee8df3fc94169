"""oracle/toy_splat.py (the CPU restatement of the reference's pure-PyTorch panorama splatting, BASELINE.json configs[0])
against outputs of the reference function itself (tests/golden/toy_*.npz, made by tests/golden/make_toy_golden.py)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import toy_splat

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "toy_*.npz")))


def test_fixtures_exist():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_restatement_matches_the_reference_function(path):
    g = np.load(path)
    t = lambda k: torch.from_numpy(g[k])
    image, depth, centre, radii = toy_splat.surface_splatting(t("means3D"), t("scales"), t("quats"), t("colors"), t("opacities"),
                                                              t("intrins"), t("viewmat"), pixel_chunk=1000)
    assert image.shape == g["image"].shape and depth.shape == g["depth"].shape
    # same operations in the same order on the same float32 inputs: tolerance covers only the pixel chunking of the
    # reductions (none along the compositing axis) -- 1e-6 absolute on values of order 1
    np.testing.assert_allclose(centre.numpy(), g["centre"], rtol=0, atol=1e-4)      # pixels, values up to ~250
    np.testing.assert_allclose(radii.numpy(), g["radii"], rtol=0, atol=1e-4)
    assert np.array_equal(np.isnan(image.numpy()), np.isnan(g["image"]))
    np.testing.assert_allclose(np.nan_to_num(image.numpy()), np.nan_to_num(g["image"]), rtol=0, atol=2e-6)
    np.testing.assert_allclose(depth.numpy(), g["depth"], rtol=1e-5, atol=1e-5)
    assert float((g["image"].sum(-1) > 0).mean()) > 0.01                            # the fixture renders something


def test_chunking_does_not_change_the_result():
    args = toy_splat.make_inputs(40, 96, 24, seed=3)
    a = toy_splat.surface_splatting(*args, pixel_chunk=96 * 24)
    b = toy_splat.surface_splatting(*args, pixel_chunk=100)
    assert torch.equal(torch.nan_to_num(a[0]), torch.nan_to_num(b[0])) and torch.equal(a[1], b[1])


def test_front_surfel_occludes_the_one_behind():
    """Two coincident-direction surfels, opacity 1: the pixel at their centre shows the nearer one's colour and range."""
    W, H = 64, 32
    means = torch.tensor([[0.0, 0.0, 4.0], [0.0, 0.0, 2.0]])
    scales = torch.tensor([[0.5, 0.5, 0.0], [0.25, 0.25, 0.0]])
    quats = torch.tensor([[1.0, 0.0, 0.0, 0.0]] * 2)
    colors = torch.tensor([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]])
    opac = torch.ones(2, 1)
    intrins = torch.tensor([[1.0, 0.0, W / 2], [0.0, 1.0, H / 2], [0.0, 0.0, 1.0]])
    image, depth, centre, radii = toy_splat.surface_splatting(means, scales, quats, colors, opac, intrins, torch.eye(4))
    px = image[H // 2, W // 2]
    assert px[1] > 0.99 and px[0] < 1e-3
    assert abs(float(depth[H // 2, W // 2]) - 2.0) < 1e-3
    assert torch.allclose(centre[:, :2], torch.tensor([[W / 2, H / 2]] * 2), atol=1e-4)   # sorted: the near one first
