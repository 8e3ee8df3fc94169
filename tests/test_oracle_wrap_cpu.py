"""CPU: the wrap-around mode of the C oracle (oracle/gsl_oracle.c, `wrap` of orc_params) -- the restatement the product's
GSL_FLAG_WRAP_AZIMUTH mode is checked against on the GPU (tests/test_wrap_gpu.py).  Properties that need no GPU:
  * with no surfel near the +-180 degree seam the mode renders what the reference semantics render;
  * seam surfels stop being binned into whole tile rows, and keep a radius like everybody else's;
  * a 180 degree yaw of the sensor shifts the periodic panorama by W/2 columns."""
import math

import numpy as np
import torch

import common
from gs_lidar_b200 import synth


def _phi(scene):
    V = scene.viewmatrix.t()
    pv = scene.means3D @ V[:3, :3].t() + V[:3, 3]
    return torch.atan2(pv[:, 0], pv[:, 2])


def test_wrap_off_the_seam_equals_the_reference_semantics():
    scene = synth.make_scene(3000, H=34, W=258, seed=141, footprint_px=1.5)
    keep = (_phi(scene).abs() < math.radians(160.0)).view(-1, 1)
    scene = scene._replace(mask=scene.mask & keep)
    a, _ = common.run_oracle(scene, None, wrap=False)
    b, _ = common.run_oracle(scene, None, wrap=True)
    assert float((a["radii"] != b["radii"]).mean()) < 2e-3
    for k in ("out_color", "out_depth", "out_feature"):
        assert np.abs(a[k] - b[k]).max() <= 1e-4 * max(np.abs(a[k]).max(), 1e-9), k


def test_wrap_keeps_seam_splats_small_and_shifts_under_a_half_turn():
    scene = synth.make_scene(3000, H=34, W=256, seed=142, footprint_px=2.0)
    near_seam = (_phi(scene).abs() > math.radians(178.5)).view(-1)
    assert int(near_seam.sum()) > 5
    ref, _ = common.run_oracle(scene, None, wrap=False)
    wrp, _ = common.run_oracle(scene, None, wrap=True)
    vis = (ref["radii"] > 0) & (wrp["radii"] > 0) & near_seam.numpy()
    # reference semantics: a seam splat's AABB spans the picture; wrap mode: a few pixels like everybody else's
    assert np.median(ref["radii"][vis]) > 100 and np.median(wrp["radii"][vis]) < 20
    assert wrp["R"] < ref["R"]
    # half turn: same position, sensor yawed by exactly 180 degrees -> the picture shifts by W/2 columns
    flip = torch.diag(torch.tensor([-1.0, 1.0, -1.0, 1.0]))
    vm_b = (flip @ scene.viewmatrix.t()).t().contiguous()
    turned, _ = common.run_oracle(scene._replace(viewmatrix=vm_b, projmatrix=vm_b), None, wrap=True)
    W = scene.W
    for k in ("out_depth", "out_color"):
        a, b = np.roll(wrp[k], W // 2, axis=-1), turned[k]
        err = np.abs(a - b) / (np.abs(a) + 1e-2 * np.abs(a).max() + 1e-12)
        # (the 16-pixel tile grid does not shift with the picture when W/2 is not a multiple of 16: here it is)
        assert np.median(err) < 1e-5 and (err > 1e-3).mean() < 5e-3, (k, np.median(err), err.max())
