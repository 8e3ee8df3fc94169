"""CPU: tests/glue_oracle.py -- the PyTorch restatement the fused glue kernels are checked against on the GPU -- pinned
against the reference's OWN GaussianModel accessors (scene/gaussian_model.py:139-186): against stored outputs
(tests/golden/glue_getters.npz, written by tests/golden/make_glue_golden.py from the unmodified class) and, where the
reference's Python files are staged (oracle/_ref/py), against the live class."""
import os
import types
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import glue_oracle as GO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "glue_getters.npz")


def _model_from(g):
    pc = SimpleNamespace()
    for n in GO.RAW + ("_features_dc", "_features_rest"):
        setattr(pc, n, torch.from_numpy(g[n]))
    pc.T, pc.velocity_decay, pc.active_sh_degree = float(g["T"]), float(g["velocity_decay"]), 3
    return pc


def _close(a, b, tol=2e-6):
    a, b = a.double(), torch.as_tensor(b).double()
    assert a.shape == b.shape
    assert float((a - b).abs().max() / (b.abs().max() + 1e-30)) < tol


def test_restated_glue_equals_the_reference_accessors_stored():
    g = dict(np.load(GOLD))
    pc = _model_from(g)
    ts, shift = float(g["timestamp"]), float(g["time_shift"])
    means3D, opacity, scales, rotations, marginal_t, _ = GO.reference_glue(pc, ts, None, dynamic=False)
    _close(means3D, g["get_xyz_SHM"])
    _close(opacity, g["get_opacity"])
    _close(scales, g["get_scaling"])
    _close(rotations, g["get_rotation"])
    _close(marginal_t, g["get_marginal_t"])
    _close(GO.get_features(pc), g["get_features"], tol=0.0 + 1e-12)
    # time-shifted branch of render() (gaussian_renderer/__init__.py:69-75) and the dynamic opacity (:77-79)
    means3D_s, opacity_s, _, _, marginal_s, _ = GO.reference_glue(pc, ts, shift, dynamic=True)
    _close(means3D_s, torch.from_numpy(g["get_xyz_SHM_shifted"]) + torch.from_numpy(g["get_inst_velocity"]) * shift)
    _close(marginal_s, g["get_marginal_t_shifted"])
    _close(opacity_s, torch.from_numpy(g["get_opacity"]) * torch.from_numpy(g["get_marginal_t_shifted"]))


def test_restated_glue_equals_the_live_reference_class():
    from oracle import ref_python
    if not ref_python.available():
        pytest.skip("reference Python files not staged (oracle/_ref/py)")
    dummy = types.ModuleType("no_rasterizer")
    dummy.GaussianRasterizationSettings = dummy.GaussianRasterizer = object
    ns = ref_python.load(dummy)
    raw = GO.make_model(1501, seed=5)
    args = SimpleNamespace(sh_degree=3, time_duration=[-0.5, 0.5], no_time_split=True, t_grad=True, contract=False, t_init=0.1,
                           big_point_threshold=0.1, cycle=raw.T, velocity_decay=raw.velocity_decay, random_init_point=0)
    pc = ns.GaussianModel(args)
    for n in GO.RAW + ("_features_dc", "_features_rest"):
        setattr(pc, n, getattr(raw, n))
    for ts, shift, dyn in ((0.07, None, True), (-0.2, 0.05, True), (0.3, -0.04, False)):
        means3D, opacity, scales, rotations, marginal_t, mask = GO.reference_glue(raw, ts, shift, dynamic=dyn)
        t_eff = ts if shift is None else ts - shift
        want_xyz = pc.get_xyz_SHM(t_eff) + (pc.get_inst_velocity * shift if shift is not None else 0.0)
        want_marg = pc.get_marginal_t(t_eff)
        want_opa = pc.get_opacity * want_marg if dyn else pc.get_opacity
        _close(means3D.detach(), want_xyz.detach())
        _close(marginal_t.detach(), want_marg.detach())
        _close(opacity.detach(), want_opa.detach())
        _close(scales.detach(), pc.get_scaling.detach())
        _close(rotations.detach(), pc.get_rotation.detach())
        want_mask = want_opa[:, 0] > 1 / 255
        if dyn:
            want_mask = want_mask & (want_marg[:, 0] > 0.05)
        assert torch.equal(mask, want_mask)
