"""GPU: the fused render() glue (gs_lidar_b200.renderer, csrc/gsl_glue.cu) against the PyTorch restatement of the
reference's lines (tests/glue_oracle.py): values to 1e-6 relative, gradients (torch.autograd of the restatement) to 1e-5;
and render() as a whole against the restated glue feeding the same rasterizer."""
import math
from types import SimpleNamespace

import pytest
import torch

import common
import glue_oracle as GO
from gs_lidar_b200 import renderer, synth

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    if a.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


@pytest.mark.parametrize("dynamic,time_shift,use_mask", [(False, None, False), (True, None, True), (True, 0.03, False),
                                                         (False, -0.02, True)])
def test_glue_forward_and_backward_match_torch_restatement(dynamic, time_shift, use_mask):
    P = 20011  # odd on purpose
    pc = GO.make_model(P, seed=3, device="cuda")
    mask = (torch.rand(P, generator=torch.Generator().manual_seed(1)) > 0.2).cuda() if use_mask else None
    ts = 0.11
    ref = GO.reference_glue(pc, ts, time_shift, dynamic, mask)
    got = renderer.activate_surfels(pc, ts, time_shift, dynamic, mask)
    for name, a, b in zip(("means3D", "opacity", "scales", "rotations", "marginal_t"), got, ref):
        assert rel(a, b) < 2e-6, name
    # the mask may differ only where opacity / marginal sit within rounding of the thresholds
    diff = got[5] != ref[5]
    near = ((ref[1][:, 0] - 1 / 255).abs() < 1e-6) | ((ref[4][:, 0] - 0.05).abs() < 1e-6)
    assert bool((~diff | near).all())
    # gradients: random cotangents on the four differentiable outputs
    g = torch.Generator().manual_seed(7)
    cots = [torch.randn(x.shape, generator=g).cuda() for x in ref[:4]]
    loss_ref = sum((x * c).sum() for x, c in zip(ref[:4], cots))
    g_ref = torch.autograd.grad(loss_ref, [getattr(pc, n) for n in GO.RAW], allow_unused=True)
    loss_got = sum((x * c).sum() for x, c in zip(got[:4], cots))
    g_got = torch.autograd.grad(loss_got, [getattr(pc, n) for n in GO.RAW])
    for n, a, b in zip(GO.RAW, g_got, g_ref):
        if b is None:  # parameter unused by the reference graph in this mode: the fused op must return zeros
            assert float(a.abs().max()) == 0.0, n
        else:
            assert rel(a, b) < 1e-5, n


def _camera(scene, timestamp=0.05):
    return SimpleNamespace(image_height=scene.H, image_width=scene.W, world_view_transform=scene.viewmatrix,
                           full_proj_transform=scene.projmatrix, camera_center=scene.campos, vfov=scene.vfov,
                           hfov=scene.hfov, timestamp=timestamp, towards="forward", FoVx=1.0, FoVy=1.0)


def test_render_matches_reference_glue_plus_rasterizer():
    """render() end to end: same dict as gaussian_renderer/__init__.py:render(), checked against the restated glue
    feeding this package's rasterizer (which is itself checked against the reference CUDA in test_parity_gpu.py)."""
    from gs_lidar_b200 import GaussianRasterizationSettings, GaussianRasterizer
    P = 30000
    scene = synth.make_scene(P, seed=5).to("cuda")
    pc = GO.make_model(P, seed=6, device="cuda")
    with torch.no_grad():  # place the model's surfels where the synthetic scene has them
        pc._xyz.copy_(scene.means3D)
        pc._scaling.copy_(scene.scales.log())
        pc._velocity.mul_(0.2)
    pc.get_features = GO.get_features(pc)
    pipe = SimpleNamespace(neg_fov=True, debug=False, scale_factor=scene.scale_factor, dynamic=True, median_depth=False,
                           compute_cov3D_python=False, convert_SHs_python=False)
    cam = _camera(scene)
    t_scale = torch.exp(pc._scaling_t).detach()
    other = [t_scale, pc._velocity.detach()]
    prior = torch.sigmoid(torch.randn(1, scene.H, scene.W, generator=torch.Generator().manual_seed(2))).cuda()
    pkg = renderer.render(cam, pc, pipe, scene.bg, env_map=lambda towards: prior, other=other, time_shift=0.01)
    # restated reference path
    m3, op, sc, rot, mt, msk = GO.reference_glue(pc, cam.timestamp, 0.01, True, None)
    st = GaussianRasterizationSettings(scene.H, scene.W, math.tan(-0.5), math.tan(-0.5), scene.bg, 1.0, scene.viewmatrix,
                                       scene.projmatrix, 3, scene.campos, False, False, scene.vfov, scene.hfov, scene.scale_factor)
    sp = torch.zeros((P, 4), device="cuda", requires_grad=True)
    contrib, img, feat, depth, alpha, radii = GaussianRasterizer(st)(
        means3D=m3, means2D=sp, shs=pc.get_features, features=torch.cat(other, 1), opacities=op, scales=sc, rotations=rot,
        mask=msk)
    # The fused glue agrees with the torch restatement to ~1 ulp (previous test); the rasterizer turns an ulp on a
    # position into ~1e-5 .. 1e-4 on a pixel (and may flip a threshold for an isolated surfel), hence 1e-3 here.
    TOL = 1e-3
    assert float((pkg["radii"] != radii).double().mean()) < 1e-3
    assert float((pkg["contrib"] != contrib).double().mean()) < 1e-2
    assert rel(pkg["intensity_sh"], img[2:3]) < TOL and rel(pkg["depth_mean"], depth[0:1]) < TOL
    assert rel(pkg["alpha"], alpha) < TOL and rel(pkg["feature"], feat[:4]) < TOL
    nrm = feat[4:] / (feat[4:].norm(dim=0, keepdim=True) + 1e-8)
    assert float((pkg["normal"] - nrm).abs().median()) < 1e-5
    rd = (prior + (1 - prior) * img[3:4]).clamp(0, 1)
    assert rel(pkg["raydrop"], rd) < TOL
    assert set(pkg) == {"viewspace_points", "visibility_filter", "radii", "contrib", "depth", "depth_mean", "depth_median",
                        "distortion", "depth_square", "alpha", "feature", "normal", "intensity_sh", "raydrop"}
    # gradients w.r.t. the raw parameters through render()
    w = torch.randn(pkg["depth"].shape, generator=torch.Generator().manual_seed(3)).cuda()
    loss = (pkg["depth"] * w).sum() + pkg["intensity_sh"].sum() + (pkg["alpha"] * w).sum()
    leaves = [getattr(pc, n) for n in GO.RAW] + [pc._features_dc, pc._features_rest]  # render() takes dc/rest unconcatenated
    g_got = torch.autograd.grad(loss, leaves, retain_graph=False)
    loss_ref = (depth[0:1] * w).sum() + img[2:3].sum() + (alpha * w).sum()
    g_ref = torch.autograd.grad(loss_ref, leaves, allow_unused=True)
    for a, b in zip(g_got, g_ref):
        if b is None:
            continue
        elem, norm = common.grad_err(a, b)
        assert norm < 1e-3, (norm, elem)


def test_render_rejects_dead_reference_paths():
    scene = synth.make_scene(100, seed=9).to("cuda")
    pc = GO.make_model(100, seed=9, device="cuda")
    pc.get_features = GO.get_features(pc)
    pipe = SimpleNamespace(neg_fov=True, debug=False, scale_factor=0.1, dynamic=False, median_depth=False,
                           compute_cov3D_python=True, convert_SHs_python=False)
    with pytest.raises(RuntimeError, match="dead paths"):
        renderer.render(_camera(scene), pc, pipe, scene.bg)
