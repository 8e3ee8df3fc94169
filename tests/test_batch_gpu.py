"""GPU: gs_lidar_b200.batch.render_frames (frames of one surfel set on alternating streams with per-stream workspaces)
returns, frame by frame, exactly what a plain GaussianRasterizer call returns for that camera."""
import pytest
import torch

from gs_lidar_b200 import GaussianRasterizer, batch, synth

pytestmark = pytest.mark.gpu


def _cams(scene, n):
    base = synth.settings_for(scene)
    out = []
    for k in range(n):
        c = synth.make_scene(16, H=scene.H, W=scene.W, seed=0, view_yaw_deg=3.0 * k - 4.0, view_shift=(0.02 * k, 0.0, -0.01 * k)).to("cuda")
        out.append(base._replace(viewmatrix=c.viewmatrix, projmatrix=c.projmatrix, campos=c.campos))
    return out


@pytest.mark.parametrize("streams", [1, 2, 3])
def test_frames_equal_single_calls(streams):
    scene = synth.make_scene(40000, seed=61).to("cuda")
    cams = _cams(scene, 7)
    kw = dict(means3D=scene.means3D, opacities=scene.opacities, shs=scene.shs, features=scene.features, scales=scene.scales,
              rotations=scene.rotations, mask=scene.mask)
    for _ in range(2):  # second round: workspaces come from the per-stream pools
        got = batch.render_frames(cams, streams=streams, **kw)
        torch.cuda.synchronize()
        assert len(got) == len(cams)
        with torch.no_grad():
            for st, g in zip(cams, got):
                want = GaussianRasterizer(st)(means2D=torch.zeros((scene.means3D.shape[0], 4), device="cuda"), **kw)
                for a, b in zip(g, want):
                    assert torch.equal(a, b)


def test_consume_callback_sees_every_frame_on_its_stream():
    scene = synth.make_scene(20000, seed=62).to("cuda")
    cams = _cams(scene, 5)
    sums = torch.zeros(5, device="cuda")

    def consume(i, out):
        sums[i] = out[3][0].sum()  # mean-depth plane

    r = batch.render_frames(cams, consume=consume, means3D=scene.means3D, opacities=scene.opacities, shs=scene.shs,
                            features=scene.features, scales=scene.scales, rotations=scene.rotations, mask=scene.mask)
    torch.cuda.synchronize()
    assert r is None and bool((sums > 0).all())
