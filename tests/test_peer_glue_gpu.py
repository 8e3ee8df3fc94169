"""Frame-parallel training of DYNAMIC scenes: render()'s glue sits in front of the rasterizer and the
Jacobian of its motion model depends on the frame's timestamp, so the fused peer-memory exchange applies that part of the
glue's VJP to every rank's rows BEFORE they are summed (gsl_peer_glue, include/gsl_b200.h).  What comes out must be the
sum over the ranks of the raw-parameter gradients each rank's own autograd graph gives (render() without an exchange).
One-GPU tests: the complete fused path with a single rank, and two emulated ranks with different poses AND timestamps."""
import ctypes as C
from types import SimpleNamespace

import pytest
import torch

import common
import glue_oracle as GO
from gs_lidar_b200 import renderer, synth

pytestmark = pytest.mark.gpu
TOL = 1e-4
LEAVES = GO.RAW + ("_features_dc", "_features_rest")


def _camera(scene, timestamp):
    return SimpleNamespace(image_height=scene.H, image_width=scene.W, world_view_transform=scene.viewmatrix,
                           full_proj_transform=scene.projmatrix, camera_center=scene.campos, vfov=scene.vfov,
                           hfov=scene.hfov, timestamp=timestamp, towards="forward", FoVx=1.0, FoVy=1.0)


def _model_on(scene, seed):
    P = scene.means3D.shape[0]
    pc = GO.make_model(P, seed=seed, device="cuda")
    with torch.no_grad():  # the model's surfels where the synthetic scene has them
        pc._xyz.copy_(scene.means3D)
        pc._scaling.copy_(scene.scales.log())
        pc._velocity.mul_(0.2)
    return pc


def _weights(scene, seed=3):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g).cuda()
    return dict(depth=r(1, scene.H, scene.W), alpha=r(1, scene.H, scene.W), intensity_sh=r(1, scene.H, scene.W),
                feature=r(4, scene.H, scene.W), normal=r(3, scene.H, scene.W))


def _step(scene, pc, pipe, timestamp, time_shift, w):
    """One training step's forward + backward through render(); returns the raw-parameter gradients (+ the proxy)."""
    for n in LEAVES:
        getattr(pc, n).grad = None
    other = [torch.exp(pc._scaling_t).detach(), pc._velocity.detach()]  # train.py:167-169
    pkg = renderer.render(_camera(scene, timestamp), pc, pipe, scene.bg, other=other, time_shift=time_shift)
    loss = sum((pkg[k] * w[k]).sum() for k in w)
    loss.backward()
    out = {n: getattr(pc, n).grad.detach().clone() for n in LEAVES}
    out["viewspace_points"] = pkg["viewspace_points"].grad.detach().clone()
    return out


def _pipe(scene, dynamic):
    return SimpleNamespace(neg_fov=True, debug=False, scale_factor=scene.scale_factor, dynamic=dynamic, median_depth=False,
                           compute_cov3D_python=False, convert_SHs_python=False)


@pytest.mark.parametrize("dynamic,time_shift", [(True, None), (True, 0.02), (False, None)])
def test_single_rank_fused_exchange_with_the_glue_folded_in_equals_plain_autograd(dynamic, time_shift):
    from gs_lidar_b200 import parallel
    P = 20000
    scene = synth.make_scene(P, seed=81).to("cuda")
    pc = _model_on(scene, seed=82)
    pipe, w = _pipe(scene, dynamic), _weights(scene)
    want = _step(scene, pc, pipe, 0.07, time_shift, w)
    ex = parallel.PeerExchange(force=True)
    try:
        for it in range(2):  # buffers reused
            with ex:
                got = _step(scene, pc, pipe, 0.07, time_shift, w)
        torch.cuda.synchronize()
        assert int(ex._err[0]) == 0
        assert ex._S == 12  # 4 feature channels + the two quads of glue gradients
        for k in want:
            elem, norm = common.grad_err(got[k], want[k])
            common.report("peer_glue_single_rank", dict(tensor=k, dynamic=dynamic, time_shift=time_shift, elem=elem, norm=norm))
            assert norm < TOL and elem < 10 * TOL, (k, elem, norm)
    finally:
        ex.close()


def test_two_emulated_ranks_with_different_timestamps_equal_the_sum_of_the_ranks_autograd_gradients():
    from gs_lidar_b200 import parallel
    from gs_lidar_b200 import _lib as L
    lib = L.load()
    P = 20000
    frames = [synth.make_scene(P, seed=83, view_yaw_deg=y, view_shift=sh)
              for y, sh in ((0.0, (0.0, 0.0, 0.0)), (3.0, (0.2, -0.1, 0.1)))]
    scenes = [frames[0].to("cuda"), frames[1]._replace(means3D=frames[0].means3D).to("cuda")]
    stamps = (0.03, 0.12)
    pc = _model_on(scenes[0], seed=84)
    pipe, w = _pipe(scenes[0], True), _weights(scenes[0])
    per_rank = [_step(scenes[r], pc, pipe, stamps[r], None, w) for r in (0, 1)]
    want = {k: per_rank[0][k] + per_rank[1][k] for k in per_rank[0]}
    # the time-dependent gradients of the two ranks really differ (else the test would not see a missing fold)
    assert float((per_rank[0]["_velocity"] - per_rank[1]["_velocity"]).abs().max()) > 0

    class Rank(parallel.PeerExchange):
        def __init__(self, r):
            super().__init__(sync=False)
            self.r = r

        def world_size(self):
            return 2

        def rank(self):
            return self.r

        def prepare(self, P_, S, M, device):
            if self.pkey is None:
                self.setup(P_, self.rows_channels(S), device, buffers=bufs)
            return super().prepare(P_, S, M, device)

    nbytes = lib.gsl_peer_buffer_bytes(P, lib.gsl_peer_rows_channels(4, 1), 2)
    bufs = []
    for _ in range(2):
        q = C.c_void_p()
        L.check(lib.gsl_peer_alloc(nbytes, C.byref(q), None), "gsl_peer_alloc")
        bufs.append(q.value)
    try:
        ranks = [Rank(0), Rank(1)]
        for r in (0, 1):
            # pushes this rank's rows (glue folded in, its own timestamp) and SH factors; what backward returns here is
            # not summed yet (sync=False: the test drives the exchange steps below)
            with ranks[r]:
                _step(scenes[r], pc, pipe, stamps[r], None, w)
            assert ranks[r]._S == 12
        st = torch.cuda.current_stream()
        sp = C.c_void_p(st.cuda_stream)
        d_sh = [torch.zeros((P, 16, 4), device="cuda") for _ in range(2)]

        def barrier(phase):
            for r in (0, 1):
                L.check(lib.gsl_peer_signal(C.byref(ranks[r].ctx), phase, sp), "gsl_peer_signal")
            for r in (0, 1):
                L.check(lib.gsl_peer_wait(C.byref(ranks[r].ctx), phase, sp), "gsl_peer_wait")

        barrier(0)  # also pushes the camera centres and the frames' timestamps
        for r in (0, 1):
            ranks[r].launch_reduce(P, 0, P, st)
            # the SH basis of rank g's factor is evaluated where rank g rasterized the surfel (xyz + v coef(its timestamp)):
            # the kernel rebuilds that position from the raw parameters, the means3D argument is ignored
            ranks[r].launch_expand(P, 3, 16, pc._xyz.detach(), d_sh[r], 0, P, st, sparse=True)
        barrier(2)
        got = [ranks[r].unpack(P) for r in (0, 1)]
        torch.cuda.synchronize()
        assert all(int(ranks[r]._err[0]) == 0 for r in (0, 1))
        for k in got[0]:
            assert torch.equal(got[0][k], got[1][k]), k  # bit-identical sums on both ranks
        g = got[0]
        feats, extras = ranks[0].split_glue(g, 4)
        assert feats.shape == (P, 4)
        # what is left of the glue's VJP is frame-independent: sigmoid', exp', normalize' on the SUMS
        os_ = torch.sigmoid(pc._opacity.detach())
        raw = {
            "_xyz": g["means3D"],
            "_velocity": extras["velocity"],
            "_t": extras["t"],
            "_scaling_t": extras["scaling_t"],
            "_opacity": g["opacities"] * os_ * (1 - os_),
            "_scaling": g["scales"] * torch.exp(pc._scaling.detach()),
        }
        q = pc._rotation.detach()
        nrm = q.norm(dim=1, keepdim=True)
        n = q / nrm
        raw["_rotation"] = (g["rotations"] - n * (n * g["rotations"]).sum(1, keepdim=True)) / nrm
        raw["viewspace_points"] = g["means2D"]
        assert torch.equal(d_sh[0], d_sh[1])
        raw["_features_dc"], raw["_features_rest"] = d_sh[0][:, :1], d_sh[0][:, 1:]
        for k, v in raw.items():
            elem, norm = common.grad_err(v, want[k])
            common.report("peer_glue_two_ranks", dict(tensor=k, elem=elem, norm=norm))
            assert norm < TOL and elem < 10 * TOL, (k, elem, norm)
    finally:
        torch.cuda.synchronize()
        for r in ranks:
            r.ctx = None
        for q_ in bufs:
            lib.gsl_peer_free(q_)
