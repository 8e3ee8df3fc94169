"""CPU: host-side logic -- synthetic scene generator, settings tuple, frame sharding, gradient bucket,
and the world_size-2 gloo path of the data-parallel gradient exchange."""
import math
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gs_lidar_b200 import parallel, synth
from gs_lidar_b200 import GaussianRasterizationSettings


def test_settings_tuple_matches_reference_fields():
    # gaussian_renderer/diff_gaussian_rasterization_2d.py:194-209
    assert GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
        "sh_degree", "campos", "prefiltered", "debug", "vfov", "hfov", "scale_factor")


def test_scene_is_deterministic_and_shaped_like_render_inputs():
    a, b = synth.make_scene(500, seed=7), synth.make_scene(500, seed=7)
    for x, y in zip(a, b):
        if isinstance(x, torch.Tensor):
            assert torch.equal(x, y)
    assert a.means3D.shape == (500, 3) and a.shs.shape == (500, 16, 4) and a.features.shape == (500, 4)
    assert a.opacities.shape == (500, 1) and a.mask.dtype == torch.bool and a.mask.shape == (500, 1)
    assert a.viewmatrix.shape == (4, 4) and a.bg.tolist() == [0.0, 0.0, 0.0, 1.0]
    c = synth.make_scene(500, seed=8)
    assert not torch.equal(a.means3D, c.means3D)


def test_scene_pose_places_surfels_in_the_panorama():
    s = synth.make_scene(2000, view_yaw_deg=30.0, view_shift=(0.2, 0.0, -0.1))
    V = s.viewmatrix.t()
    pv = s.means3D @ V[:3, :3].t() + V[:3, 3]
    r = pv.norm(dim=1)
    theta = torch.atan2(torch.sqrt(pv[:, 0] ** 2 + pv[:, 2] ** 2), -pv[:, 1])
    lo = math.pi / 2 - math.radians(2.0)
    hi = math.pi / 2 + math.radians(24.9)
    d = hi - lo
    assert float(r.min()) > 0.29 and float(r.max()) < 8.1
    assert float(theta.min()) > lo - 0.06 * d and float(theta.max()) < hi + 0.06 * d
    cam = torch.linalg.inv(V)[:3, 3]
    assert torch.allclose(cam, s.campos, atol=1e-6)


def test_pattern_cotangents_are_exact():
    c = synth.pattern_cotangents(18, 258, 4)
    assert c["color"].shape == (4, 18, 258) and c["feature"].shape == (7, 18, 258)
    assert torch.equal(c["color"] * 8, (c["color"] * 8).round())
    assert float(c["depth"].abs().max()) <= 1.0


def test_shard_frames_partitions_round_robin():
    frames = 51  # KITTI-360 10750-10800
    seen = []
    for r in range(8):
        mine = parallel.shard_frames(frames, r, 8)
        assert all(f % 8 == r for f in mine)
        seen += mine
    assert sorted(seen) == list(range(frames))
    with pytest.raises(ValueError):
        parallel.shard_frames(10, 8, 8)


def test_grad_bucket_layout():
    shapes = parallel.surfel_grad_shapes(100, 4, 16)
    b = parallel.GradBucket(shapes, "cpu")
    assert b.nbytes == 4 * 100 * (3 + 4 + 1 + 3 + 4 + 4 + 64)
    g = {k: torch.randn(v) for k, v in shapes.items()}
    b.load(g)
    for k in shapes:
        assert torch.equal(b.views[k], g[k])
    b.accumulate(g)
    assert torch.allclose(b.views["shs"], 2 * g["shs"])
    assert b.all_reduce() is None  # not initialised -> no-op


def _dp_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P, S, M = 64, 4, 16
        shapes = parallel.surfel_grad_shapes(P, S, M)
        bucket = parallel.GradBucket(shapes, "cpu")
        # every rank "renders" its own frames: the per-frame gradient is a deterministic function of the frame id
        total = torch.zeros_like(bucket.flat)
        frames = parallel.shard_frames(5, rank, world)
        for f in frames:
            g = {k: torch.full(v, float(f + 1)) for k, v in shapes.items()}
            bucket.accumulate(g)
        bucket.all_reduce()
        expect = float(sum(f + 1 for f in range(5)))
        ok = bool(torch.all(bucket.flat == expect))
        ret[rank] = (ok, frames)
    finally:
        dist.destroy_process_group()


def test_data_parallel_gradient_sum_gloo_world2():
    world = 2
    port = 29600 + (os.getpid() % 200)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_dp_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0][0] and ret[1][0]
    assert ret[0][1] == [0, 2, 4] and ret[1][1] == [1, 3]


def test_bench_byte_model_matches_survey():
    import bench
    # P=V=1M, R=2M, N=67,980, S=4, K=M=16 -> ~326 + ~692 + 80 MB
    f, b, r = bench.algorithmic_bytes(10**6, 10**6, 2 * 10**6, 66 * 1030, 4, 16, 16)
    assert abs(f / 1e6 - 326) < 3 and abs(b / 1e6 - 692) < 3 and r == 80 * 10**6


def test_gradient_exchange_buffers_alias_one_flat_allreduce_payload():
    # layout contract of parallel.GradientExchange.prepare (the backward pass writes through these views)
    ex = parallel.GradientExchange()
    assert ex.world_size() == 1
    ex.prepare(100, 4, 16, "cpu")
    widths = dict(means3D=3, means2D=4, opacities=1, scales=3, rotations=4, features=4)
    off = 0
    for k in ex.NAMES:
        v = ex.views[k]
        assert tuple(v.shape) == (100, widths[k]) and v.data_ptr() == ex.flat.data_ptr() + 4 * off
        if k in ("means2D", "rotations"):
            assert (v.data_ptr() - ex.flat.data_ptr()) % 16 == 0  # the kernels store these 16 bytes at a time
        off += 100 * widths[k]
    assert ex.flat.numel() == off  # no gaps: one zero-fill range, one all-reduce payload
    assert ex.local.numel() == ex.stride == 4 * 100 + 4 and ex.gathered.numel() == ex.stride
    gathered_before = ex.gathered
    ex.prepare(100, 4, 16, "cpu")
    assert ex.gathered is gathered_before  # the persistent comm buffers are cached for the same (P, world, device)
    ex.prepare(50, 0, 16, "cpu")
    assert ex.views["features"].shape == (50, 0)


def test_glue_restatement_is_differentiable_and_masks_like_render():
    # tests/glue_oracle.py restates gaussian_renderer/__init__.py:64-115 in torch; sanity of the checker itself on CPU
    import glue_oracle as GO
    pc = GO.make_model(500, seed=1)
    m3, op, sc, rot, mt, msk = GO.reference_glue(pc, 0.1, 0.02, True, None)
    assert m3.shape == (500, 3) and op.shape == (500, 1) and rot.shape == (500, 4) and msk.dtype == torch.bool
    assert torch.allclose(rot.norm(dim=1), torch.ones(500), atol=1e-6)
    assert bool((msk == ((op[:, 0] > 1 / 255) & (mt[:, 0] > 0.05))).all())
    g = torch.autograd.grad(m3.sum() + op.sum(), [pc._xyz, pc._velocity, pc._t, pc._scaling_t, pc._opacity])
    assert all(torch.isfinite(x).all() for x in g)


def test_peer_exchange_row_ranges_cover_the_surfels_on_tile_boundaries():
    from gs_lidar_b200 import parallel
    for P in (1, 255, 256, 5003, 1000000):
        for chunks in (1, 3, 4, 8):
            r = parallel.PeerExchange(chunks=chunks, sync=False).ranges(P)
            assert r[0][0] == 0 and r[-1][1] == P and len(r) <= chunks
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))          # contiguous
            assert all(rb % 256 == 0 for rb, _ in r)                    # ranges start on a 256-surfel tile


def test_peer_exchange_rows_carry_the_glue_gradients_behind_the_features():
    """parallel.PeerExchange with render()'s glue folded in (gsl_peer_glue): the exchange moves
    4 ceil(S/4) + 8 'feature' channels and split_glue() takes them apart again."""
    import ctypes as C
    from gs_lidar_b200 import parallel
    from gs_lidar_b200 import _lib as L
    lib = L.load()
    ex = parallel.PeerExchange(sync=False)
    assert [ex.rows_channels(S) for S in (0, 2, 4)] == [0, 2, 4]
    P = 11
    raw = [torch.zeros(P, w) for w in (3, 3, 1, 1, 1, 3, 4)]  # xyz, velocity, t, scaling_t, opacity, scaling, rotation
    p = L.gsl_glue_params(P, 0.25, 0.05, 0.2, 1.0, 1)
    ex.set_glue(dict(p=p, raw=raw, extras=None))
    g = ex._glue[0]
    assert (g.timestamp, g.time_shift, g.cycle, g.dynamic) == (0.25, C.c_float(0.05).value, C.c_float(0.2).value, 1)
    assert g.xyz == raw[0].data_ptr() and g.velocity == raw[1].data_ptr() and g.opacity == raw[4].data_ptr()
    for S in (0, 2, 4):
        n = ex.rows_channels(S)
        assert n == lib.gsl_peer_rows_channels(S, 1) == 4 * ((S + 3) // 4) + 8
        assert lib.gsl_peer_row_width(n) == 24  # 96-byte rows: three geometry quads + features + two glue quads
        f = torch.arange(P * n, dtype=torch.float32).view(P, n)
        feats, extras = ex.split_glue(dict(features=f), S)
        q = 4 * ((S + 3) // 4)
        assert feats.shape == (P, S) and torch.equal(feats, f[:, :S])
        assert torch.equal(extras["velocity"], f[:, q:q + 3]) and torch.equal(extras["t"], f[:, q + 3:q + 4])
        assert torch.equal(extras["scaling_t"], f[:, q + 4:q + 5])
    ex.set_glue(None)
    assert ex.rows_channels(4) == 4 and ex.split_glue(dict(features=torch.zeros(P, 4)), 4)[1] is None
    # the buffer grows with the row width only in the staging / result areas
    assert lib.gsl_peer_buffer_bytes(100000, 12, 2) > lib.gsl_peer_buffer_bytes(100000, 4, 2)


def test_binning_tile_groups_partition_the_tile_grid():
    """gsl_bin_groups: the counting pass of the binning takes at most 1024 consecutive tile ids at a time;
    the groups must cover every tile exactly once, in ascending tile-id order, as whole rows or pieces of one row."""
    import ctypes as C
    from gs_lidar_b200 import _lib as L
    lib = L.load()
    for W, H in ((1030, 66), (515, 66), (2048, 128), (1040, 272), (2040, 500), (16500, 24), (32767, 40), (17, 32767), (16, 16),
                 (16384, 16), (16385, 33)):
        gx, gy = (W + 15) // 16, (H + 15) // 16
        n = lib.gsl_bin_groups(W, H, None, 0)
        assert n >= 1
        buf = (C.c_int32 * (6 * n))()
        assert lib.gsl_bin_groups(W, H, buf, n) == n
        groups = [tuple(buf[6 * i:6 * i + 6]) for i in range(n)]
        nxt = 0
        for t0, nt, y0, y1, x0, x1 in groups:
            assert t0 == nxt and 1 <= nt <= 1024
            assert 0 <= y0 < y1 <= gy and 0 <= x0 < x1 <= gx
            assert nt == (y1 - y0) * (x1 - x0) and t0 == y0 * gx + x0
            assert (x0 == 0 and x1 == gx) or y1 == y0 + 1  # whole rows, or a piece of ONE row
            nxt = t0 + nt
        assert nxt == gx * gy
        if gx * gy <= 1024:
            assert n == 1
    assert lib.gsl_bin_groups(0, 66, None, 0) == -1 and b"bin_groups" in lib.gsl_last_error()
