"""Frame-parallel gradient exchange over peer memory (csrc/gsl_peer.cu, parallel.PeerExchange): the sum this package's
own kernels form must be what an all-reduce of the dense gradients gives (the reference is single-GPU).
One-GPU tests emulate two ranks in one process (both exchange buffers on the same device, no barriers: the test orders
the pieces); the two-process test needs two GPUs and is skipped otherwise."""
import ctypes as C
import os
import socket

import pytest
import torch

import common
from gs_lidar_b200 import synth

pytestmark = pytest.mark.gpu
TOL_GRAD = 1e-4
NAMES = ("means3D", "means2D", "opacities", "scales", "rotations", "features", "shs")


def _two_frames(P, seed):
    frames = [synth.make_scene(P, seed=seed, view_yaw_deg=y, view_shift=sh)
              for y, sh in ((0.0, (0.0, 0.0, 0.0)), (3.0, (0.2, -0.1, 0.1)))]
    return [frames[0], frames[1]._replace(means3D=frames[0].means3D)]  # replicated surfels, two cameras


def test_barrier_of_one_rank_returns_and_bad_contexts_are_refused():
    from gs_lidar_b200 import _lib as L
    lib = L.load()
    ptr = C.c_void_p()
    n = lib.gsl_peer_buffer_bytes(1000, 4, 1)
    assert lib.gsl_peer_row_width(4) == 16 and lib.gsl_peer_row_width(10) == 24
    assert n >= 4096 + 1024 * 16 + 2 * 1024 * 64 and n % 256 == 0
    # per-source factor tables and staging grow with the number of ranks, the result area does not
    assert lib.gsl_peer_buffer_bytes(1000, 4, 8) > n + 7 * 1024 * 16
    L.check(lib.gsl_peer_alloc(n, C.byref(ptr), None), "gsl_peer_alloc")
    try:
        err = torch.zeros(1, dtype=torch.int32).pin_memory()
        ctx = L.gsl_peer_ctx()
        ctx.rank, ctx.world, ctx.epoch, ctx.error_flag = 0, 1, 7, err.data_ptr()
        ctx.buf[0] = ptr.value
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for phase in range(4):
            assert lib.gsl_peer_barrier(C.byref(ctx), phase, st) == 0, L.last_error()
        torch.cuda.synchronize()
        assert int(err[0]) == 0
        ctx.world = 9
        assert lib.gsl_peer_barrier(C.byref(ctx), 0, st) == L.GSL_EINVAL and b"rank" in lib.gsl_last_error()
        ctx.world = 2
        assert lib.gsl_peer_barrier(C.byref(ctx), 0, st) == L.GSL_EINVAL and b"not mapped" in lib.gsl_last_error()
        ctx.world = 1
        assert lib.gsl_peer_barrier(C.byref(ctx), 5, st) == L.GSL_EINVAL
        assert lib.gsl_peer_reduce(C.byref(ctx), 1000, 4, 100, 1000, st) == L.GSL_EINVAL  # ranges start on a 256-row tile
    finally:
        lib.gsl_peer_free(ptr)


@pytest.mark.parametrize("P,split,sparse", [(20000, None, False), (5003, 2048, False), (20000, None, True), (5003, 2048, True)])
def test_emulated_two_ranks_equal_the_sum_of_dense_gradients(P, split, sparse):
    from gs_lidar_b200 import parallel
    from gs_lidar_b200 import _lib as L
    lib = L.load()
    scenes = [s.to("cuda") for s in _two_frames(P, seed=71)]
    cot = {k: v.cuda() for k, v in synth.make_cotangents(scenes[0].H, scenes[0].W, 4, seed=72).items()}
    dense = [common.run_ours(sc, cot, export=False)[2] for sc in scenes]
    expect = {k: dense[0][k] + dense[1][k] for k in dense[0]}

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
                self.setup(P_, S, device, buffers=bufs)
            return super().prepare(P_, S, M, device)

    nbytes = lib.gsl_peer_buffer_bytes(P, 4, 2)
    bufs = []
    for _ in range(2):
        q = C.c_void_p()
        L.check(lib.gsl_peer_alloc(nbytes, C.byref(q), None), "gsl_peer_alloc")
        bufs.append(q.value)
    try:
        ranks = [Rank(0), Rank(1)]
        for r in (0, 1):
            # the rank's backward pushes its packed rows (+ bits) into the staging area of the tiles' owners and its SH
            # factors into both factor tables; without the exchange steps (sync=False) what it returns is not summed
            with ranks[r]:
                common.run_ours(scenes[r], cot, export=False)
        st = torch.cuda.current_stream()
        sp = C.c_void_p(st.cuda_stream)
        # sparse: the fused step's expansion kernel, which writes only rows with a factor into a zero-filled tensor
        d_sh = [(torch.zeros if sparse else torch.empty)((P, 16, 4), device="cuda") for _ in range(2)]

        def barrier(phase):  # all ranks signal, then all wait: nothing ever spins in this single-stream emulation
            for r in (0, 1):
                L.check(lib.gsl_peer_signal(C.byref(ranks[r].ctx), phase, sp), "gsl_peer_signal")
            for r in (0, 1):
                L.check(lib.gsl_peer_wait(C.byref(ranks[r].ctx), phase, sp), "gsl_peer_wait")

        barrier(0)  # also pushes the camera centres
        for k, (rb, re) in enumerate(((0, P),) if split is None else ((0, split), (split, P))):
            if k > 0:
                barrier(1)
            for r in (0, 1):
                ranks[r].launch_reduce(P, rb, re, st)  # rank r sums the tiles it owns and pushes the sums to both
                ranks[r].launch_expand(P, 3, 16, scenes[0].means3D, d_sh[r], rb, re, st, sparse=sparse)
        barrier(2)
        got = [ranks[r].unpack(P) for r in (0, 1)]
        torch.cuda.synchronize()
        assert all(int(ranks[r]._err[0]) == 0 for r in (0, 1))
        for r in (0, 1):
            got[r]["shs"] = d_sh[r]
        for k in NAMES:
            assert torch.equal(got[0][k], got[1][k]), k          # bit-identical on every rank
            elem, norm = common.grad_err(got[0][k], expect[k])
            assert elem < TOL_GRAD and norm < TOL_GRAD, (k, elem, norm)
    finally:
        torch.cuda.synchronize()
        for q in bufs:
            lib.gsl_peer_free(q)


@pytest.mark.parametrize("P,chunks,early,low", [(20000, 1, True, True), (5003, 3, True, False), (70000, 4, False, False),
                                                (5003, 1, True, True)])
def test_single_rank_exchange_equals_the_plain_backward(P, chunks, early, low):
    """The complete fused path (gsl_backward_surfels_exchange: pushes, barriers, reduce, expand, unpack, zero-fill under
    the compositor, row ranges on the side stream) with one rank: the 'sum' must be the plain backward's gradients."""
    from gs_lidar_b200 import parallel
    scene = synth.make_scene(P, seed=75).to("cuda")
    cot = {k: v.cuda() for k, v in synth.make_cotangents(scene.H, scene.W, 4, seed=76).items()}
    dense = common.run_ours(scene, cot, export=False)[2]
    ex = parallel.PeerExchange(force=True, chunks=chunks)
    parallel.PeerExchange.set_schedule(early_factors=early, expand_low_priority=low)
    try:
        for it in range(3):  # the buffers are reused: stale rows / factors of earlier steps must never leak
            sc = scene if it != 1 else scene._replace(opacities=scene.opacities * 0.5)
            with ex:
                got = common.run_ours(sc, cot, export=False)[2]
        torch.cuda.synchronize()
        assert int(ex._err[0]) == 0
        for k in NAMES:
            elem, norm = common.grad_err(got[k], dense[k])
            assert elem < TOL_GRAD and norm < TOL_GRAD, (k, elem, norm)
        culled = dense["shs"].abs().sum(dim=(1, 2)) == 0
        assert float(got["shs"][culled].abs().sum()) == 0.0  # untouched surfels: exact zeros from the zero-fill
    finally:
        parallel.PeerExchange.set_schedule()
        ex.close()


def test_a_rank_that_never_arrives_gives_nan_gradients_and_an_error_not_a_hang():
    """Fused step of rank 0 of 2 while rank 1 never runs: the in-kernel waits time out (100 ms here), the error flag is
    raised, every gradient the step returns is NaN (never silently wrong) and check() / the next backward raise."""
    from gs_lidar_b200 import parallel
    from gs_lidar_b200 import _lib as L
    lib = L.load()
    P = 5003
    scene = synth.make_scene(P, seed=77).to("cuda")
    cot = {k: v.cuda() for k, v in synth.make_cotangents(scene.H, scene.W, 4, seed=78).items()}
    nbytes = lib.gsl_peer_buffer_bytes(P, 4, 2)
    bufs = []
    for _ in range(2):
        q = C.c_void_p()
        L.check(lib.gsl_peer_alloc(nbytes, C.byref(q), None), "gsl_peer_alloc")
        bufs.append(q.value)

    class Lonely(parallel.PeerExchange):
        def world_size(self):
            return 2

        def rank(self):
            return 0

        def prepare(self, P_, S, M, device):
            if self.pkey is None:
                self.setup(P_, S, device, buffers=bufs)
            return super().prepare(P_, S, M, device)

    ex = Lonely()
    lib.gsl_peer_set_timeout_ms(100)
    try:
        with ex:
            got = common.run_ours(scene, cot, export=False)[2]
        torch.cuda.synchronize()
        assert int(ex._err[0]) != 0
        touched = got["opacities"].isnan().any() or got["shs"].isnan().any()
        assert bool(touched)
        for k in ("means3D", "opacities", "scales", "rotations"):
            assert not bool(torch.isfinite(got[k]).all()), k
        with pytest.raises(RuntimeError, match="did not reach a barrier"):
            ex.check()
        with pytest.raises(RuntimeError, match="did not reach a barrier"):
            with ex:
                common.run_ours(scene, cot, export=False)
    finally:
        lib.gsl_peer_set_timeout_ms(20000)
        torch.cuda.synchronize()
        ex.ctx = None  # the buffers are this test's, not the exchange's
        for q in bufs:
            lib.gsl_peer_free(q)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, P, q):
    import torch.distributed as dist
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from gs_lidar_b200 import parallel
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), NCCL_DEBUG="WARN")
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        scene = _two_frames(P, seed=71)[rank % 2].to(dev)
        cot = {k: v.to(dev) for k, v in synth.make_cotangents(scene.H, scene.W, 4, seed=72).items()}
        dense = common.run_ours(scene, cot, export=False)[2]
        expect = {}
        for k in NAMES:
            t = dense[k].clone()
            dist.all_reduce(t)
            expect[k] = t
        ex = parallel.PeerExchange()
        worst = 0.0
        for it in range(3):  # several steps: epochs advance, buffers are reused
            with ex:
                got = common.run_ours(scene, cot, export=False)[2]
            for k in NAMES:
                elem, norm = common.grad_err(got[k], expect[k])
                worst = max(worst, elem, norm)
            # every rank must hold the same bits
            for k in NAMES:
                ref = got[k].clone()
                dist.broadcast(ref, 0)
                if not torch.equal(ref, got[k]):
                    worst = float("inf")
        torch.cuda.synchronize()
        err = int(ex._err[0])
        ex.close()
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, worst, err, None))
    except Exception as e:  # noqa
        import traceback
        q.put((rank, float("inf"), -1, traceback.format_exc()))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_processes_over_nvlink_equal_the_nccl_sum():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 20000, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
    for rank, worst, err, tb in res:
        assert tb is None, tb
        assert err == 0, "barrier time-out on rank %d" % rank
        assert worst < TOL_GRAD, (rank, worst)
