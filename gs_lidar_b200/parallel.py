"""Frame-parallel (data-parallel) helpers: one process per GPU, frames sharded across ranks, surfel
gradients summed with an all-reduce (NCCL over NVLink on GPUs, gloo on CPU for the tests).

The reference has no distributed code at all (SURVEY.md section 2.1 "Parallelism strategies": none);
this module is the natural sharding BASELINE.json's north_star names: a rasterizer call depends only
on (replicated surfel parameters, one camera), so frames are independent and the only exchange step
is the sum of per-parameter gradients before the optimizer step (where train.py:325 -> :375 sits).
No collective is used on the inference path (frames are simply sharded).
"""
from typing import Dict, List, Sequence

import torch
import torch.distributed as dist


def shard_frames(num_frames: int, rank: int, world_size: int, equal_steps: bool = False) -> List[int]:
    """Frame ids rendered by `rank`: round-robin, frame_id % world_size == rank.

    Inference needs nothing more (no collective).  TRAINING with a gradient exchange (GradientExchange / PeerExchange)
    requires every rank to run the SAME number of backward passes -- each one is a barrier among the ranks -- so pass
    equal_steps=True there: the last num_frames % world_size frames are dropped and all shards have equal length
    (with an uneven shard the ranks that finish early leave the others waiting at a barrier until its time-out)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size %d/%d" % (rank, world_size))
    if equal_steps:
        num_frames -= num_frames % world_size
    return list(range(rank, num_frames, world_size))


class GradBucket:
    """One flat fp32 buffer holding all per-surfel gradients, all-reduced in a single call.

    Layout: tensors are packed back to back in the order given at construction; `views` alias the
    flat buffer so the rasterizer's backward outputs can be copied (or accumulated) in place.
    """

    def __init__(self, shapes: Dict[str, Sequence[int]], device):
        self.names = list(shapes.keys())
        self.shapes = {k: tuple(int(x) for x in v) for k, v in shapes.items()}
        sizes = [int(torch.Size(self.shapes[k]).numel()) for k in self.names]
        self.offsets = {}
        off = 0
        for k, n in zip(self.names, sizes):
            self.offsets[k] = (off, n)
            off += n
        self.flat = torch.zeros(off, dtype=torch.float32, device=device)
        self.views = {k: self.flat[o:o + n].view(self.shapes[k]) for k, (o, n) in self.offsets.items()}

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def zero_(self):
        self.flat.zero_()

    def accumulate(self, grads: Dict[str, torch.Tensor]):
        for k, g in grads.items():
            if g is not None and k in self.views:
                self.views[k].add_(g.reshape(self.shapes[k]))

    def load(self, grads: Dict[str, torch.Tensor]):
        for k, g in grads.items():
            if g is not None and k in self.views:
                self.views[k].copy_(g.reshape(self.shapes[k]))

    def all_reduce(self, group=None, async_op=False):
        """Sum over ranks (in place).  No-op when torch.distributed is not initialised."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def surfel_grad_shapes(P: int, S: int, M: int) -> Dict[str, Sequence[int]]:
    """Shapes of the gradients the rasterizer returns for its tensor inputs (a3 in SURVEY.md section 8a)."""
    shapes = dict(means3D=(P, 3), means2D=(P, 4), opacities=(P, 1), scales=(P, 3), rotations=(P, 4))
    if S > 0:
        shapes["features"] = (P, S)
    if M > 0:
        shapes["shs"] = (P, M, 4)
    return shapes


class GradientExchange:
    """Frame-parallel gradient exchange fused into the rasterizer's backward pass (SH colour path).

    A dense all-reduce moves 4*(19+S... ) + 16*M bytes per surfel, 80 % of it SH gradient.  The SH gradient one
    frame gives a surfel is the outer product basis(view direction) x dL_dRGB (backward.cu:17-134), so with this
    exchange active the backward pass
      1. extracts the clamp-masked 16-byte factor dL_dRGB right after the backward compositor
         (gsl_backward_composite) and starts its all-gather (camera centre appended) -- it travels while the
         per-surfel backward kernel runs,
      2. writes its non-SH gradients straight into one flat fp32 buffer (gsl_backward_surfels under
         GSL_FLAG_BWD_SH_FACTORED, no per-tensor copies) and all-reduces it,
      3. rebuilds dL_dsh = sum over ranks of basis x dL_dRGB on the device (gsl_sh_expand) while that all-reduce is
         on the wire.
    The gradients the autograd op then returns are already summed over the ranks -- do not all-reduce them
    again.  Same result as an all-reduce of the dense gradients up to fp32 summation order.

        ex = parallel.GradientExchange()            # default process group
        with ex:                                    # or ex.enable() / ex.disable()
            loss.backward()
    """

    # 16-byte-stored tensors first (sizes are multiples of 16 B), scalar-stored ones after: every tensor is aligned
    # for its stores and the buffer has no gaps (one zero-fill, one all-reduce payload)
    NAMES = ("means2D", "rotations", "means3D", "scales", "opacities", "features")

    def __init__(self, group=None):
        self.group = group
        self.key = None

    packed = False  # True: the backward kernels write the exchange's own packed format (PeerExchange)

    def flags(self):
        from . import _lib as L
        return L.GSL_FLAG_BWD_SH_FACTORED

    def run_surfels(self, P, launch):
        launch(0, P)

    # -- communication (overridable: the tests emulate the ranks in one process) ----------------------
    def world_size(self):
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.group)
        return 1

    def _all_reduce(self, flat):
        """Starts the sum of `flat` over the ranks; returns a handle with .wait() (or None when already done)."""
        return dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def _all_gather(self, out, local):
        return dist.all_gather_into_tensor(out, local, group=self.group, async_op=True)

    # -- buffers -------------------------------------------------------------------------------------
    def prepare(self, P, S, M, device):
        G = self.world_size()
        widths = dict(means3D=3, means2D=4, opacities=1, scales=3, rotations=4, features=S)
        starts, off = {}, 0
        for k in self.NAMES:
            starts[k] = off
            off += P * widths[k]
        # a fresh flat buffer per backward: the gradients handed to autograd are views of it and nothing else keeps
        # it alive, so AccumulateGrad adopts them instead of copying (the caching allocator makes this cheap)
        self.flat = torch.empty(off, dtype=torch.float32, device=device)
        self.views = {k: self.flat[starts[k]:starts[k] + P * widths[k]].view(P, widths[k]) for k in self.NAMES}
        key = (P, G, str(device))
        if self.key != key:
            self.stride = 4 * P + 4                       # dL_dRGB (P,4) + camera centre (3) + pad
            self.local = torch.zeros(self.stride, dtype=torch.float32, device=device)
            self.gathered = torch.zeros(G * self.stride, dtype=torch.float32, device=device)
            self.key = key
        return self

    def start_gather(self, P, campos, D=None, M=None, means3D=None):
        """After the backward compositor wrote the SH factor into `local`: append the camera centre and start the
        all-gather, so that it runs while the per-surfel backward kernel is still computing."""
        self.local[4 * P:4 * P + 3].copy_(campos.reshape(3))
        self._h_gather = self._all_gather(self.gathered, self.local)

    def finish(self, P, D, M, means3D):
        """Starts the all-reduce of the non-SH gradients, rebuilds the SH gradient from the gathered factors while it
        is on the wire, and returns the summed gradients (dict, `shs` included)."""
        from . import _lib as L
        import ctypes as C
        G = self.world_size()
        h_reduce = self._all_reduce(self.flat)
        if self._h_gather is not None:
            self._h_gather.wait()
        self._h_gather = None
        campos_all = self.gathered.view(G, self.stride)[:, 4 * P:4 * P + 3].contiguous()
        d_sh = torch.empty((P, M, 4), dtype=torch.float32, device=means3D.device)
        stream = C.c_void_p(torch.cuda.current_stream(means3D.device).cuda_stream)
        L.check(L.load().gsl_sh_expand(P, D, M, G, means3D.data_ptr(), campos_all.data_ptr(), self.gathered.data_ptr(),
                                       self.stride, d_sh.data_ptr(), stream), "gsl_sh_expand")
        if h_reduce is not None:
            h_reduce.wait()
        out = dict(self.views)
        out["shs"] = d_sh
        self.views = None  # drop our references: autograd owns the gradients now
        self.flat_nbytes = self.flat.numel() * 4
        self.flat = None
        return out

    # -- activation ------------------------------------------------------------------------------------
    def enable(self):
        from . import diff_gaussian_rasterization_2d as G
        G._exchange = self
        return self

    def disable(self):
        from . import diff_gaussian_rasterization_2d as G
        if G._exchange is self:
            G._exchange = None

    __enter__ = enable

    def __exit__(self, *exc):
        self.disable()
        return False


class _DeviceMemory:
    """Zero-copy torch view of device memory this package allocated itself (torch.as_tensor reads the interface)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = dict(shape=(int(nbytes),), typestr="|u1", data=(int(ptr), False), version=3)


class PeerExchange(GradientExchange):
    """GradientExchange without a collective library on the data path: the ranks of one NVLink domain (<= 8) map each
    other's exchange buffers (CUDA IPC) and all traffic is remote stores issued by this package's own kernels
    (csrc/gsl_peer.cu, GSL_FLAG_BWD_PEER_ROWS):

      the per-surfel backward kernel runs in `chunks` row ranges; for every 256-surfel tile it pushes the packed 64-byte
      gradient rows of the surfels a pixel touched (+ row bits) to the rank that owns the tile and the non-zero 16-byte
      SH factors to every rank.  Behind every range, on a side stream: barrier -> gsl_peer_reduce (the owner sums its
      tiles in rank order and pushes the sums to every rank) -> gsl_peer_sh_expand (dL_dsh from the local factor
      tables) -- the exchange of range c travels while range c+1 is computed.  Then one barrier and gsl_peer_unpack
      writes the dense gradient tensors.

    Every rank ends up with bit-identical sums.  torch.distributed is used once, at set-up, to exchange the IPC handles.
    Same interface as GradientExchange (`with PeerExchange(): loss.backward()`)."""

    packed = True

    def __init__(self, group=None, sync=True, chunks=1, force=False):
        super().__init__(group)
        self.sync = sync          # False: single-process emulation (tests) -- no barriers, caller orders the pieces
        self.force = force        # run the exchange even with a single rank (tests: the whole path on one GPU)
        self.chunks = max(1, min(int(chunks), 32))
        self.pkey = None
        self.epoch = 0
        self.ctx = None
        self._own = None
        self._opened = []
        self.side = None
        self._glue = None

    def set_glue(self, fold):
        """`fold`: the record renderer.activate_surfels keeps of one render() call (glue parameters + raw tensors), or
        None.  With it the NEXT backward applies the frame-dependent part of the glue's VJP (motion model, marginal) to
        this rank's rows before they are summed (gsl_peer_glue in include/gsl_b200.h) and exchanges dL/dvelocity, dL/dt,
        dL/dscaling_t as 8 more row channels; the rasterizer's backward calls this itself for calls that came through
        gs_lidar_b200.renderer.render().  S <= 4."""
        from . import _lib as L
        if fold is None:
            self._glue = None
            return
        p, raw = fold["p"], fold["raw"]  # raw: xyz, velocity, t, scaling_t, opacity, scaling, rotation
        g = L.gsl_peer_glue()
        g.timestamp, g.time_shift, g.cycle, g.velocity_decay = p.timestamp, p.time_shift, p.cycle, p.velocity_decay
        g.dynamic = p.dynamic
        g.xyz = raw[0].data_ptr()
        g.velocity, g.t, g.scaling_t, g.opacity = raw[1].data_ptr(), raw[2].data_ptr(), raw[3].data_ptr(), raw[4].data_ptr()
        self._glue = (g, raw)  # keeps the struct and the tensors alive

    def rows_channels(self, S):
        """Channels of the exchanged rows' feature part: S, or 4 ceil(S / 4) + 8 with the glue folded in."""
        return S if self._glue is None else 4 * ((S + 3) // 4) + 8

    @staticmethod
    def set_schedule(early_factors=False, expand_low_priority=False):
        """Schedule of the fused step (process-wide, same on every rank; see gsl_peer_set_option in include/gsl_b200.h):
        early_factors -- SH factors pushed from a side stream under the per-surfel kernel instead of by that kernel;
        expand_low_priority -- the SH expansion on a lowest-priority stream.  Off by default (measured slower)."""
        from . import _lib as L
        lib = L.load()
        L.check(lib.gsl_peer_set_option(L.GSL_PEER_OPT_EARLY_FACTORS, int(bool(early_factors))), "gsl_peer_set_option")
        L.check(lib.gsl_peer_set_option(L.GSL_PEER_OPT_EXPAND_LOW_PRIORITY, int(bool(expand_low_priority))),
                "gsl_peer_set_option")

    def rank(self):
        return dist.get_rank(self.group) if (dist.is_available() and dist.is_initialized()) else 0

    def flags(self):
        from . import _lib as L
        return L.GSL_FLAG_BWD_SH_FACTORED | L.GSL_FLAG_BWD_PEER_ROWS

    # -- set-up: allocate, exchange handles, map ------------------------------------------------------------------
    def _exchange_handles(self, handle_bytes):
        if self.world_size() == 1:
            return [handle_bytes]
        out = [None] * self.world_size()
        dist.all_gather_object(out, handle_bytes, group=self.group)
        return out

    def setup(self, P, S, device, buffers=None):
        """Allocates this rank's exchange buffer and maps the peers' (`buffers`: already-mapped device pointers of all
        ranks, for single-process emulation)."""
        from . import _lib as L
        import ctypes as C
        lib = L.load()
        self.close()
        G, r = self.world_size(), self.rank()
        if G > L.GSL_PEER_MAX:
            raise RuntimeError("PeerExchange: %d ranks, at most %d (one NVLink domain)" % (G, L.GSL_PEER_MAX))
        self.nbytes = int(lib.gsl_peer_buffer_bytes(P, S, G))
        ctx = L.gsl_peer_ctx()
        ctx.rank, ctx.world, ctx.epoch = r, G, 0
        with torch.cuda.device(device):
            if buffers is None:
                ptr, handle = C.c_void_p(), L.gsl_peer_handle()
                L.check(lib.gsl_peer_alloc(self.nbytes, C.byref(ptr), C.byref(handle)), "gsl_peer_alloc")
                self._own = ptr.value
                handles = self._exchange_handles(bytes(handle.reserved))
                for g in range(G):
                    if g == r:
                        ctx.buf[g] = self._own
                        continue
                    h = L.gsl_peer_handle()
                    C.memmove(C.byref(h), handles[g], 64)
                    q = C.c_void_p()
                    L.check(lib.gsl_peer_open(C.byref(h), C.byref(q)), "gsl_peer_open (rank %d)" % g)
                    self._opened.append(q.value)
                    ctx.buf[g] = q.value
            else:
                for g in range(G):
                    ctx.buf[g] = buffers[g]
            self._err = torch.zeros(1, dtype=torch.int32).pin_memory()
            ctx.error_flag = self._err.data_ptr()
            self._mem = _DeviceMemory(ctx.buf[r] + L.GSL_PEER_CAMPOS_OFFSET, 16)
            self._campos = torch.as_tensor(self._mem, device=device).view(torch.float32)[:3]
            self.side = torch.cuda.Stream(device=device, priority=-1)
            self._events = [torch.cuda.Event() for _ in range(self.chunks)]
        self.ctx = ctx
        self.ctx_ptr = C.addressof(ctx)
        self.pkey = (P, S, G, str(device))

    def close(self):
        from . import _lib as L
        if self.ctx is None:
            return
        lib = L.load()
        torch.cuda.synchronize()
        for q in self._opened:
            lib.gsl_peer_close(q)
        self._opened = []
        if self._own is not None:
            lib.gsl_peer_free(self._own)
            self._own = None
        self.ctx = None
        self.pkey = None

    # -- per backward ---------------------------------------------------------------------------------------------
    def check(self, synchronize=True):
        """Raises if a rank missed a barrier of a step enqueued so far (its gradients are NaN on the ranks that noticed:
        the kernels that write them poison them on a time-out, so a failed step can never be applied silently).  With
        synchronize=True the current stream is drained first, which makes the check exact for the step just enqueued --
        call it before optimizer.step() when a late error report (at the next backward) is not good enough."""
        if self.ctx is None:
            return
        if synchronize:
            torch.cuda.current_stream(self._campos.device).synchronize()
        if int(self._err[0]) != 0:
            raise RuntimeError("PeerExchange: a rank did not reach a barrier (flag slot %d) within the time-out; the "
                               "gradients of that step are NaN / invalid (every rank must run the same number of steps, "
                               "see shard_frames(equal_steps=True))" % (int(self._err[0]) - 1))

    def prepare(self, P, S, M, device):
        """S: the rasterizer's feature channels; the rows carry rows_channels(S) of them (set_glue)."""
        import ctypes as C
        from . import _lib as L
        G = self.world_size()
        if self._glue is not None and S > 4:
            raise RuntimeError("PeerExchange: the glue's VJP can be folded into the rows for S <= 4 feature channels only")
        S = self.rows_channels(S)
        if self.pkey != (P, S, G, str(device)):
            self.setup(P, S, device)
        self.check(synchronize=False)
        self.epoch += 1
        self.ctx.parity = self.epoch & 1
        self.ctx.glue = C.pointer(self._glue[0]) if self._glue is not None else C.POINTER(L.gsl_peer_glue)()
        self._S = S
        self.views = None
        return self

    @staticmethod
    def _sp(stream):
        import ctypes as C
        return C.c_void_p(stream.cuda_stream)

    def _barrier(self, slot, ticket, stream):
        """Flag slot 0: first range of a step (also pushes the camera centres), 1: later ranges, 2: end of the step;
        tickets grow monotonically per slot."""
        if self.sync:
            from . import _lib as L
            import ctypes as C
            self.ctx.epoch = ticket & 0xFFFFFFFF
            L.check(L.load().gsl_peer_barrier(C.byref(self.ctx), slot, self._sp(stream)), "gsl_peer_barrier")

    def launch_expand(self, P, D, M, means3D, d_sh, row_begin, row_end, stream, sparse=False):
        """sparse=True: d_sh is zero-filled by the caller and only rows with a factor are written (the fused step's kernel)."""
        from . import _lib as L
        import ctypes as C
        fn = L.load().gsl_peer_sh_expand_sparse if sparse else L.load().gsl_peer_sh_expand
        L.check(fn(C.byref(self.ctx), P, self._S, D, M, row_begin, row_end, means3D.data_ptr(), d_sh.data_ptr(),
                   self._sp(stream)), "gsl_peer_sh_expand")

    def launch_reduce(self, P, row_begin, row_end, stream):
        from . import _lib as L
        import ctypes as C
        L.check(L.load().gsl_peer_reduce(C.byref(self.ctx), P, self._S, row_begin, row_end, self._sp(stream)),
                "gsl_peer_reduce")

    def start_gather(self, P, campos, D=None, M=None, means3D=None):
        """Publishes the camera centre (the first barrier of the step pushes it to every rank) and allocates dL_dsh."""
        self._campos.copy_(campos.reshape(3))
        self._d_sh = torch.empty((P, M, 4), dtype=torch.float32, device=means3D.device)
        self._expand_args = (D, M, means3D)

    def ranges(self, P):
        step = ((P + self.chunks - 1) // self.chunks + 255) // 256 * 256
        return [(rb, min(P, rb + step)) for rb in range(0, P, max(step, 256))]

    def run_surfels(self, P, launch):
        """Piecewise path (sync=False, tests / scripts): the per-surfel backward kernel for all rows; the caller drives
        barrier / reduce / expand / unpack itself.  With sync=True the rasterizer's backward makes ONE call instead,
        gsl_backward_surfels_exchange, which runs the whole sequence (row ranges, side stream) inside the library."""
        launch(0, P)

    def alloc_outputs(self, P, S, M, device):
        """Fresh dense gradient tensors (one allocation for the non-SH ones, NAMES order, + dL_dsh).  With the glue folded
        in, `features` is (P, rows_channels(S)): split_glue() separates the exchanged glue gradients from it."""
        S = self.rows_channels(S)
        widths = dict(means3D=3, means2D=4, opacities=1, scales=3, rotations=4, features=S)
        flat = torch.empty(P * (15 + S), dtype=torch.float32, device=device)
        out, off = {}, 0
        for k in self.NAMES:
            out[k] = flat[off:off + P * widths[k]].view(P, widths[k])
            off += P * widths[k]
        out["shs"] = torch.empty((P, M, 4), dtype=torch.float32, device=device)
        self.flat_nbytes = flat.numel() * 4
        return out

    def unpack(self, P):
        """Summed packed rows of the own buffer -> fresh dense gradient tensors (one allocation, NAMES order)."""
        from . import _lib as L
        import ctypes as C
        S = self._S
        dev = self._campos.device
        widths = dict(means3D=3, means2D=4, opacities=1, scales=3, rotations=4, features=S)
        flat = torch.empty(P * (15 + S), dtype=torch.float32, device=dev)
        out, off = {}, 0
        for k in self.NAMES:
            out[k] = flat[off:off + P * widths[k]].view(P, widths[k])
            off += P * widths[k]
        g = L.gsl_bwd_outputs()
        g.dL_dmeans3D, g.dL_dmeans2D, g.dL_dscales = out["means3D"].data_ptr(), out["means2D"].data_ptr(), out["scales"].data_ptr()
        g.dL_drotations, g.dL_dopacity = out["rotations"].data_ptr(), out["opacities"].data_ptr()
        g.dL_dfeatures = out["features"].data_ptr() if S > 0 and P > 0 else None
        L.check(L.load().gsl_peer_unpack(C.byref(self.ctx), P, S, C.byref(g), self._sp(torch.cuda.current_stream(dev))),
                "gsl_peer_unpack")
        self.flat_nbytes = flat.numel() * 4
        return out

    def split_glue(self, out, S):
        """(dL_dfeatures (P,S), glue gradients or None) from the `features` tensor the exchange returned: with the glue
        folded in its columns are [features | padding to whole quads | dL/dvelocity (3), dL/dt, dL/dscaling_t, 0, 0, 0]."""
        f = out["features"]
        if self._glue is None:
            return f, None
        q = 4 * ((S + 3) // 4)
        extras = dict(velocity=f[:, q:q + 3], t=f[:, q + 3:q + 4], scaling_t=f[:, q + 4:q + 5])
        return f[:, :S], extras

    def finish(self, P, D, M, means3D):
        out = self.unpack(P)
        out["shs"] = self._d_sh
        self._d_sh = None
        self._expand_args = None
        return out
