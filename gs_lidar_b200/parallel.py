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


def shard_frames(num_frames: int, rank: int, world_size: int) -> List[int]:
    """Frame ids rendered by `rank`: round-robin, frame_id % world_size == rank."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size %d/%d" % (rank, world_size))
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

    def start_gather(self, P, campos):
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
