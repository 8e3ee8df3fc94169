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
