"""Chamfer distance of two 3-D point sets: drop-in for the reference's chamfer/chamfer3D/dist_chamfer_3D.py
(`chamfer_3DFunction` :40-85, `chamfer_3DDist` :88-95) and chamfer/fscore.py, over this package's C-ABI
(`gsl_chamfer_forward` / `gsl_chamfer_backward`, csrc/gsl_chamfer.cu).  SURVEY.md 8f next-4: the only other native CUDA on
GS-LiDAR's training path (train.py:256-267), a brute-force nearest-neighbour search over two ~34k-point sweeps.

    from gs_lidar_b200.chamfer import chamfer_3DDist
    dist1, dist2, idx1, idx2 = chamfer_3DDist()(pred_lidar[None], gt_lidar[None])   # squared distances

CUDA tensors only (no CPU fallback), float32, shapes (B, N, 3) and (B, M, 3).
"""
import ctypes as C

import torch
from torch import nn
from torch.autograd import Function

from . import _lib as L


def _check(xyz, name):
    if xyz.dim() != 3 or xyz.shape[2] != 3:
        raise AssertionError("Wrong last dimension for the chamfer distance 's input! Check with .size()")
    if not xyz.is_cuda:
        raise RuntimeError("gs_lidar_b200.chamfer runs on CUDA tensors only (no CPU fallback): %s" % name)


class chamfer_3DFunction(Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        _check(xyz1, "xyz1")
        _check(xyz2, "xyz2")
        if xyz1.shape[0] != xyz2.shape[0]:
            raise RuntimeError("chamfer: batch sizes differ (%d vs %d)" % (xyz1.shape[0], xyz2.shape[0]))
        lib = L.load()
        a, b = xyz1.detach().float().contiguous(), xyz2.detach().float().contiguous()
        B, n, m = a.shape[0], a.shape[1], b.shape[1]
        dev = a.device
        with torch.cuda.device(dev):
            dist1 = torch.zeros((B, n), dtype=torch.float32, device=dev)
            dist2 = torch.zeros((B, m), dtype=torch.float32, device=dev)
            idx1 = torch.zeros((B, n), dtype=torch.int32, device=dev)
            idx2 = torch.zeros((B, m), dtype=torch.int32, device=dev)
            scratch = torch.empty(max(int(lib.gsl_chamfer_scratch_bytes(B, n, m)), 8), dtype=torch.uint8, device=dev)
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            L.check(lib.gsl_chamfer_forward(B, n, a.data_ptr(), m, b.data_ptr(), dist1.data_ptr(), idx1.data_ptr(),
                                            dist2.data_ptr(), idx2.data_ptr(), scratch.data_ptr(), st), "gsl_chamfer_forward")
        ctx.save_for_backward(a, b, idx1, idx2)
        ctx.mark_non_differentiable(idx1, idx2)
        return dist1, dist2, idx1, idx2

    @staticmethod
    def backward(ctx, graddist1, graddist2, gradidx1, gradidx2):
        a, b, idx1, idx2 = ctx.saved_tensors
        lib = L.load()
        B, n, m = a.shape[0], a.shape[1], b.shape[1]
        dev = a.device
        g1 = (torch.zeros((B, n), device=dev) if graddist1 is None else graddist1).float().contiguous()
        g2 = (torch.zeros((B, m), device=dev) if graddist2 is None else graddist2).float().contiguous()
        with torch.cuda.device(dev):
            gxyz1, gxyz2 = torch.empty_like(a), torch.empty_like(b)
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            L.check(lib.gsl_chamfer_backward(B, n, a.data_ptr(), m, b.data_ptr(), g1.data_ptr(), idx1.data_ptr(), g2.data_ptr(),
                                             idx2.data_ptr(), gxyz1.data_ptr(), gxyz2.data_ptr(), st), "gsl_chamfer_backward")
        return gxyz1, gxyz2


class chamfer_3DDist(nn.Module):
    def forward(self, input1, input2):
        return chamfer_3DFunction.apply(input1.contiguous(), input2.contiguous())


def fscore(dist1, dist2, threshold=0.001):
    """F-score of two point clouds from their squared nearest-neighbour distances (chamfer/fscore.py:4-18):
    returns (fscore, precision of set 1, precision of set 2), each (B,)."""
    p1 = torch.mean((dist1 < threshold).float(), dim=1)
    p2 = torch.mean((dist2 < threshold).float(), dim=1)
    f = 2 * p1 * p2 / (p1 + p2)
    f[torch.isnan(f)] = 0
    return f, p1, p2
