#!/usr/bin/env python
"""Chamfer distance: this repo's kernels against the reference's own (oracle/_ref/libchamfer_ref.so,
the unmodified chamfer/chamfer3D/chamfer3D.cu) on the same GPU, at the size GS-LiDAR's loss uses (two ~34k-point sweeps,
train.py:256-267).  Prints one JSON line.  CUDA events, warm-up, inputs 0.8 MB (L2-resident for both arms).

    python scripts/bench_chamfer.py [--n 34000] [--m 33000] [--iters 50]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def timed(fn, iters, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=34000)
    ap.add_argument("--m", type=int, default=33000)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--iters", type=int, default=50)
    a = ap.parse_args()
    from gs_lidar_b200.chamfer import chamfer_3DDist
    import oracle
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(a.batch, a.n, 3, generator=g) * torch.tensor([20.0, 20.0, 2.0])).cuda()
    y = (torch.randn(a.batch, a.m, 3, generator=g) * torch.tensor([20.0, 20.0, 2.0])).cuda()
    w1, w2 = torch.rand(a.batch, a.n, generator=g).cuda(), torch.rand(a.batch, a.m, generator=g).cuda()
    op = chamfer_3DDist()
    res = {"what": "chamfer distance, B=%d, n=%d, m=%d, fp32" % (a.batch, a.n, a.m), "unit": "ms"}

    def ours_fwd():
        with torch.no_grad():
            op(x, y)

    xg, yg = x.clone().requires_grad_(True), y.clone().requires_grad_(True)

    def ours_fb():
        xg.grad = None
        yg.grad = None
        d1, d2, _, _ = op(xg, yg)
        torch.autograd.backward([d1, d2], [w1, w2])

    res["ours_forward_ms"] = timed(ours_fwd, a.iters)
    res["ours_forward_backward_ms"] = timed(ours_fb, a.iters)
    pairs = 2.0 * a.batch * a.n * a.m
    res["ours_pairs_per_s"] = pairs / (res["ours_forward_ms"] * 1e-3)
    # brute force: ~9 FP32-pipe instructions per (query, target) pair (3 sub, 1 mul, 2 fma, compare, 2 selects)
    res["fp32_issue_bound_ms"] = pairs * 9 / 32 / (148 * 4 * 1.965e9) * 1e3
    if os.path.exists(oracle.REF_CHAMFER_SO):
        ref = oracle.RefChamfer()
        outs = ref.forward(x, y)

        def ref_fwd():
            ref.forward(x, y, outs)

        def ref_fb():
            d1, d2, i1, i2 = ref.forward(x, y, outs)
            ref.backward(x, y, w1, w2, i1, i2)

        res["reference_forward_ms"] = timed(ref_fwd, max(3, a.iters // 5), warmup=2)
        res["reference_forward_backward_ms"] = timed(ref_fb, max(3, a.iters // 5), warmup=2)
        res["speedup_forward"] = res["reference_forward_ms"] / res["ours_forward_ms"]
        res["speedup_forward_backward"] = res["reference_forward_backward_ms"] / res["ours_forward_backward_ms"]
        d1, d2, i1, i2 = op(x, y)
        res["parity"] = {"dist_max_rel": float(max(((d1 - outs[0]).abs() / outs[0].clamp_min(1e-30)).max(),
                                                   ((d2 - outs[1]).abs() / outs[1].clamp_min(1e-30)).max())),
                         "idx_mismatch": int((i1 != outs[2]).sum() + (i2 != outs[3]).sum())}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
