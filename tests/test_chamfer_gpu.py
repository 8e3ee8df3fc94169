"""gs_lidar_b200.chamfer against a brute-force oracle, torch.autograd and -- where the compiled
reference travelled with the repo (oracle/_ref/libchamfer_ref.so) -- the reference's own kernels
(chamfer/chamfer3D/chamfer3D.cu:9-138,167-221)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402  (test infrastructure)

pytestmark = [pytest.mark.gpu]
HAVE_REF_CHAMFER = os.path.exists(oracle.REF_CHAMFER_SO)


def oracle_nn(a, b):
    """(B,N,3), (B,M,3) -> squared distance and index of the nearest b for every a; lowest index among equal minima.
    Same float32 expression as chamfer3D.cu:36-40 ((x2-x1)^2 + (y2-y1)^2 + (z2-z1)^2), evaluated without FMA contraction,
    so distances agree to an ulp or two and indices wherever the two nearest candidates differ by more than that."""
    d = ((b[:, None, :, :] - a[:, :, None, :]) ** 2)
    d = (d[..., 0] + d[..., 1]) + d[..., 2]
    dist, idx = d.min(dim=2)
    return dist, idx, d


def sweeps(B, n, m, seed):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(B, n, 3, generator=g) * torch.tensor([20.0, 20.0, 2.0])
    b = torch.randn(B, m, 3, generator=g) * torch.tensor([20.0, 20.0, 2.0])
    return a.cuda(), b.cuda()


@pytest.mark.parametrize("B,n,m", [(1, 1, 1), (1, 7, 1500), (2, 1030, 999), (1, 5000, 4097), (3, 257, 1)])
def test_forward_matches_brute_force(B, n, m):
    from gs_lidar_b200.chamfer import chamfer_3DDist
    a, b = sweeps(B, n, m, seed=n + m)
    d1, d2, i1, i2 = chamfer_3DDist()(a, b)
    for (dist, idx, q, t) in ((d1, i1, a, b), (d2, i2, b, a)):
        od, oi, full = oracle_nn(q, t)
        assert dist.shape == od.shape and idx.dtype == torch.int32
        assert torch.allclose(dist, od, rtol=1e-5, atol=1e-6)
        # the chosen neighbour is a nearest one: its oracle distance equals the minimum up to rounding
        chosen = torch.gather(full, 2, idx.long()[..., None])[..., 0]
        assert torch.allclose(chosen, od, rtol=1e-5, atol=1e-6)
        assert float((idx.long() != oi).float().mean()) < 1e-3


def test_ties_resolve_to_the_lowest_index():
    from gs_lidar_b200.chamfer import chamfer_3DDist
    a = torch.zeros(1, 3, 3).cuda()
    b = torch.tensor([[[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0], [-1.0, 0, 0]] * 700]).cuda()   # 2800 targets, all at distance 1
    d1, d2, i1, i2 = chamfer_3DDist()(a, b)
    assert torch.equal(i1, torch.zeros_like(i1)) and torch.equal(d1, torch.ones_like(d1))
    assert torch.equal(i2, torch.zeros_like(i2))


def test_backward_matches_autograd_of_the_definition():
    from gs_lidar_b200.chamfer import chamfer_3DDist
    a, b = sweeps(2, 700, 650, seed=5)
    a.requires_grad_(True); b.requires_grad_(True)
    d1, d2, i1, i2 = chamfer_3DDist()(a, b)
    g = torch.Generator().manual_seed(6)
    w1, w2 = torch.rand(d1.shape, generator=g).cuda(), torch.rand(d2.shape, generator=g).cuda()
    ((d1 * w1).sum() + (d2 * w2).sum()).backward()
    ga, gb = a.grad.clone(), b.grad.clone()
    a2, b2 = a.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    # the definition with the op's own matches: dist1[j] = |a_j - b_idx1[j]|^2
    e1 = ((a2 - torch.gather(b2, 1, i1.long()[..., None].expand(-1, -1, 3))) ** 2).sum(-1)
    e2 = ((b2 - torch.gather(a2, 1, i2.long()[..., None].expand(-1, -1, 3))) ** 2).sum(-1)
    ((e1 * w1).sum() + (e2 * w2).sum()).backward()
    assert torch.allclose(ga, a2.grad, rtol=1e-4, atol=1e-5) and torch.allclose(gb, b2.grad, rtol=1e-4, atol=1e-5)


def test_lidar_sized_sweeps_and_fscore():
    from gs_lidar_b200.chamfer import chamfer_3DDist, fscore
    a, b = sweeps(1, 34000, 33000, seed=9)
    d1, d2, i1, i2 = chamfer_3DDist()(a, b)
    assert bool(torch.isfinite(d1).all()) and int(i1.max()) < 33000 and int(i2.max()) < 34000
    # spot-check 512 queries of each direction against brute force
    sel = torch.randperm(34000, generator=torch.Generator().manual_seed(1))[:512].cuda()
    od, _, _ = oracle_nn(a[:, sel], b)
    assert torch.allclose(d1[:, sel], od, rtol=1e-5, atol=1e-6)
    f, p1, p2 = fscore(d1, d2, threshold=1.0)
    assert f.shape == (1,) and 0 <= float(f) <= 1


def test_validation():
    from gs_lidar_b200.chamfer import chamfer_3DDist
    with pytest.raises(AssertionError, match="Wrong last dimension"):
        chamfer_3DDist()(torch.zeros(1, 4, 2).cuda(), torch.zeros(1, 4, 3).cuda())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        chamfer_3DDist()(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3))


@pytest.mark.skipif(not HAVE_REF_CHAMFER, reason="compiled reference chamfer kernels not present (oracle/_ref)")
@pytest.mark.parametrize("B,n,m", [(1, 1, 1), (1, 513, 511), (2, 1030, 999), (1, 5000, 4097), (1, 34000, 33000)])
def test_matches_the_reference_kernels(B, n, m):
    """Distances to 1e-6 relative (bit-identical in practice: same expression, same contraction), indices equal except
    for ties, gradients to 1e-5 (the reference's scatter uses atomics in no fixed order)."""
    from gs_lidar_b200.chamfer import chamfer_3DDist
    a, b = sweeps(B, n, m, seed=3 * n + m)
    ref = oracle.RefChamfer()
    rd1, rd2, ri1, ri2 = ref.forward(a, b)
    a.requires_grad_(True); b.requires_grad_(True)
    d1, d2, i1, i2 = chamfer_3DDist()(a, b)
    for ours, theirs in ((d1, rd1), (d2, rd2)):
        assert torch.allclose(ours, theirs, rtol=1e-6, atol=0)
    for (ours, theirs, od, q, t) in ((i1, ri1, d1, a, b), (i2, ri2, d2, b, a)):
        diff = ours != theirs
        if bool(diff.any()):  # only exact ties may differ, and then both are at the minimum distance
            bb, jj = diff.nonzero(as_tuple=True)
            dt = ((t.detach()[bb, theirs[bb, jj].long()] - q.detach()[bb, jj]) ** 2)
            dt = (dt[:, 0] + dt[:, 1]) + dt[:, 2]
            assert torch.allclose(dt, od.detach()[bb, jj], rtol=1e-6, atol=0)
        assert float(diff.float().mean()) < 1e-4
    g = torch.Generator().manual_seed(8)
    w1, w2 = torch.rand(d1.shape, generator=g).cuda(), torch.rand(d2.shape, generator=g).cuda()
    ((d1 * w1).sum() + (d2 * w2).sum()).backward()
    rg1, rg2 = ref.backward(a.detach(), b.detach(), w1, w2, ri1, ri2)
    if not bool((i1 != ri1).any() or (i2 != ri2).any()):
        assert torch.allclose(a.grad, rg1, rtol=1e-5, atol=1e-5) and torch.allclose(b.grad, rg2, rtol=1e-5, atol=1e-5)
