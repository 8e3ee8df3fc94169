"""CUDA-graph replay of the forward / backward pass (set_cuda_graphs): replayed steps must give the
bits the un-graphed path gives, with per-frame cameras changing between replays, cotangents arriving in new tensors
("staged" backward), instance-count overflow after capture, and gradients zeroed in place."""
import pytest
import torch

import common
from gs_lidar_b200 import synth

pytestmark = pytest.mark.gpu
NAMES = ("means3D", "means2D", "opacities", "scales", "rotations", "features", "shs")


def _leaves(scene):
    P = scene.means3D.shape[0]
    lv = dict(means3D=scene.means3D.clone(), means2D=torch.zeros((P, 4), device="cuda"), opacities=scene.opacities.clone(),
              shs=scene.shs.clone(), features=scene.features.clone(), scales=scene.scales.clone(),
              rotations=scene.rotations.clone())
    for v in lv.values():
        v.requires_grad_(True)
    return lv


def _step(G, scene, lv, cot, set_to_none=True):
    for v in lv.values():
        if set_to_none:
            v.grad = None
        elif v.grad is not None:
            v.grad.zero_()
    out = G.GaussianRasterizer(synth.settings_for(scene))(mask=scene.mask, **lv)
    torch.autograd.backward([out[1], out[2], out[3], out[4]], [cot["color"], cot["feature"], cot["depth"], cot["alpha"]])
    maps = [o.detach().clone() for o in out]
    return maps, {k: lv[k].grad.detach().clone() for k in NAMES}


def _cams(P, n):
    return [synth.make_scene(16, seed=5, view_yaw_deg=1.5 * k, view_shift=(0.05 * k, 0.0, 0.01 * k)) for k in range(n)]


@pytest.fixture
def graphs():
    import gs_lidar_b200.diff_gaussian_rasterization_2d as G
    G.set_cuda_graphs(True)
    yield G
    G.set_cuda_graphs(False)


def test_replayed_steps_equal_the_plain_path_for_changing_cameras(graphs):
    G = graphs
    P = 30000
    base = synth.make_scene(P, seed=5).to("cuda")
    cot = {k: v.cuda() for k, v in synth.make_cotangents(base.H, base.W, 4, seed=6).items()}
    cams = _cams(P, 4)
    scenes = [base._replace(viewmatrix=c.viewmatrix.cuda(), projmatrix=c.projmatrix.cuda(), campos=c.campos.cuda()) for c in cams]
    lv = _leaves(base)
    G.set_cuda_graphs(False)
    want = [_step(G, sc, lv, cot) for sc in scenes]
    G.set_cuda_graphs(True)
    got = [_step(G, sc, lv, cot) for sc in scenes]  # call 1 is eager (new signature), 2.. replay
    assert len(G._graph_cache) == 1 and next(iter(G._graph_cache.values())).fwd_graph is not None
    for (wm, wg), (gm, gg) in zip(want, got):
        for a, b in zip(wm, gm):
            assert torch.equal(a, b)  # maps, contributor counts, radii: same kernels, same inputs -> same bits
        for k in NAMES:
            elem, norm = common.grad_err(gg[k], wg[k])
            assert norm < 1e-5 and elem < 1e-4, (k, elem, norm)  # float atomics: order differs run to run


def test_new_cotangent_tensors_take_the_staged_backward_and_inplace_zeroing_is_safe(graphs):
    G = graphs
    P = 20000
    scene = synth.make_scene(P, seed=7).to("cuda")
    cot = {k: v.cuda() for k, v in synth.make_cotangents(scene.H, scene.W, 4, seed=8).items()}
    lv = _leaves(scene)
    G.set_cuda_graphs(False)
    _, want = _step(G, scene, lv, cot)
    G.set_cuda_graphs(True)
    for it in range(5):
        c = cot if it < 3 else {k: v.clone() for k, v in cot.items()}   # other addresses from step 3 on
        _, got = _step(G, scene, lv, c, set_to_none=(it % 2 == 0))        # odd steps zero .grad in place
        for k in NAMES:
            elem, norm = common.grad_err(got[k], want[k])
            assert norm < 1e-5 and elem < 1e-4, (it, k, elem, norm)
    entry = next(iter(G._graph_cache.values()))
    assert {m[0] for m in entry.bwd} == {"direct", "staged"}


def test_overflow_after_capture_regrows_and_recaptures(graphs):
    G = graphs
    P = 20000
    scene = synth.make_scene(P, seed=9).to("cuda")
    lv = _leaves(scene)
    with torch.no_grad():
        rast = G.GaussianRasterizer(synth.settings_for(scene))
        args = dict(means3D=lv["means3D"], means2D=lv["means2D"], opacities=lv["opacities"], shs=lv["shs"],
                    features=lv["features"], rotations=lv["rotations"], mask=scene.mask)
        for _ in range(3):
            small = [o.clone() for o in rast(scales=lv["scales"], **args)]
        entry = next(iter(G._graph_cache.values()))
        cap = entry.ws.r_capacity
        # same tensor, 6x larger splats: several times the instances, beyond the captured capacity
        lv["scales"].data.mul_(6.0)
        grown = [o.clone() for o in rast(scales=lv["scales"], **args)]
        assert entry.R > cap and entry.ws.r_capacity >= entry.R
        G.set_cuda_graphs(False)
        want = rast(scales=lv["scales"], **args)
        for a, b in zip(want, grown):
            assert torch.equal(a, b)
        assert not torch.equal(small[1], grown[1])


def test_outputs_are_static_and_a_busy_signature_runs_ungraphed(graphs):
    G = graphs
    P = 10000
    scene = synth.make_scene(P, seed=11).to("cuda")
    cot = {k: v.cuda() for k, v in synth.make_cotangents(scene.H, scene.W, 4, seed=12).items()}
    lv = _leaves(scene)
    rast = G.GaussianRasterizer(synth.settings_for(scene))
    for _ in range(2):
        _step(G, scene, lv, cot)
    o1 = rast(mask=scene.mask, **lv)
    torch.autograd.backward([o1[1]], [cot["color"]])       # frees the entry for the next forward
    with torch.no_grad():
        o2 = rast(mask=scene.mask, **lv)                      # replays into the same static buffers
    assert o1[1].data_ptr() == o2[1].data_ptr()
    o3 = rast(mask=scene.mask, **lv)                          # forward with grad ...
    with torch.no_grad():
        rast(mask=scene.mask, **lv)                           # ... busy entry: this one runs un-graphed
    torch.autograd.backward([o3[1]], [cot["color"]])       # still valid


def test_deferred_count_gives_the_same_steps_and_reports_an_overflow_at_the_next_call():
    """set_cuda_graphs(True, deferred_count=True): the host does not wait for the instance count inside the forward.  Steps
    must equal the plain path; a step whose instances outgrow the captured workspace (its kernels write nothing) makes
    the NEXT call of the signature raise, after which the signature works again with a larger workspace."""
    import gs_lidar_b200.diff_gaussian_rasterization_2d as G
    P = 20000
    scene = synth.make_scene(P, seed=13).to("cuda")
    cot = {k: v.cuda() for k, v in synth.make_cotangents(scene.H, scene.W, 4, seed=14).items()}
    lv = _leaves(scene)
    want_maps, want_grads = _step(G, scene, lv, cot)
    G.set_cuda_graphs(True, deferred_count=True)
    try:
        for it in range(5):
            maps, grads = _step(G, scene, lv, cot)
            for a, b in zip(maps, want_maps):
                assert torch.equal(a, b), it
            for k in NAMES:
                assert common.grad_err(grads[k], want_grads[k])[1] < 1e-5, (it, k)
        entry = next(iter(G._graph_cache.values()))
        assert entry.count_event is not None  # the deferred path was taken
        cap = entry.ws.r_capacity
        lv["scales"].data.mul_(6.0)  # same tensors, 6x larger splats: several times the instances
        with torch.no_grad():
            rast = G.GaussianRasterizer(synth.settings_for(scene))
            rast(mask=scene.mask, **lv)  # does not fit; nobody has looked yet
            with pytest.raises(RuntimeError, match="more than its binning workspace held"):
                rast(mask=scene.mask, **lv)
            grown = [o.clone() for o in rast(mask=scene.mask, **lv)]  # re-sized: works again
            entry = next(iter(G._graph_cache.values()))
            assert entry.ws.r_capacity > cap
            G.set_cuda_graphs(False)
            plain = rast(mask=scene.mask, **lv)
            for a, b in zip(plain, grown):
                assert torch.equal(a, b)
    finally:
        G.set_cuda_graphs(False)
