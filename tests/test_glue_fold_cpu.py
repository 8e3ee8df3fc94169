"""CPU: the algebra behind gsl_peer_glue ("Dynamic scenes").  Frame-parallel ranks render different
timestamps, and the Jacobian of render()'s glue (tests/glue_oracle.py restates gaussian_renderer/__init__.py:64-115 and
scene/gaussian_model.py:139-186, pinned against the reference accessors by test_glue_oracle_cpu.py) depends on the
timestamp.  The product applies the frame-dependent part of the glue's VJP to every rank's (dL/dmeans3D, dL/dopacity)
BEFORE the sum over the ranks and the frame-independent part (sigmoid', exp', normalize') AFTER it.  Here the same split is
restated in PyTorch (`fold` mirrors glue_fold of csrc/gsl_preprocess.cu line by line) and checked against autograd of the
restated glue, summed over frames -- and the naive order (sum first, one frame's Jacobian after) is shown to be wrong."""
import numpy as np
import torch

import glue_oracle as GO


def fold(pc, timestamp, time_shift, dynamic, g_means3D, g_opacity):
    """(dL/dxyz, dL/dvelocity, dL/dt, dL/dscaling_t, dL/d sigmoid(opacity)) of ONE frame: glue_fold, gsl_preprocess.cu."""
    shift = 0.0 if time_shift is None else time_shift
    ts = timestamp - shift
    a = 1 / pc.T * np.pi * 2
    sig = torch.exp(pc._scaling_t)
    ph = (ts - pc._t) * a
    coef = torch.sin(ph) / a
    ev = torch.exp(-sig / pc.T / 2 * pc.velocity_decay)
    if shift != 0.0:
        coef = coef + ev * shift
    gvv = (g_means3D * pc._velocity).sum(1, keepdim=True)
    g_t = -torch.cos(ph) * gvv
    g_sig = gvv * shift * ev * (-pc.velocity_decay / pc.T / 2) if shift != 0.0 else torch.zeros_like(gvv)
    gop = g_opacity
    if dynamic:
        d = pc._t - ts
        mt = torch.exp(-0.5 * d * d / (sig * sig))
        os_ = torch.sigmoid(pc._opacity)
        g_mt = gop * os_
        gop = gop * mt
        g_t = g_t + g_mt * mt * (-d / (sig * sig))
        g_sig = g_sig + g_mt * mt * d * d / (sig * sig * sig)
    return g_means3D, g_means3D * coef, g_t, g_sig * sig, gop


def after_sum(pc, s_xyz, s_vel, s_t, s_sigt, s_gop, s_scales, s_rot):
    """The frame-independent rest of the glue's VJP, applied to the sums (k_glue_bwd with dynamic = 0)."""
    os_ = torch.sigmoid(pc._opacity)
    q = pc._rotation
    nrm = q.norm(dim=1, keepdim=True)
    n = q / nrm
    return {"_xyz": s_xyz, "_velocity": s_vel, "_t": s_t, "_scaling_t": s_sigt, "_opacity": s_gop * os_ * (1 - os_),
            "_scaling": s_scales * torch.exp(pc._scaling), "_rotation": (s_rot - n * (n * s_rot).sum(1, keepdim=True)) / nrm}


def _frames():
    # (timestamp, time_shift) per rank: different timestamps, one rank with the training loop's random shift (train.py:172)
    return [(0.03, None), (0.12, 0.02), (-0.2, None), (0.31, -0.015)]


def _cotangents(P, seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    return r(P, 3), r(P, 1), r(P, 3), r(P, 4)


def _f64(pc):
    for n in GO.RAW:
        setattr(pc, n, getattr(pc, n).detach().double().requires_grad_(True))
    return pc


def test_fold_before_the_sum_equals_the_sum_of_the_frames_autograd_gradients():
    for dynamic in (True, False):
        P = 257
        pc = _f64(GO.make_model(P, seed=11))
        want = {n: torch.zeros_like(getattr(pc, n)) for n in GO.RAW}
        sums = [0, 0, 0, 0, 0, 0, 0]
        for k, (ts, shift) in enumerate(_frames()):
            cm, co, cs, cr = _cotangents(P, 100 + k)  # this frame's rasterizer-input gradients
            m3, op, sc, rot, _, _ = GO.reference_glue(pc, ts, shift, dynamic, None)
            loss = (m3 * cm).sum() + (op * co).sum() + (sc * cs).sum() + (rot * cr).sum()
            for n, g in zip(GO.RAW, torch.autograd.grad(loss, [getattr(pc, n) for n in GO.RAW], allow_unused=True)):
                if g is not None:
                    want[n] += g
            with torch.no_grad():
                parts = fold(pc, ts, shift, dynamic, cm, co) + (cs, cr)
                sums = [a + b for a, b in zip(sums, parts)]
        with torch.no_grad():
            got = after_sum(pc, *sums)
        for n in GO.RAW:
            err = float((got[n] - want[n]).norm() / (want[n].norm() + 1e-300))
            assert err < 1e-12, (dynamic, n, err)


def test_summing_first_and_applying_one_frames_jacobian_is_wrong_for_different_timestamps():
    """What an exchange of rasterizer-INPUT gradients followed by one rank's glue backward computes (round 1)."""
    P = 257
    pc = _f64(GO.make_model(P, seed=12))
    want = {n: torch.zeros_like(getattr(pc, n)) for n in GO.RAW}
    s_m, s_o = 0, 0
    for k, (ts, shift) in enumerate(_frames()[:2]):
        cm, co, cs, cr = _cotangents(P, 200 + k)
        m3, op, _, _, _, _ = GO.reference_glue(pc, ts, None, True, None)
        loss = (m3 * cm).sum() + (op * co).sum()
        for n, g in zip(GO.RAW, torch.autograd.grad(loss, [getattr(pc, n) for n in GO.RAW], allow_unused=True)):
            if g is not None:
                want[n] += g
        s_m, s_o = s_m + cm, s_o + co
    ts0 = _frames()[0][0]
    m3, op, _, _, _, _ = GO.reference_glue(pc, ts0, None, True, None)
    naive = torch.autograd.grad((m3 * s_m).sum() + (op * s_o).sum(), [pc._xyz, pc._velocity, pc._t])
    assert float((naive[0] - want["_xyz"]).abs().max()) < 1e-12          # dL/dxyz is frame-independent
    assert float((naive[1] - want["_velocity"]).norm() / want["_velocity"].norm()) > 1e-2
    assert float((naive[2] - want["_t"]).norm() / want["_t"].norm()) > 1e-2
