"""Drop-in replacement for GS-LiDAR's gaussian_renderer/diff_gaussian_rasterization_2d.py.

Same public names, keyword arguments, return tuple, dtypes, shapes and autograd contract as the
reference module (reference: gaussian_renderer/diff_gaussian_rasterization_2d.py:28-267), so that
`gaussian_renderer/__init__.py:10` can switch with one import line:

    from gs_lidar_b200.diff_gaussian_rasterization_2d import GaussianRasterizationSettings, GaussianRasterizer

Host side only: argument marshalling, a grow-only workspace pool (replaces the reference's
per-call geomBuffer/binningBuffer/imgBuffer tensors, rasterize_points.cu:84-91) and the
torch.autograd.Function.  All computation happens in libgsl_b200.so (hand-written sm_100a CUDA);
there is no PyTorch or CPU fallback.
"""
import ctypes as C
import threading
from typing import NamedTuple

import torch
import torch.nn as nn

from . import _lib as L

_lib = L.load()

NUM_CHANNELS = 4  # cuda_rasterizer/config.h:12


def cpu_deep_copy_tuple(input_tuple):
    copied_tensors = [item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple]
    return tuple(copied_tensors)


# --------------------------------------------------------------------------------------------------
# workspace pool
# --------------------------------------------------------------------------------------------------
class _Workspace:
    """Three device scratch chunks + a pinned int32[2] for the instance count."""

    def __init__(self, device):
        self.device = device
        self.geom = torch.zeros(0, dtype=torch.uint8, device=device)
        self.binning = torch.empty(0, dtype=torch.uint8, device=device)
        self.image = torch.empty(0, dtype=torch.uint8, device=device)
        self.r_capacity = 0
        self.host = torch.zeros(2, dtype=torch.int32).pin_memory()
        self.event = torch.cuda.Event()
        self.key = None  # (P, S) the geom chunk was last laid out for

    def ensure(self, params, r_capacity):
        sz = L.gsl_ws_sizes()
        L.check(_lib.gsl_workspace_sizes(C.byref(params), int(r_capacity), C.byref(sz)), "gsl_workspace_sizes")
        key = (params.P, params.S)
        if self.geom.numel() < sz.geom_bytes or self.key != key:
            # the packed gradient accumulators inside the geom chunk must start all-zero and the
            # layout depends on (P, S): re-zero on any relayout.
            n = max(int(sz.geom_bytes), self.geom.numel())
            if self.geom.numel() < n:
                self.geom = torch.zeros(n + n // 8, dtype=torch.uint8, device=self.device)
            else:
                self.geom.zero_()
            self.key = key
        if self.image.numel() < sz.image_bytes:
            self.image = torch.empty(int(sz.image_bytes), dtype=torch.uint8, device=self.device)
        if self.binning.numel() < sz.binning_bytes or self.r_capacity < r_capacity:
            self.binning = torch.empty(int(sz.binning_bytes), dtype=torch.uint8, device=self.device)
        self.r_capacity = int(r_capacity)

    def as_struct(self):
        ws = L.gsl_workspace()
        ws.geom = self.geom.data_ptr()
        ws.geom_bytes = self.geom.numel()
        ws.binning = self.binning.data_ptr() if self.binning.numel() else None
        ws.binning_bytes = self.binning.numel()
        ws.image = self.image.data_ptr()
        ws.image_bytes = self.image.numel()
        ws.r_capacity = self.r_capacity
        ws.num_rendered_host = self.host.data_ptr()
        return ws


class _Pool:
    """Free workspaces per (device, stream): a workspace is handed back while the GPU work of its call may still be queued,
    so it may only be reused by a later call on the SAME stream (stream order then protects it)."""

    def __init__(self):
        self.free = {}
        self.lock = threading.Lock()
        self.r_hint = {}

    def acquire(self, device):
        key = (device, torch.cuda.current_stream(device).cuda_stream)
        with self.lock:
            lst = self.free.setdefault(key, [])
            if lst:
                return lst.pop()
        ws = _Workspace(device)
        ws.pool_key = key
        return ws

    def release(self, ws):
        with self.lock:
            self.free.setdefault(getattr(ws, "pool_key", (ws.device, 0)), []).append(ws)


_pool = _Pool()

# frame-parallel training: a gs_lidar_b200.parallel.GradientExchange that the backward pass feeds directly
_exchange = None
# renderer.render() leaves the record of its glue call here right before it calls the rasterizer; the forward below attaches
# it to its autograd node, so that a fused PeerExchange can fold the glue's frame-dependent VJP into the rows it sums
_pending_fold = None

# Opt-in extension, OFF by default (= the reference's semantics): treat a 360-degree panorama as periodic in azimuth.
# The reference clamps tile rects at the image border (auxiliary.h:47-55), so a splat on the +-180 degree seam gets an
# AABB spanning the whole width and is binned into every tile of its rows; with wrap-around it keeps its true footprint
# and its low-pass distance is measured to the nearest periodic image.
_wrap_azimuth = False


def set_wrap_azimuth(flag: bool):
    global _wrap_azimuth
    _wrap_azimuth = bool(flag)


# Opt-in: CUDA-graph replay of the forward and backward pass.  A call signature -- device, stream,
# sizes / fov / flags and the addresses of the surfel tensors -- that is seen a second time gets its own workspace, its own
# STATIC output and gradient buffers and one captured graph per pass; from then on a pass is one graph launch (plus one
# one-warp kernel that gathers the per-frame view matrix / camera centre / background, which may live in different tensors
# every call).  Consequences the caller accepts by switching this on (as with torch.cuda.graphs):
#   * the tensors a call returns (maps, radii, and the gradients of its backward) are views of the static buffers of its
#     signature and are OVERWRITTEN by the next call with the same signature -- clone what must outlive a step;
#   * zero gradients with set_to_none=True (the PyTorch default): an optimizer that zeroes `.grad` in place keeps an alias of
#     the static gradient buffer alive; the backward detects that and returns copies instead (correct, but slower);
#   * a second forward with the same signature before the backward of the first one runs un-graphed.
#   * deferred_count=True (off by default): the one host wait a forward has -- the instance count, needed only to detect that
#     the binning workspace was too small -- is taken at the NEXT call of the signature instead of inside this one, so the host
#     never blocks on work it has just issued (a loop with a prefetching loader then stays ahead of the GPU).  The workspace
#     keeps 25 % of headroom over the last count; should a step still exceed it (its kernels then write nothing), the next
#     backward / forward of the signature raises instead of returning: the results of that one step were invalid.
_cuda_graphs = False
_deferred_count = False
_GRAPH_CACHE_MAX = 4
_graph_cache = {}      # signature -> _GraphEntry (insertion order = LRU order)
_graph_seen = {}       # signature -> number of eager calls so far


def set_cuda_graphs(flag: bool, deferred_count: bool = False):
    """Switches CUDA-graph replay on or off (see the comment above); switching off frees the captured graphs."""
    global _cuda_graphs, _deferred_count
    _cuda_graphs = bool(flag)
    _deferred_count = bool(flag) and bool(deferred_count)
    if not _cuda_graphs:
        for e in list(_graph_cache.values()):
            e.destroy()
        _graph_cache.clear()
        _graph_seen.clear()


class _GraphEntry:
    """Static buffers + captured graphs of one call signature."""

    def __init__(self, device):
        self.device = device
        self.ws = _Workspace(device)
        self.cam = torch.zeros(24, dtype=torch.float32, device=device)  # view matrix [0,16), camera centre [16,19), bg [20,24)
        self.maps = None
        self.radii = None
        self.fin = None
        self.fout = None
        self.wss = None
        self.fwd_graph = None
        self.bwd = {}            # mode key -> dict(graph=..., keep=...)
        self.pending = False     # a forward with grad is waiting for its backward
        self.generation = 0
        self.flat = None         # static gradient buffer (+ views)
        self.views = None
        self.d_cov3D = None
        self.cot = None          # static cotangent planes of the staged backward graph
        self.ex_out = None       # static outputs of the fused exchange
        self.R = 0
        self.count_event = None  # deferred count: recorded behind the replay whose count has not been looked at yet
        self.count_key = None

    def drop_graphs(self, backward_only=False):
        if not backward_only and self.fwd_graph is not None:
            _lib.gsl_graph_destroy(self.fwd_graph)
            self.fwd_graph = None
        for b in self.bwd.values():
            _lib.gsl_graph_destroy(b["graph"])
        self.bwd = {}

    def destroy(self):
        try:
            torch.cuda.synchronize(self.device)
            self.drop_graphs()
        except Exception:
            pass


_capture_stream = None  # set while a pass is being captured: the library-owned stream the launches must go to


def _capture(dev, fn):
    """Captures the launches fn() makes (on the stream _stream_ptr hands out) into an instantiated graph; returns its handle."""
    global _capture_stream
    cs = C.c_void_p()
    L.check(_lib.gsl_graph_begin(C.byref(cs)), "gsl_graph_begin")
    _capture_stream = cs
    try:
        fn()
    except Exception:
        _lib.gsl_graph_end(cs, None)  # leave capture mode
        raise
    finally:
        _capture_stream = None
    handle = C.c_void_p()
    L.check(_lib.gsl_graph_end(cs, C.byref(handle)), "gsl_graph_end")
    return handle


class _Holder:
    """Keeps a workspace attached to one forward call until its backward ran (or it is dropped)."""

    def __init__(self, ws, entry=None):
        self.ws = ws
        self.entry = entry  # graph mode: the workspace belongs to a _GraphEntry, not to the pool

    def release(self):
        if self.ws is not None:
            if self.entry is None:
                _pool.release(self.ws)
            else:
                self.entry.pending = False
            self.ws = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def _ptr(t):
    return t.data_ptr() if (t is not None and t.numel() > 0) else None


def _f32c(t):
    if t.dtype == torch.float32 and t.is_contiguous():
        return t
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _make_params(settings, P, S, M):
    p = L.gsl_params()
    p.P, p.S, p.D, p.M = int(P), int(S), int(settings.sh_degree), int(M)
    p.W, p.H = int(settings.image_width), int(settings.image_height)
    p.tanfovx, p.tanfovy = float(settings.tanfovx), float(settings.tanfovy)
    p.scale_modifier = float(settings.scale_modifier)
    p.vfov_min, p.vfov_max = float(settings.vfov[0]), float(settings.vfov[1])
    p.hfov_min, p.hfov_max = float(settings.hfov[0]), float(settings.hfov[1])
    p.scale_factor = float(settings.scale_factor)
    p.prefiltered = int(bool(settings.prefiltered))
    p.flags = L.GSL_FLAG_DEBUG_SYNC if settings.debug else 0
    if _wrap_azimuth:
        p.flags |= L.GSL_FLAG_WRAP_AZIMUTH
    return p


def _stream_ptr(device):
    if _capture_stream is not None:
        return _capture_stream
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, features, opacities, scales, rotations,
                        cov3Ds_precomp, mask, raster_settings, sh_rest=None):
    if sh_rest is None:
        sh_rest = torch.empty(0, dtype=torch.float32, device=means3D.device)
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, features, opacities, scales,
                                     rotations, cov3Ds_precomp, mask, raster_settings, sh_rest)


def _forward_impl(means3D, sh, colors_precomp, features, opacities, scales, rotations, mask, settings, sh_rest):
    """Runs the forward pass; returns (outputs..., holder, params, keepalive inputs)."""
    if means3D.dim() != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    if not means3D.is_cuda:
        raise RuntimeError("gs_lidar_b200 runs on CUDA tensors only (no CPU fallback)")
    dev = means3D.device
    P = means3D.shape[0]
    S = features.shape[1] if features.dim() == 2 else 0
    M = sh.shape[1] if (sh.numel() != 0 and sh.dim() == 3) else 0
    split = sh_rest.numel() != 0
    if split:
        if M != 1 or sh_rest.dim() != 3 or sh_rest.shape[0] != P or sh_rest.shape[2] != NUM_CHANNELS:
            raise RuntimeError("shs_rest (P, M-1, %d) goes with shs = the DC coefficient (P, 1, %d)"
                               % (NUM_CHANNELS, NUM_CHANNELS))
        M = 1 + sh_rest.shape[1]
    H, W = int(settings.image_height), int(settings.image_width)
    if scales.numel() == 0 or rotations.numel() == 0:
        if P > 0:
            raise RuntimeError("scales and rotations are required: like the reference kernels "
                               "(forward.cu:237) the cov3D_precomp path is not computed")

    inputs = dict(
        background=_f32c(settings.bg), means3D=_f32c(means3D), shs=_f32c(sh),
        shs_rest=_f32c(sh_rest) if split else sh_rest, colors_precomp=_f32c(colors_precomp),
        features=_f32c(features), opacities=_f32c(opacities), scales=_f32c(scales), rotations=_f32c(rotations),
        mask=mask.contiguous().view(torch.uint8) if mask.dtype == torch.bool else mask.to(torch.uint8).contiguous(),
        viewmatrix=_f32c(settings.viewmatrix), projmatrix=_f32c(settings.projmatrix), campos=_f32c(settings.campos))
    params = _make_params(settings, P, S, M)

    entry = _graph_entry_for(dev, params, inputs, settings) if (_cuda_graphs and P > 0 and not settings.debug) else None
    if entry is not None:
        return _forward_graph(entry, dev, params, inputs, P, S, H, W)

    with torch.cuda.device(dev):
        # one allocation for all maps (planes of H*W 4-byte elements), one for radii
        n_planes = 2 + NUM_CHANNELS + (S + 3) + 4 + 1
        maps = torch.empty((n_planes, H, W), dtype=torch.float32, device=dev)
        out_contrib = maps[0:2].view(torch.int32)
        out_color = maps[2:2 + NUM_CHANNELS]
        out_feature = maps[2 + NUM_CHANNELS:2 + NUM_CHANNELS + S + 3]
        out_depth = maps[n_planes - 5:n_planes - 1]
        out_alpha = maps[n_planes - 1:n_planes]
        radii = torch.empty((P,), dtype=torch.int32, device=dev)

        fin = L.gsl_fwd_inputs()
        for k, t in inputs.items():
            setattr(fin, k, _ptr(t))
        fin.cov3D_precomp = None
        fout = L.gsl_fwd_outputs()
        fout.out_contrib, fout.out_color, fout.out_feature = out_contrib.data_ptr(), out_color.data_ptr(), out_feature.data_ptr()
        fout.out_depth, fout.out_alpha, fout.radii = out_depth.data_ptr(), out_alpha.data_ptr(), _ptr(radii)

        ws = _pool.acquire(dev)
        holder = _Holder(ws)
        hint = _pool.r_hint.get((dev, P, W, H), max(4 * P, 1024))
        ws.ensure(params, max(ws.r_capacity, hint))
        wss = ws.as_struct()
        st = _stream_ptr(dev)
        L.check(_lib.gsl_forward_preprocess(C.byref(params), C.byref(fin), C.byref(fout), C.byref(wss), st),
                "gsl_forward_preprocess")
        # The render stage is enqueued speculatively against the current capacity; the instance count arrives on
        # the host while the compositing kernels are still queued (no pipeline bubble in the middle of the
        # forward pass like the reference's blocking copy, rasterizer_impl.cu:314-315).
        r_host = C.c_int32(0)
        for attempt in range(3):
            rc = _lib.gsl_forward_render(C.byref(params), C.byref(fin), C.byref(fout), C.byref(wss), st)
            if rc != L.GSL_ENOSPACE:
                L.check(rc, "gsl_forward_render")
            L.check(_lib.gsl_wait_num_rendered(C.byref(wss), C.byref(r_host), st), "gsl_wait_num_rendered")
            R = int(r_host.value)
            if R <= ws.r_capacity:
                break
            ws.ensure(params, R + R // 4)
            wss = ws.as_struct()
        else:
            raise RuntimeError("gs_lidar_b200: could not size the binning workspace for %d instances" % R)
        _pool.r_hint[(dev, P, W, H)] = max(R + R // 4, 1024)
    outs = (out_contrib, out_color, out_feature, out_depth, out_alpha, radii)
    inputs["_fin"] = fin  # the backward passes the same pointer struct again
    return outs, holder, params, inputs, R


_BIG_INPUTS = ("means3D", "shs", "shs_rest", "colors_precomp", "features", "opacities", "scales", "rotations", "mask")


def _graph_entry_for(dev, params, inputs, settings):
    """The _GraphEntry of this call's signature, or None while the signature is new (the first call runs un-graphed and
    leaves the instance-count hint the entry's workspace is sized with) or its entry is busy."""
    sig = (dev.index, torch.cuda.current_stream(dev).cuda_stream, bytes(params),
           tuple(inputs[k].data_ptr() if inputs[k].numel() else 0 for k in _BIG_INPUTS))
    entry = _graph_cache.get(sig)
    if entry is None:
        n = _graph_seen.get(sig, 0)
        _graph_seen[sig] = n + 1
        if n == 0:
            if len(_graph_seen) > 64:
                _graph_seen.clear()
            return None
        while len(_graph_cache) >= _GRAPH_CACHE_MAX:  # least recently used signature goes
            old = next(iter(_graph_cache))
            _graph_cache.pop(old).destroy()
        entry = _GraphEntry(dev)
        _graph_cache[sig] = entry
    else:
        _graph_cache[sig] = _graph_cache.pop(sig)  # most recently used
    if entry.pending:
        return None
    return entry


def _check_deferred_count(entry, dev, P, W, H, blocking=True):
    """Deferred count of the previous replay (set_cuda_graphs(deferred_count=True)): waits for its event (long complete unless
    the host is a whole step ahead), updates the capacity hint, and raises if that replay did not fit."""
    ev = entry.count_event
    if ev is None or entry.count_key is None:
        return
    if not blocking and not ev.query():
        return
    ev.synchronize()
    entry.count_key = None
    R = int(entry.ws.host[0])
    if R < 0:
        return
    if R + R // 8 > _pool.r_hint.get((dev, P, W, H), 0):
        _pool.r_hint[(dev, P, W, H)] = max(R + R // 4, 1024)
    if R > entry.ws.r_capacity:
        torch.cuda.current_stream(dev).synchronize()
        entry.drop_graphs()
        entry.ws.r_capacity = 0
        entry.R = 0
        raise RuntimeError("gs_lidar_b200 (CUDA-graph mode, deferred count): the previous step of this call signature produced "
                           "%d tile instances, more than its binning workspace held; its outputs and gradients were "
                           "invalid.  The workspace has been re-sized; repeat the step (or use set_cuda_graphs(True) without "
                           "deferred_count, which re-runs such a forward transparently)." % R)
    entry.R = R


def _forward_graph(entry, dev, params, inputs, P, S, H, W):
    """Forward pass by graph replay into the entry's static buffers (see the comment at `_cuda_graphs`)."""
    with torch.cuda.device(dev):
        st = _stream_ptr(dev)
        L.check(_lib.gsl_stage_camera(inputs["viewmatrix"].data_ptr(), inputs["campos"].data_ptr(),
                                      inputs["background"].data_ptr(), entry.cam.data_ptr(), st), "gsl_stage_camera")
        if entry.maps is None:
            n_planes = 2 + NUM_CHANNELS + (S + 3) + 4 + 1
            entry.maps = torch.empty((n_planes, H, W), dtype=torch.float32, device=dev)
            entry.radii = torch.empty((P,), dtype=torch.int32, device=dev)
            fin = L.gsl_fwd_inputs()
            for k, t in inputs.items():
                setattr(fin, k, _ptr(t))
            fin.cov3D_precomp = None
            cam = entry.cam.data_ptr()
            fin.viewmatrix, fin.campos, fin.background = cam, cam + 16 * 4, cam + 20 * 4
            fin.projmatrix = cam  # only gsl_mark_visible reads it
            entry.fin = fin  # holds ADDRESSES only: the graph is replayed only for calls whose tensors sit at them
        maps, n_planes = entry.maps, entry.maps.shape[0]
        out_contrib = maps[0:2].view(torch.int32)
        out_color = maps[2:2 + NUM_CHANNELS]
        out_feature = maps[2 + NUM_CHANNELS:2 + NUM_CHANNELS + S + 3]
        out_depth = maps[n_planes - 5:n_planes - 1]
        out_alpha = maps[n_planes - 1:n_planes]
        if entry.fout is None:
            fout = L.gsl_fwd_outputs()
            fout.out_contrib, fout.out_color, fout.out_feature = out_contrib.data_ptr(), out_color.data_ptr(), out_feature.data_ptr()
            fout.out_depth, fout.out_alpha, fout.radii = out_depth.data_ptr(), out_alpha.data_ptr(), _ptr(entry.radii)
            entry.fout = fout
        ws = entry.ws
        r_host = C.c_int32(0)
        R = 0
        for attempt in range(3):
            if entry.fwd_graph is None:
                hint = _pool.r_hint.get((dev, P, W, H), max(4 * P, 1024))
                ws.ensure(params, max(ws.r_capacity, hint))
                entry.wss = ws.as_struct()
                entry.params = params

                def launches():
                    cs = _stream_ptr(dev)  # the capture stream
                    L.check(_lib.gsl_forward_preprocess(C.byref(params), C.byref(entry.fin), C.byref(entry.fout),
                                                        C.byref(entry.wss), cs), "gsl_forward_preprocess")
                    L.check(_lib.gsl_forward_render(C.byref(params), C.byref(entry.fin), C.byref(entry.fout),
                                                    C.byref(entry.wss), cs), "gsl_forward_render")

                entry.fwd_graph = _capture(dev, launches)
            _check_deferred_count(entry, dev, P, W, H)  # the previous replay of this signature, if it was not looked at
            ws.host[0] = -1  # the sentinel gsl_wait_num_rendered polls; the graph's copy node overwrites it
            L.check(_lib.gsl_graph_launch(entry.fwd_graph, st), "gsl_graph_launch (forward)")
            if _deferred_count and entry.R > 0 and entry.R + entry.R // 8 <= ws.r_capacity:
                # trusted capacity (the last count fits with 12 % to spare): the count of THIS replay is read later
                if entry.count_event is None:
                    entry.count_event = torch.cuda.Event()
                entry.count_event.record(torch.cuda.current_stream(dev))
                entry.count_key = (dev, P, W, H)
                R = entry.R
                break
            L.check(_lib.gsl_wait_num_rendered(C.byref(entry.wss), C.byref(r_host), st), "gsl_wait_num_rendered")
            R = int(r_host.value)
            if R <= ws.r_capacity:
                break
            # the binning chunk was too small (nothing was written): grow it, capture again
            torch.cuda.current_stream(dev).synchronize()
            entry.drop_graphs()
            _pool.r_hint[(dev, P, W, H)] = max(R + R // 4, 1024)
            ws.r_capacity = 0
        else:
            raise RuntimeError("gs_lidar_b200: could not size the binning workspace for %d instances" % R)
        if R + R // 8 > _pool.r_hint.get((dev, P, W, H), 0):
            _pool.r_hint[(dev, P, W, H)] = max(R + R // 4, 1024)
        entry.R = R
        entry.generation += 1
    holder = _Holder(ws, entry)
    outs = (out_contrib, out_color, out_feature, out_depth, out_alpha, entry.radii)
    ins = dict(inputs)
    ins["_fin"] = entry.fin
    return outs, holder, params, ins, R


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, features, opacities, scales, rotations,
                cov3Ds_precomp, mask, raster_settings, sh_rest):
        args = (means3D, sh, colors_precomp, features, opacities, scales, rotations, mask, raster_settings, sh_rest)
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args[:-2] + args[-1:])  # copy them before they can be corrupted
            try:
                outs, holder, params, inputs, R = _forward_impl(*args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            outs, holder, params, inputs, R = _forward_impl(*args)
        contrib, color, feature, depth, alpha, radii = outs

        global _pending_fold
        ctx.fold, _pending_fold = _pending_fold, None
        ctx.raster_settings = raster_settings
        ctx.num_rendered = R
        ctx.gsl_params = params
        ctx.means2D_shape = tuple(means2D.shape)
        ctx.cov_shape = tuple(cov3Ds_precomp.shape)
        needs_grad = any(ctx.needs_input_grad)  # all False under torch.no_grad()
        ctx.graph_generation = holder.entry.generation if holder.entry is not None else 0
        if needs_grad and holder.entry is not None:
            holder.entry.pending = True
        if needs_grad:
            ctx.holder = holder
            # the contiguous fp32 tensors the kernels read again in backward go through save_for_backward like in the
            # reference (diff_gaussian_rasterization_2d.py:122), so an in-place change between forward and backward is
            # detected by autograd's version counters; `_fin` is the pointer struct over exactly these tensors
            ctx.input_names = [k for k, v in inputs.items() if isinstance(v, torch.Tensor)]
            ctx.fin = inputs["_fin"]
            ctx.save_for_backward(contrib, radii, *[inputs[k] for k in ctx.input_names])
        else:
            holder.release()
            ctx.holder = None
        ctx.mark_non_differentiable(contrib, radii)
        return contrib, color, feature, depth, alpha, radii

    @staticmethod
    def backward(ctx, grad_out_contrib, grad_out_color, grad_out_feature, grad_depth, grad_alpha, _):
        holder = ctx.holder
        if holder is None or holder.ws is None:
            raise RuntimeError("gs_lidar_b200: the workspace of this forward call was already released "
                               "(backward called twice, or forward ran without grad)")
        params, settings = ctx.gsl_params, ctx.raster_settings
        saved = ctx.saved_tensors
        contrib, radii = saved[0], saved[1]
        inputs = dict(zip(ctx.input_names, saved[2:]))
        inputs["_fin"] = ctx.fin
        dev = contrib.device
        P, S, M = params.P, params.S, params.M
        H, W = params.H, params.W

        def cot(g, shape):
            if g is None:
                return torch.zeros(shape, dtype=torch.float32, device=dev)
            return _f32c(g)

        g_color = cot(grad_out_color, (NUM_CHANNELS, H, W))
        g_feature = cot(grad_out_feature, (S + 3, H, W))
        g_depth = cot(grad_depth, (4, H, W))
        g_alpha = cot(grad_alpha, (1, H, W))
        entry = holder.entry  # graph mode: static buffers + captured graphs of this call signature
        if entry is not None and entry.count_key is not None:
            _check_deferred_count(entry, dev, P, W, H, blocking=False)  # raises if this step's forward did not fit
        if entry is not None and ctx.graph_generation != entry.generation:
            raise RuntimeError("gs_lidar_b200 (CUDA-graph mode): the state of this forward call was overwritten by a later "
                               "forward with the same signature; run backward before the next forward or switch "
                               "set_cuda_graphs(False)")
        if entry is not None:
            # The replayed backward writes the entry's STATIC gradient buffers.  If a leaf still holds one of them as its
            # .grad (a second backward that accumulates, or gradients zeroed in place instead of set to None) replaying would
            # overwrite what autograd is about to accumulate into: this backward then runs un-graphed into fresh tensors.
            spans = [(t.data_ptr(), t.data_ptr() + t.numel() * 4) for t in (entry.flat, entry.d_cov3D) if t is not None]
            if entry.ex_out is not None:
                spans += [(t.data_ptr(), t.data_ptr() + t.numel() * 4) for t in entry.ex_out.values()]
            for t in saved[2:]:
                g = t.grad if t.is_leaf else None
                if g is not None and any(lo <= g.data_ptr() < hi for lo, hi in spans):
                    entry = None
                    break

        split = inputs["shs_rest"].numel() != 0  # SH as two parameter tensors (dc, rest)
        ex = _exchange if (_exchange is not None and P > 0 and
                           (_exchange.world_size() > 1 or getattr(_exchange, "force", False))) else None
        if ex is not None and M == 0:
            # the fused exchange rides on the SH-factor algebra; returning local-only gradients here would let the replicas
            # diverge without any error
            raise RuntimeError("gs_lidar_b200: a gradient exchange is active but this call uses colors_precomp (no SH): "
                               "disable the exchange for this backward and all-reduce the gradients yourself")
        fold = getattr(ctx, "fold", None) if (ex is not None and ex.packed) else None
        if ex is not None and ex.packed:
            ex.set_glue(fold)
        if fold is not None:
            entry = None  # the frame's timestamp is a kernel argument of the folded VJP: not replayed from a graph
        with torch.cuda.device(dev):
            e = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
            # dL_dcov3D is all-zero (forward.cu never reads cov3D_precomp); only materialised when the caller passed one
            if entry is not None and ctx.cov_shape == (P, 6):
                if entry.d_cov3D is None:
                    entry.d_cov3D = e(P, 6)
                d_cov3D = entry.d_cov3D
            else:
                d_cov3D = e(P, 6) if ctx.cov_shape == (P, 6) else None
            fused = False
            if ex is None:
                # one allocation for all dense gradients; the returned tensors are contiguous views of it
                # 16-byte-stored tensors first (their sizes are multiples of 16 B), scalar-stored ones after: every
                # tensor is aligned for its stores and the buffer has no gaps (the C side zero-fills it as one range)
                widths = (4, NUM_CHANNELS, 4, M * NUM_CHANNELS, 3, 3, 1, S)
                if entry is not None:
                    if entry.flat is None:
                        entry.flat = e(P * sum(widths))
                    flat = entry.flat
                else:
                    flat = e(P * sum(widths))
                views, off = [], 0
                for w in widths:
                    views.append(flat[off:off + P * w].view(P, w))
                    off += P * w
                d_means2D, d_colors, d_rot, d_sh, d_means3D, d_scales, d_opacity, d_features = views
                d_sh_rest = None
                if split:  # the flat range holds dL/d(dc) (P,1,4) followed by dL/d(rest) (P,M-1,4)
                    d_sh, d_sh_rest = d_sh.view(-1)[:P * NUM_CHANNELS], d_sh.view(-1)[P * NUM_CHANNELS:]
                    d_sh, d_sh_rest = d_sh.view(P, 1, NUM_CHANNELS), d_sh_rest.view(P, M - 1, NUM_CHANNELS)
                else:
                    d_sh = d_sh.view(P, M, NUM_CHANNELS)
            else:
                # gradients go straight into the exchange's flat buffers; dL_dcolors receives the SH factor dL_dRGB
                ex.prepare(P, S, M, dev)
                d_sh = d_sh_rest = None
                fused = ex.packed and ex.sync  # the whole exchange is one C call writing the summed dense gradients
                if not fused:
                    entry = None  # the piecewise / NCCL exchanges are host-driven: not capturable
                if fused:
                    if entry is not None:
                        if entry.ex_out is None:
                            entry.ex_out = ex.alloc_outputs(P, S, M, dev)
                        v = dict(entry.ex_out)
                    else:
                        v = ex.alloc_outputs(P, S, M, dev)
                    d_means3D, d_means2D, d_opacity = v["means3D"], v["means2D"], v["opacities"]
                    d_scales, d_rot, d_features, d_sh = v["scales"], v["rotations"], v["features"], v["shs"]
                    d_colors = None
                elif ex.packed:  # emulation (tests): the pieces are driven one by one, the tensors come from ex.finish
                    d_means3D = d_means2D = d_opacity = d_scales = d_rot = d_features = d_colors = None
                else:
                    v = ex.views
                    d_means3D, d_means2D, d_opacity = v["means3D"], v["means2D"], v["opacities"]
                    d_scales, d_rot, d_features = v["scales"], v["rotations"], v["features"]
                    d_colors = e(P, NUM_CHANNELS)  # scratch: the factor itself goes to ex.local (see below)
                params = L.gsl_params.from_buffer_copy(params)
                params.flags |= ex.flags()

            fin = inputs["_fin"]
            ffwd = L.gsl_fwd_outputs()
            ffwd.out_contrib, ffwd.radii = contrib.data_ptr(), _ptr(radii)
            gin = L.gsl_bwd_inputs()
            gin.dL_dout_color, gin.dL_dout_depth = g_color.data_ptr(), g_depth.data_ptr()
            gin.dL_dout_alpha, gin.dL_dout_feature = g_alpha.data_ptr(), g_feature.data_ptr()
            gout = L.gsl_bwd_outputs()
            gout.dL_dmeans3D, gout.dL_dmeans2D, gout.dL_dsh = _ptr(d_means3D), _ptr(d_means2D), _ptr(d_sh)
            gout.dL_dsh_rest = _ptr(d_sh_rest)
            if ex is not None and ex.packed:
                gout.peer = ex.ctx_ptr
            gout.dL_dcolors, gout.dL_dfeatures, gout.dL_dopacity = _ptr(d_colors), _ptr(d_features), _ptr(d_opacity)
            gout.dL_dscales, gout.dL_drotations, gout.dL_dcov3D = _ptr(d_scales), _ptr(d_rot), _ptr(d_cov3D)
            wss = holder.ws.as_struct()

            def run():
                if ex is None:
                    L.check(_lib.gsl_backward(C.byref(params), C.byref(fin), C.byref(ffwd), C.byref(gin), C.byref(gout),
                                              C.byref(wss), _stream_ptr(dev)), "gsl_backward")
                    return
                # frame-parallel: compositor + SH factor, start the all-gather, then the per-surfel kernel under it
                factor_out = None if ex.packed else ex.local.data_ptr()  # packed: the per-surfel kernel pushes the factors
                L.check(_lib.gsl_backward_composite(C.byref(params), C.byref(fin), C.byref(ffwd), C.byref(gin),
                                                    C.byref(gout), C.byref(wss), factor_out, _stream_ptr(dev)),
                        "gsl_backward_composite")
                if fused:
                    L.check(_lib.gsl_backward_surfels_exchange(C.byref(params), C.byref(fin), C.byref(ffwd), C.byref(gout),
                                                               C.byref(wss), ex.epoch & 0xFFFFFFFF, ex.chunks,
                                                               _stream_ptr(dev)), "gsl_backward_surfels_exchange")
                    return
                ex.start_gather(P, inputs["campos"], params.D, M, inputs["means3D"])
                ex.run_surfels(P, lambda rb, re: L.check(
                    _lib.gsl_backward_surfels_rows(C.byref(params), C.byref(fin), C.byref(ffwd), C.byref(gout),
                                                   C.byref(wss), rb, re, _stream_ptr(dev)), "gsl_backward_surfels_rows"))

            if entry is not None:
                # cotangents: the graph captured at the first backward reads them where they were then ("direct"); when
                # autograd hands over other tensors they are copied into static planes first ("staged")
                eager_run = run
                ptrs = (g_color.data_ptr(), g_depth.data_ptr(), g_alpha.data_ptr(), g_feature.data_ptr())
                exkey = None if ex is None else id(ex)
                direct = entry.bwd.get(("direct", exkey))
                if direct is None or direct["ptrs"] == ptrs:
                    mode, keep = ("direct", exkey), (g_color, g_depth, g_alpha, g_feature)
                else:
                    mode = ("staged", exkey)
                    if entry.cot is None:
                        entry.cot = e(4 + 4 + 1 + S + 3, H, W)
                    c = entry.cot
                    st_c, st_d, st_a, st_f = c[0:4], c[4:8], c[8:9], c[9:]
                    st_c.copy_(g_color); st_d.copy_(g_depth); st_a.copy_(g_alpha); st_f.copy_(g_feature)
                    gin.dL_dout_color, gin.dL_dout_depth = st_c.data_ptr(), st_d.data_ptr()
                    gin.dL_dout_alpha, gin.dL_dout_feature = st_a.data_ptr(), st_f.data_ptr()
                    keep = None

                def run():  # noqa: F811 -- replay (capture first) instead of the individual launches
                    b = entry.bwd.get(mode)
                    if b is None:
                        b = dict(graph=_capture(dev, eager_run), ptrs=ptrs, keep=keep, gin=gin, gout=gout, ffwd=ffwd,
                                 params=params, wss=wss)
                        entry.bwd[mode] = b
                    L.check(_lib.gsl_graph_launch(b["graph"], _stream_ptr(dev)), "gsl_graph_launch (backward)")

            try:
                if settings.debug:
                    cpu_args = cpu_deep_copy_tuple((g_color, g_depth, g_alpha, g_feature, contrib, radii) +
                                                   tuple(v for k, v in inputs.items() if k != "_fin"))
                    try:
                        run()
                    except Exception as ex_:
                        torch.save(cpu_args, "snapshot_bw.dump")
                        print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                        raise ex_
                else:
                    run()
            except Exception:
                # the packed gradient accumulators may be left dirty (the per-surfel kernel re-zeroes what it consumed,
                # and it may not have run): force a re-zero before this workspace is used again
                holder.ws.key = None
                raise
            if ex is not None and not fused:
                g = ex.finish(P, params.D, M, inputs["means3D"])
                d_means3D, d_means2D, d_opacity = g["means3D"], g["means2D"], g["opacities"]
                d_scales, d_rot, d_features, d_sh = g["scales"], g["rotations"], g["features"], g["shs"]
                v = g
            if ex is not None and split:
                # the exchanges rebuild dL_dsh as ONE (P,M,4) tensor from the factors; the two parameter tensors receive views
                d_sh, d_sh_rest = d_sh[:, :1], d_sh[:, 1:]
            if fold is not None:
                # the exchanged rows carried dL/dvelocity, dL/dt, dL/dscaling_t of every rank's own timestamp behind the
                # features; the glue's backward (renderer._ActivateSurfels) takes them from the record
                d_features, fold["extras"] = ex.split_glue(v, S)
        if not _KEEP_WORKSPACE_AFTER_BACKWARD:
            holder.release()

        grad_cov = d_cov3D
        grads = (d_means3D, d_means2D, d_sh if M > 0 else None, d_colors if inputs["colors_precomp"].numel() else None,
                 d_features, d_opacity, d_scales, d_rot, grad_cov, None, None, d_sh_rest)
        if entry is not None:
            # fresh view objects of the static buffers: autograd adopts them as .grad without a copy
            grads = tuple(None if g is None else g.view(g.shape) for g in grads)
        return grads


# Set True to allow a second backward through the same graph (retain_graph=True); the workspace is
# then only returned to the pool when the autograd node is freed.
_KEEP_WORKSPACE_AFTER_BACKWARD = False


def set_keep_workspace_after_backward(flag: bool):
    global _KEEP_WORKSPACE_AFTER_BACKWARD
    _KEEP_WORKSPACE_AFTER_BACKWARD = bool(flag)


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool
    vfov: tuple
    hfov: tuple
    scale_factor: float


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        # Mark visible points (based on frustum culling for camera) with a boolean
        with torch.no_grad():
            rs = self.raster_settings
            pos = _f32c(positions)
            P = pos.shape[0]
            present = torch.zeros((P,), dtype=torch.bool, device=pos.device)
            if P > 0:
                vm, pm = _f32c(rs.viewmatrix), _f32c(rs.projmatrix)
                with torch.cuda.device(pos.device):
                    L.check(_lib.gsl_mark_visible(P, pos.data_ptr(), vm.data_ptr(), pm.data_ptr(),
                                                  present.data_ptr(), _stream_ptr(pos.device)), "gsl_mark_visible")
        return present

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, features=None, scales=None,
                rotations=None, cov3D_precomp=None, mask=None, shs_rest=None):
        """Same keyword arguments as the reference (diff_gaussian_rasterization_2d.py:224-267).  Extension:
        `shs_rest` -- pass GaussianModel._features_dc as `shs` (P,1,4) and _features_rest as `shs_rest` (P,M-1,4)
        and the kernels read/write the two parameter tensors directly, without the concatenated copy
        `get_features` builds per call (scene/gaussian_model.py:167-171) and its split in backward."""
        raster_settings = self.raster_settings

        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')

        if ((scales is None or rotations is None) and cov3D_precomp is None) or (
                (scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')

        dev = means3D.device
        empty = lambda: torch.empty(0, dtype=torch.float32, device=dev)
        if shs_rest is not None and shs is None:
            raise Exception('shs_rest needs shs (the DC coefficient)')
        if shs is None:
            shs = empty()
        if colors_precomp is None:
            colors_precomp = empty()
        if features is None:
            features = torch.empty_like(means3D[..., :0])
        if scales is None:
            scales = empty()
        if rotations is None:
            rotations = empty()
        if cov3D_precomp is None:
            cov3D_precomp = empty()
        if mask is None:
            mask = torch.ones_like(means3D[:, :1], dtype=torch.bool)

        # Invoke the CUDA rasterization routine
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, features, opacities, scales, rotations,
                                   cov3D_precomp, mask, raster_settings, shs_rest)
