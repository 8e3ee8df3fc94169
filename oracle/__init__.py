"""oracle/ -- TEST INFRASTRUCTURE ONLY.

Two checkers for the panoramic surfel rasterizer:

  * CpuOracle  -- ctypes front-end of oracle/gsl_oracle.c, a plain-C restatement of the reference
                  algorithm (numpy in, numpy out; runs anywhere).
  * RefCuda    -- ctypes front-end of oracle/_ref/libgslidar_ref.so, the UNMODIFIED reference CUDA
                  rasterizer compiled from /root/reference by oracle/build_ref.sh (torch CUDA tensors
                  in/out; needs a GPU).  This is the parity reference and the GPU baseline.
  * RefChamfer -- the same for the reference's Chamfer kernels (oracle/_ref/libchamfer_ref.so, the unmodified
                  chamfer/chamfer3D/chamfer3D.cu compiled against oracle/ref_stub/ATen/ATen.h).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference / the parity
gate) may import this package; gs_lidar_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libgsl_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libgslidar_ref.so")
REF_CHAMFER_SO = os.path.join(HERE, "_ref", "libchamfer_ref.so")


def build_oracle(force=False):
    src = os.path.join(HERE, "gsl_oracle.c")
    if not force and os.path.exists(ORACLE_SO) and os.path.getmtime(ORACLE_SO) >= os.path.getmtime(src):
        return ORACLE_SO
    cmd = ["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-Wall", "-o", ORACLE_SO, src, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcc failed for the oracle:\n" + r.stderr)
    return ORACLE_SO


def build_ref():
    """Compiles the reference CUDA rasterizer if /root/reference is present; else keeps the prebuilt .so."""
    r = subprocess.run(["bash", os.path.join(HERE, "build_ref.sh")], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/build_ref.sh failed:\n" + r.stdout + r.stderr)
    # the reference's Python callers of the rasterizer (render(), GaussianModel, ...), staged unmodified for the drop-in tests
    from . import ref_python
    ref_python.stage()
    return REF_SO if os.path.exists(REF_SO) else None


class orc_params(C.Structure):
    _fields_ = [("P", C.c_int), ("S", C.c_int), ("D", C.c_int), ("M", C.c_int), ("W", C.c_int), ("H", C.c_int),
                ("vfov_min", C.c_float), ("vfov_max", C.c_float), ("hfov_min", C.c_float), ("hfov_max", C.c_float),
                ("scale_factor", C.c_float), ("tanfovx", C.c_float), ("tanfovy", C.c_float), ("wrap", C.c_int)]


def _np(a, dtype):
    return np.ascontiguousarray(np.asarray(a), dtype=dtype)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class CpuOracle:
    """numpy front-end of gsl_oracle.c.  All arrays use the reference's layouts."""

    def __init__(self, threads=None):
        self.lib = C.CDLL(build_oracle())
        self.lib.orc_binning.restype = C.c_int64
        self.lib.orc_num_threads.restype = C.c_int
        if threads:
            self.lib.orc_set_num_threads(int(threads))

    @property
    def threads(self):
        return int(self.lib.orc_num_threads())

    @staticmethod
    def params(P, S, D, M, W, H, vfov, hfov, scale_factor, tanfovx=-0.5463024898437905, tanfovy=-0.5463024898437905, wrap=False):
        """wrap: the azimuth wrap-around mode of this repo (GSL_FLAG_WRAP_AZIMUTH), restated in gsl_oracle.c"""
        return orc_params(P, S, D, M, W, H, vfov[0], vfov[1], hfov[0], hfov[1], scale_factor, tanfovx, tanfovy, int(bool(wrap)))

    def preprocess(self, p, means3D, scales, rotations, opacities, shs, colors_precomp, mask, viewmatrix, campos):
        P = p.P
        out = dict(radii=np.zeros(P, np.int32), means2D=np.zeros((P, 2), np.float32), depths=np.zeros(P, np.float32),
                   transMat=np.zeros((P, 9), np.float32), normal_opacity=np.zeros((P, 4), np.float32),
                   rgb=np.zeros((P, 4), np.float32), clamped=np.zeros((P, 4), np.uint8),
                   tiles_touched=np.zeros(P, np.uint32))
        a = dict(means3D=_np(means3D, np.float32), scales=_np(scales, np.float32), rotations=_np(rotations, np.float32),
                 opacities=_np(opacities, np.float32), shs=None if shs is None else _np(shs, np.float32),
                 colors=None if colors_precomp is None else _np(colors_precomp, np.float32),
                 mask=_np(mask, np.uint8), vm=_np(viewmatrix, np.float32), campos=_np(campos, np.float32))
        self.lib.orc_preprocess(C.byref(p), _p(a["means3D"]), _p(a["scales"]), _p(a["rotations"]), _p(a["opacities"]),
                                _p(a["shs"]), _p(a["colors"]), _p(a["mask"]), _p(a["vm"]), _p(a["campos"]),
                                _p(out["radii"]), _p(out["means2D"]), _p(out["depths"]), _p(out["transMat"]),
                                _p(out["normal_opacity"]), _p(out["rgb"]), _p(out["clamped"]), _p(out["tiles_touched"]))
        return out

    def binning(self, p, radii, means2D, depths, tiles_touched):
        P = p.P
        tiles = ((p.W + 15) // 16) * ((p.H + 15) // 16)
        radii, means2D, depths = _np(radii, np.int32), _np(means2D, np.float32), _np(depths, np.float32)
        tt = _np(tiles_touched, np.uint32)
        offsets = np.zeros(max(P, 1), np.uint32)
        ranges = np.zeros((tiles, 2), np.uint32)
        R = self.lib.orc_binning(C.byref(p), _p(radii), _p(means2D), _p(depths), _p(tt), _p(offsets), None, None, None)
        keys = np.zeros(max(R, 1), np.uint64)
        vals = np.zeros(max(R, 1), np.uint32)
        self.lib.orc_binning(C.byref(p), _p(radii), _p(means2D), _p(depths), _p(tt), _p(offsets), _p(keys), _p(vals),
                             _p(ranges))
        return dict(R=int(R), point_offsets=offsets[:P], point_list_keys=keys[:R], point_list=vals[:R], ranges=ranges)

    def render_forward(self, p, ranges, point_list, means2D, colors, features, transMat, depths, normal_opacity, bg):
        N = p.W * p.H
        out = dict(final_T=np.zeros((3, p.H, p.W), np.float32), n_contrib=np.zeros((2, p.H, p.W), np.int32),
                   out_color=np.zeros((4, p.H, p.W), np.float32), out_feature=np.zeros((p.S + 3, p.H, p.W), np.float32),
                   out_depth=np.zeros((4, p.H, p.W), np.float32))
        feats = _np(features, np.float32) if p.S > 0 else np.zeros(1, np.float32)
        a = [_np(ranges, np.uint32), _np(point_list, np.uint32), _np(means2D, np.float32), _np(colors, np.float32), feats,
             _np(transMat, np.float32), _np(depths, np.float32), _np(normal_opacity, np.float32), _np(bg, np.float32)]
        self.lib.orc_render_forward(C.byref(p), *[_p(x) for x in a], _p(out["final_T"]), _p(out["n_contrib"]),
                                    _p(out["out_color"]), _p(out["out_feature"]), _p(out["out_depth"]))
        assert N >= 0
        return out

    def render_backward(self, p, ranges, point_list, bg, means2D, normal_opacity, transMat, colors, depths, features,
                        final_T, n_contrib, dL_dpix, dL_ddepth, dL_dmask, dL_dfeat):
        P, S = p.P, p.S
        out = dict(dL_dtransMat=np.zeros((P, 9), np.float32), dL_dmean2D=np.zeros((P, 4), np.float32),
                   dL_dopacity=np.zeros((P, 1), np.float32), dL_dcolors=np.zeros((P, 4), np.float32),
                   dL_dfeatures=np.zeros((P, max(S, 1)), np.float32), dL_dnormals=np.zeros((P, 3), np.float32))
        feats = _np(features, np.float32) if S > 0 else np.zeros(1, np.float32)
        a = [_np(ranges, np.uint32), _np(point_list, np.uint32), _np(bg, np.float32), _np(means2D, np.float32),
             _np(normal_opacity, np.float32), _np(transMat, np.float32), _np(colors, np.float32), _np(depths, np.float32),
             feats, _np(final_T, np.float32), _np(n_contrib, np.int32), _np(dL_dpix, np.float32), _np(dL_ddepth, np.float32),
             _np(dL_dmask, np.float32), _np(dL_dfeat, np.float32)]
        self.lib.orc_render_backward(C.byref(p), *[_p(x) for x in a], _p(out["dL_dtransMat"]), _p(out["dL_dmean2D"]),
                                     _p(out["dL_dopacity"]), _p(out["dL_dcolors"]), _p(out["dL_dfeatures"]),
                                     _p(out["dL_dnormals"]))
        out["dL_dfeatures"] = out["dL_dfeatures"][:, :S]
        return out

    def preprocess_backward(self, p, means3D, scales, rotations, shs, clamped, viewmatrix, campos, radii, transMat,
                            dL_dtransMat, dL_dnormals, dL_dcolors, dL_dmean2D):
        P, M = p.P, p.M
        out = dict(dL_dmeans2D=_np(dL_dmean2D, np.float32).copy(), dL_dmeans3D=np.zeros((P, 3), np.float32),
                   dL_dsh=np.zeros((P, max(M, 1), 4), np.float32), dL_dscales=np.zeros((P, 3), np.float32),
                   dL_drotations=np.zeros((P, 4), np.float32))
        a = [_np(means3D, np.float32), _np(scales, np.float32), _np(rotations, np.float32),
             None if shs is None else _np(shs, np.float32), _np(clamped, np.uint8), _np(viewmatrix, np.float32),
             _np(campos, np.float32), _np(radii, np.int32), _np(transMat, np.float32), _np(dL_dtransMat, np.float32),
             _np(dL_dnormals, np.float32), _np(dL_dcolors, np.float32)]
        self.lib.orc_preprocess_backward(C.byref(p), *[_p(x) for x in a], _p(out["dL_dmeans2D"]), _p(out["dL_dmeans3D"]),
                                         _p(out["dL_dsh"]), _p(out["dL_dscales"]), _p(out["dL_drotations"]))
        out["dL_dsh"] = out["dL_dsh"][:, :M]
        return out

    # ---- whole pipeline ------------------------------------------------------------------------
    def forward(self, p, means3D, scales, rotations, opacities, shs, colors_precomp, features, mask, viewmatrix, campos,
                bg):
        pre = self.preprocess(p, means3D, scales, rotations, opacities, shs, colors_precomp, mask, viewmatrix, campos)
        b = self.binning(p, pre["radii"], pre["means2D"], pre["depths"], pre["tiles_touched"])
        colors = pre["rgb"] if colors_precomp is None else _np(colors_precomp, np.float32)
        r = self.render_forward(p, b["ranges"], b["point_list"], pre["means2D"], colors, features, pre["transMat"],
                                pre["depths"], pre["normal_opacity"], bg)
        st = dict(pre)
        st.update(b)
        st.update(r)
        st["colors"] = colors
        st["out_alpha"] = 1.0 - r["final_T"][0:1]
        return st

    def backward(self, p, st, means3D, scales, rotations, shs, features, viewmatrix, campos, bg, dL_dcolor, dL_ddepth,
                 dL_dalpha, dL_dfeature):
        rb = self.render_backward(p, st["ranges"], st["point_list"], bg, st["means2D"], st["normal_opacity"],
                                  st["transMat"], st["colors"], st["depths"], features, st["final_T"], st["n_contrib"],
                                  dL_dcolor, dL_ddepth, dL_dalpha, dL_dfeature)
        pb = self.preprocess_backward(p, means3D, scales, rotations, shs, st["clamped"], viewmatrix, campos, st["radii"],
                                      st["transMat"], rb["dL_dtransMat"], rb["dL_dnormals"], rb["dL_dcolors"],
                                      rb["dL_dmean2D"])
        out = dict(pb)
        out.update(dL_dcolors=rb["dL_dcolors"], dL_dfeatures=rb["dL_dfeatures"], dL_dopacity=rb["dL_dopacity"],
                   dL_dtransMat=rb["dL_dtransMat"], dL_dnormals=rb["dL_dnormals"])
        return out

    def mark_visible(self, means3D, viewmatrix, projmatrix):
        pts = _np(means3D, np.float32)
        out = np.zeros(pts.shape[0], np.uint8)
        self.lib.orc_mark_visible(pts.shape[0], _p(pts), _p(_np(viewmatrix, np.float32)), _p(_np(projmatrix, np.float32)),
                                  _p(out))
        return out.astype(bool)


class RefCuda:
    """torch front-end of the UNMODIFIED reference CUDA rasterizer (oracle/_ref/libgslidar_ref.so).

    Runs on the legacy default stream like the reference; the caller must be on torch's default stream.
    """

    STATE_NAMES = ["depths", "clamped", "internal_radii", "means2D", "transMat", "normal_opacity", "rgb",
                   "tiles_touched", "point_offsets", "point_list", "point_list_keys", "point_list_unsorted",
                   "point_list_keys_unsorted", "ranges", "accum_alpha", "dL_dtransMat", "dL_dnormals"]

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO + " missing: run oracle/build_ref.sh where /root/reference exists")
        self.lib = C.CDLL(REF_SO)
        self.lib.gslref_create.restype = C.c_void_p
        self.h = C.c_void_p(self.lib.gslref_create())

    def __del__(self):
        try:
            self.lib.gslref_destroy(self.h)
        except Exception:
            pass

    @staticmethod
    def _ptr(t):
        return C.c_void_p(t.data_ptr()) if (t is not None and t.numel() > 0) else C.c_void_p(None)

    def forward(self, scene_args, zero_fill=True, outs=None):
        """scene_args: dict with torch CUDA tensors + scalars (see tests/common.py:ref_args)."""
        import torch
        a = scene_args
        P, S, H, W = a["P"], a["S"], a["H"], a["W"]
        dev = a["means3D"].device
        if outs is None:
            outs = dict(out_contrib=torch.empty((2, H, W), dtype=torch.int32, device=dev),
                        out_color=torch.empty((4, H, W), device=dev), out_feature=torch.empty((S + 3, H, W), device=dev),
                        out_depth=torch.empty((4, H, W), device=dev), out_T=torch.empty((1, H, W), device=dev),
                        radii=torch.empty((max(P, 1),), dtype=torch.int32, device=dev))
        f = C.c_float
        R = self.lib.gslref_forward(
            self.h, P, S, a["D"], a["M"], self._ptr(a["bg"]), W, H, self._ptr(a["means3D"]), self._ptr(a["shs"]),
            self._ptr(a["colors_precomp"]), self._ptr(a["features"]), self._ptr(a["opacities"]), self._ptr(a["scales"]),
            f(1.0), self._ptr(a["rotations"]), C.c_void_p(None), self._ptr(a["mask"]), self._ptr(a["viewmatrix"]),
            self._ptr(a["projmatrix"]), self._ptr(a["campos"]), f(a["tanfovx"]), f(a["tanfovy"]), 0,
            self._ptr(outs["out_contrib"]), self._ptr(outs["out_color"]), self._ptr(outs["out_feature"]),
            self._ptr(outs["out_depth"]), self._ptr(outs["out_T"]), self._ptr(outs["radii"]), 0, f(a["vfov"][0]),
            f(a["vfov"][1]), f(a["hfov"][0]), f(a["hfov"][1]), f(a["scale_factor"]), int(zero_fill))
        if R < 0:
            raise RuntimeError("reference forward failed")
        outs["R"] = int(R)
        return outs

    def backward(self, scene_args, fwd, cot, zero_fill=True, grads=None):
        import torch
        a = scene_args
        P, S, M, H, W = a["P"], a["S"], a["M"], a["H"], a["W"]
        dev = a["means3D"].device
        if grads is None:
            e = lambda *s: torch.empty(s, device=dev)
            grads = dict(dL_dmeans2D=e(P, 4), dL_dopacity=e(P, 1), dL_dcolors=e(P, 4), dL_dmeans3D=e(P, 3),
                         dL_dcov3D=e(P, 6), dL_dsh=e(P, max(M, 1), 4), dL_dfeatures=e(P, max(S, 1)), dL_dscales=e(P, 3),
                         dL_drotations=e(P, 4))
        f = C.c_float
        rc = self.lib.gslref_backward(
            self.h, P, S, a["D"], M, fwd["R"], self._ptr(a["bg"]), W, H, self._ptr(a["means3D"]), self._ptr(a["shs"]),
            self._ptr(a["colors_precomp"]), self._ptr(a["features"]), self._ptr(a["scales"]), f(1.0),
            self._ptr(a["rotations"]), C.c_void_p(None), self._ptr(a["viewmatrix"]), self._ptr(a["projmatrix"]),
            self._ptr(a["campos"]), f(a["tanfovx"]), f(a["tanfovy"]), self._ptr(fwd["radii"]),
            self._ptr(fwd["out_contrib"]), self._ptr(cot["color"]), self._ptr(cot["depth"]), self._ptr(cot["alpha"]),
            self._ptr(cot["feature"]), self._ptr(grads["dL_dmeans2D"]), self._ptr(grads["dL_dopacity"]),
            self._ptr(grads["dL_dcolors"]), self._ptr(grads["dL_dmeans3D"]), self._ptr(grads["dL_dcov3D"]),
            self._ptr(grads["dL_dsh"]), self._ptr(grads["dL_dfeatures"]), self._ptr(grads["dL_dscales"]),
            self._ptr(grads["dL_drotations"]), 0, f(a["vfov"][0]), f(a["vfov"][1]), f(a["hfov"][0]), f(a["hfov"][1]),
            f(a["scale_factor"]), int(zero_fill))
        if rc != 0:
            raise RuntimeError("reference backward failed")
        return grads

    def state(self, P, R, H, W):
        """Copies the reference's internal GeometryState/BinningState/ImageState arrays to torch tensors."""
        import torch
        ptrs = (C.c_void_p * 17)()
        if self.lib.gslref_state(self.h, ptrs) != 0:
            raise RuntimeError("no reference state")
        N = H * W
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        spec = dict(depths=(torch.float32, (P,)), clamped=(torch.uint8, (P, 4)), internal_radii=(torch.int32, (P,)),
                    means2D=(torch.float32, (P, 2)), transMat=(torch.float32, (P, 9)),
                    normal_opacity=(torch.float32, (P, 4)), rgb=(torch.float32, (P, 4)),
                    tiles_touched=(torch.int32, (P,)), point_offsets=(torch.int32, (P,)), point_list=(torch.int32, (R,)),
                    point_list_keys=(torch.int64, (R,)), point_list_unsorted=(torch.int32, (R,)),
                    point_list_keys_unsorted=(torch.int64, (R,)), ranges=(torch.int32, (tiles, 2)),
                    accum_alpha=(torch.float32, (3, H, W)), dL_dtransMat=(torch.float32, (P, 9)),
                    dL_dnormals=(torch.float32, (P, 3)))
        out = {}
        torch.cuda.synchronize()
        for i, name in enumerate(self.STATE_NAMES):
            dt, shape = spec[name]
            t = torch.empty(shape, dtype=dt, device="cuda")
            nbytes = t.numel() * t.element_size()
            if ptrs[i] and nbytes > 0:
                rc = self.lib.gslref_copy(C.c_void_p(t.data_ptr()), C.c_void_p(ptrs[i]), C.c_size_t(nbytes))
                if rc != 0:
                    raise RuntimeError("cudaMemcpy failed for %s: %d" % (name, rc))
            out[name] = t
        assert N > 0
        return out


class RefChamfer:
    """torch front-end of the UNMODIFIED reference Chamfer kernels (oracle/_ref/libchamfer_ref.so): same call sequence as
    chamfer/chamfer3D/dist_chamfer_3D.py:40-85 (zero-filled outputs, forward, zero-filled gradients, backward).  Legacy
    default stream, like the reference."""

    def __init__(self):
        if not os.path.exists(REF_CHAMFER_SO):
            raise FileNotFoundError(REF_CHAMFER_SO + " missing: run oracle/build_ref.sh where /root/reference exists")
        self.lib = C.CDLL(REF_CHAMFER_SO)

    def forward(self, xyz1, xyz2, outs=None):
        import torch
        B, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
        dev = xyz1.device
        if outs is None:
            outs = (torch.zeros((B, n), device=dev), torch.zeros((B, m), device=dev),
                    torch.zeros((B, n), dtype=torch.int32, device=dev), torch.zeros((B, m), dtype=torch.int32, device=dev))
        else:
            for t in outs:
                t.zero_()
        d1, d2, i1, i2 = outs
        p = lambda t: C.c_void_p(t.data_ptr())
        ok = self.lib.chamferref_forward(B, n, p(xyz1), m, p(xyz2), p(d1), p(d2), p(i1), p(i2))
        if ok != 1:
            raise RuntimeError("reference chamfer forward failed")
        return d1, d2, i1, i2

    def backward(self, xyz1, xyz2, g1, g2, i1, i2):
        import torch
        B, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
        gx1, gx2 = torch.zeros_like(xyz1), torch.zeros_like(xyz2)
        p = lambda t: C.c_void_p(t.data_ptr())
        ok = self.lib.chamferref_backward(B, n, p(xyz1), m, p(xyz2), p(gx1), p(gx2), p(g1), p(g2), p(i1), p(i2))
        if ok != 1:
            raise RuntimeError("reference chamfer backward failed")
        return gx1, gx2
