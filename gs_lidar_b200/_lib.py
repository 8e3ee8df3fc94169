"""ctypes binding of libgsl_b200.so (declared in include/gsl_b200.h).

There is deliberately NO fallback here: if the CUDA library is missing the import of the
rasterizer fails loudly with instructions to build it (python -m gs_lidar_b200.build).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GSL_B200_LIB", os.path.join(HERE, "libgsl_b200.so"))  # override: dev builds only

GSL_ABI_VERSION = 4
GSL_EINVAL, GSL_ENOSPACE, GSL_ESTATE = -1, -2, -3
GSL_FLAG_DEBUG_SYNC = 1
GSL_FLAG_BWD_SH_FACTORED = 2
GSL_FLAG_WRAP_AZIMUTH = 4
GSL_FLAG_BWD_PEER_ROWS = 8
GSL_MAX_FEATURES = 10

vp = C.c_void_p


class gsl_params(C.Structure):
    _fields_ = [("P", C.c_int32), ("S", C.c_int32), ("D", C.c_int32), ("M", C.c_int32),
                ("W", C.c_int32), ("H", C.c_int32), ("tanfovx", C.c_float), ("tanfovy", C.c_float),
                ("scale_modifier", C.c_float), ("vfov_min", C.c_float), ("vfov_max", C.c_float),
                ("hfov_min", C.c_float), ("hfov_max", C.c_float), ("scale_factor", C.c_float),
                ("prefiltered", C.c_int32), ("flags", C.c_uint32)]


class gsl_ws_sizes(C.Structure):
    _fields_ = [("geom_bytes", C.c_size_t), ("binning_bytes", C.c_size_t), ("image_bytes", C.c_size_t)]


class gsl_workspace(C.Structure):
    _fields_ = [("geom", vp), ("geom_bytes", C.c_size_t), ("binning", vp), ("binning_bytes", C.c_size_t),
                ("image", vp), ("image_bytes", C.c_size_t), ("r_capacity", C.c_int64),
                ("num_rendered_host", vp)]


class gsl_fwd_inputs(C.Structure):
    _fields_ = [(n, vp) for n in ("background", "means3D", "shs", "colors_precomp", "features", "opacities",
                                  "scales", "rotations", "cov3D_precomp", "mask", "viewmatrix", "projmatrix",
                                  "campos", "shs_rest")]


class gsl_fwd_outputs(C.Structure):
    _fields_ = [(n, vp) for n in ("out_contrib", "out_color", "out_feature", "out_depth", "out_alpha", "radii")]


class gsl_bwd_inputs(C.Structure):
    _fields_ = [(n, vp) for n in ("dL_dout_color", "dL_dout_depth", "dL_dout_alpha", "dL_dout_feature")]


class gsl_bwd_outputs(C.Structure):
    _fields_ = [(n, vp) for n in ("dL_dmeans3D", "dL_dmeans2D", "dL_dsh", "dL_dcolors", "dL_dfeatures",
                                  "dL_dopacity", "dL_dscales", "dL_drotations", "dL_dcov3D", "dL_dsh_rest",
                                  "peer")]


class gsl_state_export(C.Structure):
    _fields_ = [(n, vp) for n in ("depths", "means2D", "transMat", "normal_opacity", "rgb", "clamped",
                                  "tiles_touched", "point_offsets", "point_list_keys", "point_list", "ranges",
                                  "final_T", "pixbox")]


class gsl_glue_params(C.Structure):
    _fields_ = [("P", C.c_int32), ("timestamp", C.c_float), ("time_shift", C.c_float), ("cycle", C.c_float),
                ("velocity_decay", C.c_float), ("dynamic", C.c_int32)]


class gsl_pano_params(C.Structure):
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("vfov_min", C.c_float), ("vfov_max", C.c_float),
                ("hfov_min", C.c_float), ("hfov_max", C.c_float)]


class gsl_glue_inputs(C.Structure):
    _fields_ = [(n, vp) for n in ("xyz", "velocity", "t", "scaling_t", "opacity", "scaling", "rotation", "mask")]


class gsl_glue_outputs(C.Structure):
    _fields_ = [(n, vp) for n in ("means3D", "opacity", "scales", "rotations", "marginal_t", "mask")]


class gsl_glue_inputs_grad(C.Structure):
    _fields_ = [(n, vp) for n in ("xyz", "velocity", "t", "scaling_t", "opacity", "scaling", "rotation")]


GSL_PEER_MAX = 8
GSL_PEER_CAMPOS_OFFSET = 1024


class gsl_peer_handle(C.Structure):
    _fields_ = [("reserved", C.c_ubyte * 64)]


class gsl_peer_glue(C.Structure):
    _fields_ = [("timestamp", C.c_float), ("time_shift", C.c_float), ("cycle", C.c_float), ("velocity_decay", C.c_float),
                ("dynamic", C.c_int32), ("xyz", vp), ("velocity", vp), ("t", vp), ("scaling_t", vp), ("opacity", vp)]


class gsl_peer_ctx(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("epoch", C.c_uint32), ("parity", C.c_uint32),
                ("buf", vp * GSL_PEER_MAX), ("error_flag", vp), ("glue", C.POINTER(gsl_peer_glue))]


# name -> (restype, argtypes); every symbol include/gsl_b200.h declares
SYMBOLS = {
    "gsl_abi_version": (C.c_int, []),
    "gsl_last_error": (C.c_char_p, []),
    "gsl_workspace_sizes": (C.c_int, [C.POINTER(gsl_params), C.c_int64, C.POINTER(gsl_ws_sizes)]),
    "gsl_bin_groups": (C.c_int32, [C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int32]),
    "gsl_forward_preprocess": (C.c_int, [C.POINTER(gsl_params), C.POINTER(gsl_fwd_inputs),
                                         C.POINTER(gsl_fwd_outputs), C.POINTER(gsl_workspace), vp]),
    "gsl_forward_render": (C.c_int, [C.POINTER(gsl_params), C.POINTER(gsl_fwd_inputs),
                                     C.POINTER(gsl_fwd_outputs), C.POINTER(gsl_workspace), vp]),
    "gsl_wait_num_rendered": (C.c_int, [C.POINTER(gsl_workspace), C.POINTER(C.c_int32), vp]),
    "gsl_forward": (C.c_int, [C.POINTER(gsl_params), C.POINTER(gsl_fwd_inputs), C.POINTER(gsl_fwd_outputs),
                              C.POINTER(gsl_workspace), C.POINTER(C.c_int32), vp]),
    "gsl_backward": (C.c_int, [C.POINTER(gsl_params), C.POINTER(gsl_fwd_inputs), C.POINTER(gsl_fwd_outputs),
                               C.POINTER(gsl_bwd_inputs), C.POINTER(gsl_bwd_outputs),
                               C.POINTER(gsl_workspace), vp]),
    "gsl_backward_composite": (C.c_int, [C.POINTER(gsl_params), C.POINTER(gsl_fwd_inputs), C.POINTER(gsl_fwd_outputs),
                                         C.POINTER(gsl_bwd_inputs), C.POINTER(gsl_bwd_outputs),
                                         C.POINTER(gsl_workspace), vp, vp]),
    "gsl_backward_surfels": (C.c_int, [C.POINTER(gsl_params), C.POINTER(gsl_fwd_inputs), C.POINTER(gsl_fwd_outputs),
                                       C.POINTER(gsl_bwd_outputs), C.POINTER(gsl_workspace), vp]),
    "gsl_pano_scratch_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "gsl_pano_forward": (C.c_int, [C.POINTER(gsl_pano_params), vp, vp, vp, vp, vp, vp, vp]),
    "gsl_pano_backward": (C.c_int, [C.POINTER(gsl_pano_params), vp, C.c_int32, vp, vp, vp, vp, vp]),
    "gsl_glue_forward": (C.c_int, [C.POINTER(gsl_glue_params), C.POINTER(gsl_glue_inputs), C.POINTER(gsl_glue_outputs), vp]),
    "gsl_glue_backward": (C.c_int, [C.POINTER(gsl_glue_params), C.POINTER(gsl_glue_inputs), C.POINTER(gsl_glue_outputs),
                                    C.POINTER(gsl_glue_inputs_grad), vp]),
    "gsl_mark_visible": (C.c_int, [C.c_int32, vp, vp, vp, vp, vp]),
    "gsl_sh_expand": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp, C.c_size_t, vp, vp]),
    "gsl_peer_buffer_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32]),
    "gsl_peer_row_width": (C.c_int32, [C.c_int32]),
    "gsl_peer_rows_channels": (C.c_int32, [C.c_int32, C.c_int32]),
    "gsl_peer_alloc": (C.c_int, [C.c_size_t, C.POINTER(vp), C.POINTER(gsl_peer_handle)]),
    "gsl_peer_open": (C.c_int, [C.POINTER(gsl_peer_handle), C.POINTER(vp)]),
    "gsl_peer_close": (C.c_int, [vp]),
    "gsl_peer_set_timeout_ms": (C.c_int, [C.c_uint32]),
    "gsl_peer_set_option": (C.c_int, [C.c_int32, C.c_int32]),
    "gsl_peer_free": (C.c_int, [vp]),
    "gsl_peer_barrier": (C.c_int, [C.POINTER(gsl_peer_ctx), C.c_int32, vp]),
    "gsl_peer_signal": (C.c_int, [C.POINTER(gsl_peer_ctx), C.c_int32, vp]),
    "gsl_peer_wait": (C.c_int, [C.POINTER(gsl_peer_ctx), C.c_int32, vp]),
    "gsl_peer_sh_expand": (C.c_int, [C.POINTER(gsl_peer_ctx), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_int32, vp, vp, vp]),
    "gsl_peer_sh_expand_sparse": (C.c_int, [C.POINTER(gsl_peer_ctx), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                            C.c_int32, vp, vp, vp]),
    "gsl_peer_reduce": (C.c_int, [C.POINTER(gsl_peer_ctx), C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]),
    "gsl_peer_unpack": (C.c_int, [C.POINTER(gsl_peer_ctx), C.c_int32, C.c_int32, C.POINTER(gsl_bwd_outputs), vp]),
    "gsl_backward_surfels_exchange": (C.c_int, [C.POINTER(gsl_params), C.POINTER(gsl_fwd_inputs), C.POINTER(gsl_fwd_outputs),
                                                C.POINTER(gsl_bwd_outputs), C.POINTER(gsl_workspace), C.c_uint32, C.c_int32, vp]),
    "gsl_backward_surfels_rows": (C.c_int, [C.POINTER(gsl_params), C.POINTER(gsl_fwd_inputs), C.POINTER(gsl_fwd_outputs),
                                            C.POINTER(gsl_bwd_outputs), C.POINTER(gsl_workspace), C.c_int32, C.c_int32, vp]),
    "gsl_chamfer_scratch_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "gsl_chamfer_forward": (C.c_int, [C.c_int32, C.c_int32, vp, C.c_int32, vp, vp, vp, vp, vp, vp, vp]),
    "gsl_chamfer_backward": (C.c_int, [C.c_int32, C.c_int32, vp, C.c_int32, vp, vp, vp, vp, vp, vp, vp, vp]),
    "gsl_export_state": (C.c_int, [C.POINTER(gsl_params), C.POINTER(gsl_workspace), C.c_int64,
                                   C.POINTER(gsl_state_export), vp]),
    "gsl_graph_begin": (C.c_int, [C.POINTER(vp)]),
    "gsl_graph_end": (C.c_int, [vp, C.POINTER(vp)]),
    "gsl_graph_launch": (C.c_int, [vp, vp]),
    "gsl_graph_destroy": (C.c_int, [vp]),
    "gsl_stage_camera": (C.c_int, [vp, vp, vp, vp, vp]),
    "gsl_profile_enable": (C.c_int, [C.c_int]),
    "gsl_profile_read": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int]),
    "gsl_kernel_name": (C.c_char_p, [C.c_int]),
}
GSL_K_COUNT = 14
GSL_PEER_OPT_EARLY_FACTORS, GSL_PEER_OPT_EXPAND_LOW_PRIORITY = 0, 1
# kernels of THIS repo launched per forward / backward call for images of up to 1024 tiles (one tile group; no library
# kernel is launched at any size): used by bench.py for "gpu_launches".
OWN_LAUNCHES_FWD = 1 + 4 + 1 + 2 + 1 + 1 + 1   # depth keys, sort hist/scan/scatter/buckets, preprocess, bin count/scan,
                                            # bin scatter, tile block lists, render_fwd
OWN_LAUNCHES_BWD = 1 + 1               # render_bwd, preprocess_bwd



def OWN_LAUNCHES_PEER(ranges=1, schedule="default"):
    """extra launches of a backward with the fused peer-memory exchange: k_peer_begin, k_peer_signal, k_peer_reduce_rows,
    k_peer_sh_expand_tiles, k_peer_unpack (the waiting halves of the barriers live inside the last three); with early factors
    also k_peer_factor_extract, k_peer_factor_push and a second k_peer_signal; with the low-priority expansion three
    one-warp k_peer_wait."""
    return 5 + (3 if schedule.startswith("early") else 0) + (3 if schedule == "early-low" else 0)


_lib = None


def load():
    """dlopen the library and bind every symbol; raises if the extension is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "gs_lidar_b200: %s is missing. This package has no CPU or PyTorch fallback; build the sm_100a "
            "library with `python -m gs_lidar_b200.build` (needs nvcc)." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    ver = lib.gsl_abi_version()
    if ver != GSL_ABI_VERSION:
        raise ImportError("libgsl_b200.so ABI version %d != expected %d; rebuild" % (ver, GSL_ABI_VERSION))
    _lib = lib
    return lib


def last_error():
    return load().gsl_last_error().decode("utf-8", "replace")


def check(rc, what):
    if rc == 0:
        return
    msg = last_error()
    if rc > 0:
        raise RuntimeError("%s failed with CUDA error %d: %s" % (what, rc, msg))
    raise RuntimeError("%s: %s" % (what, msg))
