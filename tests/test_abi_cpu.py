"""CPU: the C-ABI library loads, exports every symbol include/gsl_b200.h declares, the ctypes struct
layouts match the header, and argument validation fails loudly -- no compute calls (no GPU here)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gsl_b200.h")


def declared_functions():
    txt = open(HEADER).read()
    return sorted(set(re.findall(r"GSL_API\s+[\w\s\*]+?\b(gsl_\w+)\s*\(", txt)))


def test_header_declares_entry_points():
    fns = declared_functions()
    for must in ("gsl_forward", "gsl_forward_preprocess", "gsl_forward_render", "gsl_backward", "gsl_mark_visible",
                 "gsl_workspace_sizes", "gsl_last_error", "gsl_abi_version", "gsl_export_state"):
        assert must in fns


def test_library_exports_every_declared_symbol():
    from gs_lidar_b200 import _lib as L
    lib = L.load()
    for fn in declared_functions():
        assert hasattr(lib, fn), fn
        assert fn in L.SYMBOLS, "ctypes binding missing for " + fn
    assert lib.gsl_abi_version() == L.GSL_ABI_VERSION


def test_struct_layouts_match_header():
    from gs_lidar_b200 import _lib as L
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "gsl_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(gsl_params), sizeof(gsl_ws_sizes), sizeof(gsl_workspace),
         sizeof(gsl_fwd_inputs), sizeof(gsl_fwd_outputs), sizeof(gsl_bwd_inputs), sizeof(gsl_bwd_outputs),
         sizeof(gsl_state_export));
  printf("%zu %zu %zu %zu\n", offsetof(gsl_params, scale_factor), offsetof(gsl_params, flags),
         offsetof(gsl_workspace, r_capacity), offsetof(gsl_workspace, num_rendered_host));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        cfile, exe = os.path.join(d, "t.c"), os.path.join(d, "t")
        open(cfile, "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), cfile, "-o", exe])
        out = subprocess.check_output([exe]).decode().split()
    sizes = [int(x) for x in out[:8]]
    expect = [C.sizeof(t) for t in (L.gsl_params, L.gsl_ws_sizes, L.gsl_workspace, L.gsl_fwd_inputs, L.gsl_fwd_outputs,
                                    L.gsl_bwd_inputs, L.gsl_bwd_outputs, L.gsl_state_export)]
    assert sizes == expect
    offs = [int(x) for x in out[8:]]
    assert offs == [L.gsl_params.scale_factor.offset, L.gsl_params.flags.offset, L.gsl_workspace.r_capacity.offset,
                    L.gsl_workspace.num_rendered_host.offset]


def test_extension_struct_layouts_match_header():
    """The structs added after ABI 1 (glue, peer exchange) and the fields appended to the ABI-1 structs."""
    from gs_lidar_b200 import _lib as L
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "gsl_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(gsl_peer_ctx), sizeof(gsl_peer_handle), sizeof(gsl_glue_params),
         sizeof(gsl_glue_inputs), sizeof(gsl_glue_outputs), sizeof(gsl_glue_inputs_grad));
  printf("%zu %zu %zu %zu %zu %zu\n", offsetof(gsl_fwd_inputs, shs_rest), offsetof(gsl_bwd_outputs, dL_dsh_rest),
         offsetof(gsl_bwd_outputs, peer), offsetof(gsl_peer_ctx, parity), offsetof(gsl_peer_ctx, buf),
         offsetof(gsl_peer_ctx, error_flag));
  printf("%d %d %u %u\n", GSL_PEER_MAX, GSL_PEER_CAMPOS_OFFSET, GSL_FLAG_BWD_PEER_ROWS, GSL_FLAG_BWD_SH_FACTORED);
  printf("%zu %zu %zu %zu\n", sizeof(gsl_peer_glue), offsetof(gsl_peer_ctx, glue), offsetof(gsl_peer_glue, xyz),
         offsetof(gsl_peer_glue, opacity));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        cfile, exe = os.path.join(d, "t.c"), os.path.join(d, "t")
        open(cfile, "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), cfile, "-o", exe])
        out = [int(x) for x in subprocess.check_output([exe]).decode().split()]
    assert out[:6] == [C.sizeof(t) for t in (L.gsl_peer_ctx, L.gsl_peer_handle, L.gsl_glue_params, L.gsl_glue_inputs,
                                             L.gsl_glue_outputs, L.gsl_glue_inputs_grad)]
    assert out[6:12] == [L.gsl_fwd_inputs.shs_rest.offset, L.gsl_bwd_outputs.dL_dsh_rest.offset, L.gsl_bwd_outputs.peer.offset,
                         L.gsl_peer_ctx.parity.offset, L.gsl_peer_ctx.buf.offset, L.gsl_peer_ctx.error_flag.offset]
    assert out[12:16] == [L.GSL_PEER_MAX, L.GSL_PEER_CAMPOS_OFFSET, L.GSL_FLAG_BWD_PEER_ROWS, L.GSL_FLAG_BWD_SH_FACTORED]
    assert out[16:] == [C.sizeof(L.gsl_peer_glue), L.gsl_peer_ctx.glue.offset, L.gsl_peer_glue.xyz.offset,
                        L.gsl_peer_glue.opacity.offset]
    # the exchange moves S feature channels, or whole quads of them + two quads of glue gradients
    lib = L.load()
    assert [lib.gsl_peer_rows_channels(s_, 0) for s_ in (0, 3, 4)] == [0, 3, 4]
    assert [lib.gsl_peer_rows_channels(s_, 1) for s_ in (0, 3, 4)] == [8, 12, 12]
    assert lib.gsl_peer_row_width(12) == 24


def _params(**kw):
    from gs_lidar_b200 import _lib as L
    p = L.gsl_params()
    p.P, p.S, p.D, p.M, p.W, p.H = 1000, 4, 3, 16, 1030, 66
    p.vfov_min, p.vfov_max, p.hfov_min, p.hfov_max, p.scale_factor = -24.9, 2.0, -180.0, 180.0, 0.1
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def test_workspace_sizes_and_validation():
    from gs_lidar_b200 import _lib as L
    lib = L.load()
    sz = L.gsl_ws_sizes()
    assert lib.gsl_workspace_sizes(C.byref(_params()), 5000, C.byref(sz)) == 0
    # geom: 64-B record + colour + rect + box + counters + packed gradient accumulators per surfel
    assert sz.geom_bytes >= 1000 * (64 + 16 + 8 + 8 + 4 + 4 + 1 + 4 * 24)
    assert sz.binning_bytes >= 5000 * 24
    assert sz.image_bytes >= 66 * 1030 * 12
    sz2 = L.gsl_ws_sizes()
    assert lib.gsl_workspace_sizes(C.byref(_params(P=2000)), 5000, C.byref(sz2)) == 0
    assert sz2.geom_bytes > sz.geom_bytes and sz2.binning_bytes == sz.binning_bytes
    # the reference's hard cap: S + 3 <= 13 feature accumulators (forward.cu:348)
    rc = lib.gsl_workspace_sizes(C.byref(_params(S=11)), 5000, C.byref(sz))
    assert rc == L.GSL_EINVAL and b"features" in lib.gsl_last_error()
    assert lib.gsl_workspace_sizes(C.byref(_params(D=4)), 5000, C.byref(sz)) == L.GSL_EINVAL
    assert lib.gsl_workspace_sizes(C.byref(_params(W=0)), 5000, C.byref(sz)) == L.GSL_EINVAL
    assert lib.gsl_workspace_sizes(C.byref(_params(M=4)), 5000, C.byref(sz)) == L.GSL_EINVAL  # degree 3 needs 16 coeffs
    assert lib.gsl_workspace_sizes(None, 5000, C.byref(sz)) == L.GSL_EINVAL


def test_null_arguments_are_rejected_before_any_launch():
    from gs_lidar_b200 import _lib as L
    lib = L.load()
    p = _params()
    fin, fout, ws = L.gsl_fwd_inputs(), L.gsl_fwd_outputs(), L.gsl_workspace()
    rc = lib.gsl_forward_preprocess(C.byref(p), C.byref(fin), C.byref(fout), C.byref(ws), None)
    assert rc == L.GSL_EINVAL and b"means3D" in lib.gsl_last_error()
    rc = lib.gsl_mark_visible(10, None, None, None, None, None)
    assert rc == L.GSL_EINVAL


def test_peer_exchange_layout_and_validation_without_a_gpu():
    """The exchange-buffer layout is plain arithmetic and every gsl_peer_* entry point validates its context before it
    touches CUDA."""
    from gs_lidar_b200 import _lib as L
    lib = L.load()
    assert lib.gsl_peer_row_width(0) == 16 and lib.gsl_peer_row_width(4) == 16 and lib.gsl_peer_row_width(5) == 24
    P = 1000000
    one, eight = lib.gsl_peer_buffer_bytes(P, 4, 1), lib.gsl_peer_buffer_bytes(P, 4, 8)
    tiles = (P + 255) // 256
    # header + two factor tables per source rank + staging of the owned tiles per source rank + result area
    assert one >= 4096 + 2 * tiles * 256 * 16 + tiles * 256 * 64 + tiles * 256 * 64
    assert eight - one >= 7 * 2 * tiles * 256 * 16            # factor tables grow with the number of source ranks
    assert eight < one + 7 * 2 * tiles * 256 * 16 + 2 * tiles * 256 * 64  # staging does not: 1/8 of the tiles x 8 sources
    assert one % 256 == 0 and eight % 256 == 0
    assert lib.gsl_peer_buffer_bytes(0, 4, 8) >= 4096
    ctx = L.gsl_peer_ctx()
    ctx.rank, ctx.world = 0, 0
    assert lib.gsl_peer_barrier(C.byref(ctx), 0, None) == L.GSL_EINVAL and b"rank/world" in lib.gsl_last_error()
    ctx.world = 2
    assert lib.gsl_peer_reduce(C.byref(ctx), 1000, 4, 0, 1000, None) == L.GSL_EINVAL and b"not mapped" in lib.gsl_last_error()
    assert lib.gsl_peer_sh_expand(None, 1000, 4, 3, 16, 0, 1000, None, None, None) == L.GSL_EINVAL
    assert lib.gsl_peer_unpack(None, 1000, 4, None, None) == L.GSL_EINVAL
    # the peer backward needs a mapped context in gsl_bwd_outputs.peer
    p = _params()
    p.flags = L.GSL_FLAG_BWD_PEER_ROWS
    fin, ffwd, gout, ws = L.gsl_fwd_inputs(), L.gsl_fwd_outputs(), L.gsl_bwd_outputs(), L.gsl_workspace()
    assert lib.gsl_backward_surfels_exchange(C.byref(p), C.byref(fin), C.byref(ffwd), C.byref(gout), C.byref(ws), 1, 1,
                                             None) == L.GSL_EINVAL


def test_python_api_validation_matches_reference_messages():
    import torch
    from gs_lidar_b200 import GaussianRasterizer, synth
    s = synth.make_scene(8)
    rast = GaussianRasterizer(synth.settings_for(s))
    m2 = torch.zeros(8, 4)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        rast(s.means3D, m2, s.opacities, scales=s.scales, rotations=s.rotations)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        rast(s.means3D, m2, s.opacities, shs=s.shs, colors_precomp=torch.zeros(8, 4), scales=s.scales, rotations=s.rotations)
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        rast(s.means3D, m2, s.opacities, shs=s.shs)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rast(s.means3D, m2, s.opacities, shs=s.shs, scales=s.scales, rotations=s.rotations)
    with pytest.raises(RuntimeError, match=r"means3D must have dimensions \(num_points, 3\)"):
        rast(torch.zeros(8, 2), m2, s.opacities, shs=s.shs, scales=s.scales, rotations=s.rotations)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gs_lidar_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "gsl_oracle" not in txt, f
