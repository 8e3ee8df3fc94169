// gsl_api.cu -- extern "C" entry points declared in include/gsl_b200.h.
// Validation, workspace carving and stage orchestration only; all arithmetic lives in the kernels.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include <vector>
#include "gsl_common.cuh"

namespace gsl {

// ---- per-kernel profiling ----------------------------------------------------------------------
struct ProfRec { int id; cudaEvent_t e0, e1; };
static bool g_prof_on = false;
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof_pending;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_free;
static double g_prof_ms[GSL_K_COUNT] = {0};
static int64_t g_prof_n[GSL_K_COUNT] = {0};

ProfScope::ProfScope(int id_, cudaStream_t st_) : id(id_), st(st_), slot(nullptr) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec* r = new ProfRec();
  r->id = id;
  if (!g_prof_free.empty()) {
    r->e0 = g_prof_free.back().first; r->e1 = g_prof_free.back().second;
    g_prof_free.pop_back();
  } else {
    cudaEventCreate(&r->e0); cudaEventCreate(&r->e1);
  }
  cudaEventRecord(r->e0, st);
  slot = r;
}
ProfScope::~ProfScope() {
  if (!slot) return;
  ProfRec* r = (ProfRec*)slot;
  cudaEventRecord(r->e1, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_pending.push_back(*r);
  delete r;
}

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return (int)e;
}

static int validate(const gsl_params* p) {
  if (!p) return set_error(GSL_EINVAL, "params is NULL");
  if (p->P < 0) return set_error(GSL_EINVAL, "P must be >= 0 (got %d)", p->P);
  if (p->S < 0 || p->S > GSL_MAX_FEATURES)
    return set_error(GSL_EINVAL, "features must have 0..%d channels (got %d): the compositor holds S+3 <= 13 "
                     "feature accumulators like the reference", GSL_MAX_FEATURES, p->S);
  if (p->D < 0 || p->D > 3) return set_error(GSL_EINVAL, "sh_degree must be 0..3 (got %d)", p->D);
  if (p->M < 0 || (p->M > 0 && (p->D + 1) * (p->D + 1) > p->M))
    return set_error(GSL_EINVAL, "sh has %d coefficients but degree %d needs %d", p->M, p->D, (p->D + 1) * (p->D + 1));
  if (p->W <= 0 || p->H <= 0) return set_error(GSL_EINVAL, "image size must be positive (got %dx%d)", p->W, p->H);
  if (p->W > 32767 || p->H > 32767) return set_error(GSL_EINVAL, "image side > 32767 not supported");
  if (p->flags & GSL_FLAG_WRAP_AZIMUTH) {
    if (fabsf((p->hfov_max - p->hfov_min) - 360.f) > 1e-3f)
      return set_error(GSL_EINVAL, "azimuth wrap-around needs a 360 degree hfov (got %g .. %g)", p->hfov_min, p->hfov_max);
    if (p->W > 16 * GSL_BIN_GROUP_TILES)
      return set_error(GSL_EINVAL, "azimuth wrap-around supports images of up to %d pixels in width", 16 * GSL_BIN_GROUP_TILES);
  }
  return 0;
}

static int validate_inputs(const gsl_params* p, const gsl_fwd_inputs* in) {
  if (!in) return set_error(GSL_EINVAL, "inputs is NULL");
  if (p->P == 0) return 0;
  if (!in->means3D) return set_error(GSL_EINVAL, "means3D must have dimensions (num_points, 3)");
  if (!in->scales || !in->rotations)
    return set_error(GSL_EINVAL, "scales and rotations are required (the cov3D_precomp path is not computed, as in the reference)");
  if (!in->opacities || !in->mask || !in->viewmatrix || !in->campos || !in->background)
    return set_error(GSL_EINVAL, "opacities, mask, viewmatrix, campos and background are required");
  if ((in->shs == nullptr) == (in->colors_precomp == nullptr))
    return set_error(GSL_EINVAL, "Please provide exactly one of either SHs or precomputed colors!");
  if (in->shs && p->M == 0) return set_error(GSL_EINVAL, "shs given but M == 0");
  if (in->shs_rest && (!in->shs || p->M < 2)) return set_error(GSL_EINVAL, "shs_rest needs shs (coefficient 0) and M >= 2");
  if (p->S > 0 && !in->features) return set_error(GSL_EINVAL, "features is NULL but S = %d", p->S);
  return 0;
}

static int validate_ws(const gsl_params* p, const gsl_workspace* ws, bool need_binning) {
  if (!ws) return set_error(GSL_EINVAL, "workspace is NULL");
  gsl_ws_sizes sz;
  GeomView g = geom_view(nullptr, p->P, p->S);
  ImageView im = image_view(nullptr, p->W, p->H, p->P);
  sz.geom_bytes = g.bytes;
  sz.image_bytes = im.bytes;
  if (!ws->geom || ws->geom_bytes < sz.geom_bytes)
    return set_error(GSL_ENOSPACE, "geom chunk too small: %zu < %zu", ws->geom_bytes, sz.geom_bytes);
  if (!ws->image || ws->image_bytes < sz.image_bytes)
    return set_error(GSL_ENOSPACE, "image chunk too small: %zu < %zu", ws->image_bytes, sz.image_bytes);
  if (need_binning) {
    BinView b = bin_view(nullptr, ws->r_capacity);
    if (!ws->binning || ws->binning_bytes < b.bytes)
      return set_error(GSL_ENOSPACE, "binning chunk too small for capacity %lld: %zu < %zu",
                       (long long)ws->r_capacity, ws->binning_bytes, b.bytes);
  }
  if (!ws->num_rendered_host) return set_error(GSL_EINVAL, "workspace.num_rendered_host (pinned int[2]) is NULL");
  return 0;
}

// One side stream + fork/join events per host thread and device (the surfel sort runs on it).
constexpr int SIDE_CHUNK_EVENTS = 32;
struct SideStream {
  cudaStream_t stream;  // highest priority: few CTAs that must slip in between the main stream's (sort, NVLink pushes)
  cudaStream_t low;     // lowest priority: bulk work that should only fill what the main stream leaves idle
  cudaEvent_t fork, join, join_low;
  cudaEvent_t chunk[SIDE_CHUNK_EVENTS];
};
static SideStream* side_stream() {
  static thread_local SideStream aux[64];
  static thread_local bool have[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!have[dev]) {
    // highest priority: the sort's few CTAs must slip in between the preprocess kernel's CTAs, not queue behind them
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&aux[dev].stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) return nullptr;
    if (cudaStreamCreateWithPriority(&aux[dev].low, cudaStreamNonBlocking, prio_lo) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&aux[dev].join_low, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&aux[dev].fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&aux[dev].join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    for (int k = 0; k < SIDE_CHUNK_EVENTS; ++k)
      if (cudaEventCreateWithFlags(&aux[dev].chunk[k], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    have[dev] = true;
  }
  return &aux[dev];
}

static unsigned long long g_peer_timeout_ns = 20000000000ull;
// schedule of the fused exchange step (gsl_peer_set_option; defaults = what measured best, profiles/r02_exchange_schedules.md)
static int g_peer_early_factors = 0;  // GSL_PEER_OPT_EARLY_FACTORS
static int g_peer_expand_low = 0;     // GSL_PEER_OPT_EXPAND_LOW_PRIORITY
unsigned long long peer_timeout_ns() { return g_peer_timeout_ns; }

static int debug_sync(const gsl_params* p, cudaStream_t st, const char* stage) {
  if (!(p->flags & GSL_FLAG_DEBUG_SYNC)) return 0;
  return check_cuda(cudaStreamSynchronize(st), stage);
}

}  // namespace gsl

using namespace gsl;

extern "C" {

GSL_API int gsl_abi_version(void) { return GSL_ABI_VERSION; }

GSL_API const char* gsl_last_error(void) { return g_err; }

GSL_API int gsl_workspace_sizes(const gsl_params* p, int64_t r_capacity, gsl_ws_sizes* out) {
  int rc = validate(p);
  if (rc) return rc;
  if (!out) return set_error(GSL_EINVAL, "out is NULL");
  if (r_capacity < 0 || r_capacity > 0x7fffffffLL) return set_error(GSL_EINVAL, "r_capacity out of range");
  out->geom_bytes = geom_view(nullptr, p->P, p->S).bytes;
  out->image_bytes = image_view(nullptr, p->W, p->H, p->P).bytes;
  out->binning_bytes = bin_view(nullptr, r_capacity).bytes;
  return 0;
}

GSL_API int32_t gsl_bin_groups(int32_t W, int32_t H, int32_t* groups, int32_t capacity) {
  if (W <= 0 || H <= 0 || W > 32767 || H > 32767 || capacity < 0 || (capacity > 0 && !groups)) {
    set_error(GSL_EINVAL, "bin_groups: bad arguments");
    return -1;
  }
  return bin_groups_describe(W, H, groups, capacity);
}

GSL_API int gsl_forward_preprocess(const gsl_params* p, const gsl_fwd_inputs* in, gsl_fwd_outputs* out,
                           gsl_workspace* ws, void* stream) {
  int rc = validate(p);
  if (rc) return rc;
  if ((rc = validate_inputs(p, in))) return rc;
  if (!out || (p->P > 0 && !out->radii)) return set_error(GSL_EINVAL, "outputs / radii is NULL");
  if ((rc = validate_ws(p, ws, false))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  GeomView g = geom_view(ws->geom, p->P, p->S);
  if (p->P > 0) {
    // fork: depth keys + surfel sort on a side stream UNDER the preprocess kernel; join before binning
    SideStream* aux = side_stream();
    if (!aux) return set_error(GSL_EINVAL, "could not create the side stream");
    cudaEventRecord(aux->fork, st);
    cudaStreamWaitEvent(aux->stream, aux->fork, 0);
    if ((rc = launch_depth_keys(*p, *in, g, aux->stream))) return rc;
    if ((rc = launch_surfel_sort(*p, g, aux->stream))) return rc;
    cudaEventRecord(aux->join, aux->stream);
    if ((rc = launch_preprocess(*p, *in, *out, g, st))) return rc;
    cudaStreamWaitEvent(st, aux->join, 0);
    return debug_sync(p, st, "preprocess + surfel sort");
  }
  if ((rc = launch_preprocess(*p, *in, *out, g, st))) return rc;
  return debug_sync(p, st, "preprocess");
}

GSL_API int gsl_forward_render(const gsl_params* p, const gsl_fwd_inputs* in, gsl_fwd_outputs* out,
                       gsl_workspace* ws, void* stream) {
  int rc = validate(p);
  if (rc) return rc;
  if ((rc = validate_inputs(p, in))) return rc;
  if (!out || !out->out_contrib || !out->out_color || !out->out_feature || !out->out_depth || !out->out_alpha)
    return set_error(GSL_EINVAL, "an output image pointer is NULL");
  if ((rc = validate_ws(p, ws, true))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  GeomView g = geom_view(ws->geom, p->P, p->S);
  ImageView im = image_view(ws->image, p->W, p->H, p->P);
  BinView b = bin_view(ws->binning, ws->r_capacity);
  rc = launch_binning(*p, g, im, b, ws->r_capacity, ws->num_rendered_host, st);
  if (rc == GSL_ENOSPACE)
    return set_error(GSL_ENOSPACE, "binning capacity %lld < num_rendered %d", (long long)ws->r_capacity,
                     ws->num_rendered_host[0]);
  if (rc) return rc;
  if ((rc = debug_sync(p, st, "binning"))) return rc;
  if ((rc = launch_render_forward(*p, *in, *out, g, im, b, ws->r_capacity, st))) return rc;
  return debug_sync(p, st, "render_forward");
}

GSL_API int gsl_wait_num_rendered(gsl_workspace* ws, int32_t* num_rendered, void* stream) {
  if (!ws || !ws->num_rendered_host) return set_error(GSL_EINVAL, "workspace / num_rendered_host is NULL");
  int rc = wait_num_rendered(ws->num_rendered_host, (cudaStream_t)stream);
  if (rc) return rc;
  if (num_rendered) *num_rendered = ws->num_rendered_host[0];
  return 0;
}

GSL_API int gsl_forward(const gsl_params* p, const gsl_fwd_inputs* in, gsl_fwd_outputs* out, gsl_workspace* ws,
                int32_t* num_rendered, void* stream) {
  int rc = gsl_forward_preprocess(p, in, out, ws, stream);
  if (rc) return rc;
  rc = gsl_forward_render(p, in, out, ws, stream);
  if (rc && rc != GSL_ENOSPACE) return rc;
  int32_t R = 0;
  int rc2 = gsl_wait_num_rendered(ws, &R, stream);
  if (rc2) return rc2;
  if (num_rendered) *num_rendered = R;
  if (rc == GSL_ENOSPACE || (int64_t)R > ws->r_capacity)
    return set_error(GSL_ENOSPACE, "binning capacity %lld < num_rendered %d", (long long)ws->r_capacity, R);
  return 0;
}

static int validate_peer(const gsl_peer_ctx* c) {
  if (!c) return set_error(GSL_EINVAL, "peer: ctx is NULL");
  if (c->world < 1 || c->world > GSL_PEER_MAX || c->rank < 0 || c->rank >= c->world)
    return set_error(GSL_EINVAL, "peer: bad rank/world %d/%d (at most %d ranks)", c->rank, c->world, GSL_PEER_MAX);
  for (int g = 0; g < c->world; ++g)
    if (!c->buf[g]) return set_error(GSL_EINVAL, "peer: buffer of rank %d is not mapped", g);
  if (c->glue && (!c->glue->xyz || !c->glue->velocity || !c->glue->t || !c->glue->scaling_t || !c->glue->opacity || !(c->glue->cycle > 0.f)))
    return set_error(GSL_EINVAL, "peer: glue needs xyz, velocity, t, scaling_t, opacity and a positive cycle");
  return 0;
}

static int backward_validate(const gsl_params* p, const gsl_fwd_inputs* in, const gsl_fwd_outputs* fwd,
                             const gsl_bwd_outputs* gout, gsl_workspace* ws) {
  int rc = validate(p);
  if (rc) return rc;
  if ((rc = validate_inputs(p, in))) return rc;
  if (!fwd || !gout) return set_error(GSL_EINVAL, "fwd / grad outputs is NULL");
  if (p->P == 0) return 0;
  if (p->flags & GSL_FLAG_BWD_PEER_ROWS) {
    if (!in->shs) return set_error(GSL_EINVAL, "GSL_FLAG_BWD_PEER_ROWS needs the SH colour path");
    if ((rc = validate_peer(gout->peer))) return rc;
    if (gout->peer->glue && p->S > 4)
      return set_error(GSL_EINVAL, "GSL_FLAG_BWD_PEER_ROWS with the glue's VJP folded in supports S <= 4 (got %d)", p->S);
  } else
  if (!gout->dL_dmeans3D || !gout->dL_dmeans2D || !gout->dL_dcolors || !gout->dL_dopacity ||
      !gout->dL_dscales || !gout->dL_drotations || (p->S > 0 && !gout->dL_dfeatures) ||
      (in->shs && !gout->dL_dsh && !(p->flags & GSL_FLAG_BWD_SH_FACTORED)) ||
      (in->shs_rest && !gout->dL_dsh_rest && !(p->flags & GSL_FLAG_BWD_SH_FACTORED)))
    return set_error(GSL_EINVAL, "a gradient output pointer is NULL");
  if (!fwd->radii || !fwd->out_contrib) return set_error(GSL_ESTATE, "forward outputs (radii, out_contrib) missing");
  return validate_ws(p, ws, true);
}

// true while the zero-fill forked by gsl_backward_composite has not been joined by gsl_backward_surfels (per thread)
static thread_local bool g_prezero_pending = false;

GSL_API int gsl_backward_composite(const gsl_params* p, const gsl_fwd_inputs* in, const gsl_fwd_outputs* fwd,
                           const gsl_bwd_inputs* gin, gsl_bwd_outputs* gout, gsl_workspace* ws, float* sh_factor_out,
                           void* stream) {
  int rc = backward_validate(p, in, fwd, gout, ws);
  if (rc) return rc;
  if (!gin) return set_error(GSL_EINVAL, "grad inputs is NULL");
  g_prezero_pending = false;
  if (p->P == 0) return 0;
  if (!gin->dL_dout_color || !gin->dL_dout_depth || !gin->dL_dout_alpha || !gin->dL_dout_feature)
    return set_error(GSL_EINVAL, "a cotangent pointer is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  GeomView g = geom_view(ws->geom, p->P, p->S);
  ImageView im = image_view(ws->image, p->W, p->H, p->P);
  BinView b = bin_view(ws->binning, ws->r_capacity);
  // fork: the dense gradient outputs are zero-filled on the side stream while the compositor (which touches only
  // the packed accumulators, and is not DRAM-bound) runs; the per-surfel kernel then writes non-zero rows only
  SideStream* aux = side_stream();
#ifdef GSL_NO_PREZERO
  aux = nullptr;
#endif
  {
    // Captured into a CUDA graph the fork costs more than it hides (measured at 1M surfels: 0.925 ms / step with the
    // forked fill against 0.895 ms with the per-surfel kernel writing its own zeros; on plain streams 0.884 against 0.888),
    // so a captured single-GPU backward skips it.  The fused exchange keeps it: its kernels write non-zero rows only.
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusActive &&
        !(p->flags & GSL_FLAG_BWD_PEER_ROWS))
      aux = nullptr;
  }
  if (p->flags & GSL_FLAG_BWD_PEER_ROWS) {
    // the dense outputs given here (the targets of gsl_backward_surfels_exchange) are zero-filled on the side stream;
    // without them (the piecewise calls) there is nothing to fill
    if (!gout->dL_dsh && !gout->dL_dmeans3D && !gout->dL_dcov3D) aux = nullptr;
    if (sh_factor_out) return set_error(GSL_EINVAL, "GSL_FLAG_BWD_PEER_ROWS: the SH factors are pushed by the per-surfel kernel; sh_factor_out must be NULL");
  }
  if (aux) {
    cudaEventRecord(aux->fork, st);
    cudaStreamWaitEvent(aux->stream, aux->fork, 0);
    if ((rc = launch_zero_outputs(*p, *in, *gout, aux->stream))) return rc;
    cudaEventRecord(aux->join, aux->stream);
    g_prezero_pending = true;
  }
  if ((rc = launch_render_backward(*p, *in, *fwd, *gin, g, im, b, ws->r_capacity, st))) return rc;
  if ((rc = debug_sync(p, st, "render_backward"))) return rc;
  if (sh_factor_out) {
    if ((rc = launch_extract_sh_factor(*p, g, sh_factor_out, st))) return rc;
    return debug_sync(p, st, "extract_sh_factor");
  }
  return 0;
}

GSL_API int gsl_backward_surfels(const gsl_params* p, const gsl_fwd_inputs* in, const gsl_fwd_outputs* fwd,
                         gsl_bwd_outputs* gout, gsl_workspace* ws, void* stream) {
  return gsl_backward_surfels_rows(p, in, fwd, gout, ws, 0, p ? p->P : 0, stream);
}

GSL_API int gsl_backward_surfels_rows(const gsl_params* p, const gsl_fwd_inputs* in, const gsl_fwd_outputs* fwd,
                              gsl_bwd_outputs* gout, gsl_workspace* ws, int32_t row_begin, int32_t row_end, void* stream) {
  int rc = backward_validate(p, in, fwd, gout, ws);
  if (rc) return rc;
  if (row_begin < 0 || row_end > p->P || row_begin > row_end || (row_begin & 255))
    return set_error(GSL_EINVAL, "backward_surfels_rows: bad row range [%d, %d)", row_begin, row_end);
  if ((row_begin != 0 || row_end != p->P) && !(p->flags & GSL_FLAG_BWD_PEER_ROWS))
    return set_error(GSL_EINVAL, "backward_surfels_rows: partial ranges need GSL_FLAG_BWD_PEER_ROWS");
  const bool prezeroed = g_prezero_pending;
  g_prezero_pending = false;
  if (p->P == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  GeomView g = geom_view(ws->geom, p->P, p->S);
  if (prezeroed) {
    SideStream* aux = side_stream();
    if (!aux) return set_error(GSL_ESTATE, "side stream lost between the two backward stages");
    cudaStreamWaitEvent(st, aux->join, 0);
  }
  if ((rc = launch_preprocess_backward(*p, *in, *fwd, *gout, g, prezeroed, row_begin, row_end, st))) return rc;
  return debug_sync(p, st, "preprocess_backward");
}

// The whole second half of a frame-parallel backward pass in one call (see include/gsl_b200.h).
GSL_API int gsl_backward_surfels_exchange(const gsl_params* p, const gsl_fwd_inputs* in, const gsl_fwd_outputs* fwd,
                                  gsl_bwd_outputs* gout, gsl_workspace* ws, uint32_t step, int32_t chunks, void* stream) {
  int rc = backward_validate(p, in, fwd, gout, ws);
  if (rc) return rc;
  if (!(p->flags & GSL_FLAG_BWD_PEER_ROWS)) return set_error(GSL_EINVAL, "backward_surfels_exchange needs GSL_FLAG_BWD_PEER_ROWS");
  if (p->P == 0) return 0;
  if (!gout->dL_dmeans3D || !gout->dL_dmeans2D || !gout->dL_dopacity || !gout->dL_dscales || !gout->dL_drotations ||
      (p->S > 0 && !gout->dL_dfeatures) || !gout->dL_dsh)
    return set_error(GSL_EINVAL, "backward_surfels_exchange: a dense gradient output pointer is NULL");
  (void)step; (void)chunks;  // see include/gsl_b200.h: the ticket is a device-side step counter, one row range
  SideStream* aux = side_stream();
  if (!aux) return set_error(GSL_ESTATE, "no side stream");
  cudaStream_t st = (cudaStream_t)stream;
  GeomView g = geom_view(ws->geom, p->P, p->S);
  // The dense outputs are zero-filled on the side stream under the backward compositor (gsl_backward_composite); the
  // expansion below runs on that stream behind the fill, the unpack on this stream waits for its event here.
  bool prezeroed = g_prezero_pending;
  g_prezero_pending = false;
  if (prezeroed) {
    cudaStreamWaitEvent(st, aux->join, 0);
  } else {
    if ((rc = launch_zero_outputs(*p, *in, *gout, st))) return rc;
    prezeroed = true;
  }
  const gsl_peer_ctx* ctx = gout->peer;
  const int P = p->P;
  const int S_rows = peer_rows_S(p->S, ctx);  // with the glue folded in: dL_dfeatures is (P, S_rows), see gsl_peer_glue
  // One step (gsl_peer.cuh); no kernel of it blocks the stream waiting for other ranks before it has launched:
  //   k_peer_begin            (one warp) step counter++, camera centre -> every rank
  //   k_peer_factor_extract   SH factors of this rank -> own factor table (before the accumulators are re-zeroed)
  //   k_preprocess_bwd        VJP; packed rows -> tile owners        | side stream, behind an event:
  //   k_peer_signal           (one warp) publishes "pushed"          | k_peer_factor_push: own factors -> every rank
  //   k_peer_reduce_rows      every CTA waits for "pushed" of all    | k_peer_signal: publishes "factors"
  //                           ranks; sums of my tiles -> every rank; | k_peer_sh_expand_tiles: waits for "factors",
  //                           last CTA publishes "summed"            | dL_dsh from the local factor tables
  //   k_peer_unpack           every CTA waits for "summed"; dense tensors
  // The factor tables are half of the step's NVLink bytes and are complete before the per-surfel kernel starts: pushed
  // from the side stream they cross the links UNDER that kernel, and the expansion no longer waits for anybody's rows.
  const bool early = g_peer_early_factors != 0;
  if (early && ctx->glue) return set_error(GSL_EINVAL, "backward_surfels_exchange: the early-factor schedule does not carry the glue's VJP");
  if ((rc = launch_peer_begin(ctx, in->campos, st))) return rc;
  if (early) {
    if ((rc = launch_peer_factor_extract(ctx, *p, g, st))) return rc;
  } else {
    if ((rc = launch_preprocess_backward(*p, *in, *fwd, *gout, g, false, 0, P, st, true, false))) return rc;
    if ((rc = launch_peer_signal_fused(ctx, PEER_SLOT_PUSHED, st))) return rc;
  }
  cudaEventRecord(aux->chunk[0], st);
  cudaStream_t xs = aux->stream;  // the stream of the expansion
  const bool low = early && g_peer_expand_low;
  if (early) {
    cudaStreamWaitEvent(aux->stream, aux->chunk[0], 0);
    {
      ProfScope prof(GSL_K_PEER_FACTORS, aux->stream);
      if ((rc = launch_peer_factor_push(ctx, *p, aux->stream))) return rc;
      if ((rc = launch_peer_signal_fused(ctx, PEER_SLOT_FACTORS, aux->stream))) return rc;
    }
    cudaEventRecord(aux->join, aux->stream);
    if (low) {
      // The expansion must not be resident before this rank's own "factors" flag is out (its CTAs would spin on a flag
      // that a kernel of this very GPU has yet to publish), and it should not hold SMs while it waits for the other ranks:
      // behind the signal's event, one warp waits for everybody's factors, then the bulk kernel starts.
      xs = aux->low;
      cudaStreamWaitEvent(xs, aux->join, 0);
      if ((rc = launch_peer_wait_fused(ctx, PEER_SLOT_FACTORS, xs))) return rc;
    }
  } else {
    cudaStreamWaitEvent(xs, aux->chunk[0], 0);
  }
  {
    ProfScope prof(GSL_K_PEER_EXPAND, xs);
    if ((rc = launch_peer_sh_expand(ctx, P, S_rows, p->D, p->M, 0, P, prezeroed, in->means3D, gout->dL_dsh, xs, true,
                                    early ? PEER_SLOT_FACTORS : PEER_SLOT_PUSHED)))
      return rc;
  }
  cudaEventRecord(aux->join_low, xs);
  if (early) {
    if ((rc = launch_preprocess_backward(*p, *in, *fwd, *gout, g, false, 0, P, st, true, true))) return rc;
    if ((rc = launch_peer_signal_fused(ctx, PEER_SLOT_PUSHED, st))) return rc;
  }
  {
    ProfScope prof(GSL_K_PEER_REDUCE, st);
    if (low && (rc = launch_peer_wait_fused(ctx, PEER_SLOT_PUSHED, st))) return rc;
    if ((rc = launch_peer_reduce_rows(ctx, P, S_rows, 0, P, st, true))) return rc;
  }
  {
    ProfScope prof(GSL_K_PEER_UNPACK, st);
    if (low && (rc = launch_peer_wait_fused(ctx, PEER_SLOT_SUMMED, st))) return rc;
    if ((rc = launch_peer_unpack(ctx, P, S_rows, prezeroed, *gout, st, true))) return rc;
  }
  if (early) cudaStreamWaitEvent(st, aux->join, 0);  // the factor pushes of this rank have left
  cudaStreamWaitEvent(st, aux->join_low, 0);         // dL_dsh complete
  return debug_sync(p, st, "backward_surfels_exchange");
}

GSL_API int gsl_backward(const gsl_params* p, const gsl_fwd_inputs* in, const gsl_fwd_outputs* fwd,
                 const gsl_bwd_inputs* gin, gsl_bwd_outputs* gout, gsl_workspace* ws, void* stream) {
  int rc = gsl_backward_composite(p, in, fwd, gin, gout, ws, nullptr, stream);
  if (rc) return rc;
  return gsl_backward_surfels(p, in, fwd, gout, ws, stream);
}

GSL_API int gsl_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, void* stream) {
  if (P < 0) return set_error(GSL_EINVAL, "P must be >= 0");
  if (P > 0 && (!means3D || !viewmatrix || !projmatrix || !present))
    return set_error(GSL_EINVAL, "mark_visible: NULL pointer");
  return launch_mark_visible(P, means3D, viewmatrix, projmatrix, present, (cudaStream_t)stream);
}

GSL_API int gsl_sh_expand(int32_t P, int32_t D, int32_t M, int32_t G, const float* means3D, const float* campos_all,
                  const float* drgb_all, size_t drgb_stride, float* dL_dsh, void* stream) {
  if (P < 0 || D < 0 || D > 3 || M < 0 || G < 1) return set_error(GSL_EINVAL, "sh_expand: bad sizes");
  if (M > 0 && (D + 1) * (D + 1) > M) return set_error(GSL_EINVAL, "sh_expand: degree %d needs %d coefficients", D, (D + 1) * (D + 1));
  if (P > 0 && M > 0 && (!means3D || !campos_all || !drgb_all || !dL_dsh))
    return set_error(GSL_EINVAL, "sh_expand: NULL pointer");
  if (drgb_stride < (size_t)4 * (size_t)P) return set_error(GSL_EINVAL, "sh_expand: drgb_stride < 4 P");
  return launch_sh_expand(P, D, M, G, means3D, campos_all, drgb_all, drgb_stride, dL_dsh, (cudaStream_t)stream);
}

// ---- peer-memory gradient exchange (gsl_peer.cu) -----------------------------------------------------------------
GSL_API size_t gsl_peer_buffer_bytes(int64_t P, int32_t S, int32_t world) {
  return peer_layout((size_t)(P < 0 ? 0 : P), S, world).total;
}
GSL_API int32_t gsl_peer_row_width(int32_t S) { return peer_row_width(S); }
GSL_API int32_t gsl_peer_rows_channels(int32_t S, int32_t with_glue) { return with_glue ? 4 * ((S + 3) / 4) + 8 : S; }

GSL_API int gsl_peer_alloc(size_t bytes, void** dptr, gsl_peer_handle* handle) {
  static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(gsl_peer_handle), "handle size");
  if (!dptr || bytes == 0) return set_error(GSL_EINVAL, "peer_alloc: bad arguments");
  void* p = nullptr;
  int rc = check_cuda(cudaMalloc(&p, bytes), "peer_alloc cudaMalloc");
  if (rc) return rc;
  rc = check_cuda(cudaMemset(p, 0, bytes), "peer_alloc memset");
  if (!rc && handle) {
    cudaIpcMemHandle_t h;
    rc = check_cuda(cudaIpcGetMemHandle(&h, p), "cudaIpcGetMemHandle");
    if (!rc) {
      memset(handle, 0, sizeof(*handle));
      memcpy(handle, &h, sizeof(h));
    }
  }
  if (rc) {
    cudaFree(p);
    return rc;
  }
  *dptr = p;
  return 0;
}

GSL_API int gsl_peer_open(const gsl_peer_handle* handle, void** dptr) {
  if (!handle || !dptr) return set_error(GSL_EINVAL, "peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  return check_cuda(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
}

GSL_API int gsl_peer_set_option(int32_t option, int32_t value) {
  switch (option) {
    case GSL_PEER_OPT_EARLY_FACTORS: g_peer_early_factors = value ? 1 : 0; return 0;
    case GSL_PEER_OPT_EXPAND_LOW_PRIORITY: g_peer_expand_low = value ? 1 : 0; return 0;
    default: return set_error(GSL_EINVAL, "peer_set_option: unknown option %d", option);
  }
}

GSL_API int gsl_peer_set_timeout_ms(uint32_t ms) {
  g_peer_timeout_ns = (unsigned long long)(ms ? ms : 1u) * 1000000ull;
  return 0;
}

GSL_API int gsl_peer_close(void* dptr) { return dptr ? check_cuda(cudaIpcCloseMemHandle(dptr), "cudaIpcCloseMemHandle") : 0; }
GSL_API int gsl_peer_free(void* dptr) { return dptr ? check_cuda(cudaFree(dptr), "peer_free") : 0; }

static int peer_barrier(const gsl_peer_ctx* ctx, int32_t phase, int mode, void* stream) {
  int rc = validate_peer(ctx);
  if (rc) return rc;
  if (phase < 0 || phase > 3) return set_error(GSL_EINVAL, "peer_barrier: phase must be 0..3");
  return launch_peer_barrier(ctx, phase, mode, (cudaStream_t)stream);
}
GSL_API int gsl_peer_barrier(const gsl_peer_ctx* ctx, int32_t phase, void* stream) { return peer_barrier(ctx, phase, 3, stream); }
GSL_API int gsl_peer_signal(const gsl_peer_ctx* ctx, int32_t phase, void* stream) { return peer_barrier(ctx, phase, 1, stream); }
GSL_API int gsl_peer_wait(const gsl_peer_ctx* ctx, int32_t phase, void* stream) { return peer_barrier(ctx, phase, 2, stream); }

static int validate_rows(const char* what, int32_t P, int32_t S, int32_t row_begin, int32_t row_end) {
  if (P < 0 || S < 0 || S > 12 || row_begin < 0 || row_end > P || row_begin > row_end || (row_begin & 255) ||
      ((row_end & 255) && row_end != P))
    return set_error(GSL_EINVAL, "%s: bad sizes / row range [%d, %d) of %d", what, row_begin, row_end, P);
  return 0;
}

static int peer_sh_expand(const gsl_peer_ctx* ctx, int32_t P, int32_t S, int32_t D, int32_t M, int32_t row_begin,
                          int32_t row_end, const float* means3D, float* dL_dsh, void* stream, bool sparse);
GSL_API int gsl_peer_sh_expand(const gsl_peer_ctx* ctx, int32_t P, int32_t S, int32_t D, int32_t M, int32_t row_begin,
                               int32_t row_end, const float* means3D, float* dL_dsh, void* stream) {
  return peer_sh_expand(ctx, P, S, D, M, row_begin, row_end, means3D, dL_dsh, stream, false);
}
GSL_API int gsl_peer_sh_expand_sparse(const gsl_peer_ctx* ctx, int32_t P, int32_t S, int32_t D, int32_t M, int32_t row_begin,
                                      int32_t row_end, const float* means3D, float* dL_dsh, void* stream) {
  if (M > 16) return set_error(GSL_EINVAL, "peer_sh_expand_sparse: at most 16 SH coefficients");
  return peer_sh_expand(ctx, P, S, D, M, row_begin, row_end, means3D, dL_dsh, stream, true);
}
static int peer_sh_expand(const gsl_peer_ctx* ctx, int32_t P, int32_t S, int32_t D, int32_t M, int32_t row_begin,
                          int32_t row_end, const float* means3D, float* dL_dsh, void* stream, bool sparse) {
  int rc = validate_peer(ctx);
  if (rc) return rc;
  if ((rc = validate_rows("peer_sh_expand", P, S, row_begin, row_end))) return rc;
  if (D < 0 || D > 3 || M < 0) return set_error(GSL_EINVAL, "peer_sh_expand: bad sizes");
  if (M > 0 && (D + 1) * (D + 1) > M) return set_error(GSL_EINVAL, "peer_sh_expand: degree %d needs %d coefficients", D, (D + 1) * (D + 1));
  if (P > 0 && M > 0 && (!means3D || !dL_dsh)) return set_error(GSL_EINVAL, "peer_sh_expand: NULL pointer");
  return launch_peer_sh_expand(ctx, P, S, D, M, row_begin, row_end, sparse, means3D, dL_dsh, (cudaStream_t)stream);
}

GSL_API int gsl_peer_reduce(const gsl_peer_ctx* ctx, int32_t P, int32_t S, int32_t row_begin, int32_t row_end, void* stream) {
  int rc = validate_peer(ctx);
  if (rc) return rc;
  if ((rc = validate_rows("peer_reduce", P, S, row_begin, row_end))) return rc;
  return launch_peer_reduce_rows(ctx, P, S, row_begin, row_end, (cudaStream_t)stream);
}

GSL_API int gsl_peer_unpack(const gsl_peer_ctx* ctx, int32_t P, int32_t S, const gsl_bwd_outputs* out, void* stream) {
  int rc = validate_peer(ctx);
  if (rc) return rc;
  if (P < 0 || S < 0 || S > 12) return set_error(GSL_EINVAL, "peer_unpack: bad sizes");
  if (P > 0 && (!out || !out->dL_dmeans3D || !out->dL_dmeans2D || !out->dL_dscales || !out->dL_drotations ||
                !out->dL_dopacity || (S > 0 && !out->dL_dfeatures)))
    return set_error(GSL_EINVAL, "peer_unpack: an output pointer is NULL");
  return launch_peer_unpack(ctx, P, S, false, *out, (cudaStream_t)stream);
}

// ---- Chamfer distance (gsl_chamfer.cu) ------------------------------------------------------------------------------
static int validate_chamfer(int32_t b, int32_t n, int32_t m, const void* xyz1, const void* xyz2) {
  if (b < 0 || n < 0 || m < 0 || b > 65535) return set_error(GSL_EINVAL, "chamfer: bad sizes b=%d n=%d m=%d", b, n, m);
  if ((size_t)b * (size_t)n > 0x7fffffffull || (size_t)b * (size_t)m > 0x7fffffffull)
    return set_error(GSL_EINVAL, "chamfer: b * n and b * m must fit 31 bits");
  if (b > 0 && ((n > 0 && !xyz1) || (m > 0 && !xyz2))) return set_error(GSL_EINVAL, "chamfer: a point set pointer is NULL");
  return 0;
}

GSL_API size_t gsl_chamfer_scratch_bytes(int32_t b, int32_t n, int32_t m) {
  if (b < 0 || n < 0 || m < 0) return 0;
  return 8 * (size_t)b * ((size_t)n + (size_t)m);
}

GSL_API int gsl_chamfer_forward(int32_t b, int32_t n, const float* xyz1, int32_t m, const float* xyz2, float* dist1,
                                int32_t* idx1, float* dist2, int32_t* idx2, void* scratch, void* stream) {
  int rc = validate_chamfer(b, n, m, xyz1, xyz2);
  if (rc) return rc;
  if (b == 0 || (n == 0 && m == 0)) return 0;
  if ((n > 0 && (!dist1 || !idx1)) || (m > 0 && (!dist2 || !idx2)) || !scratch)
    return set_error(GSL_EINVAL, "chamfer_forward: an output / scratch pointer is NULL");
  return launch_chamfer_forward(b, n, xyz1, m, xyz2, dist1, idx1, dist2, idx2, scratch, (cudaStream_t)stream);
}

GSL_API int gsl_chamfer_backward(int32_t b, int32_t n, const float* xyz1, int32_t m, const float* xyz2,
                                 const float* gdist1, const int32_t* idx1, const float* gdist2, const int32_t* idx2,
                                 float* gxyz1, float* gxyz2, void* stream) {
  int rc = validate_chamfer(b, n, m, xyz1, xyz2);
  if (rc) return rc;
  if (b == 0 || (n == 0 && m == 0)) return 0;
  if ((n > 0 && (!gdist1 || !idx1 || !gxyz1)) || (m > 0 && (!gdist2 || !idx2 || !gxyz2)))
    return set_error(GSL_EINVAL, "chamfer_backward: a pointer is NULL");
  return launch_chamfer_backward(b, n, xyz1, m, xyz2, gdist1, idx1, gdist2, idx2, gxyz1, gxyz2, (cudaStream_t)stream);
}

static int validate_glue(const gsl_glue_params* p, const gsl_glue_inputs* in) {
  if (!p || !in) return set_error(GSL_EINVAL, "glue: params / inputs is NULL");
  if (p->P < 0) return set_error(GSL_EINVAL, "glue: P must be >= 0");
  if (!(p->cycle > 0.f)) return set_error(GSL_EINVAL, "glue: cycle (GaussianModel.T) must be positive");
  if (p->P > 0 && (!in->xyz || !in->velocity || !in->t || !in->scaling_t || !in->opacity || !in->scaling || !in->rotation))
    return set_error(GSL_EINVAL, "glue: a raw parameter pointer is NULL");
  return 0;
}

GSL_API int gsl_glue_forward(const gsl_glue_params* p, const gsl_glue_inputs* in, const gsl_glue_outputs* out, void* stream) {
  int rc = validate_glue(p, in);
  if (rc) return rc;
  if (!out || (p->P > 0 && (!out->means3D || !out->opacity || !out->scales || !out->rotations || !out->mask)))
    return set_error(GSL_EINVAL, "glue: an output pointer is NULL");
  return launch_glue_forward(*p, *in, *out, (cudaStream_t)stream);
}

GSL_API int gsl_glue_backward(const gsl_glue_params* p, const gsl_glue_inputs* in, const gsl_glue_outputs* gout,
                      const gsl_glue_inputs_grad* gin, void* stream) {
  int rc = validate_glue(p, in);
  if (rc) return rc;
  if (!gout || !gin) return set_error(GSL_EINVAL, "glue: cotangents / gradient outputs is NULL");
  if (p->P > 0 && (!gin->xyz || !gin->velocity || !gin->t || !gin->scaling_t || !gin->opacity || !gin->scaling || !gin->rotation))
    return set_error(GSL_EINVAL, "glue: a gradient output pointer is NULL");
  return launch_glue_backward(*p, *in, *gout, *gin, (cudaStream_t)stream);
}

// ---- panorama post-ops (gsl_postops.cu) ---------------------------------------------------------------------------------
GSL_API size_t gsl_pano_scratch_bytes(int32_t H, int32_t W) {
  const long long n = (long long)(H > 0 ? H : 0) * (W > 0 ? W : 0);
  return (size_t)((n + 255) / 256 + 1) * sizeof(int32_t);
}

static int validate_pano(const gsl_pano_params* p, const float* range) {
  if (!p) return set_error(GSL_EINVAL, "pano: params is NULL");
  if (p->H < 0 || p->W < 0 || (long long)p->H * p->W > 0x7fffffffLL) return set_error(GSL_EINVAL, "pano: bad image size %d x %d", p->H, p->W);
  if ((long long)p->H * p->W > 0 && !range) return set_error(GSL_EINVAL, "pano: range is NULL");
  return 0;
}

GSL_API int gsl_pano_forward(const gsl_pano_params* p, const float* range, float* points, int32_t* index, int32_t* count,
                             float* normals, void* scratch, void* stream) {
  int rc = validate_pano(p, range);
  if (rc) return rc;
  if (!points && !normals) return set_error(GSL_EINVAL, "pano_forward: no output requested");
  if (points && (!count || !scratch)) return set_error(GSL_EINVAL, "pano_forward: points need count and scratch");
  return launch_pano_forward(*p, range, points, index, count, normals, scratch, (cudaStream_t)stream);
}

GSL_API int gsl_pano_backward(const gsl_pano_params* p, const float* range, int32_t K, const float* g_points,
                              const int32_t* index, const float* g_normals, float* g_range, void* stream) {
  int rc = validate_pano(p, range);
  if (rc) return rc;
  if (!g_range && (long long)p->H * p->W > 0) return set_error(GSL_EINVAL, "pano_backward: g_range is NULL");
  if (K < 0 || (K > 0 && g_points && !index)) return set_error(GSL_EINVAL, "pano_backward: g_points need the point index");
  return launch_pano_backward(*p, range, K, g_points, index, g_normals, g_range, (cudaStream_t)stream);
}

// ---- CUDA-graph capture of a step -----------------------------------------------------------
// The forward's instance count is polled AFTER all launches and every kernel argument of a step is constant once the
// workspace, the outputs and the inputs keep their addresses, so a forward (gsl_forward_preprocess + gsl_forward_render)
// and a backward (gsl_backward, or gsl_backward_composite + gsl_backward_surfels_exchange) can each be captured once and
// replayed: one launch per pass instead of ~14 / ~4-9, no per-launch host cost, no launch jitter between the ranks.
__global__ void k_stage_camera(const float* __restrict__ vm, const float* __restrict__ campos, const float* __restrict__ bg,
                               float* __restrict__ dst) {
  const int t = threadIdx.x;
  if (t < 16) dst[t] = vm[t];
  else if (t < 19) dst[t] = campos[t - 16];
  else if (t >= 20 && t < 24) dst[t] = bg[t - 20];
}

GSL_API int gsl_stage_camera(const float* viewmatrix, const float* campos, const float* background, float* dst, void* stream) {
  if (!viewmatrix || !campos || !background || !dst) return set_error(GSL_EINVAL, "stage_camera: NULL pointer");
  k_stage_camera<<<1, 32, 0, (cudaStream_t)stream>>>(viewmatrix, campos, background, dst);
  return check_cuda(cudaGetLastError(), "k_stage_camera launch");
}

// The legacy default stream (what torch uses unless told otherwise) cannot be captured, so the capture runs on a stream the
// library owns (one per host thread and device); the instantiated graph can be launched on any stream, the default included.
static cudaStream_t capture_stream() {
  static thread_local cudaStream_t cap[64];
  static thread_local bool have[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!have[dev]) {
    // middle priority: the captured "main stream" kernels then rank above the low side stream (bulk work that should only fill
    // idle SMs) and below the high one (torch's own streams have the lowest priority, so on plain streams low == main)
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&cap[dev], cudaStreamNonBlocking, (prio_lo + prio_hi) / 2) != cudaSuccess) return nullptr;
    have[dev] = true;
  }
  return cap[dev];
}

GSL_API int gsl_graph_begin(void** capture_stream_out) {
  if (!capture_stream_out) return set_error(GSL_EINVAL, "graph_begin: NULL argument");
  if (g_prof_on) return set_error(GSL_ESTATE, "graph capture: switch per-kernel profiling off first (gsl_profile_enable(0))");
  if (!side_stream()) return set_error(GSL_ESTATE, "graph capture: could not create the side stream");
  cudaStream_t cap = capture_stream();
  if (!cap) return set_error(GSL_ESTATE, "graph capture: could not create the capture stream");
  int rc = check_cuda(cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal), "cudaStreamBeginCapture");
  if (rc) return rc;
  *capture_stream_out = (void*)cap;
  return 0;
}

GSL_API int gsl_graph_end(void* stream, void** graph_exec) {
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture((cudaStream_t)stream, &graph);
  if (e != cudaSuccess || !graph) {
    cudaGetLastError();
    return check_cuda(e != cudaSuccess ? e : cudaErrorUnknown, "cudaStreamEndCapture");
  }
  if (!graph_exec) {  // abort: the caller only wanted the stream out of capture mode
    cudaGraphDestroy(graph);
    return 0;
  }
#ifndef GSL_GRAPH_NO_PRIO
  {  // the side stream's priority is not captured with its kernels: re-apply it to the surfel sort's nodes
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    size_t n = 0;
    if (cudaGraphGetNodes(graph, nullptr, &n) == cudaSuccess && n > 0) {
      std::vector<cudaGraphNode_t> nodes(n);
      if (cudaGraphGetNodes(graph, nodes.data(), &n) == cudaSuccess) {
        for (size_t i = 0; i < n; ++i) {
          cudaGraphNodeType type;
          if (cudaGraphNodeGetType(nodes[i], &type) != cudaSuccess || type != cudaGraphNodeTypeKernel) continue;
          cudaKernelNodeParams kp;
          if (cudaGraphKernelNodeGetParams(nodes[i], &kp) != cudaSuccess) continue;
          if (is_sort_kernel(kp.func) || is_depth_keys_kernel(kp.func)) {
            cudaLaunchAttributeValue v;
            memset(&v, 0, sizeof(v));
            v.priority = prio_hi;
            cudaGraphKernelNodeSetAttribute(nodes[i], cudaLaunchAttributePriority, &v);
          }
        }
      }
    }
    cudaGetLastError();
  }
#endif
  cudaGraphExec_t exec = nullptr;
  e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return check_cuda(e, "cudaGraphInstantiate");
  *graph_exec = (void*)exec;
  return 0;
}

GSL_API int gsl_graph_launch(void* graph_exec, void* stream) {
  if (!graph_exec) return set_error(GSL_EINVAL, "graph_launch: NULL graph");
  return check_cuda(cudaGraphLaunch((cudaGraphExec_t)graph_exec, (cudaStream_t)stream), "cudaGraphLaunch");
}

GSL_API int gsl_graph_destroy(void* graph_exec) {
  return graph_exec ? check_cuda(cudaGraphExecDestroy((cudaGraphExec_t)graph_exec), "cudaGraphExecDestroy") : 0;
}

GSL_API int gsl_profile_enable(int on) { g_prof_on = on != 0; return 0; }

GSL_API int gsl_profile_read(double* total_ms, int64_t* launches, int reset) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof_pending) {
    cudaEventSynchronize(r.e1);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) { g_prof_ms[r.id] += ms; g_prof_n[r.id] += 1; }
    g_prof_free.push_back({r.e0, r.e1});
  }
  g_prof_pending.clear();
  for (int i = 0; i < GSL_K_COUNT; ++i) {
    if (total_ms) total_ms[i] = g_prof_ms[i];
    if (launches) launches[i] = g_prof_n[i];
    if (reset) { g_prof_ms[i] = 0; g_prof_n[i] = 0; }
  }
  return 0;
}

GSL_API const char* gsl_kernel_name(int id) {
  static const char* names[GSL_K_COUNT] = {"k_preprocess_fwd", "k_bin_(count|scan)", "k_bin_scatter",
                                           "k_sort_(hist|scan|scatter|buckets)", "k_tile_blists", "k_render_fwd",
                                           "k_render_bwd", "k_preprocess_bwd", "k_glue_fwd", "k_glue_bwd",
                                           "k_peer_reduce_rows", "k_peer_sh_expand", "k_peer_unpack", "k_peer_factor_push"};
  return (id >= 0 && id < GSL_K_COUNT) ? names[id] : "?";
}

// ---- state export in the reference's layouts (tests only) ----------------------------------------
__global__ void k_export_geom(int P, const float4* __restrict__ rec, const float4* __restrict__ rgb,
                              const uint8_t* __restrict__ clamped, const short4* __restrict__ pixbox,
                              gsl_state_export d) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  float4 a = rec[4 * (size_t)i], b = rec[4 * (size_t)i + 1], c = rec[4 * (size_t)i + 2], e = rec[4 * (size_t)i + 3];
  if (d.depths) d.depths[i] = e.w;
  if (d.means2D) { d.means2D[2 * i] = c.y; d.means2D[2 * i + 1] = c.z; }
  if (d.transMat) {
    float* t = d.transMat + 9 * (size_t)i;
    t[0] = a.x; t[1] = a.y; t[2] = a.z; t[3] = a.w; t[4] = b.x; t[5] = b.y; t[6] = b.z; t[7] = b.w; t[8] = c.x;
  }
  if (d.normal_opacity) {
    float* t = d.normal_opacity + 4 * (size_t)i;
    t[0] = e.x; t[1] = e.y; t[2] = e.z; t[3] = c.w;
  }
  if (d.rgb) reinterpret_cast<float4*>(d.rgb)[i] = rgb[i];
  if (d.clamped) {
    uint8_t cl = clamped[i];
    for (int k = 0; k < 4; ++k) d.clamped[4 * (size_t)i + k] = (cl >> k) & 1;
  }
  if (d.pixbox) {
    short4 pb = pixbox[i];
    d.pixbox[4 * (size_t)i] = pb.x; d.pixbox[4 * (size_t)i + 1] = pb.y;
    d.pixbox[4 * (size_t)i + 2] = pb.z; d.pixbox[4 * (size_t)i + 3] = pb.w;
  }
}

GSL_API int gsl_export_state(const gsl_params* p, const gsl_workspace* ws, int64_t R, const gsl_state_export* dst,
                     void* stream) {
  int rc = validate(p);
  if (rc) return rc;
  if (!ws || !dst) return set_error(GSL_EINVAL, "workspace / dst is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  GeomView g = geom_view(ws->geom, p->P, p->S);
  ImageView im = image_view(ws->image, p->W, p->H, p->P);
  if (p->P > 0) {
    if (dst->point_offsets) {  // the render path does not need the tiles_touched scan; run it for the export
      if ((rc = launch_scan(*p, g, nullptr, st))) return rc;
    }
    k_export_geom<<<(p->P + 255) / 256, 256, 0, st>>>(p->P, g.rec, g.rgb, g.clamped, g.pixbox, *dst);
    if (dst->tiles_touched)
      cudaMemcpyAsync(dst->tiles_touched, g.tiles, (size_t)p->P * 4, cudaMemcpyDeviceToDevice, st);
    if (dst->point_offsets)
      cudaMemcpyAsync(dst->point_offsets, g.offs, (size_t)p->P * 4, cudaMemcpyDeviceToDevice, st);
  }
  if (R > 0 && ws->binning) {
    if (R > ws->r_capacity) return set_error(GSL_EINVAL, "R exceeds the binning capacity");
    BinView b = bin_view(ws->binning, ws->r_capacity);
    if (dst->point_list_keys) {
      if ((rc = launch_export_keys(*p, g, im, b.vals_b, dst->point_list_keys, st))) return rc;
    }
    if (dst->point_list) cudaMemcpyAsync(dst->point_list, b.vals_b, (size_t)R * 4, cudaMemcpyDeviceToDevice, st);
  }
  const size_t tiles = (size_t)((p->W + GSL_BLOCK_X - 1) / GSL_BLOCK_X) * ((p->H + GSL_BLOCK_Y - 1) / GSL_BLOCK_Y);
  if (dst->ranges) cudaMemcpyAsync(dst->ranges, im.ranges, tiles * 8, cudaMemcpyDeviceToDevice, st);
  if (dst->final_T)
    cudaMemcpyAsync(dst->final_T, im.final_T, (size_t)3 * p->W * p->H * 4, cudaMemcpyDeviceToDevice, st);
  return check_cuda(cudaGetLastError(), "export_state");
}

}  // extern "C"
