// gsl_common.cuh -- shared definitions for the sm_100a surfel rasterizer kernels.
//
// HBM layout of the three scratch chunks (all offsets 256-B aligned), replacing the reference's
// GeometryState / ImageState / BinningState (cuda_rasterizer/rasterizer_impl.h:26-63):
//
//   geom chunk (per surfel i):
//     rec[4*i .. 4*i+3]  float4 x4 = one 64-B record gathered by the compositors:
//         rec0 = (Tu.x, Tu.y, Tu.z, Tv.x)      Tu,Tv,Tw = rows of the reference's transMat
//         rec1 = (Tv.y, Tv.z, Tw.x, Tw.y)      (forward.cu:238-241)
//         rec2 = (Tw.z, xy.x, xy.y, opacity)   xy = means2D (forward.cu:252-253)
//         rec3 = (n.x, n.y, n.z, depth)        normal_opacity.xyz + depths (forward.cu:282-285)
//     rgb[i]      float4   SH colour (forward.cu:275-278); unused with colors_precomp
//     rect[i]     ushort4  tile rect (min.x, min.y, max.x, max.y) of getRect (auxiliary.h:47-55)
//     pixbox[i]   short4   conservative pixel box of the surfel's support (this design)
//     tiles[i]    u32      tiles_touched;  offs[i] u32 inclusive scan (state export only) /
//                          scratch of the surfel sort
//     skey_a/b, sval_a/b u32  surfel depth sort: keys, bucket-scattered (key, id), ids in (depth, id) order
//     clamped[i]  u8       bit c set when SH channel c was clamped (forward.cu:64-67)
//     grad[i]     float[32] packed gradient accumulators, kept all-zero between steps:
//         [0..8] dL_dtransMat, [9..10] dL_dmean2D.xy, [11] dL_dopacity,
//         [12..15] dL_dcolor, [16..18] dL_dnormal, [19] pad, [20..20+S) dL_dfeature, pad to 32
//     ctrl        u32[64]  [0]=R, [1]=overflow, [2]=finished CTAs of k_bin_scan, [8]=max(~depth key), [9]=max(depth key)
//   image chunk: final_T (3N f32: T, M1, M2), ranges (tiles x uint2),
//     bdesc uint4[tiles*8]  per 8x4 pixel block (tile, b): (start, end) of its block list inside plane b,
//                     .z = number of leading entries the backward pass has to walk (written by the forward)
//     hist u32[min(tiles, 1024)][ceil(P/256)], bintotal u32[min(tiles, 1024)]  counting pass of the binning
//   binning chunk (capacity Rcap): vals_b u32 = point_list (sorted surfel ids);
//     blist uint2[8][Rcap]  BLOCK LISTS: plane b holds, for tile t at [ranges[t].x, ...), in list order, the
//                     (surfel id, list position) of every entry of t whose conservative pixel box overlaps
//                     8x4 pixel block b = (row/4)*2 + col/8 of the tile -- the list of block (t, b)
//     pairmask u32[8][Rcap] per block-list entry: which of the block's 32 pixels the entry contributed to
//                     in the forward pass (the backward pass evaluates exactly those pairs)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/gsl_b200.h"

#define GSL_BLOCK_X 16
#define GSL_BLOCK_Y 16
#define GSL_MY_PI 3.14159265  // auxiliary.h:17 (a double literal, NOT M_PI)

namespace gsl {

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// 32 floats (128 B) per surfel whatever S is: the backward compositor's warp reduction leaves component c of
// the record in lane c, so one 32-lane reduction instruction flushes a whole record.
inline __host__ __device__ int grad_stride(int /*S*/) { return 32; }

// The binning is "depth-sort the surfels once, then one stable counting pass per tile" (gsl_binning.cu).  The
// counting pass keeps a [tile][256 surfels] bitmap in shared memory (36 bytes per tile), so it takes this many
// consecutive 16x16 tiles at a time; larger images are binned group after group.
#define GSL_BIN_GROUP_TILES 1024
// most buckets of the surfel depth sort (gsl_sort.cu): ~32-64 surfels per bucket up to 4M surfels
#define GSL_SORT_MAX_BUCKETS 65536
inline __host__ __device__ int tile_count(int W, int H) {
  return ((W + GSL_BLOCK_X - 1) / GSL_BLOCK_X) * ((H + GSL_BLOCK_Y - 1) / GSL_BLOCK_Y);
}
// groups of a gx x gy tile grid: whole tile rows, or pieces of one row when a row alone exceeds a group
inline int bin_group_count(int gx, int gy) {
  if (gx <= GSL_BIN_GROUP_TILES) {
    const int rows = GSL_BIN_GROUP_TILES / gx;
    return (gy + rows - 1) / rows;
  }
  return gy * ((gx + GSL_BIN_GROUP_TILES - 1) / GSL_BIN_GROUP_TILES);
}

struct GeomView {
  float4* rec;
  float4* rgb;
  ushort4* rect;
  short4* pixbox;
  uint32_t* tiles;
  uint32_t* offs;
  uint8_t* clamped;
  float* grad;
  uint32_t* ctrl;
  uint32_t* scan_state;  // per-CTA sums of the tiles_touched scan
  uint32_t* skey_a;      // surfel depth sort (gsl_sort.cu): depth-bit keys,
  uint32_t* skey_b;      //   keys / ids scattered into their buckets,
  uint32_t* sval_a;
  uint32_t* sval_b;      //   surfel ids in (depth, id) order
  uint32_t* sort_buckets;  // count / start arrays of the bucket sort, 2 x (GSL_SORT_MAX_BUCKETS + 64)
  uint8_t* touched;      // 1 = some pixel composited this surfel (set by k_render_fwd, consumed and cleared by the backward):
                         // the per-surfel backward reads the 128-byte accumulator of a surfel only when it is set
  size_t bytes;
};

struct ImageView {
  float* final_T;   // 3N
  uint2* ranges;    // tiles
  uint4* bdesc;     // tiles * 8
  uint32_t* hist;   // [tiles of a group][ncta]: instances of tile t emitted by surfel chunk c, then its exclusive prefix
  uint32_t* bintotal;  // [tiles of a group]
  size_t ncta;      // surfel chunks of 256 (depth-rank order)
  size_t bytes;
};

struct BinView {
  uint32_t* vals_b;     // point_list
  uint2* blist;
  uint32_t* pairmask;
  size_t plane_stride;  // entries per plane of blist / pairmask
  size_t bytes;
};

template <typename T>
inline void carve(char*& p, T*& out, size_t count) {
  size_t a = align_up((size_t)(uintptr_t)p, 256);
  out = (T*)a;
  p = (char*)(a + count * sizeof(T));
}

inline GeomView geom_view(void* base, int P, int S) {
  GeomView g;
  char* p = (char*)base;
  size_t Pp = (size_t)(P > 0 ? P : 1);
  carve(p, g.rec, 4 * Pp);
  carve(p, g.rgb, Pp);
  carve(p, g.rect, Pp);
  carve(p, g.pixbox, Pp);
  carve(p, g.tiles, Pp);
  carve(p, g.offs, Pp);
  carve(p, g.clamped, Pp);
  carve(p, g.grad, Pp * grad_stride(S));
  carve(p, g.ctrl, 64);
  carve(p, g.scan_state, 2 * ((Pp + 1023) / 1024 + 1));
  carve(p, g.skey_a, Pp);
  carve(p, g.skey_b, Pp);
  carve(p, g.sval_a, Pp);
  carve(p, g.sval_b, Pp);
  carve(p, g.sort_buckets, 2 * (GSL_SORT_MAX_BUCKETS + 64));
  carve(p, g.touched, Pp);
  g.bytes = (size_t)(p - (char*)base) + 256;
  return g;
}

inline ImageView image_view(void* base, int W, int H, int P) {
  ImageView v;
  char* p = (char*)base;
  size_t N = (size_t)W * H;
  size_t tiles = (size_t)tile_count(W, H);
  carve(p, v.final_T, 3 * N);
  carve(p, v.ranges, tiles + 1);
  carve(p, v.bdesc, 8 * tiles + 8);
  v.ncta = ((size_t)(P > 0 ? P : 1) + 255) / 256;
  const size_t group = tiles < (size_t)GSL_BIN_GROUP_TILES ? tiles : (size_t)GSL_BIN_GROUP_TILES;
  carve(p, v.hist, group * v.ncta);
  carve(p, v.bintotal, group + 1);
  v.bytes = (size_t)(p - (char*)base) + 256;
  return v;
}

inline BinView bin_view(void* base, int64_t Rcap) {
  BinView b;
  char* p = (char*)base;
  size_t R = (size_t)(Rcap > 0 ? Rcap : 1);
  carve(p, b.vals_b, R);
  b.plane_stride = align_up(R, 64) + 64;
  carve(p, b.blist, 8 * b.plane_stride);
  carve(p, b.pairmask, 8 * b.plane_stride);
  b.bytes = (size_t)(p - (char*)base) + 256;
  return b;
}

// FOV constants, evaluated in double then rounded exactly as forward.cu:221-226 does per thread.
struct Fov {
  float VFOV_min, VFOV_max, HFOV_min, HFOV_max;
};
inline Fov make_fov(const gsl_params& p) {
  Fov f;
  f.VFOV_max = (float)(GSL_MY_PI / 2 - p.vfov_min * GSL_MY_PI / 180);
  f.VFOV_min = (float)(GSL_MY_PI / 2 - p.vfov_max * GSL_MY_PI / 180);
  f.HFOV_max = (float)(p.hfov_max * GSL_MY_PI / 180);
  f.HFOV_min = (float)(p.hfov_min * GSL_MY_PI / 180);
  return f;
}

// Launch constants of the two compositing kernels.
struct RenderParams {
  int W, H, gx, gy, S;
  float VFOV_min, VFOV_max, HFOV_min, HFOV_max;
  float near_, far_, far_over_range;  // 2*sf, 300*sf, far/(far-near) (forward.cu:366-367,453)
  float wrapW;                        // azimuth wrap-around mode: period in pixels (= W), else 0
  uint32_t r_capacity;
};
RenderParams make_render_params(const gsl_params& p, int64_t r_capacity);

// Optional per-kernel event timing (gsl_api.cu); a no-op unless gsl_profile_enable(1) was called.
struct ProfScope {
  int id;
  cudaStream_t st;
  void* slot;
  ProfScope(int id, cudaStream_t st);
  ~ProfScope();
};

// error plumbing (gsl_api.cu)
int set_error(int code, const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

// stage launchers (one per .cu file)
int launch_preprocess(const gsl_params& p, const gsl_fwd_inputs& in, gsl_fwd_outputs& out,
                      const GeomView& g, cudaStream_t st);
int launch_scan(const gsl_params& p, const GeomView& g, int32_t* r_host, cudaStream_t st);
int launch_depth_keys(const gsl_params& p, const gsl_fwd_inputs& in, const GeomView& g, cudaStream_t st);
int launch_surfel_sort(const gsl_params& p, const GeomView& g, cudaStream_t st);
int bin_groups_describe(int W, int H, int32_t* out, int capacity);
uint32_t sort_num_buckets_host(int P);
bool is_sort_kernel(const void* func);        // kernels that run with the side stream's (high) priority
bool is_depth_keys_kernel(const void* func);
int launch_binning(const gsl_params& p, const GeomView& g, const ImageView& im, const BinView& b,
                   int64_t r_capacity, int32_t* r_host, cudaStream_t st);
int launch_export_keys(const gsl_params& p, const GeomView& g, const ImageView& im, const uint32_t* point_list,
                       uint64_t* keys_out, cudaStream_t st);
int wait_num_rendered(int32_t* r_host, cudaStream_t st);
int launch_render_forward(const gsl_params& p, const gsl_fwd_inputs& in, gsl_fwd_outputs& out,
                          const GeomView& g, const ImageView& im, const BinView& b,
                          int64_t r_capacity, cudaStream_t st);
int launch_render_backward(const gsl_params& p, const gsl_fwd_inputs& in, const gsl_fwd_outputs& fwd,
                           const gsl_bwd_inputs& gin, const GeomView& g, const ImageView& im,
                           const BinView& b, int64_t r_capacity, cudaStream_t st);
int launch_extract_sh_factor(const gsl_params& p, const GeomView& g, float* out, cudaStream_t st);
int launch_zero_outputs(const gsl_params& p, const gsl_fwd_inputs& in, gsl_bwd_outputs& gout, cudaStream_t st);
int launch_preprocess_backward(const gsl_params& p, const gsl_fwd_inputs& in,
                               const gsl_fwd_outputs& fwd, gsl_bwd_outputs& gout, const GeomView& g,
                               bool prezeroed, int row0, int row1, cudaStream_t st, bool fused = false,
                               bool factors_done = false);
int launch_sh_expand(int P, int D, int M, int G, const float* means3D, const float* campos_all, const float* drgb_all,
                     size_t drgb_stride, float* dL_dsh, cudaStream_t st);
int launch_glue_forward(const gsl_glue_params& p, const gsl_glue_inputs& in, const gsl_glue_outputs& out, cudaStream_t st);
int launch_glue_backward(const gsl_glue_params& p, const gsl_glue_inputs& in, const gsl_glue_outputs& gout,
                         const gsl_glue_inputs_grad& gin, cudaStream_t st);
int launch_pano_forward(const gsl_pano_params& p, const float* range, float* points, int32_t* index, int32_t* count,
                        float* normals, void* scratch, cudaStream_t st);
int launch_pano_backward(const gsl_pano_params& p, const float* range, int K, const float* g_points, const int32_t* index,
                         const float* g_normals, float* g_range, cudaStream_t st);
int launch_mark_visible(int P, const float* means3D, const float* viewmatrix,
                        const float* projmatrix, uint8_t* present, cudaStream_t st);

// ---- peer-memory gradient exchange (gsl_peer.cu) -------------------------------------------------------------
constexpr int PEER_MAX = GSL_PEER_MAX;
// Header of an exchange buffer (first PEER_HEADER bytes):
//   [0, 256)      flags u32[8 slots][PEER_MAX]: slot s, word g = the last ticket rank g published on slot s.
//                 slots 0..3: host-ticketed barriers (gsl_peer_barrier / _signal / _wait); slot 4: "rows + factors of
//                 this step pushed", slot 5: "sums of my tiles pushed", slot 6: "SH factors of this step pushed" (side stream,
//                 under the per-surfel kernel) -- the in-kernel barriers of the fused step
//   [512, 516)    step counter of the fused step (device side: incremented by k_peer_begin, identical on all ranks)
//   [516, 520)    error word: != 0 after a barrier time-out; the kernels that write gradients then write NaN
//   [520, 552)    finished-CTA counters (one per signalling kernel) for "last CTA publishes the flag"
//   [1024, 1036)  this rank's camera centre;  [1280, 1536) float4[2 parities][PEER_MAX] camera centres of all ranks
constexpr size_t PEER_HEADER = 4096;
constexpr int PEER_SLOT_PUSHED = 4, PEER_SLOT_SUMMED = 5, PEER_SLOT_FACTORS = 6, PEER_SLOTS = 8;
constexpr size_t PEER_STEP_OFF = 512, PEER_ERROR_OFF = 516, PEER_DONE_OFF = 520;
constexpr size_t PEER_CAMPOS_OFF = GSL_PEER_CAMPOS_OFFSET;
constexpr size_t PEER_CAMPOS_ALL_OFF = 1280;  // float4[2 parities][PEER_MAX]: camera centres of all ranks
static_assert(PEER_SLOTS * PEER_MAX * 4 <= PEER_STEP_OFF && PEER_DONE_OFF + 32 <= PEER_CAMPOS_OFF, "header layout");
constexpr size_t PEER_GLUE_ALL_OFF = 1536;  // float4[2 parities][PEER_MAX]: (timestamp - time_shift, time_shift, 0, 0) of all ranks (gsl_peer_glue)
static_assert(PEER_CAMPOS_ALL_OFF + 2 * PEER_MAX * 16 <= PEER_GLUE_ALL_OFF && PEER_GLUE_ALL_OFF + 2 * PEER_MAX * 16 <= PEER_HEADER,
              "header too small");
// What the SH expansion needs to rebuild the position a rank rasterized: xyz + velocity * coef(that rank's timestamp)
struct GlueMean {
  const float* xyz;   // null: no glue, the caller's means3D is used for every rank
  const float* vel;
  const float* t0;
  const float* sigt;
  float a, inv2T_decay;
};
inline GlueMean make_glue_mean(const gsl_peer_ctx* c) {
  GlueMean m = {nullptr, nullptr, nullptr, nullptr, 0.f, 0.f};
  if (c && c->glue) {
    m.xyz = c->glue->xyz; m.vel = c->glue->velocity; m.t0 = c->glue->t; m.sigt = c->glue->scaling_t;
    m.a = (float)(1.0 / (double)c->glue->cycle * 3.141592653589793 * 2.0);
    m.inv2T_decay = c->glue->velocity_decay / c->glue->cycle / 2.f;
  }
  return m;
}
struct PeerView {  // gsl_peer_ctx by value, as the kernels take it
  int rank, world;
  uint32_t epoch;
  int parity;  // which factor table (step & 1)
  int dev_step;  // 1: epoch / parity come from the device-side step counter (fused step; constant kernel arguments,
                 // so the step can be replayed from a CUDA graph)
  char* own;  // buf[rank]
  char* buf[PEER_MAX];
  int* error_host;  // pinned host flag (may be null)
  unsigned long long timeout_ns;  // a rank that never arrives at a barrier: report (error word + host flag), do not hang
};
unsigned long long peer_timeout_ns();  // gsl_peer_set_timeout_ms (gsl_api.cu), 20 s by default
inline PeerView make_view(const gsl_peer_ctx* c, bool dev_step = false) {
  PeerView v;
  v.rank = c->rank;
  v.world = c->world;
  v.epoch = c->epoch;
  for (int g = 0; g < PEER_MAX; ++g) v.buf[g] = g < c->world ? (char*)c->buf[g] : nullptr;
  v.own = (char*)c->buf[c->rank];
  v.parity = (int)(c->parity & 1u);
  v.dev_step = dev_step ? 1 : 0;
  v.error_host = (int*)c->error_flag;
  v.timeout_ns = peer_timeout_ns();
  return v;
}
// Byte offsets inside an exchange buffer for P surfels (tiles of 256), S feature channels and `world` ranks.
struct PeerLayout {
  int tiles, tiles_per_rank;  // ceil(P / 256); tiles a rank owns (tile t belongs to rank t % world)
  size_t off_fmeta;      // uint2 [2 parities][world][tiles * 8]: (factor bits, non-zero factors of the tile before this word), per source rank
  size_t off_factor;     // float4 [2 parities][world][tiles * 256]: SH factors per source rank, the non-zero ones of a tile packed to its front
  size_t off_stagebits;  // u32 [world][tiles_per_rank * 8]: row bits of the tiles this rank owns, per source rank
  size_t off_stage;      // float [world][tiles_per_rank * 256][rw]: packed rows of the tiles this rank owns, per source rank
  size_t off_rowbits;    // u32 [tiles * 8]: OR of the row bits over the ranks (written by the tile owners)
  size_t off_rows;       // float [tiles * 256][rw]: summed rows (written by the tile owners)
  size_t total;
};
int peer_row_width(int S);
// "feature" channels of the packed rows: S, or S rounded up to whole quads + two quads of glue gradients (gsl_peer_glue)
inline int peer_rows_S(int S, const gsl_peer_ctx* c) { return (c && c->glue) ? 4 * ((S + 3) / 4) + 8 : S; }
PeerLayout peer_layout(size_t P, int S, int world);
int launch_peer_barrier(const gsl_peer_ctx* c, int phase, int mode, cudaStream_t st);
int launch_peer_begin(const gsl_peer_ctx* c, const float* campos, cudaStream_t st);
int launch_peer_signal_fused(const gsl_peer_ctx* c, int slot, cudaStream_t st);
int launch_peer_wait_fused(const gsl_peer_ctx* c, int slot, cudaStream_t st);
// fused: true = one step of gsl_backward_surfels_exchange (device-side step counter, in-kernel flag waits / signals)
int launch_peer_sh_expand(const gsl_peer_ctx* c, int P, int S, int D, int M, int row0, int row1, bool prezeroed,
                          const float* means3D, float* dL_dsh, cudaStream_t st, bool fused = false,
                          int wait_slot = PEER_SLOT_PUSHED);
// fused step: SH factors of this rank out of the packed accumulators into the own factor table (main stream, before the
// per-surfel kernel re-zeroes them), then from there into every other rank's table (side stream, under that kernel)
int launch_peer_factor_extract(const gsl_peer_ctx* c, const gsl_params& p, const GeomView& g, cudaStream_t st);
int launch_peer_factor_push(const gsl_peer_ctx* c, const gsl_params& p, cudaStream_t st);
int launch_chamfer_forward(int b, int n, const float* xyz1, int m, const float* xyz2, float* dist1, int* idx1, float* dist2,
                           int* idx2, void* scratch, cudaStream_t st);
int launch_chamfer_backward(int b, int n, const float* xyz1, int m, const float* xyz2, const float* gdist1, const int* idx1,
                            const float* gdist2, const int* idx2, float* gxyz1, float* gxyz2, cudaStream_t st);
int launch_peer_reduce_rows(const gsl_peer_ctx* c, int P, int S, int row_begin, int row_end, cudaStream_t st, bool fused = false);
int launch_peer_unpack(const gsl_peer_ctx* c, int P, int S, bool prezeroed, const gsl_bwd_outputs& out, cudaStream_t st,
                       bool fused = false);

}  // namespace gsl
