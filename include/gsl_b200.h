/*
 * gsl_b200.h -- C-ABI of the B200-native panoramic 2D-Gaussian-surfel rasterizer.
 *
 * This is the drop-in boundary for the ONE hot path of GS-LiDAR
 * (diff-gaussian-rasterization-2d).  Every entry point cites the reference interface it
 * replaces (paths relative to /root/reference):
 *
 *   gsl_forward*      <- RasterizeGaussiansCUDA            rasterize_points.cu:35-139
 *                        CudaRasterizer::Rasterizer::forward  cuda_rasterizer/rasterizer.h:28-63
 *   gsl_backward      <- RasterizeGaussiansBackwardCUDA    rasterize_points.cu:141-247
 *                        CudaRasterizer::Rasterizer::backward cuda_rasterizer/rasterizer.h:65-103
 *   gsl_mark_visible  <- markVisible                       rasterize_points.cu:249-267
 *                        CudaRasterizer::Rasterizer::markVisible cuda_rasterizer/rasterizer.h:22-27
 *   gsl_workspace_*   <- GeometryState/ImageState/BinningState::fromChunk + required<T>()
 *                        cuda_rasterizer/rasterizer_impl.h:26-70, rasterizer_impl.cu:159-208
 *
 * Plain pointers and sizes only: no torch / pybind / C++ types cross this boundary.  All data
 * pointers are DEVICE pointers (float32 contiguous, same layouts as the reference tensors)
 * unless a parameter name ends in _host.  `stream` is a cudaStream_t passed as void*.
 * Every function returns 0 on success, a negative GSL_E* code on a validation error, or a
 * positive cudaError_t; gsl_last_error() returns a thread-local message for the last failure.
 * Nothing here falls back to a CPU path.
 */
#ifndef GSL_B200_H_
#define GSL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define GSL_API __attribute__((visibility("default")))
#else
#define GSL_API
#endif

#define GSL_ABI_VERSION 4
#define GSL_NUM_CHANNELS 4 /* cuda_rasterizer/config.h:12 */
#define GSL_TILE 16        /* cuda_rasterizer/config.h:13-14 */
#define GSL_MAX_FEATURES 10 /* forward.cu:348: F[13] holds S features + 3 normal channels */

#define GSL_EINVAL -1   /* bad argument (shape, null pointer, S > 10, ...) */
#define GSL_ENOSPACE -2 /* a workspace buffer is smaller than gsl_workspace_sizes() demands */
#define GSL_ESTATE -3   /* backward called without a matching forward */

/* flags */
#define GSL_FLAG_DEBUG_SYNC 1u /* raster_settings.debug: sync + check after every stage */
#define GSL_FLAG_WRAP_AZIMUTH 4u /* opt-in, NOT the reference's semantics (it clamps rects, auxiliary.h:47-55): the
                                    panorama is periodic in azimuth -- a splat on the +-180 deg seam keeps its true
                                    footprint (modular tile columns) and its low-pass distance wraps.  Needs
                                    hfov_max - hfov_min = 360 and an image of <= 16384 pixels in width. */
#define GSL_FLAG_BWD_SH_FACTORED 2u /* gsl_backward: write the clamp-masked dL_dRGB factor into dL_dcolors and do
                                       not write dL_dsh (frame-parallel training rebuilds it with gsl_sh_expand) */

#define GSL_FLAG_BWD_PEER_ROWS 8u /* backward for the peer-memory gradient exchange (gsl_peer_*): implies SH_FACTORED;
                                    gsl_backward_surfels[_rows] writes no dense tensor but pushes packed gradient rows
                                    and SH factors into the exchange buffers of gsl_bwd_outputs.peer (remote stores). */

/* Mirrors the scalar arguments of Rasterizer::forward / ::backward
 * (rasterizer.h:31-63) plus GaussianRasterizationSettings (diff_gaussian_rasterization_2d.py:194-209). */
typedef struct gsl_params {
  int32_t P;          /* surfels */
  int32_t S;          /* extra feature channels (features.shape[1]), 0..10 */
  int32_t D;          /* active SH degree 0..3 */
  int32_t M;          /* stored SH coefficients per surfel (0 when colors_precomp is used) */
  int32_t W, H;       /* image_width, image_height */
  float tanfovx, tanfovy; /* carried for API parity; unused by the math like in the reference */
  float scale_modifier;   /* carried for API parity; the reference ignores it (forward.cu:85) */
  float vfov_min, vfov_max, hfov_min, hfov_max; /* degrees */
  float scale_factor;
  int32_t prefiltered;
  uint32_t flags;
} gsl_params;

/* Byte sizes of the three scratch chunks (the reference's geomBuffer / binningBuffer / imgBuffer). */
typedef struct gsl_ws_sizes {
  size_t geom_bytes;    /* function of P */
  size_t binning_bytes; /* function of the instance capacity R_cap */
  size_t image_bytes;   /* function of W*H */
} gsl_ws_sizes;

/* Caller-owned scratch.  The wrapper keeps these alive from forward to backward exactly as the
 * reference keeps geomBuffer/binningBuffer/imgBuffer in ctx (diff_gaussian_rasterization_2d.py:122). */
typedef struct gsl_workspace {
  void* geom;    size_t geom_bytes;
  void* binning; size_t binning_bytes;
  void* image;   size_t image_bytes;
  int64_t r_capacity;     /* instance capacity the binning chunk was sized for */
  int32_t* num_rendered_host; /* PINNED host int[2]: [0]=R (tile instances; -1 while in flight), [1]=overflow flag */
} gsl_workspace;

typedef struct gsl_fwd_inputs {
  const float* background;     /* (4) */
  const float* means3D;        /* (P,3) */
  const float* shs;            /* (P,M,4) or NULL; with shs_rest: coefficient 0 only, (P,1,4) */
  const float* colors_precomp; /* (P,4) or NULL */
  const float* features;       /* (P,S) or NULL when S==0 */
  const float* opacities;      /* (P,1) */
  const float* scales;         /* (P,3) */
  const float* rotations;      /* (P,4) */
  const float* cov3D_precomp;  /* ignored, like the reference */
  const uint8_t* mask;         /* (P,1) bool */
  const float* viewmatrix;     /* (4,4) transposed world->camera, as scene/cameras.py:62 */
  const float* projmatrix;     /* (4,4), only used by gsl_mark_visible */
  const float* campos;         /* (3) */
  /* ABI 2, appended so that ABI-1 initialisers keep their meaning: */
  const float* shs_rest;       /* NULL, or coefficients 1..M-1 as (P,M-1,4): GaussianModel._features_dc /
                                  _features_rest taken without the torch.cat of get_features (gaussian_model.py:167-171) */
} gsl_fwd_inputs;

typedef struct gsl_fwd_outputs {
  int32_t* out_contrib; /* (2,H,W): last contributor, median contributor */
  float* out_color;     /* (4,H,W) */
  float* out_feature;   /* (S+3,H,W) */
  float* out_depth;     /* (4,H,W): mean, median, distortion, mean of squares */
  float* out_alpha;     /* (1,H,W) = 1 - T  (the reference returns T and Python does 1-T) */
  int32_t* radii;       /* (P) */
} gsl_fwd_outputs;

typedef struct gsl_bwd_inputs {
  const float* dL_dout_color;   /* (4,H,W) */
  const float* dL_dout_depth;   /* (4,H,W) */
  const float* dL_dout_alpha;   /* (1,H,W)  (fed as dL_dout_mask in the reference) */
  const float* dL_dout_feature; /* (S+3,H,W) */
} gsl_bwd_inputs;

struct gsl_peer_ctx;
typedef struct gsl_bwd_outputs {
  float* dL_dmeans3D;  /* (P,3) */
  float* dL_dmeans2D;  /* (P,4): densification proxy in .xy, zeros in .zw (backward.cu:700-711) */
  float* dL_dsh;       /* (P,M,4) or NULL; (P,1,4) when the input came as shs + shs_rest */
  float* dL_dcolors;   /* (P,4) */
  float* dL_dfeatures; /* (P,S) or NULL */
  float* dL_dopacity;  /* (P,1) */
  float* dL_dscales;   /* (P,3) */
  float* dL_drotations;/* (P,4) */
  float* dL_dcov3D;    /* (P,6) all zero, like the reference; may be NULL */
  float* dL_dsh_rest;  /* ABI 2: (P,M-1,4) when the input came as shs + shs_rest, else ignored */
  const struct gsl_peer_ctx* peer; /* ABI 3, GSL_FLAG_BWD_PEER_ROWS only: the mapped exchange buffers (see below) */
} gsl_bwd_outputs;

GSL_API int gsl_abi_version(void);
GSL_API const char* gsl_last_error(void);

/* Sizes of the scratch chunks for P surfels, r_capacity tile instances and W*H pixels. */
GSL_API int gsl_workspace_sizes(const gsl_params* p, int64_t r_capacity, gsl_ws_sizes* out);

/* How the binning walks the 16x16 tiles of a W x H image: groups of at most 1024 consecutive tile ids, whole tile rows
 * or -- beyond 16384 pixels of width -- pieces of one row (the counting pass that replaces the reference's 64-bit key sort,
 * rasterizer_impl.cu:338-344, holds a shared-memory bitmap per tile).  Returns the number of groups and writes up to
 * `capacity` of them as (first tile id, tiles, first tile row, end row, first tile column, end column); -1 on bad
 * arguments.  Informational: the forward pass does this itself. */
GSL_API int32_t gsl_bin_groups(int32_t W, int32_t H, int32_t* groups, int32_t capacity);

/* Stage 1 of the forward pass: per-surfel preprocess with the depth sort of the surfels on a side stream under
 * it.  Never blocks the host. */
GSL_API int gsl_forward_preprocess(const gsl_params* p, const gsl_fwd_inputs* in, gsl_fwd_outputs* out,
                           gsl_workspace* ws, void* stream);

/* Stage 2: binning (tile lists bit-identical to the reference's sorted list, tile ranges, block lists) and
 * per-tile compositing.  Enqueued SPECULATIVELY against ws->r_capacity: if the device-side instance count R
 * exceeds it the kernels write nothing and the overflow flag is raised; the caller then learns R with
 * gsl_wait_num_rendered(), grows the binning chunk and calls this function again.  An asynchronous copy of
 * (R, overflow) into ws->num_rendered_host is enqueued right after the counting kernels, long before the
 * compositing finishes.  Images of more than 1024 tiles are binned in groups of 1024 consecutive tiles by the same
 * kernels (no library sort, no host wait at any size). */
GSL_API int gsl_forward_render(const gsl_params* p, const gsl_fwd_inputs* in, gsl_fwd_outputs* out,
                       gsl_workspace* ws, void* stream);

/* Waits (polling the pinned word, not the whole stream) until the instance count of the forward enqueued on
 * `stream` with `ws` has arrived on the host and returns it.  Replaces the blocking cudaMemcpy of
 * rasterizer_impl.cu:314-315, but sits AFTER all launches of the forward pass instead of in their middle. */
GSL_API int gsl_wait_num_rendered(gsl_workspace* ws, int32_t* num_rendered, void* stream);

/* Blocking convenience with the semantics of Rasterizer::forward: runs both stages, waits for
 * R, returns it in *num_rendered.  Returns GSL_ENOSPACE (with *num_rendered set) if the
 * binning chunk is too small for R. */
GSL_API int gsl_forward(const gsl_params* p, const gsl_fwd_inputs* in, gsl_fwd_outputs* out,
                gsl_workspace* ws, int32_t* num_rendered, void* stream);

/* Backward pass for the forward that last used `ws`.  Writes every element of every output
 * (no pre-zeroing needed by the caller). */
GSL_API int gsl_backward(const gsl_params* p, const gsl_fwd_inputs* in, const gsl_fwd_outputs* fwd,
                 const gsl_bwd_inputs* gin, gsl_bwd_outputs* gout, gsl_workspace* ws, void* stream);

/* Frame-parallel training (no counterpart in the reference, which is single-GPU): the SH gradient one frame gives
 * a surfel is basis(view direction) x dL_dRGB, so ranks exchange the 16-byte factor (all-gather of the dL_dcolors
 * written under GSL_FLAG_BWD_SH_FACTORED) instead of all-reducing 16*M bytes per surfel, and each rank rebuilds
 *   dL_dsh[i] = sum_g basis(normalize(means3D[i] - campos_all[g])) x drgb_all[g * drgb_stride + 4 i .. +3].
 * campos_all: (G,3) device floats; drgb_all: G blocks of >= 4*P floats, drgb_stride floats apart; dL_dsh (P,M,4). */
GSL_API int gsl_sh_expand(int32_t P, int32_t D, int32_t M, int32_t G, const float* means3D, const float* campos_all,
                  const float* drgb_all, size_t drgb_stride, float* dL_dsh, void* stream);

/* gsl_backward in two stages, for callers that start communication between them: _composite runs the backward
 * compositor (and forks the zero-fill of the dense outputs onto a side stream); when sh_factor_out is non-NULL it
 * also writes the clamp-masked dL_dRGB factor (P,4) there, so that a frame-parallel caller can all-gather it while
 * _surfels (the per-surfel VJP, which joins the zero-fill) is still running.  Must be called in this order by the
 * same host thread. */
GSL_API int gsl_backward_composite(const gsl_params* p, const gsl_fwd_inputs* in, const gsl_fwd_outputs* fwd,
                           const gsl_bwd_inputs* gin, gsl_bwd_outputs* gout, gsl_workspace* ws, float* sh_factor_out,
                           void* stream);
GSL_API int gsl_backward_surfels(const gsl_params* p, const gsl_fwd_inputs* in, const gsl_fwd_outputs* fwd,
                         gsl_bwd_outputs* gout, gsl_workspace* ws, void* stream);

/* ---- render() glue: the per-surfel element-wise work gaussian_renderer/__init__.py:64-115 and
 * scene/gaussian_model.py:139-186 run in PyTorch before the rasterizer, as one kernel per direction. ---- */
typedef struct gsl_glue_params {
  int32_t P;
  float timestamp;      /* viewpoint_camera.timestamp */
  float time_shift;     /* render(time_shift=...), 0 when None */
  float cycle;          /* GaussianModel.T (args.cycle) */
  float velocity_decay; /* GaussianModel.velocity_decay */
  int32_t dynamic;      /* pipe.dynamic */
} gsl_glue_params;
typedef struct gsl_glue_inputs {   /* raw (pre-activation) parameters of GaussianModel */
  const float* xyz;        /* _xyz (P,3) */
  const float* velocity;   /* _velocity (P,3) */
  const float* t;          /* _t (P,1) */
  const float* scaling_t;  /* _scaling_t (P,1), log */
  const float* opacity;    /* _opacity (P,1), logit */
  const float* scaling;    /* _scaling (P,3), log */
  const float* rotation;   /* _rotation (P,4), unnormalised */
  const uint8_t* mask;     /* optional user mask (P), may be NULL */
} gsl_glue_inputs;
typedef struct gsl_glue_outputs {  /* what render() feeds the rasterizer (also used for their cotangents) */
  float* means3D;    /* (P,3) */
  float* opacity;    /* (P,1) */
  float* scales;     /* (P,3) */
  float* rotations;  /* (P,4) */
  float* marginal_t; /* (P,1), may be NULL */
  uint8_t* mask;     /* (P) prefilter mask */
} gsl_glue_outputs;
typedef struct gsl_glue_inputs_grad {
  float* xyz; float* velocity; float* t; float* scaling_t; float* opacity; float* scaling; float* rotation;
} gsl_glue_inputs_grad;
GSL_API int gsl_glue_forward(const gsl_glue_params* p, const gsl_glue_inputs* in, const gsl_glue_outputs* out, void* stream);
/* gout: cotangents of (means3D, opacity, scales, rotations); NULL members count as zero. */
GSL_API int gsl_glue_backward(const gsl_glue_params* p, const gsl_glue_inputs* in, const gsl_glue_outputs* gout,
                      const gsl_glue_inputs_grad* gin, void* stream);

/* ---- panorama post-ops: utils/graphics_utils.py:96-118 pano_to_lidar and :121-149 depth_to_normal of
 * the reference (called on every rendered range image, train.py:261-262,306), as one pass over the (H, W) range image.
 * Pixel (row, col) looks along normalize(sin th sin ph, -cos th, sin th cos ph), th = (90 - vfov_max + row / H * (vfov_max -
 * vfov_min)) deg, ph = (hfov_min + col / W * (hfov_max - hfov_min)) deg. ---- */
typedef struct gsl_pano_params {
  int32_t H, W;
  float vfov_min, vfov_max, hfov_min, hfov_max; /* degrees, as args.vfov / args.hfov */
} gsl_pano_params;
/* scratch bytes for gsl_pano_forward */
GSL_API size_t gsl_pano_scratch_bytes(int32_t H, int32_t W);
/* range (H*W) -> points (capacity H*W x 3): direction * range of the pixels with range > 0 in row-major order, index (H*W):
 * the pixel of every point (needed by the backward pass; may be NULL), *count (device int32): how many; normals (3,H,W):
 * normalize(cross(p[y+1,x] - p[y-1,x], p[y,x+1] - p[y,x-1])), zero on the one-pixel border.  points (with index, count) or
 * normals may be NULL to compute only the other. */
GSL_API int gsl_pano_forward(const gsl_pano_params* p, const float* range, float* points, int32_t* index, int32_t* count,
                             float* normals, void* scratch, void* stream);
/* g_range (H*W, overwritten) = d/d range of <g_points, points> + <g_normals, normals>; either cotangent may be NULL. */
GSL_API int gsl_pano_backward(const gsl_pano_params* p, const float* range, int32_t K, const float* g_points,
                              const int32_t* index, const float* g_normals, float* g_range, void* stream);

/* Pinhole frustum test, present[i] = in_frustum(means3D[i]) (auxiliary.h:157-180). */
GSL_API int gsl_mark_visible(int32_t P, const float* means3D, const float* viewmatrix,
                     const float* projmatrix, uint8_t* present, void* stream);

/* Test/diagnostic export of the internal state in the REFERENCE's layouts so parity tests can
 * compare bit-for-bit (rasterizer_impl.h:26-63).  Any pointer may be NULL. */
typedef struct gsl_state_export {
  float* depths;          /* (P) */
  float* means2D;         /* (P,2) */
  float* transMat;        /* (P,9) */
  float* normal_opacity;  /* (P,4) */
  float* rgb;             /* (P,4) */
  uint8_t* clamped;       /* (P,4) */
  uint32_t* tiles_touched;/* (P) */
  uint32_t* point_offsets;/* (P) inclusive scan */
  uint64_t* point_list_keys; /* (R) sorted keys */
  uint32_t* point_list;      /* (R) sorted surfel ids */
  uint32_t* ranges;          /* (tiles,2) */
  float* final_T;            /* (3,H,W): T, M1, M2 */
  int16_t* pixbox;           /* (P,4) conservative pixel box x0,y0,x1,y1 (this design only) */
} gsl_state_export;
GSL_API int gsl_export_state(const gsl_params* p, const gsl_workspace* ws, int64_t R,
                     const gsl_state_export* dst, void* stream);

/* ---- Frame-parallel gradient exchange over NVLink peer memory (no counterpart in the reference, which is single-GPU;
 * SURVEY.md 8e).  Every rank owns one exchange buffer that all ranks of the node map (CUDA IPC).  Under
 * GSL_FLAG_BWD_PEER_ROWS the per-surfel backward kernel PUSHES its results into the ranks' buffers with remote stores
 * (packed gradient rows of a 256-surfel tile -> the rank that owns the tile, tile % world; SH factors -> every rank),
 * the owners sum their tiles and push the sums to everybody, and each rank rebuilds dL_dsh and the dense gradient
 * tensors locally -- no collective library and no remote load on the data path.  All calls take row ranges (multiples
 * of 256), so the exchange of one range can run on a side stream while the backward kernel computes the next one.
 * Per step and rank (ticket = a number that grows with every barrier on a flag slot, same sequence on all ranks):
 *   camera centre -> own buffer + GSL_PEER_CAMPOS_OFFSET;  gsl_backward_composite(sh_factor_out = NULL);
 *   for each row range:  gsl_backward_surfels_rows();  gsl_peer_barrier(slot 0 on the first range, else 1);
 *                        gsl_peer_reduce();  [side stream, after the barrier] gsl_peer_sh_expand();
 *   gsl_peer_barrier(slot 2);  gsl_peer_unpack();  join.
 * After the last barrier every rank holds bit-identical sums and may start the next step. */
#define GSL_PEER_MAX 8 /* ranks of one NVLink domain */
#define GSL_PEER_CAMPOS_OFFSET 1024 /* float[3] at this byte offset of the own buffer: this rank's camera centre
                                       (pushed to every rank's table by a barrier on flag slot 0) */
typedef struct gsl_peer_handle { unsigned char reserved[64]; } gsl_peer_handle; /* cudaIpcMemHandle_t */
/* Optional: the per-surfel glue of render() (gsl_glue_forward below) sits IN FRONT of the rasterizer, and the Jacobian of
 * its motion model depends on the FRAME's timestamp (means3D = xyz + v sin((t - t0) a) / a, opacity * marginal(t)): for
 * ranks rendering different timestamps sum_g J_g^T G_g != J^T sum_g G_g, so the time-dependent part of the glue's VJP has
 * to be applied BEFORE the sum.  With `glue` set in the context the per-surfel backward kernel does that: the packed rows
 * carry 8 more floats after the feature quads -- (dL/dvelocity.xyz, dL/dt | dL/dscaling_t, 0, 0, 0) -- and the opacity
 * slot holds dL/d sigmoid(opacity) (the marginal already applied); dL_dmeans3D is dL/dxyz as it is.  The exchange then
 * moves S_rows = 4 ceil(S / 4) + 8 "feature" channels (gsl_peer_rows_channels): pass S_rows wherever a gsl_peer_* call
 * takes S, and a (P, S_rows) dL_dfeatures to gsl_peer_unpack / gsl_backward_surfels_exchange -- columns [0, S) are
 * dL_dfeatures, [4 ceil(S/4), +3) dL/dvelocity, then dL/dt and dL/dscaling_t.  S <= 4.  The remaining, frame-independent
 * part of the glue's VJP (sigmoid', exp', normalize') is applied to the sums (gsl_glue_backward with dynamic = 0; its
 * velocity / t / scaling_t outputs are then replaced by the exchanged ones).  The SH expansion evaluates rank g's basis at
 * the position rank g rasterized, xyz + velocity * coef(timestamp_g): every rank's (timestamp - time_shift, time_shift) is
 * pushed next to its camera centre, and the `means3D` argument of gsl_peer_sh_expand* is ignored when glue is set. */
typedef struct gsl_peer_glue {
  float timestamp, time_shift, cycle, velocity_decay; /* as gsl_glue_params */
  int32_t dynamic;
  const float* xyz;       /* (P,3) raw parameters, as gsl_glue_inputs */
  const float* velocity;  /* (P,3) */
  const float* t;         /* (P,1) */
  const float* scaling_t; /* (P,1) */
  const float* opacity;   /* (P,1) */
} gsl_peer_glue;
typedef struct gsl_peer_ctx {
  int32_t rank, world;       /* world <= GSL_PEER_MAX */
  uint32_t epoch;            /* the ticket of the next barrier call; barriers wait for flags >= ticket (wrap-safe) */
  uint32_t parity;           /* step & 1: which of the two SH-factor tables this step pushes into / expands from (two
                                tables, so that a fast rank's next step never overwrites factors a slow rank still reads) */
  void* buf[GSL_PEER_MAX];   /* exchange buffer of every rank as mapped into THIS process; buf[rank] is the own one */
  int32_t* error_flag;       /* device-visible int (pinned host memory): set to 1 + slot when a barrier timed out */
  const gsl_peer_glue* glue; /* NULL: rows of rasterizer-input gradients; else see gsl_peer_glue (appended field) */
} gsl_peer_ctx;
/* channels the exchange moves for S feature channels: S, or 4 ceil(S / 4) + 8 with the glue's VJP folded in */
GSL_API int32_t gsl_peer_rows_channels(int32_t S, int32_t with_glue);
GSL_API size_t gsl_peer_buffer_bytes(int64_t P, int32_t S, int32_t world);
/* floats per packed row: [means2D.xy scales.xy | rotations | means3D opacity | features (S), zero padded] */
GSL_API int32_t gsl_peer_row_width(int32_t S);
GSL_API int gsl_peer_alloc(size_t bytes, void** dptr, gsl_peer_handle* handle); /* cudaMalloc, zero, export */
GSL_API int gsl_peer_open(const gsl_peer_handle* handle, void** dptr);          /* map a peer's buffer (other process) */
GSL_API int gsl_peer_close(void* dptr);
/* How long a barrier (kernel or in-kernel wait) waits for a rank that does not arrive before it raises the error flag and
 * lets the step finish with NaN gradients; process-wide, 20000 ms by default. */
GSL_API int gsl_peer_set_timeout_ms(uint32_t ms);
/* Schedule of the fused step (gsl_backward_surfels_exchange), process-wide.  Both options are OFF by default: after the
 * backward compositor every kernel of the step is HBM- or NVLink-bound and the GPU is busy until the step ends, so moving
 * work between streams re-orders it without shortening it (measured on 2 B200s, profiles/r02_exchange_schedules.md:
 * 1.03 ms / step by default, 1.06 with early factors, 1.07 with the low-priority expansion as well).
 *   GSL_PEER_OPT_EARLY_FACTORS: the SH factors are extracted right after the backward compositor and pushed from the
 *     library's high-priority side stream UNDER the per-surfel kernel; 0: that kernel pushes them together with its rows.
 *   GSL_PEER_OPT_EXPAND_LOW_PRIORITY (only with early factors): the SH expansion runs on a lowest-priority stream behind
 *     a one-warp wait for every rank's factors; the waits of the sum and unpack kernels are parked in one-warp kernels too.
 * Every rank must use the same options. */
enum { GSL_PEER_OPT_EARLY_FACTORS = 0, GSL_PEER_OPT_EXPAND_LOW_PRIORITY = 1 };
GSL_API int gsl_peer_set_option(int32_t option, int32_t value);
GSL_API int gsl_peer_free(void* dptr);
/* Signal "this rank reached ticket ctx->epoch on flag slot `slot` (0..3)" to all ranks and wait for all of them. */
GSL_API int gsl_peer_barrier(const gsl_peer_ctx* ctx, int32_t slot, void* stream);
/* The two halves of the barrier (signal: publish + flag stores to all ranks; wait: spin on the own flag slots), for
 * callers that put work between them.  Slot 0's signal also pushes the camera centre into every rank's table. */
GSL_API int gsl_peer_signal(const gsl_peer_ctx* ctx, int32_t slot, void* stream);
GSL_API int gsl_peer_wait(const gsl_peer_ctx* ctx, int32_t slot, void* stream);
/* gsl_sh_expand for the surfels [row_begin, row_end) over the factor tables and camera centres the ranks pushed into
 * the own buffer (local reads); dL_dsh is the full (P,M,4) tensor. */
GSL_API int gsl_peer_sh_expand(const gsl_peer_ctx* ctx, int32_t P, int32_t S, int32_t D, int32_t M, int32_t row_begin,
                               int32_t row_end, const float* means3D, float* dL_dsh, void* stream);
/* The same over a dL_dsh the caller has ZERO-FILLED: only the rows some rank has a factor for are written (about half of
 * the surfels have none), by dense warps over the compacted rows of each 256-surfel tile.  M <= 16.  This is the kernel
 * the fused step (gsl_backward_surfels_exchange) runs. */
GSL_API int gsl_peer_sh_expand_sparse(const gsl_peer_ctx* ctx, int32_t P, int32_t S, int32_t D, int32_t M, int32_t row_begin,
                                      int32_t row_end, const float* means3D, float* dL_dsh, void* stream);
/* Sum of the staged packed rows [row_begin, row_end) (multiples of 256, or row_end = P): this rank sums the tiles it
 * owns (every world-th) in rank order and pushes sums + OR-ed row bits into every rank's result area; complete after
 * the next barrier. */
GSL_API int gsl_peer_reduce(const gsl_peer_ctx* ctx, int32_t P, int32_t S, int32_t row_begin, int32_t row_end, void* stream);
/* Summed packed rows + OR-ed row bits of the own buffer -> dense dL_dmeans3D, dL_dmeans2D, dL_dscales, dL_drotations,
 * dL_dopacity, dL_dfeatures of `out` (every element written). */
GSL_API int gsl_peer_unpack(const gsl_peer_ctx* ctx, int32_t P, int32_t S, const gsl_bwd_outputs* out, void* stream);
/* gsl_backward_surfels for the surfels [row_begin, row_end) only (GSL_FLAG_BWD_PEER_ROWS; row_begin a multiple of 256). */
GSL_API int gsl_backward_surfels_rows(const gsl_params* p, const gsl_fwd_inputs* in, const gsl_fwd_outputs* fwd,
                              gsl_bwd_outputs* gout, gsl_workspace* ws, int32_t row_begin, int32_t row_end, void* stream);
/* The whole second half of a frame-parallel backward pass in one call (after gsl_backward_composite), as ONE fused step:
 * per-surfel VJP with its pushes, sum of the owned tiles, SH expansion (library side stream) and unpack (other schedules:
 * gsl_peer_set_option), with the barriers INSIDE the kernels (the last CTA of a producing kernel publishes a flag in every rank's buffer, every CTA of
 * a consuming kernel waits for the flags of all ranks) and the ticket taken from a device-side step counter in the
 * exchange buffer -- every kernel argument is constant from step to step, so the call can be captured in a CUDA graph.
 * Every rank must make the same sequence of calls on the same exchange buffers.  `step` and `chunks` are accepted for
 * ABI compatibility and ignored (one row range measured best).  gout: peer = the mapped exchange buffers; the dense
 * non-SH pointers and dL_dsh receive the gradients summed over the ranks.  If a rank misses a barrier (20 s), the error
 * flag is raised and the gradients of the step are NaN on the ranks that noticed -- never silently wrong. */
GSL_API int gsl_backward_surfels_exchange(const gsl_params* p, const gsl_fwd_inputs* in, const gsl_fwd_outputs* fwd,
                                  gsl_bwd_outputs* gout, gsl_workspace* ws, uint32_t step, int32_t chunks, void* stream);

/* ---- Chamfer distance of two point sets.  Replaces chamfer_cuda_forward / chamfer_cuda_backward
 * of chamfer/chamfer3D/chamfer3D.cu:142-165,199-230 (pybind names `forward` / `backward`, chamfer_cuda.cpp) of the
 * reference: xyz1 (b,n,3), xyz2 (b,m,3) float32 -> dist1 (b,n) / dist2 (b,m) squared distance to the nearest neighbour
 * in the other set, idx1 / idx2 int32 its index (lowest index among equal minima, like the reference's scan).
 * scratch: 8 * b * (n + m) bytes of device memory.  The backward zero-fills and then accumulates gxyz1 (b,n,3), gxyz2
 * (b,m,3) (the reference's Python wrapper passes torch.zeros). */
GSL_API size_t gsl_chamfer_scratch_bytes(int32_t b, int32_t n, int32_t m);
GSL_API int gsl_chamfer_forward(int32_t b, int32_t n, const float* xyz1, int32_t m, const float* xyz2, float* dist1,
                                int32_t* idx1, float* dist2, int32_t* idx2, void* scratch, void* stream);
GSL_API int gsl_chamfer_backward(int32_t b, int32_t n, const float* xyz1, int32_t m, const float* xyz2,
                                 const float* gdist1, const int32_t* idx1, const float* gdist2, const int32_t* idx2,
                                 float* gxyz1, float* gxyz2, void* stream);

/* ---- CUDA-graph replay of a pass (the reference re-allocates, zero-fills and synchronises in every
 * step, rasterize_points.cu:77-91,186-197, train.py:379).  Every kernel argument of a pass is constant once the
 * workspace, the output / gradient buffers and the inputs keep their addresses, and the forward's instance count is
 * polled after all launches, so a caller can capture a pass once and replay it with one launch:
 *     gsl_graph_begin(&cs);  gsl_forward_preprocess(..., cs); gsl_forward_render(..., cs);  gsl_graph_end(cs, &g);
 *     per step:  ws.num_rendered_host[0] = -1;  gsl_graph_launch(g, stream);  gsl_wait_num_rendered(..., stream);
 * gsl_graph_begin hands out a library-owned stream in capture mode (`cs`; the legacy default stream cannot be captured) on
 * which the pass is issued; the instantiated graph runs on any stream.
 * (same for gsl_backward, or gsl_backward_composite + gsl_backward_surfels_exchange: the fused exchange takes its tickets
 * from a device-side step counter).  If the instance count exceeds ws.r_capacity after a replay nothing was written: grow
 * the binning chunk, capture again.  Capture is thread-local; per-kernel profiling must be off; GSL_FLAG_DEBUG_SYNC is not
 * capturable.  gsl_graph_end(stream, NULL) aborts a capture.  gsl_stage_camera gathers the three small per-frame inputs
 * (view matrix 16 floats, camera centre 3, background 4) into one 24-float device buffer -- [0,16) [16,19) [20,24) -- so
 * that a captured pass can read them from fixed addresses whatever tensors the caller holds them in. */
GSL_API int gsl_graph_begin(void** capture_stream);
GSL_API int gsl_graph_end(void* stream, void** graph_exec);
GSL_API int gsl_graph_launch(void* graph_exec, void* stream);
GSL_API int gsl_graph_destroy(void* graph_exec);
GSL_API int gsl_stage_camera(const float* viewmatrix, const float* campos, const float* background, float* dst, void* stream);

/* Per-kernel device timing (CUDA events on the launching stream), for bench.py's roofline block.
 * Kernel ids index the arrays of gsl_profile_read. */
enum {
  GSL_K_PREPROCESS_FWD = 0,
  GSL_K_SCAN = 1,      /* k_bin_count + k_bin_scan (per group of 1024 tiles) */
  GSL_K_DUPLICATE = 2, /* k_bin_scatter */
  GSL_K_SORT = 3,      /* k_depth_keys + k_sort_hist/scan/scatter/buckets: this repo's depth sort of the surfels */
  GSL_K_RANGES = 4,    /* k_tile_blists */
  GSL_K_RENDER_FWD = 5,
  GSL_K_RENDER_BWD = 6,
  GSL_K_PREPROCESS_BWD = 7,
  GSL_K_GLUE_FWD = 8,  /* k_glue_fwd (render() glue, next-1) */
  GSL_K_GLUE_BWD = 9,  /* k_glue_bwd */
  GSL_K_PEER_REDUCE = 10, /* k_peer_reduce_rows (includes its in-kernel wait for the ranks' "pushed" flags) */
  GSL_K_PEER_EXPAND = 11, /* k_peer_sh_expand_tiles, side stream (includes its in-kernel wait for the ranks' "factors" flags) */
  GSL_K_PEER_UNPACK = 12, /* k_peer_unpack (includes its in-kernel wait for the owners' "summed" flags) */
  GSL_K_PEER_FACTORS = 13, /* k_peer_factor_push + its k_peer_signal, side stream, under k_preprocess_bwd */
  GSL_K_COUNT = 14
};
GSL_API int gsl_profile_enable(int on);
/* Waits for all recorded events, then returns accumulated milliseconds and launch counts per id. */
GSL_API int gsl_profile_read(double* total_ms /*[GSL_K_COUNT]*/, int64_t* launches /*[GSL_K_COUNT]*/, int reset);
GSL_API const char* gsl_kernel_name(int id);

#ifdef __cplusplus
}
#endif
#endif /* GSL_B200_H_ */
