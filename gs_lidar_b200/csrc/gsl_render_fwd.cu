// gsl_render_fwd.cu -- per-tile front-to-back alpha compositing (semantics of forward.cu:292-505).
//
// One CTA per 16x16 tile, one thread per pixel; each warp owns an 8x4 pixel block so that the
// conservative pixel boxes computed by the preprocess cull whole warps.  Surfel records are staged
// through shared memory as float4 (one 64-B record per list entry, coalesced 16-B loads).  The
// per-pair arithmetic is gsl::eval_pair (bit-identical to the reference), the blend recursion below
// follows the reference's rounding sequence too, so the rendered maps match bit-for-bit in
// practice; the contract tested is 1e-5 relative.
#include "gsl_common.cuh"
#include "gsl_math.cuh"

namespace gsl {

constexpr int FWD_BATCH = 256;

#ifdef GSL_STATS
__device__ unsigned long long g_stats[16];
#define STAT_ADD(i, v) do { unsigned long long _s = __reduce_add_sync(0xffffffffu, (unsigned)(v)); if ((threadIdx.x & 31) == 0) atomicAdd(&g_stats[i], _s); } while (0)
#else
#define STAT_ADD(i, v)
#endif

template <int S_T>
__global__ void __launch_bounds__(256) k_render_fwd(
    RenderParams rp, const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
    const float4* __restrict__ rec, const short4* __restrict__ pixbox, const float4* __restrict__ colors,
    const float* __restrict__ features, const float* __restrict__ bg, const uint32_t* __restrict__ ctrl,
    float* __restrict__ final_T, int32_t* __restrict__ out_contrib, float* __restrict__ out_color,
    float* __restrict__ out_feature, float* __restrict__ out_depth, float* __restrict__ out_alpha) {
  const int S = (S_T >= 0) ? S_T : rp.S;
  __shared__ float4 s_rec[4][FWD_BATCH];
  __shared__ uint32_t s_id[FWD_BATCH];
  __shared__ short4 s_box[FWD_BATCH];

  const int tile = blockIdx.x;
  const int tx = tile % rp.gx, ty = tile / rp.gx;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bx0 = tx * GSL_BLOCK_X + (warp & 1) * 8, by0 = ty * GSL_BLOCK_Y + (warp >> 1) * 4;
  const int pxi = bx0 + (lane & 7), pyi = by0 + (lane >> 3);
  const bool inside = pxi < rp.W && pyi < rp.H;
  const int N = rp.W * rp.H;
  const int pix_id = rp.W * pyi + pxi;
  // warp block extents clipped to the image (for the box test)
  const int wbx1 = min(bx0 + 7, rp.W - 1), wby1 = min(by0 + 3, rp.H - 1);

  uint2 range = ranges[tile];
  if (ctrl[0] > rp.r_capacity) range = make_uint2(0, 0);
  const int total = (int)(range.y - range.x);

  const PixelRay ray = make_pixel_ray((float)pxi, (float)pyi, rp.HFOV_min, rp.HFOV_max, rp.VFOV_min,
                                      rp.VFOV_max, rp.W, rp.H);
  bool done = !inside;
  float T = 1.0f;
  int contributor = 0, last_contributor = 0, median_contributor = 0;
  float C[4] = {0.f, 0.f, 0.f, 0.f};
  float F[GSL_MAX_FEATURES];
#pragma unroll
  for (int i = 0; i < GSL_MAX_FEATURES; ++i) F[i] = 0.f;
  float Nn[3] = {0.f, 0.f, 0.f};
  float D = 0.f, D2 = 0.f, M1 = 0.f, M2 = 0.f, distortion = 0.f, median_depth = 0.f;
#ifdef GSL_STATS
  unsigned st_scan = 0, st_box = 0, st_any = 0, st_valid = 0, st_eval_lanes = 0;
#endif

  for (int base = 0; base < total; base += FWD_BATCH) {
    if (__syncthreads_count(done) == 256) break;
    const int nb = min(FWD_BATCH, total - base);
    if ((int)threadIdx.x < nb) {
      uint32_t id = point_list[range.x + base + threadIdx.x];
      s_id[threadIdx.x] = id;
      s_box[threadIdx.x] = pixbox[id];
      const float4* r4 = rec + 4 * (size_t)id;
      s_rec[0][threadIdx.x] = r4[0];
      s_rec[1][threadIdx.x] = r4[1];
      s_rec[2][threadIdx.x] = r4[2];
      s_rec[3][threadIdx.x] = r4[3];
    }
    __syncthreads();
    // The warp leaves the batch loop as a unit; done lanes are masked by the branch below.
    for (int j = 0; j < nb; ++j) {
      if (__all_sync(0xffffffffu, done)) break;
#ifdef GSL_STATS
      if (lane == 0) st_scan++;
#endif
      // warp-uniform cull: does the surfel's conservative pixel box touch this warp's 8x4 block?
      const short4 bb = s_box[j];
      const bool ovy = (int)bb.y <= wby1 && (int)bb.w >= by0;
      const bool ovx = (bb.x <= bb.z) ? ((int)bb.x <= wbx1 && (int)bb.z >= bx0)
                                      : ((int)bb.x <= wbx1 || (int)bb.z >= bx0);
      if (!(ovx && ovy)) continue;  // no pixel of this warp can get alpha >= 1/255 from it
#ifdef GSL_STATS
      if (lane == 0) st_box++;
      if (!done) st_eval_lanes++;
#endif
      if (done) continue;
      Splat s;
      {
        float4 a = s_rec[0][j], b = s_rec[1][j], c = s_rec[2][j], d = s_rec[3][j];
        s.Tux = a.x; s.Tuy = a.y; s.Tuz = a.z; s.Tvx = a.w;
        s.Tvy = b.x; s.Tvz = b.y; s.Twx = b.z; s.Twy = b.w;
        s.Twz = c.x; s.mx = c.y; s.my = c.z; s.opacity = c.w;
        s.nx = d.x; s.ny = d.y; s.nz = d.z; s.depth = d.w;
      }
      const PairEval e = eval_pair<false>(s, ray, rp.near_, rp.far_);
#ifdef GSL_STATS
      if (e.valid) st_valid++;
      { unsigned m = __ballot_sync(__activemask(), e.valid); if (m && (lane == (__ffs(__activemask()) - 1))) st_any++; }
#endif
      if (!e.valid) continue;
      const float alpha = e.alpha;
      const float test_T = GSL_FM(T, GSL_FS(1.f, alpha));
      if (test_T < 0.0001f) {
        done = true;
        continue;
      }
      const int pos = base + j + 1;  // 1-based list position == the reference's `contributor`
      const float w = GSL_FM(T, alpha);
      const float A = GSL_FS(1.f, T);
      const float m = GSL_FM(rp.far_over_range, GSL_FS(1.f, GSL_FD(rp.near_, e.depth)));
      const float mm = GSL_FM(m, m);
      const float t0 = GSL_FF(-M1, GSL_FA(m, m), GSL_FF(A, mm, M2));
      distortion = GSL_FF(w, t0, distortion);
      M1 = GSL_FF(w, m, M1);
      M2 = GSL_FF(w, mm, M2);
      if (T > 0.5f) {
        median_depth = e.depth;
        median_contributor = pos;
      }
      const uint32_t id = s_id[j];
      const float4 col = colors[id];
      C[0] = GSL_FF(T, GSL_FM(alpha, col.x), C[0]);
      C[1] = GSL_FF(T, GSL_FM(alpha, col.y), C[1]);
      C[2] = GSL_FF(T, GSL_FM(alpha, col.z), C[2]);
      C[3] = GSL_FF(T, GSL_FM(alpha, col.w), C[3]);
      if (S_T == 4) {
        const float4 f = *reinterpret_cast<const float4*>(features + 4 * (size_t)id);
        F[0] = GSL_FF(T, GSL_FM(alpha, f.x), F[0]);
        F[1] = GSL_FF(T, GSL_FM(alpha, f.y), F[1]);
        F[2] = GSL_FF(T, GSL_FM(alpha, f.z), F[2]);
        F[3] = GSL_FF(T, GSL_FM(alpha, f.w), F[3]);
      } else {
#pragma unroll
        for (int ch = 0; ch < GSL_MAX_FEATURES; ++ch)
          if (ch < S) F[ch] = GSL_FF(T, GSL_FM(alpha, features[(size_t)id * S + ch]), F[ch]);
      }
      Nn[0] = GSL_FF(T, GSL_FM(alpha, s.nx), Nn[0]);
      Nn[1] = GSL_FF(T, GSL_FM(alpha, s.ny), Nn[1]);
      Nn[2] = GSL_FF(T, GSL_FM(alpha, s.nz), Nn[2]);
      D = GSL_FF(T, GSL_FM(alpha, e.depth), D);
      D2 = GSL_FF(T, GSL_FM(alpha, GSL_FM(e.depth, e.depth)), D2);
      T = test_T;
      last_contributor = pos;
    }
    (void)contributor;
  }

#ifdef GSL_STATS
  STAT_ADD(0, st_scan); STAT_ADD(1, st_box); STAT_ADD(2, st_any); STAT_ADD(3, st_valid); STAT_ADD(4, st_eval_lanes);
#endif
  if (inside) {
    final_T[pix_id] = T;
    final_T[pix_id + N] = M1;
    final_T[pix_id + 2 * N] = M2;
    out_contrib[pix_id] = last_contributor;
    out_contrib[pix_id + N] = median_contributor;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) out_color[ch * N + pix_id] = GSL_FF(T, bg[ch], C[ch]);
#pragma unroll
    for (int ch = 0; ch < GSL_MAX_FEATURES; ++ch)
      if (ch < S) out_feature[ch * N + pix_id] = F[ch];
    out_feature[(S + 0) * N + pix_id] = Nn[0];
    out_feature[(S + 1) * N + pix_id] = Nn[1];
    out_feature[(S + 2) * N + pix_id] = Nn[2];
    out_depth[pix_id] = D;
    out_depth[N + pix_id] = median_depth;
    out_depth[2 * N + pix_id] = distortion;
    out_depth[3 * N + pix_id] = D2;
    out_alpha[pix_id] = 1.f - T;
  }
}

RenderParams make_render_params(const gsl_params& p, int64_t r_capacity) {
  RenderParams rp;
  rp.W = p.W; rp.H = p.H;
  rp.gx = (p.W + GSL_BLOCK_X - 1) / GSL_BLOCK_X;
  rp.gy = (p.H + GSL_BLOCK_Y - 1) / GSL_BLOCK_Y;
  rp.S = p.S;
  Fov f = make_fov(p);
  rp.VFOV_min = f.VFOV_min; rp.VFOV_max = f.VFOV_max; rp.HFOV_min = f.HFOV_min; rp.HFOV_max = f.HFOV_max;
  // near_n * scale_factor, far_n * scale_factor in float (forward.cu:366-367); 2*sf == sf+sf exactly
  volatile float nr = 2.0f * p.scale_factor;
  volatile float fr = 300.0f * p.scale_factor;
  volatile float range = fr - nr;
  volatile float q = fr / range;
  rp.near_ = nr; rp.far_ = fr; rp.far_over_range = q;
  rp.r_capacity = (uint32_t)(r_capacity < 0 ? 0 : (r_capacity > 0xffffffffLL ? 0xffffffffLL : r_capacity));
  return rp;
}

#ifdef GSL_STATS
extern "C" __attribute__((visibility("default"))) void gsl_stats_read(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_stats, sizeof(g_stats));
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_stats, z, sizeof(z)); }
}
#endif

int launch_render_forward(const gsl_params& p, const gsl_fwd_inputs& in, gsl_fwd_outputs& out,
                          const GeomView& g, const ImageView& im, const BinView& b, int64_t r_capacity,
                          cudaStream_t st) {
  RenderParams rp = make_render_params(p, r_capacity);
  const int tiles = rp.gx * rp.gy;
  if (tiles == 0) return 0;
  const float4* colors = in.colors_precomp ? reinterpret_cast<const float4*>(in.colors_precomp) : g.rgb;
  ProfScope prof(GSL_K_RENDER_FWD, st);
#define GSL_LAUNCH_FWD(ST)                                                                              \
  k_render_fwd<ST><<<tiles, 256, 0, st>>>(rp, im.ranges, b.vals_b, g.rec, g.pixbox, colors, in.features, \
                                          in.background, g.ctrl, im.final_T, out.out_contrib,            \
                                          out.out_color, out.out_feature, out.out_depth, out.out_alpha)
  if (p.S == 4) GSL_LAUNCH_FWD(4);
  else if (p.S == 0) GSL_LAUNCH_FWD(0);
  else GSL_LAUNCH_FWD(-1);
#undef GSL_LAUNCH_FWD
  return check_cuda(cudaGetLastError(), "k_render_fwd launch");
}

}  // namespace gsl
