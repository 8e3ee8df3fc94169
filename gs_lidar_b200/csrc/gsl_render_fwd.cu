// gsl_render_fwd.cu -- front-to-back alpha compositing (semantics of forward.cu:292-505).
//
// One warp per 8x4 pixel block (see gsl_render.cuh for the decomposition and the pipeline).  Per-pair
// arithmetic is gsl::eval_pair and the blend recursion follows the reference's rounding sequence, so the
// maps match the reference bit-for-bit in practice (contract: 1e-5 relative).  Two consecutive candidates
// are evaluated together (their ray-splat intersections are independent; only the short blend is serial).
// The forward also records, per block and list position, whether the entry contributed to any pixel of
// the block (bit-planes `used`); the backward pass walks only those.
#include "gsl_render.cuh"

namespace gsl {

#ifdef GSL_STATS
__device__ unsigned long long g_stats[16];
#define STAT_ADD(i, v) do { unsigned long long _s = __reduce_add_sync(0xffffffffu, (unsigned)(v)); if ((threadIdx.x & 31) == 0) atomicAdd(&g_stats[i], _s); } while (0)
#endif

template <int S_T>
__global__ void __launch_bounds__(32) k_render_fwd(
    RenderParams rp, const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
    const uint8_t* __restrict__ bmask, const float4* __restrict__ rec, const float4* __restrict__ colors,
    const float* __restrict__ features, const float* __restrict__ bg, const uint32_t* __restrict__ ctrl,
    uint32_t* __restrict__ used, size_t used_words, float* __restrict__ final_T, int32_t* __restrict__ out_contrib,
    float* __restrict__ out_color, float* __restrict__ out_feature, float* __restrict__ out_depth,
    float* __restrict__ out_alpha) {
  constexpr bool FEAT4 = (S_T == 4);
  const int S = (S_T >= 0) ? S_T : rp.S;
  __shared__ WarpStage stg[2];

  const int lane = threadIdx.x;
  const BlockGeom bg_ = block_geom(rp, blockIdx.x, lane);
  const int bbit = bg_.bbit;
  const bool inside = bg_.inside;
  const int N = rp.W * rp.H;
  const int pix_id = bg_.pix_id;

  uint2 range = ranges[bg_.tile];
  if (ctrl[0] > rp.r_capacity) range = make_uint2(0, 0);
  const uint32_t r0 = range.x, r1 = range.y;

  const PixelRay ray = make_pixel_ray((float)bg_.pxi, (float)bg_.pyi, rp.HFOV_min, rp.HFOV_max, rp.VFOV_min,
                                      rp.VFOV_max, rp.W, rp.H);
  bool done = !inside;
  float T = 1.0f;
  int last_contributor = 0, median_contributor = 0;
  float C[4] = {0.f, 0.f, 0.f, 0.f};
  float F[GSL_MAX_FEATURES];
#pragma unroll
  for (int i = 0; i < GSL_MAX_FEATURES; ++i) F[i] = 0.f;
  float Nn[3] = {0.f, 0.f, 0.f};
  float D = 0.f, D2 = 0.f, M1 = 0.f, M2 = 0.f, distortion = 0.f, median_depth = 0.f;
#ifdef GSL_STATS
  unsigned st_scan = 0, st_box = 0, st_any = 0, st_valid = 0, st_eval_lanes = 0;
#endif

  // blend of one evaluated candidate into this lane's pixel; returns whether it contributed
  auto blend = [&](const WarpStage& sb, int s, const Splat& sp, const PairEval& e, uint32_t wbase) -> bool {
    if (done || !e.valid) return false;
    const float alpha = e.alpha;
    const float test_T = GSL_FM(T, GSL_FS(1.f, alpha));
    if (test_T < 0.0001f) {
      done = true;
      return false;
    }
    const int pos = (int)(wbase + sb.lanepos[s] - r0) + 1;  // 1-based list position (the reference's `contributor`)
    const float wgt = GSL_FM(T, alpha);
    const float A = GSL_FS(1.f, T);
    const float mm1 = GSL_FM(rp.far_over_range, GSL_FS(1.f, GSL_FD(rp.near_, e.depth)));
    const float mm = GSL_FM(mm1, mm1);
    const float t0 = GSL_FF(-M1, GSL_FA(mm1, mm1), GSL_FF(A, mm, M2));
    distortion = GSL_FF(wgt, t0, distortion);
    M1 = GSL_FF(wgt, mm1, M1);
    M2 = GSL_FF(wgt, mm, M2);
    if (T > 0.5f) {
      median_depth = e.depth;
      median_contributor = pos;
    }
    const float4 col = sb.v[4][s];
    C[0] = GSL_FF(T, GSL_FM(alpha, col.x), C[0]);
    C[1] = GSL_FF(T, GSL_FM(alpha, col.y), C[1]);
    C[2] = GSL_FF(T, GSL_FM(alpha, col.z), C[2]);
    C[3] = GSL_FF(T, GSL_FM(alpha, col.w), C[3]);
    if (FEAT4) {
      const float4 f = sb.v[5][s];
      F[0] = GSL_FF(T, GSL_FM(alpha, f.x), F[0]);
      F[1] = GSL_FF(T, GSL_FM(alpha, f.y), F[1]);
      F[2] = GSL_FF(T, GSL_FM(alpha, f.z), F[2]);
      F[3] = GSL_FF(T, GSL_FM(alpha, f.w), F[3]);
    } else if (S > 0) {
      const float* fp = features + (size_t)sb.id[s] * S;
#pragma unroll
      for (int ch = 0; ch < GSL_MAX_FEATURES; ++ch)
        if (ch < S) F[ch] = GSL_FF(T, GSL_FM(alpha, __ldg(fp + ch)), F[ch]);
    }
    Nn[0] = GSL_FF(T, GSL_FM(alpha, sp.nx), Nn[0]);
    Nn[1] = GSL_FF(T, GSL_FM(alpha, sp.ny), Nn[1]);
    Nn[2] = GSL_FF(T, GSL_FM(alpha, sp.nz), Nn[2]);
    D = GSL_FF(T, GSL_FM(alpha, e.depth), D);
    D2 = GSL_FF(T, GSL_FM(alpha, GSL_FM(e.depth, e.depth)), D2);
    T = test_T;
    last_contributor = pos;
    return true;
  };

  if (r1 > r0) {
    const uint32_t w0 = r0 >> 5, w1 = (r1 - 1) >> 5;
    uint32_t* __restrict__ used_plane = used + (size_t)bbit * used_words;
    const uint32_t lt_mask = (1u << lane) - 1u;

    auto load_mi = [&](uint32_t w, bool& cand, uint32_t& id) {
      const uint32_t p = (w << 5) + lane;
      const bool in = (w <= w1) && p >= r0 && p < r1;
      cand = in && ((__ldg(bmask + p) >> bbit) & 1u);
      id = cand ? __ldg(point_list + p) : 0u;
    };

    // prologue: word w0 staged in buffer 0, ids of word w0+1 in flight
    bool candN, candNN;
    uint32_t idN, idNN;
    CandRegs rg;
    int cur = 0;
    uint32_t mCur;
    {
      bool c0;
      uint32_t id0;
      load_mi(w0, c0, id0);
      gather_cand<FEAT4>(c0, id0, rec, colors, features, rg);
      load_mi(w0 + 1, candN, idN);
      mCur = __ballot_sync(0xffffffffu, c0);
      if (c0) stage_cand<FEAT4>(stg[0], __popc(mCur & lt_mask), rg, id0, (uint32_t)lane);
      __syncwarp();
    }
    for (uint32_t w = w0; w <= w1; ++w) {
      // ---- later pipeline stages: records of word w+1, mask + ids of word w+2
      gather_cand<FEAT4>(candN, idN, rec, colors, features, rg);
      load_mi(w + 2, candNN, idNN);
      // ---- composite the staged candidates of word w
      const WarpStage& sb = stg[cur];
      const int cnt = __popc(mCur);
      const uint32_t wbase = w << 5;
#ifdef GSL_STATS
      if (lane == 0) { st_scan += min(32u, r1 - max(r0, wbase)); st_box += cnt; }
#endif
      uint32_t slotmask = 0;
      for (int s = 0; s < cnt; s += 2) {
        if (__all_sync(0xffffffffu, done)) break;
        const bool two = s + 1 < cnt;  // warp-uniform
        const int s1 = two ? s + 1 : s;
        const Splat spa = staged_splat(sb, s);
        const Splat spb = staged_splat(sb, s1);
        const PairEval ea = eval_pair<false>(spa, ray, rp.near_, rp.far_);
        PairEval eb = eval_pair<false>(spb, ray, rp.near_, rp.far_);
        eb.valid = eb.valid && two;
#ifdef GSL_STATS
        if (!done) { st_eval_lanes += two ? 2 : 1; st_valid += (ea.valid ? 1 : 0) + (eb.valid ? 1 : 0); }
#endif
        const bool ca = blend(sb, s, spa, ea, wbase);
        const bool cb = blend(sb, s1, spb, eb, wbase);
        if (__any_sync(0xffffffffu, ca)) slotmask |= 1u << s;
        if (__any_sync(0xffffffffu, cb)) slotmask |= 1u << s1;
      }
      if (slotmask != 0u) {
        // slot mask -> list-position mask of this word
        const bool mine = ((mCur >> lane) & 1u) && ((slotmask >> __popc(mCur & lt_mask)) & 1u);
        const uint32_t usedbits = __ballot_sync(0xffffffffu, mine);
#ifdef GSL_STATS
        if (lane == 0) st_any += __popc(usedbits);
#endif
        if (lane == 0) {
          if (w == w0 || w == w1) atomicOr(&used_plane[w], usedbits);  // word shared with the neighbouring tile
          else used_plane[w] = usedbits;
        }
      }
      if (__all_sync(0xffffffffu, done)) break;
      // ---- stage word w+1 into the other buffer, rotate
      cur ^= 1;
      mCur = __ballot_sync(0xffffffffu, candN);
      if (candN) stage_cand<FEAT4>(stg[cur], __popc(mCur & lt_mask), rg, idN, (uint32_t)lane);
      __syncwarp();
      candN = candNN;
      idN = idNN;
    }
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
  const int nblocks = ((p.W + 7) / 8) * ((p.H + 3) / 4);
  if (nblocks == 0) return 0;
  const float4* colors = in.colors_precomp ? reinterpret_cast<const float4*>(in.colors_precomp) : g.rgb;
  ProfScope prof(GSL_K_RENDER_FWD, st);
#define GSL_LAUNCH_FWD(ST)                                                                                  \
  k_render_fwd<ST><<<nblocks, 32, 0, st>>>(rp, im.ranges, b.vals_b, b.bmask, g.rec, colors, in.features,     \
                                           in.background, g.ctrl, b.used, b.used_words, im.final_T,          \
                                           out.out_contrib, out.out_color, out.out_feature, out.out_depth,  \
                                           out.out_alpha)
  if (p.S == 4) GSL_LAUNCH_FWD(4);
  else if (p.S == 0) GSL_LAUNCH_FWD(0);
  else GSL_LAUNCH_FWD(-1);
#undef GSL_LAUNCH_FWD
  return check_cuda(cudaGetLastError(), "k_render_fwd launch");
}

}  // namespace gsl
