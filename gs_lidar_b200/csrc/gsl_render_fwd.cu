// gsl_render_fwd.cu -- front-to-back alpha compositing (semantics of forward.cu:292-505).
//
// One warp per 8x4 pixel block walking that block's list (see gsl_render.cuh for the decomposition and
// the pipeline): 32 entries are staged at a time and every lane composites, in list order, the staged
// entries whose pixel box contains its pixel.  Per-pair arithmetic is gsl::eval_pair and the blend
// recursion follows the reference's rounding sequence, so the maps match the reference bit-for-bit in
// practice (contract: 1e-5 relative); `contributor` numbers are the reference's 16x16-tile list positions.
// The forward also stores, per block-list entry, the mask of pixels it contributed to (`pairmask`) and,
// per block, how many leading entries the backward pass has to walk.
#include "gsl_render.cuh"

namespace gsl {

#ifndef GSL_FWD_ILP
#define GSL_FWD_ILP 2
#endif

#ifdef GSL_STATS
__device__ unsigned long long g_stats[16];
#define STAT_ADD(i, v) do { unsigned long long _s = __reduce_add_sync(0xffffffffu, (unsigned)(v)); if ((threadIdx.x & 31) == 0) atomicAdd(&g_stats[i], _s); } while (0)
#endif

template <int S_T>
__global__ void __launch_bounds__(32) k_render_fwd(
    RenderParams rp, const uint2* __restrict__ ranges, uint4* __restrict__ bdesc, const uint2* __restrict__ blist,
    uint32_t* __restrict__ pairmask, size_t plane_stride, const float4* __restrict__ rec,
    const float4* __restrict__ colors, const float* __restrict__ features, const short4* __restrict__ pixbox,
    const float* __restrict__ bg, const uint32_t* __restrict__ ctrl, uint8_t* __restrict__ touched,
    float* __restrict__ final_T, int32_t* __restrict__ out_contrib, float* __restrict__ out_color, float* __restrict__ out_feature,
    float* __restrict__ out_depth, float* __restrict__ out_alpha) {
  constexpr bool FEAT4 = (S_T == 4);
  const int S = (S_T >= 0) ? S_T : rp.S;
  __shared__ ChunkStage stg[2];

  const int lane = threadIdx.x;
  const BlockGeom bg_ = block_geom(rp, blockIdx.x, lane);
  const bool inside = bg_.inside;
  const int N = rp.W * rp.H;
  const int pix_id = bg_.pix_id;
  const int bidx = bg_.tile * 8 + bg_.bbit;

  uint4 bd = bdesc[bidx];
  if (ctrl[0] > rp.r_capacity) bd = make_uint4(0, 0, 0, 0);
  const uint32_t bs = bd.x, be = bd.y;
  const uint32_t r0 = ranges[bg_.tile].x;
  const uint2* __restrict__ bl = blist + (size_t)bg_.bbit * plane_stride;
  uint32_t* __restrict__ pmk = pairmask + (size_t)bg_.bbit * plane_stride;

  PixelRay ray = make_pixel_ray((float)bg_.pxi, (float)bg_.pyi, rp.HFOV_min, rp.HFOV_max, rp.VFOV_min, rp.VFOV_max,
                                rp.W, rp.H);
  ray.wrapW = rp.wrapW;
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
  unsigned st_cand = 0, st_iter = 0, st_any = 0, st_valid = 0, st_eval = 0;
#endif

  // gathers of one staged candidate (lane = slot), straight into shared memory
  auto issue = [&](ChunkStage& sb, const uint2 ent, bool act) {
    if (act) {
      const float4* r4 = rec + 4 * (size_t)ent.x;
      cp_async16(&sb.v[0][lane], r4);
      cp_async16(&sb.v[1][lane], r4 + 1);
      cp_async16(&sb.v[2][lane], r4 + 2);
      cp_async16(&sb.v[3][lane], r4 + 3);
      cp_async16(&sb.v[4][lane], colors + ent.x);
      if (FEAT4) cp_async16(&sb.v[5][lane], reinterpret_cast<const float4*>(features) + ent.x);
      cp_async8(&sb.box[lane], pixbox + ent.x);
      sb.ent[lane] = ent;
    }
    cp_async_commit();
  };

  uint32_t walk = 0;  // entries (from the start of the block list) up to the last one that contributed
  if (be > bs) {
    int cur = 0;
    uint2 entN = make_uint2(0, 0);
    {
      uint2 ent0 = make_uint2(0, 0);
      if (bs + lane < be) ent0 = __ldg(bl + bs + lane);
      if (bs + 32 + lane < be) entN = __ldg(bl + bs + 32 + lane);
      issue(stg[0], ent0, bs + lane < be);
    }
    for (uint32_t c0 = bs; c0 < be; c0 += 32) {
      const int n = (int)min(32u, be - c0);
      cp_async_wait_all();
      __syncwarp();
      // ---- later pipeline stages: gathers of chunk c0+32, entries of chunk c0+64
      issue(stg[cur ^ 1], entN, c0 + 32 + lane < be);
      uint2 entNN = make_uint2(0, 0);
      if (c0 + 64 + lane < be) entNN = __ldg(bl + c0 + 64 + lane);
      // ---- which staged entries can reach my pixel
      const ChunkStage& sb = stg[cur];
      uint32_t bm = 0;
      if (lane < n) bm = box_pixel_mask(sb.box[lane], bg_.bx0, bg_.by0);
      uint32_t mine = transpose32(bm, lane);
      if (done) mine = 0;
      uint32_t contributed = 0;
#ifdef GSL_STATS
      if (lane == 0) st_cand += n;
#endif
      // One composited pair: the reference's blend recursion (forward.cu:449-487) in its rounding sequence.
      auto blend = [&](const int j, const Splat& sp, const PairEval& e) {
        const float alpha = e.alpha;
        const float test_T = GSL_FM(T, GSL_FS(1.f, alpha));
        if (test_T < 0.0001f) {
          done = true;
          mine = 0;
          return;
        }
        const int pos = (int)(sb.ent[j].y - r0) + 1;  // 1-based list position (the reference's `contributor`)
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
        const float4 col = sb.v[4][j];
        C[0] = GSL_FF(T, GSL_FM(alpha, col.x), C[0]);
        C[1] = GSL_FF(T, GSL_FM(alpha, col.y), C[1]);
        C[2] = GSL_FF(T, GSL_FM(alpha, col.z), C[2]);
        C[3] = GSL_FF(T, GSL_FM(alpha, col.w), C[3]);
        if (FEAT4) {
          const float4 f = sb.v[5][j];
          F[0] = GSL_FF(T, GSL_FM(alpha, f.x), F[0]);
          F[1] = GSL_FF(T, GSL_FM(alpha, f.y), F[1]);
          F[2] = GSL_FF(T, GSL_FM(alpha, f.z), F[2]);
          F[3] = GSL_FF(T, GSL_FM(alpha, f.w), F[3]);
        } else if (S > 0) {
          const float* fp = features + (size_t)sb.ent[j].x * S;
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
        contributed |= 1u << j;
      };
      while (__any_sync(0xffffffffu, mine != 0)) {
#ifdef GSL_STATS
        if (lane == 0) st_iter++;
#endif
        if (mine != 0) {
#if GSL_FWD_ILP == 2
          // TWO entries per turn: their pair evaluations are independent and interleave in the pipeline (the warps of
          // this kernel are few -- one per 8x4 block -- so a lane's own instruction-level parallelism is what hides the
          // latency of the divisions and the exponential); the blends stay in list order.
          const int j = __ffs(mine) - 1;
          mine &= mine - 1;
          const bool two = mine != 0;
          const int j2 = two ? __ffs(mine) - 1 : j;
          mine &= mine - 1;  // 0 stays 0
          const Splat sp = staged_splat(sb, j);
          const Splat sp2 = staged_splat(sb, j2);
          const PairEval e = eval_pair<false>(sp, ray, rp.near_, rp.far_);
          const PairEval e2 = eval_pair<false>(sp2, ray, rp.near_, rp.far_);
#ifdef GSL_STATS
          st_eval += two ? 2 : 1;
          if (e.valid) st_valid++;
          if (two && e2.valid) st_valid++;
#endif
          if (e.valid) blend(j, sp, e);
          if (two && !done && e2.valid) blend(j2, sp2, e2);
#else
          const int j = __ffs(mine) - 1;
          mine &= mine - 1;
          const Splat sp = staged_splat(sb, j);
          const PairEval e = eval_pair<false>(sp, ray, rp.near_, rp.far_);
#ifdef GSL_STATS
          st_eval++;
          if (e.valid) st_valid++;
#endif
          if (e.valid) blend(j, sp, e);
#endif
        }
      }
      // ---- per-entry masks of the pixels it contributed to
      const uint32_t pm = transpose32(contributed, lane);
      if (lane < n) {
        pmk[c0 + lane] = pm;
        if (pm != 0u) touched[sb.ent[lane].x] = 1;  // idempotent byte store: the backward skips every other surfel's accumulator
      }
      const uint32_t nz = __ballot_sync(0xffffffffu, pm != 0u);
      if (nz) walk = (c0 - bs) + (32u - (uint32_t)__clz(nz));
#ifdef GSL_STATS
      if (lane == 0) st_any += __popc(nz);
#endif
      if (__all_sync(0xffffffffu, done)) break;
      cur ^= 1;
      entN = entNN;
    }
    cp_async_wait_all();
  }
  if (lane == 0) bdesc[bidx].z = walk;
#ifdef GSL_STATS
  STAT_ADD(0, st_cand); STAT_ADD(1, st_iter); STAT_ADD(2, st_any); STAT_ADD(3, st_valid); STAT_ADD(4, st_eval);
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
  rp.wrapW = (p.flags & GSL_FLAG_WRAP_AZIMUTH) ? (float)p.W : 0.f;
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
  k_render_fwd<ST><<<nblocks, 32, 0, st>>>(rp, im.ranges, im.bdesc, b.blist, b.pairmask, b.plane_stride, g.rec, \
                                           colors, in.features, g.pixbox, in.background, g.ctrl, g.touched, im.final_T,  \
                                           out.out_contrib, out.out_color, out.out_feature, out.out_depth,   \
                                           out.out_alpha)
  if (p.S == 4) GSL_LAUNCH_FWD(4);
  else if (p.S == 0) GSL_LAUNCH_FWD(0);
  else GSL_LAUNCH_FWD(-1);
#undef GSL_LAUNCH_FWD
  return check_cuda(cudaGetLastError(), "k_render_fwd launch");
}

}  // namespace gsl
