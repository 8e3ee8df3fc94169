// gsl_render_bwd.cu -- reverse-order backward compositing (semantics of backward.cu:137-515).
//
// One warp per 8x4 pixel block walking that block's list back to front (see gsl_render.cuh): 32 entries are
// staged at a time and every lane runs the reference's per-pixel recursion over exactly the staged entries
// its pixel contributed to in the forward pass (`pairmask`), at its own pace.  Differences from the
// reference's schedule (same sums):
//   * traversal starts at the last entry that contributed to the block instead of the end of the tile list,
//     and never evaluates a (pixel, surfel) pair that did not contribute;
//   * each contributing pair adds its 20 + S gradient components to the surfel's packed 128-B accumulator
//     record with 16-byte vector reductions (red.global.add.v4.f32) instead of ~21 scalar atomics.
// Summation order differs from the reference's atomics (which are unordered anyway): the contract
// is 1e-4 relative on the final gradients.
#include "gsl_render.cuh"

namespace gsl {

#ifdef GSL_STATS
__device__ unsigned long long g_stats_bwd[16];
#define STATB_ADD(i, v) do { unsigned long long _s = __reduce_add_sync(0xffffffffu, (unsigned)(v)); if ((threadIdx.x & 31) == 0) atomicAdd(&g_stats_bwd[i], _s); } while (0)
extern "C" __attribute__((visibility("default"))) void gsl_stats_read_bwd(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_stats_bwd, sizeof(g_stats_bwd));
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_stats_bwd, z, sizeof(z)); }
}
#endif

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

__device__ __forceinline__ void red_add_f32(float* addr, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}

// Transposing butterfly: on entry every lane holds K partial sums v[0..K); on exit v[0] of lane L holds the
// warp-wide total of component butterfly_component(K, L).  Stage `off` halves the number of live values:
// lanes with bit `off` clear keep the lower half and receive the partner's lower half, the others keep and
// receive the upper half.  Shuffles: ceil(K/2) + ceil(K/4) + ... (24 for K = 24) instead of 5 K.
template <int K, int OFF>
struct Butterfly {
  static constexpr int H = (K + 1) / 2;
  __device__ __forceinline__ static void run(float (&v)[32], int lane) {
    const bool upper = (lane & OFF) != 0;
#pragma unroll
    for (int i = 0; i < H; ++i) {
      const float a = v[i];
      const float b = (i + H < K) ? v[i + H] : 0.f;
      const float send = upper ? a : b;
      const float keep = upper ? b : a;
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
    }
    Butterfly<H, OFF / 2>::run(v, lane);
  }
};
template <int K>
struct Butterfly<K, 0> {
  __device__ __forceinline__ static void run(float (&)[32], int) {}
};
// Component whose total ends up in v[0] of `lane` (-1: structural padding).  Traces the slot kept at each
// stage back from the last stage to the first.
__device__ __forceinline__ int butterfly_component(int K0, int lane) {
  int Ks[5], Hs[5];
  int k = K0;
#pragma unroll
  for (int t = 0; t < 5; ++t) { Ks[t] = k; Hs[t] = (k + 1) / 2; k = Hs[t]; }
  int i = 0;
  bool pad = false;
#pragma unroll
  for (int t = 4; t >= 0; --t) {
    if (lane & (16 >> t)) i += Hs[t];
    pad = pad || (i >= Ks[t]);
  }
  return pad ? -1 : i;
}

template <int S_T>
__global__ void __launch_bounds__(32) k_render_bwd(
    RenderParams rp, const uint2* __restrict__ ranges, const uint4* __restrict__ bdesc,
    const uint2* __restrict__ blist, const uint32_t* __restrict__ pairmask, size_t plane_stride,
    const float4* __restrict__ rec, const float4* __restrict__ colors, const float* __restrict__ bg,
    const uint32_t* __restrict__ ctrl, const float* __restrict__ final_T, const int32_t* __restrict__ n_contrib,
    const float* __restrict__ dL_dpix, const float* __restrict__ dL_ddepth, const float* __restrict__ dL_dmask,
    const float* __restrict__ dL_dfeat, float* __restrict__ grad) {
  constexpr int KS = (S_T >= 0) ? S_T : GSL_MAX_FEATURES;  // feature slots held in registers
  constexpr int K = 20 + KS;                               // live components of the packed record
  const int S = (S_T >= 0) ? S_T : rp.S;
  __shared__ ChunkStage stg[3];

  const int lane = threadIdx.x;
  const BlockGeom bg_ = block_geom(rp, blockIdx.x, lane);
  const bool inside = bg_.inside;
  const int N = rp.W * rp.H;
  const int pix_id = bg_.pix_id;

  uint4 bd = bdesc[bg_.tile * 8 + bg_.bbit];
  if (ctrl[0] > rp.r_capacity) bd = make_uint4(0, 0, 0, 0);
  const uint32_t bs = bd.x, nwalk = min(bd.z, bd.y - bd.x);
  if (nwalk == 0) return;
  const uint32_t r0 = ranges[bg_.tile].x;
  const uint2* __restrict__ bl = blist + (size_t)bg_.bbit * plane_stride;
  const uint32_t* __restrict__ pmk = pairmask + (size_t)bg_.bbit * plane_stride;

  PixelRay ray = make_pixel_ray((float)bg_.pxi, (float)bg_.pyi, rp.HFOV_min, rp.HFOV_max, rp.VFOV_min, rp.VFOV_max,
                                rp.W, rp.H);
  ray.wrapW = rp.wrapW;
  const float T_final = inside ? final_T[pix_id] : 0.f;
  float T = T_final;
  const int median_contributor = inside ? n_contrib[pix_id + N] : 0;
  const float final_D = inside ? final_T[pix_id + N] : 0.f;
  const float final_D2 = inside ? final_T[pix_id + 2 * N] : 0.f;
  const float final_A = 1.f - T_final;

  float dpix[4] = {0.f, 0.f, 0.f, 0.f};
  float dfeat[KS > 0 ? KS : 1];
#pragma unroll
  for (int i = 0; i < KS; ++i) dfeat[i] = 0.f;
  float dnorm[3] = {0.f, 0.f, 0.f};
  float dL_depth = 0.f, dL_dmedian = 0.f, dL_ddist = 0.f, dL_depth_sq = 0.f, dL_mask = 0.f;
  if (inside) {
#pragma unroll
    for (int i = 0; i < 4; ++i) dpix[i] = dL_dpix[i * N + pix_id];
#pragma unroll
    for (int i = 0; i < KS; ++i)
      if (i < S) dfeat[i] = dL_dfeat[i * N + pix_id];
#pragma unroll
    for (int i = 0; i < 3; ++i) dnorm[i] = dL_dfeat[(S + i) * N + pix_id];
    dL_depth = dL_ddepth[pix_id];
    dL_dmedian = dL_ddepth[N + pix_id];
    dL_ddist = dL_ddepth[2 * N + pix_id];
    dL_depth_sq = dL_ddepth[3 * N + pix_id];
    dL_mask = dL_dmask[pix_id];
  }
  float bg_dot = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) bg_dot += bg[i] * dpix[i];

  float accum_rec[4] = {0.f, 0.f, 0.f, 0.f}, last_color[4] = {0.f, 0.f, 0.f, 0.f};
  float accum_n[3] = {0.f, 0.f, 0.f}, last_n[3] = {0.f, 0.f, 0.f};
  float accum_depth = 0.f, last_depth = 0.f, accum_mask = 0.f, last_alpha = 0.f, last_dL_dT = 0.f;

  const float far_near_over_range = rp.far_ * rp.near_ / (rp.far_ - rp.near_);
  int my_comp = butterfly_component(K, lane);
  if (my_comp == 19 || my_comp >= 20 + S) my_comp = -1;  // padding slots of the packed record
#ifdef GSL_STATS
  unsigned st_cand = 0, st_iter = 0, st_valid = 0, st_uni = 0;
#endif

  // slot j of a chunk ending at entry index c1 (exclusive) holds entry c1 - 1 - j: ascending slots go back to front
  auto issue = [&](ChunkStage& sb, const uint2 ent, bool act) {
    if (act) {
      const float4* r4 = rec + 4 * (size_t)ent.x;
      cp_async16(&sb.v[0][lane], r4);
      cp_async16(&sb.v[1][lane], r4 + 1);
      cp_async16(&sb.v[2][lane], r4 + 2);
      cp_async16(&sb.v[3][lane], r4 + 3);
      cp_async16(&sb.v[4][lane], colors + ent.x);
      sb.ent[lane] = ent;
    }
    cp_async_commit();
  };
  // chunk c (0 = deepest) holds entries hi-1-32c ... hi-32-32c; slot j of a chunk = entry top(c) - j
  const uint32_t hi = bs + nwalk;
  const int nC = (int)((nwalk + 31u) >> 5);
  auto load_entry = [&](int c, uint32_t& pm, uint2& ent) {
    pm = 0u;
    ent = make_uint2(0, 0);
    const uint32_t back = 32u * (uint32_t)c + (uint32_t)lane;  // distance from the last walked entry
    if (c < nC && back < nwalk) {
      const uint32_t e = hi - 1u - back;
      pm = __ldg(pmk + e);
      if (pm != 0u) ent = __ldg(bl + e);
    }
  };

  // Two chunks are live at a time (A = deeper, B = next towards the sensor); a lane that has finished its pairs of A
  // goes on with its pairs of B while slower lanes are still in A -- see gsl_render_fwd.cu.
  int a = 0;
  int bufA = 0, bufB = 1, bufC = 2;  // stage buffers of A, B and of the chunk in flight
  uint32_t mineA, mineB = 0;
  // A chunk whose entries are used by many pixels of the block each (splats larger than the block) is walked in
  // LOCK-STEP instead: all lanes take the warp's lowest pending entry together, so that its pairs are summed by the
  // butterfly below -- one reduction per entry instead of one per pair, and a tree sum instead of a serial one.
  constexpr uint32_t DENSE_BITS = 32u * 10u;  // >= 10 pixels per entry on average
  bool denseA, denseB = false;
  uint32_t pmI;
  uint2 entI;
  {
    uint32_t pm0, pm1;
    uint2 e0, e1;
    load_entry(0, pm0, e0);
    load_entry(1, pm1, e1);
    load_entry(2, pmI, entI);
    issue(stg[0], e0, pm0 != 0u);
    if (nC > 1) issue(stg[1], e1, pm1 != 0u);
    cp_async_wait_all();
    __syncwarp();
    mineA = transpose32(pm0, lane);  // bit j: my pixel contributed to the entry staged in slot j
    denseA = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(pm0)) >= DENSE_BITS;
    if (nC > 1) {
      mineB = transpose32(pm1, lane);
      denseB = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(pm1)) >= DENSE_BITS;
    }
#ifdef GSL_STATS
    if (lane == 0) st_cand += __popc(__ballot_sync(0xffffffffu, pm0 != 0u)) + __popc(__ballot_sync(0xffffffffu, pm1 != 0u));
    else { __ballot_sync(0xffffffffu, pm0 != 0u); __ballot_sync(0xffffffffu, pm1 != 0u); }
#endif
  }
  uint32_t pmB2 = pmI;  // pair masks of chunk a + 2 (gathers in flight)
  if (nC > 2) issue(stg[2], entI, pmI != 0u);
  load_entry(3, pmI, entI);
  for (;;) {
    while (__any_sync(0xffffffffu, mineA != 0)) {
      // Lanes normally sit on different entries.  When (almost) every pixel of the block is on the SAME entry --
      // splats much larger than the block -- the pairs are summed by a transposing butterfly and the record is
      // updated by one coalesced reduction instead of 32 colliding ones (also a more accurate sum).
      bool inA = mineA != 0;
      if (denseA) {  // lock-step: only the warp's lowest pending entry of A, no running ahead into B
        const uint32_t pending = __reduce_or_sync(0xffffffffu, mineA);
        inA = (mineA & (pending & (0u - pending))) != 0;
      }
      const bool active = inA || (!denseA && mineB != 0);
      const uint32_t amask = __ballot_sync(0xffffffffu, active);
      // slot key: bit 5 = chunk B
      const int j = inA ? (__ffs(mineA) - 1) : (mineB != 0 ? 32 + (__ffs(mineB) - 1) : -1);
      const int j0 = __shfl_sync(0xffffffffu, j, __ffs(amask) - 1);
      const bool uniform = __popc(amask) >= 8 && __all_sync(0xffffffffu, !active || j == j0);
#ifdef GSL_STATS
      if (lane == 0) { st_iter++; if (uniform) st_uni++; }
#endif
      // packed record: [0..8] dL_dT, [9..10] dL_dmean2D, [11] dL_dopacity, [12..15] dL_dcolor,
      // [16..18] dL_dnormal, [19] pad, [20..20+S) dL_dfeature
      float g[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) g[i] = 0.f;
      uint32_t sid = 0;
      if (active) {
        if (inA) mineA &= mineA - 1;
        else mineB &= mineB - 1;
        const ChunkStage& sb = stg[inA ? bufA : bufB];
        const int js = j & 31;
        const Splat s = staged_splat(sb, js);
        const PairEval e = eval_pair<true>(s, ray, rp.near_, rp.far_);  // exact like the forward: alpha feeds every term
#ifdef GSL_STATS
        st_valid++;
#endif
        const uint2 ent = sb.ent[js];
        sid = ent.x;
        const int pos0 = (int)(ent.y - r0);  // 0-based list position == the reference's `contributor` after --

        const float alpha = e.alpha, G = e.G, depth = e.depth;
        // the transmittance is a running product over the whole list: keep its division IEEE-exact
        T = T / (1.f - alpha);
        const float inv_1ma = fast_rcp(1.f - alpha);
        const float wgt = alpha * T;
        float dL_dalpha = 0.f;
        const float4 c4 = sb.v[4][js];
        const float col[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
          last_color[ch] = col[ch];
          dL_dalpha += (col[ch] - accum_rec[ch]) * dpix[ch];
          g[12 + ch] = wgt * dpix[ch];
        }
        float dL_dr = 0.f;
        dL_dr += alpha * T * dL_depth;
        dL_dr += alpha * T * 2 * depth * dL_depth_sq;
        if (pos0 == median_contributor - 1) dL_dr += dL_dmedian;

        const float inv_depth = fast_rcp(depth);
        const float m_d = rp.far_over_range * (1 - rp.near_ * inv_depth);
        const float dmd_dd = far_near_over_range * inv_depth * inv_depth;
        const float dL_dweight = (final_D2 + m_d * m_d * final_A - 2 * m_d * final_D) * dL_ddist;
        dL_dalpha += dL_dweight - last_dL_dT;
        last_dL_dT = dL_dweight * alpha + (1 - alpha) * last_dL_dT;
        const float dL_dmd = 2.0f * (T * alpha) * (m_d * final_A - final_D) * dL_ddist;
        dL_dr += dL_dmd * dmd_dd;

#pragma unroll
        for (int ch = 0; ch < KS; ++ch)
          if (ch < S) g[20 + ch] = wgt * dfeat[ch];  // features do not feed dL_dalpha (backward.cu:395)
        const float nrm[3] = {s.nx, s.ny, s.nz};
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          accum_n[ch] = last_alpha * last_n[ch] + (1.f - last_alpha) * accum_n[ch];
          last_n[ch] = nrm[ch];
          dL_dalpha += (nrm[ch] - accum_n[ch]) * dnorm[ch];
          g[16 + ch] = wgt * dnorm[ch];
        }
        accum_depth = last_alpha * last_depth + (1.f - last_alpha) * accum_depth;
        last_depth = depth;
        dL_dalpha += (depth - accum_depth) * dL_depth;
        accum_mask = last_alpha + (1.f - last_alpha) * accum_mask;
        dL_dalpha += (1.f - accum_mask) * dL_mask;
        dL_dalpha *= T;
        last_alpha = alpha;
        dL_dalpha += (-T_final * inv_1ma) * bg_dot;

        const float dL_dG = s.opacity * dL_dalpha;
        if (e.rho3d <= e.rho2d) {
          const float ex = ray.sth * ray.sphi, ey = -ray.cth, ez = ray.sth * ray.cphi;
          const float dsx = dL_dG * -G * e.sx + dL_dr * (s.Tux * ex + s.Tvx * ey + s.Twx * ez);
          const float dsy = dL_dG * -G * e.sy + dL_dr * (s.Tuy * ex + s.Tvy * ey + s.Twy * ez);
          const float ipz = fast_rcp(e.pz);
          const float qx = dsx * ipz, qy = dsy * ipz;
          const float dpx = qx, dpy = qy, dpz = -(qx * e.sx + qy * e.sy);
          // dL_dk = l x dL_dp ; dL_dl = dL_dp x k
          const float dkx = e.ly * dpz - e.lz * dpy, dky = e.lz * dpx - e.lx * dpz, dkz = e.lx * dpy - e.ly * dpx;
          const float dlx = dpy * e.kz - dpz * e.ky, dly = dpz * e.kx - dpx * e.kz, dlz = dpx * e.ky - dpy * e.kx;
          const float rx = dL_dr * ex, ry = dL_dr * ey, rz = dL_dr * ez;
          g[0] = ray.cphi * dkx + ray.sphi_cth * dlx + rx * e.sx;
          g[1] = ray.cphi * dky + ray.sphi_cth * dly + rx * e.sy;
          g[2] = ray.cphi * dkz + ray.sphi_cth * dlz + rx;
          g[3] = ray.sth * dlx + ry * e.sx;
          g[4] = ray.sth * dly + ry * e.sy;
          g[5] = ray.sth * dlz + ry;
          g[6] = -ray.sphi * dkx + ray.cphi_cth * dlx + rz * e.sx;
          g[7] = -ray.sphi * dky + ray.cphi_cth * dly + rz * e.sy;
          g[8] = -ray.sphi * dkz + ray.cphi_cth * dlz + rz;
        } else {
          g[9] = dL_dG * (-G * 2.f * e.dx);
          g[10] = dL_dG * (-G * 2.f * e.dy);
          const float dr_d = dL_dr * inv_depth;
          g[2] = dr_d * s.Tuz;
          g[5] = dr_d * s.Tvz;
          g[8] = dr_d * s.Twz;
        }
        g[11] = G * dL_dalpha;

        if (!uniform) {
          float* gdst = grad + (size_t)sid * 32;
          red_add_v4(gdst + 0, g[0], g[1], g[2], g[3]);
          red_add_v4(gdst + 4, g[4], g[5], g[6], g[7]);
          red_add_v4(gdst + 8, g[8], g[9], g[10], g[11]);
          red_add_v4(gdst + 12, g[12], g[13], g[14], g[15]);
          red_add_v4(gdst + 16, g[16], g[17], g[18], 0.f);
          if (KS > 0) red_add_v4(gdst + 20, g[20], g[21], g[22], g[23]);
          if (KS > 4) red_add_v4(gdst + 24, g[24], g[25], g[26], g[27]);
          if (KS > 8) red_add_v4(gdst + 28, g[28], g[29], 0.f, 0.f);
        }
      }
      if (uniform) {
        const uint32_t sid0 = __shfl_sync(0xffffffffu, sid, __ffs(amask) - 1);
        Butterfly<K, 16>::run(g, lane);
        if (my_comp >= 0) red_add_f32(grad + (size_t)sid0 * 32 + my_comp, g[0]);
      }
    }
    // every lane is through chunk A: retire it, B becomes A, the chunk whose gathers were in flight becomes B
    if (a + 1 >= nC) break;
    ++a;
    mineA = mineB;
    mineB = 0;
    denseA = denseB;
    denseB = false;
    { const int t = bufA; bufA = bufB; bufB = bufC; bufC = t; }  // the retired buffer receives the next gathers
    if (a + 1 < nC) {
      cp_async_wait_all();
      __syncwarp();
      mineB = transpose32(pmB2, lane);
      denseB = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(pmB2)) >= DENSE_BITS;
#ifdef GSL_STATS
      if (lane == 0) st_cand += __popc(__ballot_sync(0xffffffffu, pmB2 != 0u));
      else __ballot_sync(0xffffffffu, pmB2 != 0u);
#endif
      pmB2 = pmI;
      if (a + 2 < nC) issue(stg[bufC], entI, pmI != 0u);
      load_entry(a + 3, pmI, entI);
    }
  }
  cp_async_wait_all();
#ifdef GSL_STATS
  STATB_ADD(0, st_cand); STATB_ADD(1, st_iter); STATB_ADD(3, st_valid); STATB_ADD(4, st_uni);
#endif
}

int launch_render_backward(const gsl_params& p, const gsl_fwd_inputs& in, const gsl_fwd_outputs& fwd,
                           const gsl_bwd_inputs& gin, const GeomView& g, const ImageView& im, const BinView& b,
                           int64_t r_capacity, cudaStream_t st) {
  RenderParams rp = make_render_params(p, r_capacity);
  const int nblocks = ((p.W + 7) / 8) * ((p.H + 3) / 4);
  if (nblocks == 0 || p.P == 0) return 0;
  const float4* colors = in.colors_precomp ? reinterpret_cast<const float4*>(in.colors_precomp) : g.rgb;
  static_assert(sizeof(float) * 32 == 128, "packed accumulator record is one 128-B line");
  ProfScope prof(GSL_K_RENDER_BWD, st);
#define GSL_LAUNCH_BWD(ST)                                                                                \
  k_render_bwd<ST><<<nblocks, 32, 0, st>>>(rp, im.ranges, im.bdesc, b.blist, b.pairmask, b.plane_stride, g.rec, \
                                           colors, in.background, g.ctrl, im.final_T, fwd.out_contrib,        \
                                           gin.dL_dout_color, gin.dL_dout_depth, gin.dL_dout_alpha,            \
                                           gin.dL_dout_feature, g.grad)
  if (p.S == 4) GSL_LAUNCH_BWD(4);
  else if (p.S == 0) GSL_LAUNCH_BWD(0);
  else GSL_LAUNCH_BWD(-1);
#undef GSL_LAUNCH_BWD
  return check_cuda(cudaGetLastError(), "k_render_bwd launch");
}

}  // namespace gsl
