// gsl_render_bwd.cu -- reverse-order backward compositing (semantics of backward.cu:137-515).
//
// One CTA per 16x16 tile, one thread per pixel, warps own 8x4 pixel blocks (same mapping as the
// forward).  Differences from the reference's schedule (results are the same sums):
//   * traversal starts at the tile's largest last_contributor instead of the end of the list;
//   * whole warps skip surfels whose conservative pixel box misses their 8x4 block;
//   * the per-pair gradient contributions of a warp are reduced with xor-shuffles and flushed by
//     six lanes with one 16-byte vector reduction each (red.global.add.v4.f32) into the packed
//     per-surfel accumulator, instead of ~21 scalar atomics per (pixel, surfel) pair.
// Summation order differs from the reference's atomics (which are unordered anyway): the contract
// is 1e-4 relative on the final gradients.
#include "gsl_common.cuh"
#include "gsl_math.cuh"

namespace gsl {

constexpr int BWD_BATCH = 128;

#ifdef GSL_STATS
__device__ unsigned long long g_stats_bwd[16];
#define STATB_ADD(i, v) do { unsigned long long _s = __reduce_add_sync(0xffffffffu, (unsigned)(v)); if ((threadIdx.x & 31) == 0) atomicAdd(&g_stats_bwd[i], _s); } while (0)
extern "C" __attribute__((visibility("default"))) void gsl_stats_read_bwd(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_stats_bwd, sizeof(g_stats_bwd));
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_stats_bwd, z, sizeof(z)); }
}
#endif

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

template <int S_T>
__global__ void __launch_bounds__(256) k_render_bwd(
    RenderParams rp, const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
    const float4* __restrict__ rec, const short4* __restrict__ pixbox, const float4* __restrict__ colors,
    const float* __restrict__ features, const float* __restrict__ bg, const uint32_t* __restrict__ ctrl,
    const float* __restrict__ final_T, const int32_t* __restrict__ n_contrib,
    const float* __restrict__ dL_dpix, const float* __restrict__ dL_ddepth, const float* __restrict__ dL_dmask,
    const float* __restrict__ dL_dfeat, float* __restrict__ grad, int gstride) {
  const int S = (S_T >= 0) ? S_T : rp.S;
  __shared__ float4 s_rec[4][BWD_BATCH];
  __shared__ float4 s_col[BWD_BATCH];
  __shared__ uint32_t s_id[BWD_BATCH];
  __shared__ short4 s_box[BWD_BATCH];
  __shared__ int s_max;

  const int tile = blockIdx.x;
  const int tx = tile % rp.gx, ty = tile / rp.gx;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bx0 = tx * GSL_BLOCK_X + (warp & 1) * 8, by0 = ty * GSL_BLOCK_Y + (warp >> 1) * 4;
  const int pxi = bx0 + (lane & 7), pyi = by0 + (lane >> 3);
  const bool inside = pxi < rp.W && pyi < rp.H;
  const int N = rp.W * rp.H;
  const int pix_id = rp.W * pyi + pxi;
  const int wbx1 = min(bx0 + 7, rp.W - 1), wby1 = min(by0 + 3, rp.H - 1);

  uint2 range = ranges[tile];
  if (ctrl[0] > rp.r_capacity) range = make_uint2(0, 0);

  const PixelRay ray = make_pixel_ray((float)pxi, (float)pyi, rp.HFOV_min, rp.HFOV_max, rp.VFOV_min,
                                      rp.VFOV_max, rp.W, rp.H);
  const float T_final = inside ? final_T[pix_id] : 0.f;
  float T = T_final;
  const int last_contributor = inside ? n_contrib[pix_id] : 0;
  const int median_contributor = inside ? n_contrib[pix_id + N] : 0;
  const float final_D = inside ? final_T[pix_id + N] : 0.f;
  const float final_D2 = inside ? final_T[pix_id + 2 * N] : 0.f;
  const float final_A = 1.f - T_final;

  float dpix[4] = {0.f, 0.f, 0.f, 0.f};
  float dfeat[GSL_MAX_FEATURES];
#pragma unroll
  for (int i = 0; i < GSL_MAX_FEATURES; ++i) dfeat[i] = 0.f;
  float dnorm[3] = {0.f, 0.f, 0.f};
  float dL_depth = 0.f, dL_dmedian = 0.f, dL_ddist = 0.f, dL_depth_sq = 0.f, dL_mask = 0.f;
  if (inside) {
#pragma unroll
    for (int i = 0; i < 4; ++i) dpix[i] = dL_dpix[i * N + pix_id];
#pragma unroll
    for (int i = 0; i < GSL_MAX_FEATURES; ++i)
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

  // start of the traversal: the deepest list position any pixel of the tile used
  if (threadIdx.x == 0) s_max = 0;
  __syncthreads();
  const int warp_max = __reduce_max_sync(0xffffffffu, last_contributor);
  if (lane == 0 && warp_max > 0) atomicMax(&s_max, warp_max);
  __syncthreads();
  const int total = min(s_max, (int)(range.y - range.x));

#ifdef GSL_STATS
  unsigned st_scan = 0, st_box = 0, st_any = 0, st_valid = 0, st_multi = 0;
#endif
  const float far_near = rp.far_ * rp.near_;
  const float range_fn = rp.far_ - rp.near_;

  // batches from the back: positions [lo, hi)
  for (int hi = total; hi > 0; hi -= BWD_BATCH) {
    const int lo = max(0, hi - BWD_BATCH);
    const int nb = hi - lo;
    __syncthreads();
    if ((int)threadIdx.x < nb) {
      uint32_t id = point_list[range.x + lo + threadIdx.x];
      s_id[threadIdx.x] = id;
      s_box[threadIdx.x] = pixbox[id];
      const float4* r4 = rec + 4 * (size_t)id;
      s_rec[0][threadIdx.x] = r4[0];
      s_rec[1][threadIdx.x] = r4[1];
      s_rec[2][threadIdx.x] = r4[2];
      s_rec[3][threadIdx.x] = r4[3];
      s_col[threadIdx.x] = colors[id];
    }
    __syncthreads();
    for (int j = nb - 1; j >= 0; --j) {
      const int pos0 = lo + j;  // 0-based list position == the reference's `contributor` after --
      if (pos0 >= warp_max) continue;  // warp-uniform
#ifdef GSL_STATS
      if (lane == 0) st_scan++;
#endif
      const short4 bb = s_box[j];
      const bool ovy = (int)bb.y <= wby1 && (int)bb.w >= by0;
      const bool ovx = (bb.x <= bb.z) ? ((int)bb.x <= wbx1 && (int)bb.z >= bx0)
                                      : ((int)bb.x <= wbx1 || (int)bb.z >= bx0);
      if (!(ovx && ovy)) continue;  // warp-uniform
#ifdef GSL_STATS
      if (lane == 0) st_box++;
#endif

      bool valid = pos0 < last_contributor;
      Splat s;
      {
        float4 a = s_rec[0][j], b = s_rec[1][j], c = s_rec[2][j], d = s_rec[3][j];
        s.Tux = a.x; s.Tuy = a.y; s.Tuz = a.z; s.Tvx = a.w;
        s.Tvy = b.x; s.Tvz = b.y; s.Twx = b.z; s.Twy = b.w;
        s.Twz = c.x; s.mx = c.y; s.my = c.z; s.opacity = c.w;
        s.nx = d.x; s.ny = d.y; s.nz = d.z; s.depth = d.w;
      }
      PairEval e;
      e.valid = false;
      if (valid) e = eval_pair<true>(s, ray, rp.near_, rp.far_);
      valid = valid && e.valid;
      const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
      if (vmask == 0) continue;
#ifdef GSL_STATS
      if (lane == 0) { st_any++; if (__popc(vmask) > 1) st_multi++; }
      if (valid) st_valid++;
#endif

      float g_dT[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      float g_m2x = 0.f, g_m2y = 0.f, g_op = 0.f;
      float g_col[4] = {0.f, 0.f, 0.f, 0.f}, g_nrm[3] = {0.f, 0.f, 0.f};
      float g_feat[GSL_MAX_FEATURES];
#pragma unroll
      for (int i = 0; i < GSL_MAX_FEATURES; ++i) g_feat[i] = 0.f;

      if (valid) {
        const float alpha = e.alpha, G = e.G, depth = e.depth;
        T = T / (1.f - alpha);
        const float wgt = alpha * T;
        float dL_dalpha = 0.f;
        const float4 c4 = s_col[j];
        const float col[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
          last_color[ch] = col[ch];
          dL_dalpha += (col[ch] - accum_rec[ch]) * dpix[ch];
          g_col[ch] = wgt * dpix[ch];
        }
        float dL_dr = 0.f;
        dL_dr += alpha * T * dL_depth;
        dL_dr += alpha * T * 2 * depth * dL_depth_sq;
        if (pos0 == median_contributor - 1) dL_dr += dL_dmedian;

        const float m_d = rp.far_over_range * (1 - rp.near_ / depth);
        const float dmd_dd = far_near / (range_fn * depth * depth);
        const float dL_dweight = (final_D2 + m_d * m_d * final_A - 2 * m_d * final_D) * dL_ddist;
        dL_dalpha += dL_dweight - last_dL_dT;
        last_dL_dT = dL_dweight * alpha + (1 - alpha) * last_dL_dT;
        const float dL_dmd = 2.0f * (T * alpha) * (m_d * final_A - final_D) * dL_ddist;
        dL_dr += dL_dmd * dmd_dd;

#pragma unroll
        for (int ch = 0; ch < GSL_MAX_FEATURES; ++ch)
          if (ch < S) g_feat[ch] = wgt * dfeat[ch];  // features do not feed dL_dalpha (backward.cu:395)
        const float nrm[3] = {s.nx, s.ny, s.nz};
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          accum_n[ch] = last_alpha * last_n[ch] + (1.f - last_alpha) * accum_n[ch];
          last_n[ch] = nrm[ch];
          dL_dalpha += (nrm[ch] - accum_n[ch]) * dnorm[ch];
          g_nrm[ch] = wgt * dnorm[ch];
        }
        accum_depth = last_alpha * last_depth + (1.f - last_alpha) * accum_depth;
        last_depth = depth;
        dL_dalpha += (depth - accum_depth) * dL_depth;
        accum_mask = last_alpha + (1.f - last_alpha) * accum_mask;
        dL_dalpha = (float)((double)dL_dalpha + (1.0 - (double)accum_mask) * (double)dL_mask);
        dL_dalpha *= T;
        last_alpha = alpha;
        dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;

        const float dL_dG = s.opacity * dL_dalpha;
        if (e.rho3d <= e.rho2d) {
          const float ex = ray.sth * ray.sphi, ey = -ray.cth, ez = ray.sth * ray.cphi;
          const float dsx = dL_dG * -G * e.sx + dL_dr * (s.Tux * ex + s.Tvx * ey + s.Twx * ez);
          const float dsy = dL_dG * -G * e.sy + dL_dr * (s.Tuy * ex + s.Tvy * ey + s.Twy * ez);
          const float qx = dsx / e.pz, qy = dsy / e.pz;
          const float dpx = qx, dpy = qy, dpz = -(qx * e.sx + qy * e.sy);
          // dL_dk = l x dL_dp ; dL_dl = dL_dp x k
          const float dkx = e.ly * dpz - e.lz * dpy, dky = e.lz * dpx - e.lx * dpz, dkz = e.lx * dpy - e.ly * dpx;
          const float dlx = dpy * e.kz - dpz * e.ky, dly = dpz * e.kx - dpx * e.kz, dlz = dpx * e.ky - dpy * e.kx;
          const float rx = dL_dr * ex, ry = dL_dr * ey, rz = dL_dr * ez;
          g_dT[0] = ray.cphi * dkx + ray.sphi_cth * dlx + rx * e.sx;
          g_dT[1] = ray.cphi * dky + ray.sphi_cth * dly + rx * e.sy;
          g_dT[2] = ray.cphi * dkz + ray.sphi_cth * dlz + rx;
          g_dT[3] = ray.sth * dlx + ry * e.sx;
          g_dT[4] = ray.sth * dly + ry * e.sy;
          g_dT[5] = ray.sth * dlz + ry;
          g_dT[6] = -ray.sphi * dkx + ray.cphi_cth * dlx + rz * e.sx;
          g_dT[7] = -ray.sphi * dky + ray.cphi_cth * dly + rz * e.sy;
          g_dT[8] = -ray.sphi * dkz + ray.cphi_cth * dlz + rz;
        } else {
          g_m2x = dL_dG * (-G * 2.f * e.dx);
          g_m2y = dL_dG * (-G * 2.f * e.dy);
          g_dT[2] = dL_dr * s.Tuz / depth;
          g_dT[5] = dL_dr * s.Tvz / depth;
          g_dT[8] = dL_dr * s.Twz / depth;
        }
        g_op = G * dL_dalpha;
      }

      // ---- warp reduction + vector reductions into the packed accumulator
      float* gdst = grad + (size_t)s_id[j] * gstride;
      if (__popc(vmask) > 1) {
#pragma unroll
        for (int i = 0; i < 9; ++i) g_dT[i] = warp_sum(g_dT[i]);
        g_m2x = warp_sum(g_m2x); g_m2y = warp_sum(g_m2y); g_op = warp_sum(g_op);
#pragma unroll
        for (int i = 0; i < 4; ++i) g_col[i] = warp_sum(g_col[i]);
#pragma unroll
        for (int i = 0; i < 3; ++i) g_nrm[i] = warp_sum(g_nrm[i]);
#pragma unroll
        for (int i = 0; i < GSL_MAX_FEATURES; ++i)
          if (i < S) g_feat[i] = warp_sum(g_feat[i]);
        if (lane == 0) red_add_v4(gdst + 0, g_dT[0], g_dT[1], g_dT[2], g_dT[3]);
        if (lane == 1) red_add_v4(gdst + 4, g_dT[4], g_dT[5], g_dT[6], g_dT[7]);
        if (lane == 2) red_add_v4(gdst + 8, g_dT[8], g_m2x, g_m2y, g_op);
        if (lane == 3) red_add_v4(gdst + 12, g_col[0], g_col[1], g_col[2], g_col[3]);
        if (lane == 4) red_add_v4(gdst + 16, g_nrm[0], g_nrm[1], g_nrm[2], 0.f);
        if (S > 0 && lane == 5) red_add_v4(gdst + 20, g_feat[0], g_feat[1], g_feat[2], g_feat[3]);
        if (S > 4 && lane == 6) red_add_v4(gdst + 24, g_feat[4], g_feat[5], g_feat[6], g_feat[7]);
        if (S > 8 && lane == 7) red_add_v4(gdst + 28, g_feat[8], g_feat[9], 0.f, 0.f);
      } else if (valid) {
        red_add_v4(gdst + 0, g_dT[0], g_dT[1], g_dT[2], g_dT[3]);
        red_add_v4(gdst + 4, g_dT[4], g_dT[5], g_dT[6], g_dT[7]);
        red_add_v4(gdst + 8, g_dT[8], g_m2x, g_m2y, g_op);
        red_add_v4(gdst + 12, g_col[0], g_col[1], g_col[2], g_col[3]);
        red_add_v4(gdst + 16, g_nrm[0], g_nrm[1], g_nrm[2], 0.f);
        if (S > 0) red_add_v4(gdst + 20, g_feat[0], g_feat[1], g_feat[2], g_feat[3]);
        if (S > 4) red_add_v4(gdst + 24, g_feat[4], g_feat[5], g_feat[6], g_feat[7]);
        if (S > 8) red_add_v4(gdst + 28, g_feat[8], g_feat[9], 0.f, 0.f);
      }
    }
  }
#ifdef GSL_STATS
  STATB_ADD(0, st_scan); STATB_ADD(1, st_box); STATB_ADD(2, st_any); STATB_ADD(3, st_valid); STATB_ADD(4, st_multi);
#endif
}

int launch_render_backward(const gsl_params& p, const gsl_fwd_inputs& in, const gsl_fwd_outputs& fwd,
                           const gsl_bwd_inputs& gin, const GeomView& g, const ImageView& im, const BinView& b,
                           int64_t r_capacity, cudaStream_t st) {
  RenderParams rp = make_render_params(p, r_capacity);
  const int tiles = rp.gx * rp.gy;
  if (tiles == 0 || p.P == 0) return 0;
  const float4* colors = in.colors_precomp ? reinterpret_cast<const float4*>(in.colors_precomp) : g.rgb;
  const int gs = grad_stride(p.S);
  ProfScope prof(GSL_K_RENDER_BWD, st);
#define GSL_LAUNCH_BWD(ST)                                                                                \
  k_render_bwd<ST><<<tiles, 256, 0, st>>>(rp, im.ranges, b.vals_b, g.rec, g.pixbox, colors, in.features,   \
                                          in.background, g.ctrl, im.final_T, fwd.out_contrib,              \
                                          gin.dL_dout_color, gin.dL_dout_depth, gin.dL_dout_alpha,          \
                                          gin.dL_dout_feature, g.grad, gs)
  if (p.S == 4) GSL_LAUNCH_BWD(4);
  else if (p.S == 0) GSL_LAUNCH_BWD(0);
  else GSL_LAUNCH_BWD(-1);
#undef GSL_LAUNCH_BWD
  return check_cuda(cudaGetLastError(), "k_render_bwd launch");
}

}  // namespace gsl
