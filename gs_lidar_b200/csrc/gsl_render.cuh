// gsl_render.cuh -- pieces shared by the forward and backward compositors.
//
// Work decomposition of both compositing kernels (B200: 148 SMs, a 66x1030 panorama is only 325 tiles
// with thousands of surfels per tile list, so "one CTA per tile" leaves the machine idle and latency-bound):
//   * one WARP per 8x4 pixel block, one thread per pixel, one warp per CTA -> ~2200 independent warps over
//     the SMs, no __syncthreads anywhere, a warp retires the moment its 32 pixels are finished;
//   * the tile|depth sorted list is exactly the reference's (16x16 tiles, bit-exact keys and positions); a
//     second-level binning (gsl_binning.cu) splits every tile list into the lists of its eight 8x4 blocks
//     using the surfels' conservative pixel boxes, so a warp only sees entries that can reach its pixels;
//   * a warp stages 32 block-list entries at a time in shared memory (cp.async gathers of the 64-B records,
//     double buffered: the gathers of chunk c+1 and the entry loads of chunk c+2 fly while chunk c is
//     composited), then every LANE walks -- in list order, at its own pace -- only the staged entries
//     whose pixel box contains its own pixel.  The lanes of a warp therefore work on different surfels in
//     the same instruction; a chunk costs max-over-lanes(entries relevant to that pixel) evaluations
//     instead of 32, and the per-pixel recursion never waits for pixels the surfel cannot reach;
//   * the forward stores, per block-list entry, the 32-bit mask of pixels it contributed to (`pairmask`);
//     the backward walks the block list back to front and evaluates exactly those (pixel, surfel) pairs.
#pragma once
#include "gsl_common.cuh"
#include "gsl_math.cuh"

namespace gsl {

// geometry of the 8x4 pixel block a warp owns
struct BlockGeom {
  int bx0, by0;   // first pixel
  int tile;       // 16x16 tile id (the reference's tile numbering)
  int bbit;       // which of the tile's 8 blocks: (row/4)*2 + col/8
  int pxi, pyi;   // this lane's pixel
  bool inside;
  int pix_id;
};

__device__ __forceinline__ BlockGeom block_geom(const RenderParams& rp, int block_id, int lane) {
  BlockGeom g;
  const int nbx = (rp.W + 7) >> 3;
  const int bxi = block_id % nbx, byi = block_id / nbx;
  g.bx0 = bxi * 8;
  g.by0 = byi * 4;
  g.tile = (g.by0 >> 4) * rp.gx + (g.bx0 >> 4);
  g.bbit = (((g.by0 & 15) >> 2) << 1) | ((g.bx0 & 15) >> 3);
  g.pxi = g.bx0 + (lane & 7);
  g.pyi = g.by0 + (lane >> 3);
  g.inside = g.pxi < rp.W && g.pyi < rp.H;
  g.pix_id = rp.W * g.pyi + g.pxi;
  return g;
}

// ---- per-warp candidate stage (double-buffered, filled with cp.async, component-major) --------------
struct ChunkStage {
  float4 v[6][32];  // 64-B record (4), colour, features (S == 4)
  uint2 ent[32];    // (surfel id, list position)
  short4 box[32];   // conservative pixel box (forward only)
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// 32x32 bit-matrix transpose across the warp: bit i of the result in lane j = bit j of `x` in lane i.
// Five butterfly steps (swap the off-diagonal k x k blocks with lane ^ k), 5 shuffles instead of 32 ballots.
__device__ __forceinline__ uint32_t transpose32(uint32_t x, int lane) {
#pragma unroll
  for (int k = 16; k >= 1; k >>= 1) {
    const uint32_t m = (k == 16) ? 0x0000ffffu : (k == 8) ? 0x00ff00ffu : (k == 4) ? 0x0f0f0f0fu
                     : (k == 2) ? 0x33333333u : 0x55555555u;
    const uint32_t t = __shfl_xor_sync(0xffffffffu, x, k);
    x = (lane & k) ? ((x & ~m) | ((t >> k) & m)) : ((x & m) | ((t << k) & ~m));
  }
  return x;
}

// Which of the 32 pixels of the 8x4 block at (bx0, by0) lie inside the (possibly azimuth-wrapped) pixel box;
// bit = row * 8 + column, i.e. the lane that owns the pixel.
__device__ __forceinline__ uint32_t box_pixel_mask(const short4 bb, int bx0, int by0) {
  const int r_lo = max((int)bb.y - by0, 0), r_hi = min((int)bb.w - by0, 3);
  if (r_lo > r_hi) return 0u;
  const int c0 = (int)bb.x - bx0, c1 = (int)bb.z - bx0;
  uint32_t colbits;
  if (bb.x <= bb.z) {
    const int lo = max(c0, 0), hi = min(c1, 7);
    colbits = (lo <= hi) ? ((0xffu >> (7 - hi)) & (0xffu << lo)) : 0u;
  } else {  // wrapped: x >= bb.x || x <= bb.z
    const uint32_t up = (c0 <= 7) ? (0xffu << max(c0, 0)) : 0u;
    const uint32_t dn = (c1 >= 0) ? (0xffu >> (7 - min(c1, 7))) : 0u;
    colbits = up | dn;
  }
  colbits &= 0xffu;
  uint32_t m = 0;
#pragma unroll
  for (int r = 0; r < 4; ++r)
    if (r >= r_lo && r <= r_hi) m |= colbits << (8 * r);
  return m;
}

__device__ __forceinline__ Splat staged_splat(const ChunkStage& sb, int s) {
  const float4 a = sb.v[0][s], b = sb.v[1][s], c = sb.v[2][s], d = sb.v[3][s];
  Splat sp;
  sp.Tux = a.x; sp.Tuy = a.y; sp.Tuz = a.z; sp.Tvx = a.w;
  sp.Tvy = b.x; sp.Tvz = b.y; sp.Twx = b.z; sp.Twy = b.w;
  sp.Twz = c.x; sp.mx = c.y; sp.my = c.z; sp.opacity = c.w;
  sp.nx = d.x; sp.ny = d.y; sp.nz = d.z; sp.depth = d.w;
  return sp;
}

}  // namespace gsl
