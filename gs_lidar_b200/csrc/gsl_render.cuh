// gsl_render.cuh -- pieces shared by the forward and backward compositors.
//
// Work decomposition of both compositing kernels (B200: 148 SMs, a 66x1030 panorama is only 325 tiles
// with thousands of surfels per tile list, so "one CTA per tile" leaves the machine idle and latency-bound):
//   * one WARP per 8x4 pixel block, one thread per pixel, one warp per CTA -> ~2200 independent warps over
//     the SMs, no __syncthreads anywhere, a warp retires the moment its 32 pixels are finished;
//   * the tile|depth sorted list is exactly the reference's (16x16 tiles, bit-exact keys), but every list
//     position carries an 8-bit block mask (gsl_binning.cu) built from the surfel's conservative pixel
//     box, so a warp only evaluates the entries that can reach one of its pixels;
//   * the forward records per block which list positions contributed to at least one of its pixels (`used`
//     bit-planes); the backward walks just those, back to front.
//   * per-warp software pipeline over 32-position words of the list:
//        [mask/used bits + surfel ids of word w+2] -> [record gathers of word w+1] -> [composite word w]
//     candidates are compacted with ballot/popc into a double-buffered shared-memory stage and read back
//     with broadcast LDS.128.
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

// One staged candidate: the 64-B record + colour + (S == 4) features, gathered by one lane.
struct CandRegs {
  float4 r0, r1, r2, r3, col, feat;
};

// Double-buffered per-warp stage (component-major so that the compaction store is conflict-free and the
// broadcast read of one candidate is six LDS.128 of the same address in every lane).
struct WarpStage {
  float4 v[6][32];
  uint32_t id[32];
  uint32_t lanepos[32];
};

template <bool WITH_FEAT4>
__device__ __forceinline__ void gather_cand(bool c, uint32_t id, const float4* __restrict__ rec,
                                            const float4* __restrict__ colors, const float* __restrict__ features,
                                            CandRegs& o) {
  if (c) {
    const float4* r4 = rec + 4 * (size_t)id;
    o.r0 = __ldg(r4);
    o.r1 = __ldg(r4 + 1);
    o.r2 = __ldg(r4 + 2);
    o.r3 = __ldg(r4 + 3);
    o.col = __ldg(colors + id);
    if (WITH_FEAT4) o.feat = __ldg(reinterpret_cast<const float4*>(features) + id);
  }
}

template <bool WITH_FEAT4>
__device__ __forceinline__ void stage_cand(WarpStage& sb, int slot, const CandRegs& r, uint32_t id, uint32_t lanepos) {
  sb.v[0][slot] = r.r0;
  sb.v[1][slot] = r.r1;
  sb.v[2][slot] = r.r2;
  sb.v[3][slot] = r.r3;
  sb.v[4][slot] = r.col;
  if (WITH_FEAT4) sb.v[5][slot] = r.feat;
  sb.id[slot] = id;
  sb.lanepos[slot] = lanepos;
}

__device__ __forceinline__ Splat staged_splat(const WarpStage& sb, int s) {
  const float4 a = sb.v[0][s], b = sb.v[1][s], c = sb.v[2][s], d = sb.v[3][s];
  Splat sp;
  sp.Tux = a.x; sp.Tuy = a.y; sp.Tuz = a.z; sp.Tvx = a.w;
  sp.Tvy = b.x; sp.Tvz = b.y; sp.Twx = b.z; sp.Twy = b.w;
  sp.Twz = c.x; sp.mx = c.y; sp.my = c.z; sp.opacity = c.w;
  sp.nx = d.x; sp.ny = d.y; sp.nz = d.z; sp.depth = d.w;
  return sp;
}

}  // namespace gsl
