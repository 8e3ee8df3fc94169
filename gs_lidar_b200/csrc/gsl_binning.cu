// gsl_binning.cu -- tile binning.  What must come out, bit for bit, is what rasterizer_impl.cu:68-142 and :310-354
// (K2-K7 in SURVEY.md) produce:
//   offsets = inclusive_scan(tiles_touched); R = offsets[P-1]
//   key = (tile_id << 32) | float_bits(depth), value = surfel id, emitted y-major/x-minor per surfel
//   stable sort over key bits [0, 32 + bits(tiles));  ranges[tile] = [first, last+1)
// Two ways to get there:
//   * fast path (images of up to GSL_FAST_BIN_MAX_TILES tiles, every BASELINE.json config): the surfels are depth-
//     sorted once (gsl_sort.cu) and ONE stable counting pass distributes their instances to the tiles
//     (k_bin_count / k_bin_scan / k_bin_bases / k_bin_scatter below) -- no 64-bit keys, no library;
//   * general path: the reference's own scheme -- scan (k_scan_*), key duplication (k_duplicate) and the library
//     64-bit radix sort (the same cub call the reference makes).
// Both end in k_tile_blists, which also builds the per-8x4-block lists of this design.
#include <cub/cub.cuh>
#include "gsl_common.cuh"

namespace gsl {

// ------------------------------------------------------------------------------------------------
// scan: 1024 elements per block, three small kernels (reduce, scan of block sums, downsweep)
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t n = __shfl_up_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) >= o) v += n;
  }
  return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, total in *total
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* smem, uint32_t* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t inc = warp_incl_scan(v);
  if (lane == 31) smem[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    uint32_t w = (lane < (blockDim.x >> 5)) ? smem[lane] : 0;
    uint32_t wi = warp_incl_scan(w);
    smem[lane] = wi - w;
    if (lane == 31) smem[32] = wi;
  }
  __syncthreads();
  uint32_t res = inc - v + smem[wid];
  *total = smem[32];
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const uint32_t* __restrict__ in, int P,
                                                              uint32_t* __restrict__ block_sums) {
  __shared__ uint32_t sm[33];
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  uint32_t s = 0;
  if (base + SCAN_ITEMS <= P) {
    uint4 v = *reinterpret_cast<const uint4*>(in + base);
    s = v.x + v.y + v.z + v.w;
  } else {
    for (int i = 0; i < SCAN_ITEMS; ++i)
      if (base + i < P) s += in[base + i];
  }
  uint32_t total;
  block_excl_scan(s, sm, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_sums(uint32_t* __restrict__ block_sums, int nblocks,
                                                    uint32_t* __restrict__ ctrl, int32_t* r_host_unused) {
  __shared__ uint32_t sm[33];
  uint32_t carry = 0;
  for (int base = 0; base < nblocks; base += 1024) {
    int i = base + threadIdx.x;
    uint32_t v = (i < nblocks) ? block_sums[i] : 0;
    uint32_t total;
    uint32_t ex = block_excl_scan(v, sm, &total);
    if (i < nblocks) block_sums[i] = ex + carry;
    carry += total;
  }
  if (threadIdx.x == 0) {
    ctrl[0] = carry;  // R
    ctrl[1] = 0;      // overflow flag
  }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_down(const uint32_t* __restrict__ in, int P,
                                                            const uint32_t* __restrict__ block_sums,
                                                            uint32_t* __restrict__ out) {
  __shared__ uint32_t sm[33];
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS] = {0, 0, 0, 0};
  if (base + SCAN_ITEMS <= P) {
    uint4 q = *reinterpret_cast<const uint4*>(in + base);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  } else {
    for (int i = 0; i < SCAN_ITEMS; ++i)
      if (base + i < P) v[i] = in[base + i];
  }
  uint32_t s = v[0] + v[1] + v[2] + v[3];
  uint32_t total;
  uint32_t ex = block_excl_scan(s, sm, &total) + block_sums[blockIdx.x];
  uint32_t o0 = ex + v[0], o1 = o0 + v[1], o2 = o1 + v[2], o3 = o2 + v[3];
  if (base + SCAN_ITEMS <= P) {
    *reinterpret_cast<uint4*>(out + base) = make_uint4(o0, o1, o2, o3);
  } else {
    uint32_t o[4] = {o0, o1, o2, o3};
    for (int i = 0; i < SCAN_ITEMS; ++i)
      if (base + i < P) out[base + i] = o[i];
  }
}

int launch_scan(const gsl_params& p, const GeomView& g, int32_t* r_host, cudaStream_t st) {
  if (p.P == 0) {
    cudaMemsetAsync(g.ctrl, 0, 8, st);
  } else {
    int nblocks = (p.P + SCAN_TILE - 1) / SCAN_TILE;
    ProfScope prof(GSL_K_SCAN, st);
    k_scan_reduce<<<nblocks, SCAN_THREADS, 0, st>>>(g.tiles, p.P, g.scan_state);
    k_scan_sums<<<1, 1024, 0, st>>>(g.scan_state, nblocks, g.ctrl, nullptr);
    k_scan_down<<<nblocks, SCAN_THREADS, 0, st>>>(g.tiles, p.P, g.scan_state, g.offs);
  }
  if (r_host) {
    r_host[0] = -1;  // sentinel for wait_num_rendered
    cudaMemcpyAsync(r_host, g.ctrl, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
  }
  return check_cuda(cudaGetLastError(), "scan launch");
}

// ------------------------------------------------------------------------------------------------
// key duplication (rasterizer_impl.cu:68-111).  One WARP per 32 surfels: lanes first take their own
// surfel, then surfels whose rect spans many tiles (azimuth-seam surfels span whole tile rows) are
// emitted cooperatively by the whole warp so no lane serialises a long loop.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_duplicate(int P, const float4* __restrict__ rec,
                                                   const ushort4* __restrict__ rect,
                                                   const uint32_t* __restrict__ tiles,
                                                   const uint32_t* __restrict__ offs, int gx,
                                                   const uint32_t* __restrict__ ctrl, uint32_t r_capacity,
                                                   uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  if (ctrl[0] > r_capacity) return;  // binning chunk too small: host re-runs after growing it
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  uint32_t n = 0, off = 0, depth_bits = 0;
  ushort4 rc = make_ushort4(0, 0, 0, 0);
  if (idx < P) {
    n = tiles[idx];
    if (n > 0) {
      off = offs[idx] - n;  // exclusive offset
      rc = rect[idx];
      depth_bits = __float_as_uint(rec[4 * (size_t)idx + 3].w);
    }
  }
  constexpr uint32_t COOP = 16;
  // small rects: each lane emits its own
  if (n > 0 && n <= COOP) {
    const uint32_t w = rc.z - rc.x;
    for (uint32_t k = 0; k < n; ++k) {
      uint32_t y = rc.y + k / w, x = rc.x + k % w;
      uint64_t key = ((uint64_t)(y * gx + x) << 32) | depth_bits;
      keys[off + k] = key;
      vals[off + k] = (uint32_t)idx;
    }
  }
  // large rects: whole warp
  uint32_t big = __ballot_sync(0xffffffffu, n > COOP);
  while (big) {
    int src = __ffs(big) - 1;
    big &= big - 1;
    uint32_t bn = __shfl_sync(0xffffffffu, n, src);
    uint32_t boff = __shfl_sync(0xffffffffu, off, src);
    uint32_t bdepth = __shfl_sync(0xffffffffu, depth_bits, src);
    uint32_t bx = __shfl_sync(0xffffffffu, (uint32_t)rc.x, src);
    uint32_t by = __shfl_sync(0xffffffffu, (uint32_t)rc.y, src);
    uint32_t bz = __shfl_sync(0xffffffffu, (uint32_t)rc.z, src);
    uint32_t bidx = (uint32_t)(idx - lane + src);
    uint32_t w = bz - bx;
    for (uint32_t k = lane; k < bn; k += 32) {
      uint32_t y = by + k / w, x = bx + k % w;
      keys[boff + k] = ((uint64_t)(y * gx + x) << 32) | bdepth;
      vals[boff + k] = bidx;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// identifyTileRanges (rasterizer_impl.cu:116-142) + second-level binning into 8x4-pixel BLOCK LISTS,
// one CTA per 16x16 tile:
//   * two warps find the tile's [first, last+1) in the sorted keys with a 32-ary search (5 probes rounds for
//     millions of instances) -> ranges[tile] (empty tiles get (0,0) like the reference's memset);
//   * the CTA then streams the tile's list; every position computes which of the tile's eight 8x4 blocks
//     the surfel's conservative pixel box overlaps, and an order-preserving compaction per block (ballot
//     ranks + running counters) appends (surfel id, list position) to blist[b][first + k].  The region of
//     block (tile, b) starts at the tile's own `first` inside plane b, so no global scan is needed.
// ------------------------------------------------------------------------------------------------
constexpr int TB_THREADS = 512;

__device__ __forceinline__ uint32_t block_mask_of(const short4 bb, uint32_t tile, int gx, int W, int H) {
  const int tx0 = (int)(tile % (uint32_t)gx) * GSL_BLOCK_X, ty0 = (int)(tile / (uint32_t)gx) * GSL_BLOCK_Y;
  uint32_t colm = 0, rowm = 0;
#pragma unroll
  for (int bc = 0; bc < 2; ++bc) {
    const int x0 = tx0 + bc * 8, x1 = min(x0 + 7, W - 1);
    const bool ov = (bb.x <= bb.z) ? ((int)bb.x <= x1 && (int)bb.z >= x0) : ((int)bb.x <= x1 || (int)bb.z >= x0);
    colm |= (ov && x0 < W) ? (1u << bc) : 0u;
  }
#pragma unroll
  for (int br = 0; br < 4; ++br) {
    const int y0 = ty0 + br * 4, y1 = min(y0 + 3, H - 1);
    const bool ov = (int)bb.y <= y1 && (int)bb.w >= y0;
    rowm |= (ov && y0 < H) ? (1u << br) : 0u;
  }
  uint32_t m = 0;
#pragma unroll
  for (int br = 0; br < 4; ++br)
    if (rowm & (1u << br)) m |= colm << (2 * br);
  return m;
}

// first index in [0, R) whose tile id (key >> 32) is >= t; executed by one full warp
__device__ __forceinline__ uint32_t warp_lower_bound_tile(const uint64_t* __restrict__ keys, uint32_t R, uint32_t t) {
  const int lane = threadIdx.x & 31;
  uint32_t lo = 0, hi = R;  // answer in [lo, hi]
  while (hi > lo) {
    const uint32_t span = hi - lo;
    if (span <= 32u) {
      const uint32_t p = lo + lane;
      const bool ge = (p < hi) ? ((uint32_t)(keys[p] >> 32) >= t) : true;
      const uint32_t bal = __ballot_sync(0xffffffffu, ge);
      return lo + (uint32_t)(__ffs(bal) - 1);
    }
    // 32 interior probes split [lo, hi) into 33 pieces
    const uint32_t p = lo + (uint32_t)(((uint64_t)span * (uint32_t)(lane + 1)) / 33u);
    const bool ge = (uint32_t)(keys[p] >> 32) >= t;
    const uint32_t bal = __ballot_sync(0xffffffffu, ge);
    const int f = bal ? (__ffs(bal) - 1) : 32;  // first probe that is >= t
    const uint32_t p_f = __shfl_sync(0xffffffffu, p, f & 31);
    const uint32_t p_prev = __shfl_sync(0xffffffffu, p, (f - 1) & 31);
    const uint32_t new_hi = (f < 32) ? p_f : hi;
    const uint32_t new_lo = (f > 0) ? p_prev + 1 : lo;
    lo = new_lo;
    hi = new_hi;
  }
  return lo;
}

__global__ void __launch_bounds__(TB_THREADS) k_tile_blists(
    const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, const short4* __restrict__ pixbox,
    const uint32_t* __restrict__ ctrl, uint32_t r_capacity, int gx, int W, int H, size_t plane_stride,
    uint2* __restrict__ ranges, uint2* __restrict__ blist, uint4* __restrict__ bdesc, bool have_ranges) {
  const uint32_t R = ctrl[0];
  const uint32_t tile = blockIdx.x;
  constexpr int NW = TB_THREADS / 32;
  constexpr int ITEMS = 4;             // list positions per thread and round (round k: base + k*TB_THREADS + tid)
  constexpr int SEGS = ITEMS * NW;     // warp-sized segments of one round, in list order
  __shared__ uint32_t s_bounds[2];
  __shared__ uint32_t s_seg[SEGS][8];  // per segment and plane: count, then exclusive prefix (incl. running total)
  __shared__ uint32_t s_run[8];
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  if (R == 0 || R > r_capacity) {
    if (threadIdx.x == 0) ranges[tile] = make_uint2(0, 0);
    if (threadIdx.x < 8) bdesc[tile * 8 + threadIdx.x] = make_uint4(0, 0, 0, 0);
    return;
  }
  if (have_ranges) {  // written by k_bin_bases
    if (threadIdx.x == 0) {
      const uint2 r = ranges[tile];
      s_bounds[0] = r.x;
      s_bounds[1] = r.y;
    }
  } else if (wv < 2) {
    const uint32_t v = warp_lower_bound_tile(keys, R, tile + wv);
    if (lane == 0) s_bounds[wv] = v;
  }
  if (threadIdx.x < 8) s_run[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t lo = s_bounds[0], hi = s_bounds[1];
  if (!have_ranges && threadIdx.x == 0) ranges[tile] = (hi > lo) ? make_uint2(lo, hi) : make_uint2(0, 0);
  const uint32_t lt = (1u << lane) - 1u;
  for (uint32_t base = lo; base < hi; base += ITEMS * TB_THREADS) {
    uint32_t id[ITEMS], m[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const uint32_t i = base + k * TB_THREADS + threadIdx.x;
      id[k] = (i < hi) ? vals[i] : 0u;
    }
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const uint32_t i = base + k * TB_THREADS + threadIdx.x;
      m[k] = (i < hi) ? block_mask_of(pixbox[id[k]], tile, gx, W, H) : 0u;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const uint32_t bal = __ballot_sync(0xffffffffu, (m[k] >> b) & 1u);
        if (lane == b) s_seg[k * NW + wv][b] = __popc(bal);
      }
    }
    __syncthreads();
    if (threadIdx.x < 8) {  // exclusive prefix over the segments of this plane, seeded with the running total
      uint32_t run = s_run[threadIdx.x];
#pragma unroll 8
      for (int sg = 0; sg < SEGS; ++sg) {
        const uint32_t c = s_seg[sg][threadIdx.x];
        s_seg[sg][threadIdx.x] = run;
        run += c;
      }
      s_run[threadIdx.x] = run;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const uint32_t i = base + k * TB_THREADS + threadIdx.x;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const uint32_t bal = __ballot_sync(0xffffffffu, (m[k] >> b) & 1u);
        if ((m[k] >> b) & 1u)
          blist[(size_t)b * plane_stride + lo + s_seg[k * NW + wv][b] + __popc(bal & lt)] = make_uint2(id[k], i);
      }
    }
    __syncthreads();  // s_seg is rewritten by the next round
  }
  __syncthreads();
  if (threadIdx.x < 8) bdesc[tile * 8 + threadIdx.x] = make_uint4(lo, lo + s_run[threadIdx.x], 0, 0);
}

// ------------------------------------------------------------------------------------------------
// Fast binning (images of up to GSL_FAST_BIN_MAX_TILES tiles).  The reference sorts R (tile | depth) keys of 64
// bits; the same permutation is obtained by
//   1. sorting the P SURFELS once by (depth bits, id)  (library radix sort over 32-bit keys; P is known on the
//      host, so nothing waits for the instance count), and
//   2. one stable counting pass that distributes the instances, emitted in that surfel order, to their tiles:
//      within a tile the instances then appear in (depth, id) order, which is exactly the order of the
//      reference's stable sort (a surfel emits a tile at most once).
// Step 2 works on chunks of 256 depth-ranked surfels (one CTA each):
//   k_bin_count    per chunk, a shared-memory bitmap [tile][256 surfels]; hist[tile][chunk] = popcount
//   k_bin_scan     per tile, exclusive scan of hist[tile][*] over the chunks + total[tile]
//   k_bin_bases    exclusive scan of total[] -> ranges[tile], R, overflow flag
//   k_bin_scatter  rebuilds the bitmap; instance of surfel thread t in tile b goes to
//                  ranges[b].x + hist[b][chunk] + popcount(bitmap[b] below t)
// ------------------------------------------------------------------------------------------------
// Visits every tile of every surfel of the chunk: surfels covering up to BIG_RECT tiles are walked by their own
// thread, larger ones (azimuth-seam surfels cover whole tile rows) by all 32 lanes of their warp together, so no
// thread serialises a long loop.  f(tile, t_src, payload) is called with t_src = chunk-local index of the surfel
// and the payload of the thread that owns it.
constexpr int BIG_RECT = 12;
template <typename F>
__device__ __forceinline__ void for_each_chunk_tile(ushort4 rc, int gx, uint32_t payload, F f) {
  const int lane = threadIdx.x & 31;
  const int w = (int)rc.z - (int)rc.x, n = w * ((int)rc.w - (int)rc.y);
  // rc.z may exceed gx in azimuth wrap-around mode: the column is x mod gx (gsl_preprocess.cu)
  if (n > 0 && n <= BIG_RECT) {
    for (int y = rc.y; y < rc.w; ++y)
      for (int x = rc.x; x < rc.z; ++x) f(y * gx + (x >= gx ? x - gx : x), (int)threadIdx.x, payload);
  }
  uint32_t big = __ballot_sync(0xffffffffu, n > BIG_RECT);
  while (big) {
    const int src = __ffs(big) - 1;
    big &= big - 1;
    const int bx = __shfl_sync(0xffffffffu, (int)rc.x, src), by = __shfl_sync(0xffffffffu, (int)rc.y, src);
    const int bw = __shfl_sync(0xffffffffu, w, src), bn = __shfl_sync(0xffffffffu, n, src);
    const uint32_t pl = __shfl_sync(0xffffffffu, payload, src);
    const int t_src = (int)threadIdx.x - lane + src;
    for (int k = lane; k < bn; k += 32) {
      const int x = bx + k % bw;
      f((by + k / bw) * gx + (x >= gx ? x - gx : x), t_src, pl);
    }
  }
}

__device__ __forceinline__ void chunk_bitmap(uint32_t* s_bits, int tiles, int gx, ushort4 rc) {
  for (int k = threadIdx.x; k < tiles * 8; k += 256) s_bits[k] = 0u;
  __syncthreads();
  for_each_chunk_tile(rc, gx, 0u,
                      [&](int tile, int t, uint32_t) { atomicOr(&s_bits[tile * 8 + (t >> 5)], 1u << (t & 31)); });
  __syncthreads();
}

__global__ void __launch_bounds__(256) k_bin_count(int P, const uint32_t* __restrict__ order,
                                                   const ushort4* __restrict__ rect, int tiles, int gx, size_t ncta,
                                                   uint32_t* __restrict__ hist) {
  extern __shared__ uint32_t s_bits[];
  const int j = blockIdx.x * 256 + threadIdx.x;
  ushort4 rc = make_ushort4(0, 0, 0, 0);
  if (j < P) rc = rect[order[j]];
  chunk_bitmap(s_bits, tiles, gx, rc);
  for (int b = threadIdx.x; b < tiles; b += 256) {
    uint32_t c = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) c += __popc(s_bits[b * 8 + w]);
    hist[(size_t)b * ncta + blockIdx.x] = c;
  }
}

__global__ void __launch_bounds__(256) k_bin_scan(uint32_t* __restrict__ hist, size_t ncta,
                                                  uint32_t* __restrict__ bintotal) {
  __shared__ uint32_t sm[33];
  uint32_t* row = hist + (size_t)blockIdx.x * ncta;
  uint32_t carry = 0;
  for (size_t base = 0; base < ncta; base += 256 * 4) {
    const size_t i0 = base + (size_t)threadIdx.x * 4;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (i0 + k < ncta) ? row[i0 + k] : 0u;
    const uint32_t s = v[0] + v[1] + v[2] + v[3];
    uint32_t total;
    uint32_t ex = block_excl_scan(s, sm, &total) + carry;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i0 + k < ncta) row[i0 + k] = ex;
      ex += v[k];
    }
    carry += total;
  }
  if (threadIdx.x == 0) bintotal[blockIdx.x] = carry;
}

__global__ void __launch_bounds__(1024) k_bin_bases(const uint32_t* __restrict__ bintotal, int tiles,
                                                    uint2* __restrict__ ranges, uint32_t* __restrict__ ctrl,
                                                    uint32_t r_capacity) {
  __shared__ uint32_t sm[33];
  uint32_t carry = 0;
  for (int base = 0; base < tiles; base += 1024) {
    const int i = base + threadIdx.x;
    const uint32_t v = (i < tiles) ? bintotal[i] : 0u;
    uint32_t total;
    const uint32_t ex = block_excl_scan(v, sm, &total) + carry;
    if (i < tiles) ranges[i] = v ? make_uint2(ex, ex + v) : make_uint2(0, 0);  // empty tiles: (0,0) like the memset
    carry += total;
  }
  if (threadIdx.x == 0) {
    ctrl[0] = carry;                          // R
    ctrl[1] = carry > r_capacity ? 1u : 0u;   // overflow: the instance buffers are too small, nothing is written
  }
}

__global__ void __launch_bounds__(256) k_bin_scatter(int P, const uint32_t* __restrict__ order,
                                                     const ushort4* __restrict__ rect, int tiles, int gx, size_t ncta,
                                                     const uint32_t* __restrict__ hist, const uint2* __restrict__ ranges,
                                                     const uint32_t* __restrict__ ctrl, uint32_t r_capacity,
                                                     uint32_t* __restrict__ point_list) {
  extern __shared__ uint32_t s_bits[];  // [tiles][8] bitmap, then [tiles] first output slot of this chunk
  if (ctrl[0] > r_capacity) return;
  uint32_t* s_base = s_bits + tiles * 8;
  const int j = blockIdx.x * 256 + threadIdx.x;
  ushort4 rc = make_ushort4(0, 0, 0, 0);
  uint32_t id = 0;
  if (j < P) {
    id = order[j];
    rc = rect[id];
  }
  for (int b = threadIdx.x; b < tiles; b += 256) s_base[b] = ranges[b].x + hist[(size_t)b * ncta + blockIdx.x];
  chunk_bitmap(s_bits, tiles, gx, rc);
  for_each_chunk_tile(rc, gx, id, [&](int tile, int t, uint32_t sid) {
    const int w = t >> 5;
    uint32_t r = __popc(s_bits[tile * 8 + w] & ((1u << (t & 31)) - 1u));
    for (int k = 0; k < w; ++k) r += __popc(s_bits[tile * 8 + k]);
    point_list[s_base[tile] + r] = sid;
  });
}

// keys of the sorted list, reconstructed for state export: (tile << 32) | depth bits of the surfel
__global__ void __launch_bounds__(256) k_export_keys(const uint2* __restrict__ ranges,
                                                     const uint32_t* __restrict__ point_list,
                                                     const float4* __restrict__ rec, uint64_t* __restrict__ keys) {
  const uint2 r = ranges[blockIdx.x];
  for (uint32_t p = r.x + threadIdx.x; p < r.y; p += 256)
    keys[p] = ((uint64_t)blockIdx.x << 32) | (uint64_t)__float_as_uint(rec[4 * (size_t)point_list[p] + 3].w);
}

__global__ void k_flag_overflow(uint32_t* ctrl, uint32_t r_capacity) {
  if (ctrl[0] > r_capacity) ctrl[1] = 1;
}

static uint32_t higher_msb(uint32_t n) {  // rasterizer_impl.cu:32-47
  uint32_t msb = sizeof(n) * 4;
  uint32_t step = msb;
  while (step > 1) {
    step /= 2;
    if (n >> msb) msb += step; else msb -= step;
  }
  if (n >> msb) msb++;
  return msb;
}

size_t sort_temp_bytes(int64_t Rcap) {
  static thread_local int64_t cached_cap = -1;
  static thread_local size_t cached_bytes = 0;
  if (Rcap == cached_cap) return cached_bytes;
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                  (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)Rcap);
  cached_cap = Rcap;
  cached_bytes = bytes + 256;
  return cached_bytes;
}

int wait_num_rendered(int32_t* r_host, cudaStream_t st) {
  volatile int32_t* v = r_host;
  while (v[0] < 0) {
    cudaError_t q = cudaStreamQuery(st);
    if (q == cudaSuccess) break;              // everything ran: the copy has landed
    if (q != cudaErrorNotReady) return check_cuda(q, "waiting for the instance count");
  }
  if (v[0] < 0) return check_cuda(cudaStreamSynchronize(st), "waiting for the instance count");
  return 0;
}

int launch_export_keys(const gsl_params& p, const GeomView& g, const ImageView& im, const uint32_t* point_list,
                       uint64_t* keys_out, cudaStream_t st) {
  const int tiles = tile_count(p.W, p.H);
  k_export_keys<<<tiles, 256, 0, st>>>(im.ranges, point_list, g.rec, keys_out);
  return check_cuda(cudaGetLastError(), "k_export_keys launch");
}

// r_host[0] must already hold R (the caller synchronised on the scan) -- cub needs the count on
// the host.  TODO(round 2): device-count sort to drop this dependency.
int launch_binning(const gsl_params& p, const GeomView& g, const ImageView& im, const BinView& b,
                   int64_t r_capacity, int32_t* r_host, cudaStream_t st) {
  const int gx = (p.W + GSL_BLOCK_X - 1) / GSL_BLOCK_X, gy = (p.H + GSL_BLOCK_Y - 1) / GSL_BLOCK_Y;
  const int tiles = gx * gy;
  if (fast_binning(p.W, p.H)) {
    // ---- surfels are already depth-sorted (launch_surfel_sort): one stable counting pass per tile
    if (p.P == 0) {
      cudaMemsetAsync(g.ctrl, 0, 8, st);
      cudaMemsetAsync(im.ranges, 0, (size_t)tiles * sizeof(uint2), st);
      cudaMemsetAsync(im.bdesc, 0, (size_t)tiles * 8 * sizeof(uint4), st);
      if (r_host) {
        r_host[0] = -1;
        cudaMemcpyAsync(r_host, g.ctrl, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
      }
      return check_cuda(cudaGetLastError(), "binning (empty)");
    }
    const int ncta = (int)im.ncta;
    const size_t smem = (size_t)tiles * 9 * sizeof(uint32_t);
    {
      ProfScope prof(GSL_K_SCAN, st);
      k_bin_count<<<ncta, 256, smem, st>>>(p.P, g.sval_b, g.rect, tiles, gx, im.ncta, im.hist);
      k_bin_scan<<<tiles, 256, 0, st>>>(im.hist, im.ncta, im.bintotal);
      k_bin_bases<<<1, 1024, 0, st>>>(im.bintotal, tiles, im.ranges, g.ctrl, (uint32_t)r_capacity);
    }
    if (r_host) {
      r_host[0] = -1;  // sentinel for wait_num_rendered
      cudaMemcpyAsync(r_host, g.ctrl, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
    }
    {
      ProfScope prof(GSL_K_DUPLICATE, st);
      k_bin_scatter<<<ncta, 256, smem, st>>>(p.P, g.sval_b, g.rect, tiles, gx, im.ncta, im.hist, im.ranges, g.ctrl,
                                            (uint32_t)r_capacity, b.vals_b);
    }
    ProfScope prof(GSL_K_RANGES, st);
    k_tile_blists<<<tiles, TB_THREADS, 0, st>>>(nullptr, b.vals_b, g.pixbox, g.ctrl, (uint32_t)r_capacity, gx, p.W, p.H,
                                               b.plane_stride, im.ranges, b.blist, im.bdesc, true);
    return check_cuda(cudaGetLastError(), "binning launch");
  }
  // ---- general path: 64-bit (tile | depth) key sort like the reference; needs R on the host
  if (r_host) {
    int rc = wait_num_rendered(r_host, st);
    if (rc) return rc;
  }

  const int64_t R = r_host[0];
  if (p.P == 0 || R == 0) {
    cudaMemsetAsync(im.ranges, 0, (size_t)tiles * sizeof(uint2), st);
    cudaMemsetAsync(im.bdesc, 0, (size_t)tiles * 8 * sizeof(uint4), st);
    return check_cuda(cudaGetLastError(), "binning (empty)");
  }
  if (R > r_capacity) {
    k_flag_overflow<<<1, 1, 0, st>>>(g.ctrl, (uint32_t)r_capacity);
    return GSL_ENOSPACE;
  }
  {
    ProfScope prof(GSL_K_DUPLICATE, st);
    k_duplicate<<<(p.P + 255) / 256, 256, 0, st>>>(p.P, g.rec, g.rect, g.tiles, g.offs, gx, g.ctrl,
                                                   (uint32_t)r_capacity, b.keys_a, b.vals_a);
  }
  int bit = (int)higher_msb((uint32_t)tiles);
  size_t tmp = b.sort_tmp_bytes;
  {
    ProfScope prof(GSL_K_SORT, st);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(b.sort_tmp, tmp, b.keys_a, b.keys_b, b.vals_a, b.vals_b,
                                                    (int)R, 0, 32 + bit, st);
    if (e != cudaSuccess) return check_cuda(e, "cub::DeviceRadixSort::SortPairs");
  }
  ProfScope prof(GSL_K_RANGES, st);
  k_tile_blists<<<tiles, TB_THREADS, 0, st>>>(b.keys_b, b.vals_b, g.pixbox, g.ctrl, (uint32_t)r_capacity, gx, p.W, p.H,
                                             b.plane_stride, im.ranges, b.blist, im.bdesc, false);
  return check_cuda(cudaGetLastError(), "binning launch");
}

}  // namespace gsl
