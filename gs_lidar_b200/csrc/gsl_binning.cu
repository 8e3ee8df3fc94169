// gsl_binning.cu -- tile binning.  What must come out, bit for bit, is what rasterizer_impl.cu:68-142 and :310-354
// (K2-K7 in SURVEY.md) produce:
//   offsets = inclusive_scan(tiles_touched); R = offsets[P-1]
//   key = (tile_id << 32) | float_bits(depth), value = surfel id, emitted y-major/x-minor per surfel
//   stable sort over key bits [0, 32 + bits(tiles));  ranges[tile] = [first, last+1)
// How: the surfels are depth-sorted once (gsl_sort.cu) and a stable counting pass distributes their instances to the
// tiles (k_bin_count / k_bin_scan / k_bin_scatter below) -- no 64-bit keys are materialised and no library
// kernel runs.  The counting pass holds a [tile][256 surfels] bitmap in shared memory, so it handles GSL_BIN_GROUP_TILES
// tiles at a time: images with more tiles (every BASELINE.json config has at most 1,024) are processed as consecutive
// GROUPS of tile ids -- whole tile rows, or pieces of one row for images wider than 16,384 pixels -- each group
// continuing the list where the previous one ended.  k_tile_blists then builds the per-8x4-block lists of this design.
// The k_scan_* kernels (the reference's tiles_touched scan) remain for gsl_export_state only: nothing on the render
// path needs the offsets.
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
// identifyTileRanges (rasterizer_impl.cu:116-142) + second-level binning into 8x4-pixel BLOCK LISTS,
// one CTA per 16x16 tile:
//   * ranges[tile] = [first, last+1) comes from k_bin_scan (empty tiles hold (0,0) like the reference's memset);
//   * the CTA streams the tile's list; every position computes which of the tile's eight 8x4 blocks
//     the surfel's conservative pixel box overlaps, and an order-preserving compaction per block (ballot
//     ranks + running counters) appends (surfel id, list position) to blist[b][first + k].  The region of
//     block (tile, b) starts at the tile's own `first` inside plane b, so no global scan is needed.
// ------------------------------------------------------------------------------------------------
#ifndef GSL_TB_THREADS
#define GSL_TB_THREADS 512
#endif
constexpr int TB_THREADS = GSL_TB_THREADS;

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

__global__ void __launch_bounds__(TB_THREADS) k_tile_blists(
    const uint32_t* __restrict__ vals, const short4* __restrict__ pixbox,
    const uint32_t* __restrict__ ctrl, uint32_t r_capacity, int gx, int W, int H, size_t plane_stride,
    uint2* __restrict__ ranges, uint2* __restrict__ blist, uint4* __restrict__ bdesc) {
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
  if (threadIdx.x == 0) {  // written by k_bin_scan
    const uint2 r = ranges[tile];
    s_bounds[0] = r.x;
    s_bounds[1] = r.y;
  }
  if (threadIdx.x < 8) s_run[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t lo = s_bounds[0], hi = s_bounds[1];
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
    if (wv < 8) {  // warp p: exclusive prefix over the segments of plane p, seeded with the running total
      static_assert(SEGS % 32 == 0, "whole segments per lane");
      constexpr int PER = SEGS / 32;
      uint32_t c[PER], sum = 0;
#pragma unroll
      for (int k = 0; k < PER; ++k) { c[k] = s_seg[lane * PER + k][wv]; sum += c[k]; }
      uint32_t ex = warp_incl_scan(sum) - sum + s_run[wv];
      __syncwarp();
#pragma unroll
      for (int k = 0; k < PER; ++k) { s_seg[lane * PER + k][wv] = ex; ex += c[k]; }
      if (lane == 31) s_run[wv] = ex;
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
// Binning.  The reference sorts R (tile | depth) keys of 64 bits; the same permutation is obtained by
//   1. sorting the P SURFELS once by (depth bits, id)  (gsl_sort.cu; P is known on the host, so nothing waits for
//      the instance count), and
//   2. a stable counting pass that distributes the instances, emitted in that surfel order, to their tiles:
//      within a tile the instances then appear in (depth, id) order, which is exactly the order of the
//      reference's stable sort (a surfel emits a tile at most once).
// Step 2 works on chunks of 256 depth-ranked surfels (one CTA each) and on one GROUP of at most
// GSL_BIN_GROUP_TILES consecutive tile ids at a time (TileGroup; one group for every image of up to 1,024 tiles):
//   k_bin_count    per chunk, a shared-memory bitmap [tile of the group][256 surfels]; hist[tile][chunk] = popcount
//   k_bin_scan     per tile, exclusive scan of hist[tile][*] over the chunks + total[tile]; its last CTA: exclusive scan
//                  of total[], continued from the previous group -> ranges[tile], R, overflow flag
//   k_bin_scatter  rebuilds the bitmap; instance of surfel thread t in tile b goes to
//                  ranges[b].x + hist[b][chunk] + popcount(bitmap[b] below t)
// ------------------------------------------------------------------------------------------------
// Consecutive tile ids [t0, t0 + nt): tile rows [y0, y1) x columns [x0, x1) of the tile grid.  Whole rows when a row
// fits a group (x0 = 0, x1 >= the grid width; in azimuth wrap-around mode rect columns run past the grid width and
// are taken modulo it), otherwise a piece of ONE row.
struct TileGroup {
  int t0, nt, y0, y1, x0, x1;
};

// Visits every tile of the group of every surfel of the chunk: surfels covering up to BIG_RECT tiles are walked by
// their own thread, larger ones (azimuth-seam surfels cover whole tile rows) by all 32 lanes of their warp together,
// so no thread serialises a long loop.  f(group-local tile, t_src, payload) is called with t_src = chunk-local index
// of the surfel and the payload of the thread that owns it.
constexpr int BIG_RECT = 12;
template <typename F>
__device__ __forceinline__ void for_each_chunk_tile(ushort4 rc, int gx, const TileGroup tg, uint32_t payload, F f) {
  const int lane = threadIdx.x & 31;
  // the part of the rect inside the group
  const int rx = max((int)rc.x, tg.x0), ry = max((int)rc.y, tg.y0);
  const int w = min((int)rc.z, tg.x1) - rx, h = min((int)rc.w, tg.y1) - ry;
  const int n = (w > 0 && h > 0) ? w * h : 0;
  // rc.z may exceed gx in azimuth wrap-around mode: the column is x mod gx (gsl_preprocess.cu)
  if (n > 0 && n <= BIG_RECT) {
    for (int y = ry; y < ry + h; ++y)
      for (int x = rx; x < rx + w; ++x) f(y * gx + (x >= gx ? x - gx : x) - tg.t0, (int)threadIdx.x, payload);
  }
  uint32_t big = __ballot_sync(0xffffffffu, n > BIG_RECT);
  while (big) {
    const int src = __ffs(big) - 1;
    big &= big - 1;
    const int bx = __shfl_sync(0xffffffffu, rx, src), by = __shfl_sync(0xffffffffu, ry, src);
    const int bw = __shfl_sync(0xffffffffu, w, src), bn = __shfl_sync(0xffffffffu, n, src);
    const uint32_t pl = __shfl_sync(0xffffffffu, payload, src);
    const int t_src = (int)threadIdx.x - lane + src;
    for (int k = lane; k < bn; k += 32) {
      const int x = bx + k % bw;
      f((by + k / bw) * gx + (x >= gx ? x - gx : x) - tg.t0, t_src, pl);
    }
  }
}

__device__ __forceinline__ void chunk_bitmap(uint32_t* s_bits, int gx, const TileGroup tg, ushort4 rc) {
  for (int k = threadIdx.x; k < tg.nt * 8; k += 256) s_bits[k] = 0u;
  __syncthreads();
  for_each_chunk_tile(rc, gx, tg, 0u,
                      [&](int tile, int t, uint32_t) { atomicOr(&s_bits[tile * 8 + (t >> 5)], 1u << (t & 31)); });
  __syncthreads();
}

__global__ void __launch_bounds__(256) k_bin_count(int P, const uint32_t* __restrict__ order,
                                                   const ushort4* __restrict__ rect, const TileGroup tg, int gx,
                                                   size_t ncta, uint32_t* __restrict__ hist) {
  extern __shared__ uint32_t s_bits[];
  const int j = blockIdx.x * 256 + threadIdx.x;
  ushort4 rc = make_ushort4(0, 0, 0, 0);
  if (j < P) rc = rect[order[j]];
  chunk_bitmap(s_bits, gx, tg, rc);
  for (int b = threadIdx.x; b < tg.nt; b += 256) {
    uint32_t c = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) c += __popc(s_bits[b * 8 + w]);
    hist[(size_t)b * ncta + blockIdx.x] = c;
  }
}

__global__ void __launch_bounds__(256) k_bin_scan(uint32_t* __restrict__ hist, size_t ncta, uint32_t* __restrict__ bintotal,
                                                  const TileGroup tg, int first, uint2* __restrict__ ranges,
                                                  uint32_t* __restrict__ ctrl, uint32_t r_capacity) {
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
  // ---- the CTA that finishes last turns the tile totals into ranges (one launch less on the critical path):
  // exclusive scan of total[], continued at the instance count the previous groups left in ctrl[0]
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    bintotal[blockIdx.x] = carry;
    __threadfence();
    const uint32_t done = atomicAdd(&ctrl[2], 1u);
    s_last = (done == gridDim.x - 1) ? 1 : 0;
    if (s_last) ctrl[2] = 0u;  // left at zero for the next launch
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  uint32_t base_count = first ? 0u : ctrl[0];
  __syncthreads();  // everybody has read ctrl[0] before thread 0 replaces it
  for (int base = 0; base < tg.nt; base += 256) {
    const int i = base + threadIdx.x;
    const uint32_t v = (i < tg.nt) ? __ldcg(bintotal + i) : 0u;
    uint32_t total;
    const uint32_t ex = block_excl_scan(v, sm, &total) + base_count;
    if (i < tg.nt) ranges[tg.t0 + i] = v ? make_uint2(ex, ex + v) : make_uint2(0, 0);  // empty tiles: (0,0) like the memset
    base_count += total;
  }
  if (threadIdx.x == 0) {
    ctrl[0] = base_count;                          // R (so far)
    ctrl[1] = base_count > r_capacity ? 1u : 0u;   // overflow: the instance buffers are too small, nothing more is written
  }
}

__global__ void __launch_bounds__(256) k_bin_scatter(int P, const uint32_t* __restrict__ order,
                                                     const ushort4* __restrict__ rect, const TileGroup tg, int gx,
                                                     size_t ncta, const uint32_t* __restrict__ hist,
                                                     const uint2* __restrict__ ranges, const uint32_t* __restrict__ ctrl,
                                                     uint32_t r_capacity, uint32_t* __restrict__ point_list) {
  extern __shared__ uint32_t s_bits[];  // [tiles of the group][8] bitmap, then [tiles] first output slot of this chunk
  if (ctrl[0] > r_capacity) return;     // ctrl[0] = end of this group's instances: they would not fit
  uint32_t* s_base = s_bits + tg.nt * 8;
  const int j = blockIdx.x * 256 + threadIdx.x;
  ushort4 rc = make_ushort4(0, 0, 0, 0);
  uint32_t id = 0;
  if (j < P) {
    id = order[j];
    rc = rect[id];
  }
  uint8_t* s_pref = reinterpret_cast<uint8_t*>(s_base + tg.nt);  // [tiles][8]: instances of the tile before each 32-surfel word
  for (int b = threadIdx.x; b < tg.nt; b += 256) s_base[b] = ranges[tg.t0 + b].x + hist[(size_t)b * ncta + blockIdx.x];
  chunk_bitmap(s_bits, gx, tg, rc);
  for (int b = threadIdx.x; b < tg.nt; b += 256) {
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      s_pref[b * 8 + w] = (uint8_t)run;  // <= 224
      run += __popc(s_bits[b * 8 + w]);
    }
  }
  __syncthreads();
  for_each_chunk_tile(rc, gx, tg, id, [&](int tile, int t, uint32_t sid) {
    const int w = t >> 5;
    const uint32_t r = (uint32_t)s_pref[tile * 8 + w] + __popc(s_bits[tile * 8 + w] & ((1u << (t & 31)) - 1u));
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

// Group g of the tile grid (bin_group_count(gx, gy) groups, consecutive tile ids in ascending order).
static TileGroup tile_group(int gx, int gy, int g) {
  TileGroup tg;
  if (gx <= GSL_BIN_GROUP_TILES) {  // whole tile rows
    const int rows = GSL_BIN_GROUP_TILES / gx;
    tg.y0 = g * rows;
    tg.y1 = tg.y0 + rows < gy ? tg.y0 + rows : gy;
    tg.x0 = 0;
    tg.x1 = 1 << 30;  // no column clipping (wrapped rects run past gx)
    tg.t0 = tg.y0 * gx;
    tg.nt = (tg.y1 - tg.y0) * gx;
  } else {                          // pieces of one row
    const int per_row = (gx + GSL_BIN_GROUP_TILES - 1) / GSL_BIN_GROUP_TILES;
    tg.y0 = g / per_row;
    tg.y1 = tg.y0 + 1;
    tg.x0 = (g % per_row) * GSL_BIN_GROUP_TILES;
    tg.x1 = tg.x0 + GSL_BIN_GROUP_TILES < gx ? tg.x0 + GSL_BIN_GROUP_TILES : gx;
    tg.t0 = tg.y0 * gx + tg.x0;
    tg.nt = tg.x1 - tg.x0;
  }
  return tg;
}

// host-side view of the grouping for include/gsl_b200.h: gsl_bin_groups (CPU tests of the partition)
int bin_groups_describe(int W, int H, int32_t* out, int capacity) {
  const int gx = (W + GSL_BLOCK_X - 1) / GSL_BLOCK_X, gy = (H + GSL_BLOCK_Y - 1) / GSL_BLOCK_Y;
  const int n = bin_group_count(gx, gy);
  for (int g = 0; g < n && g < capacity; ++g) {
    const TileGroup tg = tile_group(gx, gy, g);
    int32_t* o = out + 6 * g;
    o[0] = tg.t0; o[1] = tg.nt; o[2] = tg.y0; o[3] = tg.y1; o[4] = tg.x0; o[5] = tg.x1 < gx ? tg.x1 : gx;
  }
  return n;
}

// The surfels are already depth-sorted (launch_surfel_sort): one stable counting pass per tile group, then the block
// lists.  Nothing here needs R on the host; r_host receives (R, overflow) asynchronously.
int launch_binning(const gsl_params& p, const GeomView& g, const ImageView& im, const BinView& b,
                   int64_t r_capacity, int32_t* r_host, cudaStream_t st) {
  const int gx = (p.W + GSL_BLOCK_X - 1) / GSL_BLOCK_X, gy = (p.H + GSL_BLOCK_Y - 1) / GSL_BLOCK_Y;
  const int tiles = gx * gy;
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
  const int ngroups = bin_group_count(gx, gy);
  for (int gi = 0; gi < ngroups; ++gi) {
    const TileGroup tg = tile_group(gx, gy, gi);
    const size_t smem = (size_t)tg.nt * 9 * sizeof(uint32_t);         // bitmap [+ first slots]
    const size_t smem_scatter = smem + (size_t)tg.nt * 8;               // + prefix bytes: 44 KB for a full group
    {
      ProfScope prof(GSL_K_SCAN, st);
      k_bin_count<<<ncta, 256, smem, st>>>(p.P, g.sval_b, g.rect, tg, gx, im.ncta, im.hist);
      k_bin_scan<<<tg.nt, 256, 0, st>>>(im.hist, im.ncta, im.bintotal, tg, gi == 0 ? 1 : 0, im.ranges, g.ctrl,
                                        (uint32_t)r_capacity);
    }
    if (r_host && gi == ngroups - 1) {
      r_host[0] = -1;  // sentinel for wait_num_rendered
      cudaMemcpyAsync(r_host, g.ctrl, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
    }
    ProfScope prof(GSL_K_DUPLICATE, st);
    k_bin_scatter<<<ncta, 256, smem_scatter, st>>>(p.P, g.sval_b, g.rect, tg, gx, im.ncta, im.hist, im.ranges, g.ctrl,
                                          (uint32_t)r_capacity, b.vals_b);
  }
  ProfScope prof(GSL_K_RANGES, st);
  k_tile_blists<<<tiles, TB_THREADS, 0, st>>>(b.vals_b, g.pixbox, g.ctrl, (uint32_t)r_capacity, gx, p.W, p.H,
                                             b.plane_stride, im.ranges, b.blist, im.bdesc);
  return check_cuda(cudaGetLastError(), "binning launch");
}

}  // namespace gsl
