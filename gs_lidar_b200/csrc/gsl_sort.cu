// gsl_sort.cu -- depth sort of the surfels, hand-written, no library.
//
// What is needed is the permutation of the P surfels by (depth bits, id): the instances of a tile then come out in
// the order of the reference's stable 64-bit (tile | depth) radix sort (rasterizer_impl.cu:338-344; see
// gsl_binning.cu).  Depth keys are positive floats, i.e. their bit patterns are monotone and smoothly spread, so an
// MSD bucket sort finishes in two levels:
//   k_depth_keys (gsl_preprocess.cu)  key[i] = bits(r_i); min / max key by warp reduction + one atomic per warp;
//                   also zero-fills the histogram
//   k_sort_hist     bucket = (key - kmin) >> shift  (NB <= 65536 buckets of ~32-64 surfels), global histogram
//   k_sort_hist     ... and the atomic's return value = the surfel's arrival rank inside its bucket
//   k_sort_scan     exclusive scan of the histogram -> bucket_start[]
//   k_sort_scatter  (key, id) -> bucket_start + rank (the order inside a bucket does not matter ...)
//   k_sort_buckets  ... because every bucket is then sorted by the unique 64-bit value (key << 32 | id).  One WARP per
//                   bucket of up to 256 surfels: 1, 2, 4 or 8 elements per lane in registers, bitonic network with
//                   shuffles for partner distances below 32 and register-local exchanges above -- no shared memory,
//                   no barrier.  Larger buckets (clustered depths) are sorted by the whole CTA in shared memory, and
//                   buckets that do not fit there (degenerate inputs: thousands of surfels at identical range) by
//                   the same network in global memory -- slow but exact.
// The whole sort runs on the side stream under k_preprocess_fwd (gsl_api.cu).
#include "gsl_common.cuh"

namespace gsl {

constexpr int SORT_CAP = 2048;      // elements a bucket may hold to be sorted in shared memory
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARP_MAX = 256;  // largest bucket one warp sorts in registers (8 elements per lane)
constexpr int SORT_CTA_BUCKETS = 64;  // consecutive buckets one CTA owns at most

__host__ __device__ inline uint32_t sort_num_buckets(int P) {
  uint32_t nb = 256;
  while (nb < (uint32_t)GSL_SORT_MAX_BUCKETS && (uint64_t)nb * 64u < (uint64_t)(P > 0 ? P : 1)) nb <<= 1;
  return nb;
}

// bucket = (key - kmin) >> shift with the smallest shift that maps [kmin, kmax] into [0, nb)
struct SortDomain {
  uint32_t kmin, shift;
};
__device__ __forceinline__ SortDomain sort_domain(const uint32_t* __restrict__ ctrl, uint32_t nb) {
  SortDomain d;
  d.kmin = ~ctrl[8];  // ctrl[8] accumulates max(~key)
  const uint32_t span = ctrl[9] - d.kmin;  // kmax - kmin (0 when all keys are equal)
  const int bits_span = 32 - __clz(span | 1u);
  const int bits_nb = 31 - __clz(nb);
  d.shift = (uint32_t)max(0, bits_span - bits_nb);
  return d;
}

// histogram; the value the atomic returns is the element's (arbitrary but unique) rank inside its bucket, which
// saves the scatter pass its own atomics.  Two elements per turn: the atomics are latency, not issue.
__global__ void __launch_bounds__(256) k_sort_hist(int P, const uint32_t* __restrict__ keys, const uint32_t* __restrict__ ctrl,
                                                   uint32_t nb, uint32_t* __restrict__ count, uint32_t* __restrict__ rank) {
  const SortDomain d = sort_domain(ctrl, nb);
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += 2 * stride) {
    const int i2 = i + stride;
    const uint32_t k1 = keys[i];
    const uint32_t k2 = i2 < P ? keys[i2] : 0u;
    const uint32_t r1 = atomicAdd(&count[(k1 - d.kmin) >> d.shift], 1u);
    uint32_t r2 = 0;
    if (i2 < P) r2 = atomicAdd(&count[(k2 - d.kmin) >> d.shift], 1u);
    rank[i] = r1;
    if (i2 < P) rank[i2] = r2;
  }
}

// exclusive scan of count[0..nb) -> start[0..nb]; one CTA, every thread owns PER = nb / 1024 consecutive buckets (nb is a
// power of two >= 256).  PER <= 16: the counts stay in registers between the two halves (16-byte loads / stores); the
// single CTA is a latency chain on the sort's critical path, so every dependent round trip counts.
template <int PER>
__device__ __forceinline__ void sort_scan_body(uint32_t nb, const uint32_t* __restrict__ count, uint32_t* __restrict__ start,
                                               uint32_t* s_w) {
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  const uint32_t i0 = threadIdx.x * PER;
  uint32_t v[PER];
  uint32_t sum = 0;
  if (PER >= 4) {
#pragma unroll
    for (int k = 0; k < PER; k += 4) {
      const uint4 q = (i0 + k < nb) ? *reinterpret_cast<const uint4*>(count + i0 + k) : make_uint4(0u, 0u, 0u, 0u);
      v[k] = q.x; v[k + 1] = q.y; v[k + 2] = q.z; v[k + 3] = q.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < PER; ++k) v[k] = (i0 + k < nb) ? count[i0 + k] : 0u;
  }
#pragma unroll
  for (int k = 0; k < PER; ++k) sum += v[k];
  uint32_t inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_w[wv] = inc;
  __syncthreads();
  if (wv == 0) {
    uint32_t w = s_w[lane], winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    s_w[lane] = winc - w;
    if (lane == 31) start[nb] = winc;
  }
  __syncthreads();
  uint32_t ex = s_w[wv] + inc - sum;
  if (i0 >= nb) return;
  if (PER >= 4) {
#pragma unroll
    for (int k = 0; k < PER; k += 4) {
      uint4 q;
      q.x = ex; ex += v[k];
      q.y = ex; ex += v[k + 1];
      q.z = ex; ex += v[k + 2];
      q.w = ex; ex += v[k + 3];
      *reinterpret_cast<uint4*>(start + i0 + k) = q;
    }
  } else {
#pragma unroll
    for (int k = 0; k < PER; ++k) { start[i0 + k] = ex; ex += v[k]; }
  }
}

__global__ void __launch_bounds__(1024) k_sort_scan(uint32_t nb, const uint32_t* __restrict__ count, uint32_t* __restrict__ start) {
  __shared__ uint32_t s_w[32];
  const uint32_t per = (nb + 1023u) / 1024u;  // 1 (nb <= 1024), 2, 4, ..., GSL_SORT_MAX_BUCKETS / 1024
  if (per <= 1) { sort_scan_body<1>(nb, count, start, s_w); return; }
  if (per == 2) { sort_scan_body<2>(nb, count, start, s_w); return; }
  if (per == 4) { sort_scan_body<4>(nb, count, start, s_w); return; }
  if (per == 8) { sort_scan_body<8>(nb, count, start, s_w); return; }
  if (per == 16) { sort_scan_body<16>(nb, count, start, s_w); return; }
  if (per == 32) { sort_scan_body<32>(nb, count, start, s_w); return; }
  sort_scan_body<64>(nb, count, start, s_w);
}

__global__ void __launch_bounds__(256) k_sort_scatter(int P, const uint32_t* __restrict__ keys, const uint32_t* __restrict__ ctrl,
                                                      uint32_t nb, const uint32_t* __restrict__ start,
                                                      const uint32_t* __restrict__ rank, uint32_t* __restrict__ tmp_key,
                                                      uint32_t* __restrict__ tmp_id) {
  const SortDomain d = sort_domain(ctrl, nb);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
    const uint32_t k = keys[i];
    const uint32_t pos = start[(k - d.kmin) >> d.shift] + rank[i];
    tmp_key[pos] = k;
    tmp_id[pos] = (uint32_t)i;
  }
}

// One warp sorts a bucket of n <= 32 E elements held E per lane (element i = e * 32 + lane, padding = +inf): bitonic
// network, ascending.  Partner i ^ j is another lane for j < 32 (one 64-bit shuffle) and another register of the same
// lane for j >= 32; with the loops unrolled every direction folds into a constant or a test of the lane id.
template <int E>
__device__ __forceinline__ void warp_sort_bucket(uint32_t lo, uint32_t n, const uint32_t* __restrict__ tmp_key,
                                                 const uint32_t* __restrict__ tmp_id, uint32_t* __restrict__ order, int lane) {
  unsigned long long v[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const uint32_t i = (uint32_t)(e * 32 + lane);
    v[e] = (i < n) ? (((unsigned long long)tmp_key[lo + i] << 32) | tmp_id[lo + i]) : ~0ull;
  }
#pragma unroll
  for (int k = 2; k <= 32 * E; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int jj = j >> 5;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          if ((e & jj) == 0) {
            const unsigned long long a = v[e], b = v[e | jj];
            const bool up = ((e * 32) & k) == 0;  // k >= 64 here: the lane bits do not matter
            const bool swap = (a > b) == up;
            v[e] = swap ? b : a;
            v[e | jj] = swap ? a : b;
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const unsigned long long o = __shfl_xor_sync(0xffffffffu, v[e], j);
          const int i = e * 32 + lane;
          const bool take_min = (((i & j) == 0) == ((i & k) == 0));
          v[e] = ((v[e] < o) == take_min) ? v[e] : o;
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const uint32_t i = (uint32_t)(e * 32 + lane);
    if (i < n) order[lo + i] = (uint32_t)v[e];
  }
}

// A bucket too large for one warp, sorted by the whole CTA (must be called by all its threads).
__device__ void sort_big_bucket(unsigned long long* s_v, uint32_t lo, uint32_t n, uint32_t* __restrict__ tmp_key,
                                uint32_t* __restrict__ tmp_id, uint32_t* __restrict__ order) {
  uint32_t m = 2;
  while (m < n) m <<= 1;
  if (n <= (uint32_t)SORT_CAP) {
    for (uint32_t i = threadIdx.x; i < m; i += SORT_THREADS)
      s_v[i] = (i < n) ? (((unsigned long long)tmp_key[lo + i] << 32) | tmp_id[lo + i]) : ~0ull;
    __syncthreads();
    for (uint32_t k = 2; k <= m; k <<= 1) {
      for (uint32_t j = k >> 1; j > 0; j >>= 1) {
        for (uint32_t t = threadIdx.x; t < (m >> 1); t += SORT_THREADS) {
          const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // index with bit j clear
          const uint32_t p = i | j;
          const unsigned long long a = s_v[i], b = s_v[p];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { s_v[i] = b; s_v[p] = a; }
        }
        __syncthreads();
      }
    }
    for (uint32_t i = threadIdx.x; i < n; i += SORT_THREADS) order[lo + i] = (uint32_t)s_v[i];
    return;
  }
  // degenerate bucket: the same network on the (key, id) pairs in global memory
  for (uint32_t k = 2; k <= m; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t t = threadIdx.x; t < (m >> 1); t += SORT_THREADS) {
        const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const uint32_t p = i | j;
        if (p < n) {  // i < p; an out-of-range partner is +inf and never moves down
          const unsigned long long a = ((unsigned long long)tmp_key[lo + i] << 32) | tmp_id[lo + i];
          const unsigned long long b = ((unsigned long long)tmp_key[lo + p] << 32) | tmp_id[lo + p];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            tmp_key[lo + i] = (uint32_t)(b >> 32); tmp_id[lo + i] = (uint32_t)b;
            tmp_key[lo + p] = (uint32_t)(a >> 32); tmp_id[lo + p] = (uint32_t)a;
          }
        }
      }
      __syncthreads();
    }
  }
  for (uint32_t i = threadIdx.x; i < n; i += SORT_THREADS) order[lo + i] = tmp_id[lo + i];
}

// Every CTA owns `per` <= SORT_CTA_BUCKETS consecutive buckets: its warps take them in turn (a few CTAs per SM, so that
// the sort leaves room for the kernel it runs under); the buckets no warp can hold are queued and sorted by the whole
// CTA afterwards.  Block 0 also resets the key range for the next sort (k_depth_keys accumulates it with atomicMax).
__global__ void __launch_bounds__(SORT_THREADS) k_sort_buckets(uint32_t nb, uint32_t per, const uint32_t* __restrict__ start,
                                                               uint32_t* __restrict__ tmp_key, uint32_t* __restrict__ tmp_id,
                                                               uint32_t* __restrict__ order, uint32_t* __restrict__ ctrl) {
  __shared__ unsigned long long s_v[SORT_CAP];
  __shared__ uint32_t s_big[SORT_CTA_BUCKETS];
  __shared__ uint32_t s_nbig;
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_nbig = 0;
  if (blockIdx.x == 0 && threadIdx.x < 2) ctrl[8 + threadIdx.x] = 0u;
  __syncthreads();
  const uint32_t b0 = blockIdx.x * per, b1 = min(nb, b0 + per);
  for (uint32_t b = b0 + wv; b < b1; b += SORT_THREADS / 32) {
    const uint32_t lo = start[b], n = start[b + 1] - lo;
    if (n == 0) continue;
    if (n == 1) {
      if (lane == 0) order[lo] = tmp_id[lo];
    } else if (n <= 32) {
      warp_sort_bucket<1>(lo, n, tmp_key, tmp_id, order, lane);
    } else if (n <= 64) {
      warp_sort_bucket<2>(lo, n, tmp_key, tmp_id, order, lane);
    } else if (n <= 128) {
      warp_sort_bucket<4>(lo, n, tmp_key, tmp_id, order, lane);
    } else if (n <= (uint32_t)SORT_WARP_MAX) {
      warp_sort_bucket<8>(lo, n, tmp_key, tmp_id, order, lane);
    } else if (lane == 0) {
      s_big[atomicAdd(&s_nbig, 1u)] = b;
    }
  }
  __syncthreads();
  const uint32_t nbig = s_nbig;
  for (uint32_t q = 0; q < nbig; ++q) {
    const uint32_t b = s_big[q];
    const uint32_t lo = start[b], n = start[b + 1] - lo;
    sort_big_bucket(s_v, lo, n, tmp_key, tmp_id, order);
    __syncthreads();  // s_v is reused by the next bucket
  }
}

// CUDA-graph capture does not carry the side stream's priority over to the kernel nodes; gsl_graph_end re-applies it to
// the nodes of these kernels (they must slip in between the CTAs of k_preprocess_fwd, not queue behind them).
bool is_sort_kernel(const void* func) {
  return func == (const void*)k_sort_hist || func == (const void*)k_sort_scan || func == (const void*)k_sort_scatter ||
         func == (const void*)k_sort_buckets;
}

uint32_t sort_num_buckets_host(int P) { return sort_num_buckets(P); }

// surfel ids in (depth bits, id) order -> g.sval_b.  g.skey_a holds the keys, g.ctrl[8..9] their (~min, max); the
// histogram was zero-filled by k_depth_keys.
int launch_surfel_sort(const gsl_params& p, const GeomView& g, cudaStream_t st) {
  if (p.P == 0) return 0;
  const uint32_t nb = sort_num_buckets(p.P);
  uint32_t* count = g.sort_buckets;
  uint32_t* start = g.sort_buckets + GSL_SORT_MAX_BUCKETS + 64;
  uint32_t* rank = g.offs;  // scratch: the tiles_touched scan is only run for state exports
  ProfScope prof(GSL_K_SORT, st);
  // small grids (a few CTAs per SM): the sort is atomic / latency bound and shares the GPU with k_preprocess_fwd
#ifndef GSL_SORT_GRID
#define GSL_SORT_GRID 12
#endif
  const int blocks = min((p.P + 255) / 256, 148 * GSL_SORT_GRID);
  k_sort_hist<<<blocks, 256, 0, st>>>(p.P, g.skey_a, g.ctrl, nb, count, rank);
  k_sort_scan<<<1, 1024, 0, st>>>(nb, count, start);
  k_sort_scatter<<<blocks, 256, 0, st>>>(p.P, g.skey_a, g.ctrl, nb, start, rank, g.skey_b, g.sval_a);
  int bgrid = min((int)nb, 148 * GSL_SORT_GRID);
  uint32_t per = (nb + (uint32_t)bgrid - 1u) / (uint32_t)bgrid;
  if (per > (uint32_t)SORT_CTA_BUCKETS) {
    per = SORT_CTA_BUCKETS;
    bgrid = (int)((nb + per - 1u) / per);
  }
  k_sort_buckets<<<bgrid, SORT_THREADS, 0, st>>>(nb, per, start, g.skey_b, g.sval_a, g.sval_b, g.ctrl);
  return check_cuda(cudaGetLastError(), "surfel sort launch");
}

}  // namespace gsl
