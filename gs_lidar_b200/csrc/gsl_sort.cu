// gsl_sort.cu -- depth sort of the surfels (fast binning path), hand-written, no library.
//
// What is needed is the permutation of the P surfels by (depth bits, id): the instances of a tile then come out in
// the order of the reference's stable 64-bit (tile | depth) radix sort (rasterizer_impl.cu:338-344; see
// gsl_binning.cu).  Depth keys are positive floats, i.e. their bit patterns are monotone and smoothly spread, so an
// MSD bucket sort finishes in two levels:
//   k_depth_keys (gsl_preprocess.cu)  key[i] = bits(r_i); min / max key by warp reduction + one atomic per warp
//   k_sort_hist     bucket = (key - kmin) >> shift  (NB <= 16384 buckets of ~128 surfels), global histogram
//   k_sort_hist     ... and the atomic's return value = the surfel's arrival rank inside its bucket
//   k_sort_scan     exclusive scan of the histogram -> bucket_start[]
//   k_sort_scatter  (key, id) -> bucket_start + rank (the order inside a bucket does not matter ...)
//   k_sort_buckets  ... because every bucket is then sorted by the unique 64-bit value (key << 32 | id): bitonic
//                   network in shared memory, one CTA per bucket.  A bucket that does not fit (degenerate inputs:
//                   thousands of surfels at identical range) is sorted by the same network in global memory --
//                   slow but exact.
// The whole sort runs on the side stream under k_preprocess_fwd (gsl_api.cu).
#include "gsl_common.cuh"

namespace gsl {

constexpr int SORT_CAP = 2048;      // elements a bucket may hold to be sorted in shared memory
constexpr int SORT_THREADS = 256;

__host__ __device__ inline uint32_t sort_num_buckets(int P) {
  uint32_t nb = 256;
  while (nb < 16384u && (uint64_t)nb * 192u < (uint64_t)(P > 0 ? P : 1)) nb <<= 1;
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
// saves the scatter pass its own atomics
__global__ void __launch_bounds__(256) k_sort_hist(int P, const uint32_t* __restrict__ keys, const uint32_t* __restrict__ ctrl,
                                                   uint32_t nb, uint32_t* __restrict__ count, uint32_t* __restrict__ rank) {
  const SortDomain d = sort_domain(ctrl, nb);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x)
    rank[i] = atomicAdd(&count[(keys[i] - d.kmin) >> d.shift], 1u);
}

// exclusive scan of count[0..nb) -> start[0..nb]; one CTA, every thread owns nb / 1024 consecutive buckets
__global__ void __launch_bounds__(1024) k_sort_scan(uint32_t nb, const uint32_t* __restrict__ count, uint32_t* __restrict__ start) {
  __shared__ uint32_t s_w[32];
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  const uint32_t per = (nb + 1023u) / 1024u;  // <= 16
  const uint32_t i0 = threadIdx.x * per;
  uint32_t v[16];
  uint32_t sum = 0;
#pragma unroll
  for (uint32_t k = 0; k < 16; ++k) {
    v[k] = (k < per && i0 + k < nb) ? count[i0 + k] : 0u;
    sum += v[k];
  }
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
#pragma unroll
  for (uint32_t k = 0; k < 16; ++k) {
    if (k < per && i0 + k < nb) start[i0 + k] = ex;
    ex += v[k];
  }
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

// compare-exchange network step of the bitonic sort over m = 2^x virtual elements (indices >= n hold +inf)
// Bitonic network over M <= SORT_THREADS elements, one per thread (threads >= M idle): partners closer than a warp are
// exchanged with shuffles, only the steps with j >= 32 go through shared memory.  Fully unrolled: the directions and
// partner tests fold into constants per step.  Must be called by all threads of the CTA.
template <int M>
__device__ __forceinline__ unsigned long long bitonic_one_per_thread(unsigned long long v, unsigned long long* s_v) {
  const uint32_t i = threadIdx.x;
  const bool warp_live = (i & ~31u) < (uint32_t)M;
#pragma unroll
  for (int k = 2; k <= M; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      unsigned long long o = 0;
      if (j >= 32) {
        __syncthreads();
        if (warp_live) s_v[i] = v;
        __syncthreads();
        if (warp_live) o = s_v[i ^ j];
      } else if (warp_live) {
        o = __shfl_xor_sync(0xffffffffu, v, j);
      }
      if (warp_live) {
        // the lower partner keeps the smaller value when sorting upwards: one 64-bit compare, one select
        const bool take_min = (((i & j) == 0) == ((i & k) == 0));
        v = ((v < o) == take_min) ? v : o;
      }
    }
  }
  return v;
}

__device__ void sort_one_bucket(unsigned long long* s_v, uint32_t lo, uint32_t n, uint32_t* __restrict__ tmp_key,
                                uint32_t* __restrict__ tmp_id, uint32_t* __restrict__ order) {
  if (n == 0) return;
  if (n == 1) {
    if (threadIdx.x == 0) order[lo] = tmp_id[lo];
    return;
  }
  uint32_t m = 2;
  while (m < n) m <<= 1;
  if (n <= (uint32_t)SORT_THREADS) {
    // the common case: one element per thread, network fully unrolled for the padded size
    const uint32_t i = threadIdx.x;
    unsigned long long v = (i < n) ? (((unsigned long long)tmp_key[lo + i] << 32) | tmp_id[lo + i]) : ~0ull;
    if (m <= 32) v = bitonic_one_per_thread<32>(v, s_v);
    else if (m == 64) v = bitonic_one_per_thread<64>(v, s_v);
    else if (m == 128) v = bitonic_one_per_thread<128>(v, s_v);
    else v = bitonic_one_per_thread<256>(v, s_v);
    if (i < n) order[lo + i] = (uint32_t)v;
    return;
  }
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

// persistent: a few CTAs per SM walk the buckets, so that the sort leaves room for the kernel it runs under
__global__ void __launch_bounds__(SORT_THREADS) k_sort_buckets(uint32_t nb, const uint32_t* __restrict__ start,
                                                               uint32_t* __restrict__ tmp_key, uint32_t* __restrict__ tmp_id,
                                                               uint32_t* __restrict__ order) {
  __shared__ unsigned long long s_v[SORT_CAP];
  for (uint32_t b = blockIdx.x; b < nb; b += gridDim.x) {
    const uint32_t lo = start[b], n = start[b + 1] - lo;
    sort_one_bucket(s_v, lo, n, tmp_key, tmp_id, order);
    __syncthreads();  // s_v is reused by the next bucket
  }
}

// CUDA-graph capture does not carry the side stream's priority over to the kernel nodes; gsl_graph_end re-applies it to
// the nodes of these kernels (they must slip in between the CTAs of k_preprocess_fwd, not queue behind them).
bool is_sort_kernel(const void* func) {
  return func == (const void*)k_sort_hist || func == (const void*)k_sort_scan || func == (const void*)k_sort_scatter ||
         func == (const void*)k_sort_buckets;
}

// surfel ids in (depth bits, id) order -> g.sval_b.  g.skey_a holds the keys, g.ctrl[8..9] their (~min, max).
int launch_surfel_sort(const gsl_params& p, const GeomView& g, cudaStream_t st) {
  if (p.P == 0) return 0;
  const uint32_t nb = sort_num_buckets(p.P);
  uint32_t* count = g.sort_buckets;
  uint32_t* start = g.sort_buckets + 16384 + 64;
  uint32_t* rank = g.offs;  // scratch: the tiles_touched scan is not used on the fast binning path
  ProfScope prof(GSL_K_SORT, st);
  cudaMemsetAsync(count, 0, nb * sizeof(uint32_t), st);
  // small grids (a few CTAs per SM): the sort is atomic / latency bound and shares the GPU with k_preprocess_fwd
#ifndef GSL_SORT_GRID
#define GSL_SORT_GRID 12
#endif
  const int blocks = min((p.P + 255) / 256, 148 * GSL_SORT_GRID);
  k_sort_hist<<<blocks, 256, 0, st>>>(p.P, g.skey_a, g.ctrl, nb, count, rank);
  k_sort_scan<<<1, 1024, 0, st>>>(nb, count, start);
  k_sort_scatter<<<blocks, 256, 0, st>>>(p.P, g.skey_a, g.ctrl, nb, start, rank, g.skey_b, g.sval_a);
  k_sort_buckets<<<min((int)nb, 148 * GSL_SORT_GRID), SORT_THREADS, 0, st>>>(nb, start, g.skey_b, g.sval_a, g.sval_b);
  return check_cuda(cudaGetLastError(), "surfel sort launch");
}

}  // namespace gsl
