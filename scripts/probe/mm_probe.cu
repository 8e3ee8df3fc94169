// NVLS probe (development aid, not part of the product): multimem.st / multimem.ld_reduce / multimem.red on a multicast
// address handed in by the caller (torch symmetric memory).  Built by scripts/probe/run_mm_probe.py.
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k_st(float4* mc, float4 v, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc + i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__global__ void k_ld(const float4* mc, float4* out, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc + i) : "memory");
    out[i] = r;
  }
}
__global__ void k_p2p(float4* dst, float4 v, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) dst[i] = v;
}
__global__ void k_or(uint32_t* mc, uint32_t v, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) asm volatile("multimem.red.relaxed.sys.global.or.b32 [%0], %1;" ::"l"(mc + i), "r"(v) : "memory");
}
__global__ void k_st_weak(float4* mc, float4 v, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc + i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__global__ void k_ld_weak(const float4* mc, float4* out, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float4 r;
    asm volatile("multimem.ld_reduce.weak.global.add.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc + i) : "memory");
    out[i] = r;
  }
}
// ld_reduce + multimem.st of the sum in one pass (what a tile owner would do)
__global__ void k_ld_st(const float4* mc_in, float4* mc_out, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float4 r;
    asm volatile("multimem.ld_reduce.weak.global.add.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc_in + i) : "memory");
    asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc_out + i), "f"(r.x), "f"(r.y), "f"(r.z), "f"(r.w) : "memory");
  }
}
extern "C" {
int mm_st_weak(void* mc, float v, long n, int blocks, void* st) { k_st_weak<<<blocks, 256, 0, (cudaStream_t)st>>>((float4*)mc, make_float4(v, v, v, v), n); return (int)cudaGetLastError(); }
int mm_ld_weak(const void* mc, void* out, long n, int blocks, void* st) { k_ld_weak<<<blocks, 256, 0, (cudaStream_t)st>>>((const float4*)mc, (float4*)out, n); return (int)cudaGetLastError(); }
int mm_ld_st(const void* mc, void* mc_out, long n, int blocks, void* st) { k_ld_st<<<blocks, 256, 0, (cudaStream_t)st>>>((const float4*)mc, (float4*)mc_out, n); return (int)cudaGetLastError(); }
int mm_st(void* mc, float v, long n, void* st) { k_st<<<148 * 4, 256, 0, (cudaStream_t)st>>>((float4*)mc, make_float4(v, v, v, v), n); return (int)cudaGetLastError(); }
int mm_ld(const void* mc, void* out, long n, void* st) { k_ld<<<148 * 4, 256, 0, (cudaStream_t)st>>>((const float4*)mc, (float4*)out, n); return (int)cudaGetLastError(); }
int mm_p2p(void* dst, float v, long n, void* st) { k_p2p<<<148 * 4, 256, 0, (cudaStream_t)st>>>((float4*)dst, make_float4(v, v, v, v), n); return (int)cudaGetLastError(); }
int mm_or(void* mc, uint32_t v, int n, void* st) { k_or<<<(n + 255) / 256, 256, 0, (cudaStream_t)st>>>((uint32_t*)mc, v, n); return (int)cudaGetLastError(); }
}
