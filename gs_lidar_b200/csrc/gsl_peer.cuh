// gsl_peer.cuh -- device helpers of the peer-memory gradient exchange shared by gsl_peer.cu and gsl_preprocess.cu:
// system-scope flag stores / loads, the device-side step counter and the two in-kernel halves of a barrier
// ("the last CTA of the producing kernel publishes the flag", "every CTA of the consuming kernel waits for the flags").
#pragma once
#include "gsl_common.cuh"

namespace gsl {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Fused step: ticket and factor-table parity come from the step counter in the own header (written by k_peer_begin,
// an earlier kernel of the same stream), so every kernel argument of the step is a constant and the step can be
// captured in a CUDA graph.  Piecewise calls (tests, callers that schedule the pieces themselves) pass them by value.
__device__ __forceinline__ void peer_resolve_step(PeerView& pv) {
  if (pv.dev_step) {
    const uint32_t s = ld_relaxed_gpu(reinterpret_cast<const uint32_t*>(pv.own + PEER_STEP_OFF));
    pv.epoch = s;
    pv.parity = (int)(s & 1u);
  }
}

__device__ __forceinline__ bool peer_error(const PeerView& pv) {
  return ld_relaxed_gpu(reinterpret_cast<const uint32_t*>(pv.own + PEER_ERROR_OFF)) != 0u;
}

// Every CTA of a consuming kernel: thread g waits until rank g has published ticket >= pv.epoch on flag slot `slot` of the
// OWN buffer (local memory), then the CTA goes on.  Everything rank g stored (to any rank) before it published the flag is
// visible afterwards.  Nothing of the awaited data may have been read by this kernel before the call (the L1 is not
// coherent with remote stores).  A peer that never arrives raises the error word instead of hanging.
__device__ __forceinline__ void peer_wait_flags(const PeerView& pv, int slot) {
  if ((int)threadIdx.x < pv.world) {
    const uint32_t* f = reinterpret_cast<const uint32_t*>(pv.own) + slot * PEER_MAX + threadIdx.x;
    if ((int32_t)(ld_acquire_sys(f) - pv.epoch) < 0) {
      const unsigned long long t0 = global_timer_ns();
      while ((int32_t)(ld_acquire_sys(f) - pv.epoch) < 0) {
        if (peer_error(pv)) break;  // the step has failed already (another CTA timed out): do not wait again
        if (global_timer_ns() - t0 > pv.timeout_ns) {
          *reinterpret_cast<volatile uint32_t*>(pv.own + PEER_ERROR_OFF) = 1u + (uint32_t)slot;
          if (pv.error_host) *reinterpret_cast<volatile int*>(pv.error_host) = 1 + slot;
          break;
        }
        __nanosleep(64);
      }
    }
  }
  __syncthreads();
}

// End of a producing kernel, called by ALL threads of every CTA: each thread's remote stores are made visible system-wide,
// and the CTA that finishes last publishes ticket pv.epoch on flag slot `slot` of every rank (release).  `which` selects the
// finished-CTA counter in the own header (one per producing kernel of the step); it is left at zero for the next step.
__device__ __forceinline__ void peer_signal_when_last(const PeerView& pv, int slot, int which) {
  __shared__ int s_last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t* done = reinterpret_cast<uint32_t*>(pv.own + PEER_DONE_OFF) + which;
    const unsigned total = gridDim.x * gridDim.y * gridDim.z;
    const unsigned old = atomicAdd(done, 1u);
    s_last = (old == total - 1u) ? 1 : 0;
    if (s_last) *done = 0u;
  }
  __syncthreads();
  if (s_last && (int)threadIdx.x < pv.world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t*>(pv.buf[threadIdx.x]) + slot * PEER_MAX + pv.rank, pv.epoch);
  }
}

}  // namespace gsl
