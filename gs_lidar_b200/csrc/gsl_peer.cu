// gsl_peer.cu -- frame-parallel gradient exchange over NVLink peer memory (no library collective on the data path).
//
// The reference is single-GPU; SURVEY.md 8(e) shards the path by LiDAR frame and sums the surfel gradients of the
// ranks before the optimizer step.  With a ~0.9 ms compute step the two NCCL collectives of the first design (a 76 MB
// all-reduce + a G x 16 MB all-gather at 1M surfels) cost 0.31 + 0.21 ms on 8 B200s, most of it exposed, and P2P LOADS
// turned out latency-bound (~2 us per dependent round trip).  Here every rank keeps ONE peer-visible exchange buffer
// (cudaMalloc + CUDA IPC, mapped into the other ranks' address spaces) and all NVLink traffic is remote STORES issued
// by the kernels that produce the data:
//   * k_preprocess_bwd (gsl_preprocess.cu, GSL_FLAG_BWD_PEER_ROWS) pushes, per 256-surfel tile, the packed 64-byte
//     gradient rows of the surfels a pixel touched (~half of them) + their row bits into the staging area of the rank
//     that owns the tile (tile % world), and the non-zero 16-byte SH factors (packed to the front of the tile's
//     segment, + bits / prefix words) into every rank's factor table;
//   * barriers         -- a ticket release-stored into the flag slot this rank owns in every peer's buffer and an
//                         acquire-spin on the own slots (with a time-out that reports instead of hanging).  In the fused
//                         step (gsl_backward_surfels_exchange) the two halves live INSIDE the kernels: the last CTA of the
//                         producing kernel publishes the flag, every CTA of the consuming kernel waits for it
//                         (gsl_peer.cuh) -- no barrier kernel, no launch gap on the critical path, and the ticket is a
//                         device-side step counter, so every kernel argument is constant (CUDA-graph replayable).
//                         k_peer_barrier (one warp) remains for callers that schedule the pieces themselves;
//   * k_peer_reduce_rows -- the owner sums the staged rows of its tiles (local reads, fixed rank order) and pushes the
//                         sums + OR-ed bits into every rank's result area: every rank ends up with bit-identical sums;
//   * k_peer_sh_expand (gsl_preprocess.cu) -- gsl_sh_expand over the local factor tables;
//   * k_peer_unpack    -- summed packed rows -> the dense tensors autograd returns.
// All of it works on row ranges, so the exchange of one range runs (on a side stream) while k_preprocess_bwd computes
// the next.  Layout of an exchange buffer: PeerLayout (gsl_common.cuh).
#include "gsl_common.cuh"
#include "gsl_peer.cuh"
#include <algorithm>

namespace gsl {

// floats per packed exchange row: means2D.xy scales.xy | rot | means3D opacity | features (padded to whole 32-B sectors)
// (with the glue's VJP folded in the "features" are 4 ceil(S/4) + 8 channels, see peer_rows_S: 24 floats for S <= 4)
int peer_row_width(int S) { return S <= 4 ? 16 : 24; }

PeerLayout peer_layout(size_t P, int S, int world) {
  PeerLayout l;
  if (world < 1) world = 1;
  l.tiles = (int)((P + 255) / 256);
  l.tiles_per_rank = (l.tiles + world - 1) / world;
  const size_t rowbytes = (size_t)peer_row_width(S) * 4;
  l.off_fmeta = PEER_HEADER;
  l.off_factor = align_up(l.off_fmeta + 2 * (size_t)world * l.tiles * 8 * 8, 256);
  l.off_stagebits = align_up(l.off_factor + 2 * (size_t)world * l.tiles * 256 * 16, 256);
  l.off_stage = align_up(l.off_stagebits + (size_t)world * l.tiles_per_rank * 8 * 4, 256);
  l.off_rowbits = align_up(l.off_stage + (size_t)world * l.tiles_per_rank * 256 * rowbytes, 256);
  l.off_rows = align_up(l.off_rowbits + (size_t)l.tiles * 8 * 4, 256);
  l.total = align_up(l.off_rows + (size_t)l.tiles * 256 * rowbytes, 256);
  return l;
}

// ---- barrier ---------------------------------------------------------------------------------------------------
// Lane g tells rank g "rank `rank` reached `phase` of step `epoch`" and waits for the same news from rank g.  Everything
// this rank's stream did before (kernel boundary) is visible to a peer that has seen the flag; a peer that never
// arrives makes the wait give up after timeout_ns and raise *error (the results of the step are then undefined).
__global__ void k_peer_barrier(PeerView pv, int phase, int mode, unsigned long long timeout_ns, int* error, float glue_ts,
                               float glue_shift) {
  const int g = threadIdx.x;
  if (g >= pv.world) return;
  if ((phase == 0 || phase == 3) && (mode & 1)) {  // publish this rank's camera centre into rank g's table: k_peer_sh_expand then reads it locally
    const float* own = reinterpret_cast<const float*>(pv.own + PEER_CAMPOS_OFF);
    float* dst = reinterpret_cast<float*>(pv.buf[g] + PEER_CAMPOS_ALL_OFF) + 4 * (pv.parity * PEER_MAX + pv.rank);
    dst[0] = own[0]; dst[1] = own[1]; dst[2] = own[2];
    float* gdst = reinterpret_cast<float*>(pv.buf[g] + PEER_GLUE_ALL_OFF) + 4 * (pv.parity * PEER_MAX + pv.rank);
    gdst[0] = glue_ts; gdst[1] = glue_shift;  // (timestamp - time_shift, time_shift) of this rank's frame (gsl_peer_glue)
  }
  if (mode & 1) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t*>(pv.buf[g]) + phase * PEER_MAX + pv.rank, pv.epoch);
  }
  if (!(mode & 2)) return;
  const uint32_t* mine = reinterpret_cast<const uint32_t*>(pv.own) + phase * PEER_MAX + g;
  const unsigned long long t0 = global_timer_ns();
  while ((int32_t)(ld_acquire_sys(mine) - pv.epoch) < 0) {
    if (global_timer_ns() - t0 > timeout_ns) {
      if (error) *error = 1 + phase;
      *reinterpret_cast<volatile uint32_t*>(pv.own + PEER_ERROR_OFF) = 1u + (uint32_t)phase;  // gradients become NaN
      break;
    }
    __nanosleep(64);
  }
}

// First kernel of a fused step (one warp): advances the device-side step counter (same value on every rank: every rank
// runs the same number of steps), publishes this rank's camera centre in its own header and pushes it into every rank's
// table of the step's parity.  The pushes are ordered before the "pushed" flag of this step, which the last CTA of the
// per-surfel kernel (a later kernel of this stream) releases.
__global__ void k_peer_begin(PeerView pv, const float* __restrict__ campos, float glue_ts, float glue_shift) {
  const int g = threadIdx.x;
  uint32_t* step = reinterpret_cast<uint32_t*>(pv.own + PEER_STEP_OFF);
  const uint32_t s = *step + 1u;
  __syncwarp();
  if (g == 0) *step = s;
  if (g < 3) reinterpret_cast<float*>(pv.own + PEER_CAMPOS_OFF)[g] = campos[g];
  if (g < pv.world) {
    float* dst = reinterpret_cast<float*>(pv.buf[g] + PEER_CAMPOS_ALL_OFF) + 4 * ((int)(s & 1u) * PEER_MAX + pv.rank);
    dst[0] = campos[0]; dst[1] = campos[1]; dst[2] = campos[2];
    float* gdst = reinterpret_cast<float*>(pv.buf[g] + PEER_GLUE_ALL_OFF) + 4 * ((int)(s & 1u) * PEER_MAX + pv.rank);
    gdst[0] = glue_ts; gdst[1] = glue_shift;
    __threadfence_system();
  }
}

// Publishing half of an in-kernel barrier as its own one-warp kernel: everything this stream's earlier kernels stored (to any
// rank) is complete at the kernel boundary; lane g then release-stores the step's ticket into slot `slot` of rank g.  Used
// behind k_preprocess_bwd, whose ~4000 short-lived CTAs must not each stall on a system-scope fence (measured: +0.1 ms);
// k_peer_reduce_rows, one wave of persistent CTAs, publishes its flag itself (peer_signal_when_last).
__global__ void k_peer_signal(PeerView pv, int slot) {
  peer_resolve_step(pv);
  const int g = threadIdx.x;
  if (g >= pv.world) return;
  __threadfence_system();
  st_release_sys(reinterpret_cast<uint32_t*>(pv.buf[g]) + slot * PEER_MAX + pv.rank, pv.epoch);
}

// Waiting half as its own one-warp kernel: the stream goes on once every rank has published the step's ticket on `slot`.
// The consuming kernels wait in-kernel as well (peer_wait_flags), which is free once the flags are there; parking the wait
// in ONE warp instead of in every resident CTA of a 4000-CTA kernel leaves the SMs to the other streams of the step (the
// low-priority SH expansion fills exactly these waits).
__global__ void k_peer_wait(PeerView pv, int slot) {
  peer_resolve_step(pv);
  peer_wait_flags(pv, slot);
}

int launch_peer_wait_fused(const gsl_peer_ctx* c, int slot, cudaStream_t st) {
  k_peer_wait<<<1, 32, 0, st>>>(make_view(c, true), slot);
  return check_cuda(cudaGetLastError(), "k_peer_wait launch");
}

int launch_peer_signal_fused(const gsl_peer_ctx* c, int slot, cudaStream_t st) {
  k_peer_signal<<<1, 32, 0, st>>>(make_view(c, true), slot);
  return check_cuda(cudaGetLastError(), "k_peer_signal launch");
}

int launch_peer_begin(const gsl_peer_ctx* c, const float* campos, cudaStream_t st) {
  const float ts = c->glue ? c->glue->timestamp - c->glue->time_shift : 0.f, sh = c->glue ? c->glue->time_shift : 0.f;
  k_peer_begin<<<1, 32, 0, st>>>(make_view(c, true), campos, ts, sh);
  return check_cuda(cudaGetLastError(), "k_peer_begin launch");
}

// mode: 1 = signal, 2 = wait, 3 = both (the barrier)
int launch_peer_barrier(const gsl_peer_ctx* c, int phase, int mode, cudaStream_t st) {
  const float ts = c->glue ? c->glue->timestamp - c->glue->time_shift : 0.f, sh = c->glue ? c->glue->time_shift : 0.f;
  k_peer_barrier<<<1, 32, 0, st>>>(make_view(c), phase, mode, peer_timeout_ns(), c->error_flag, ts, sh);
  return check_cuda(cudaGetLastError(), "k_peer_barrier launch");
}

// ---- sum of the packed gradient rows -------------------------------------------------------------------------------
// Tile t (256 rows) belongs to rank t % world; every rank pushed its rows of the tile (and their bits) into the owner's
// staging area.  The owner reads the staged rows whose bit is set (local memory), sums them in rank order and pushes the
// sum row and the OR of the bits into every rank's result area -- remote STORES only, every rank gets the same bits.
// Rows nobody touched are neither read nor written (k_peer_unpack looks at the OR-ed bits).
__global__ void __launch_bounds__(256) k_peer_reduce_rows(PeerView pv, const PeerLayout pl, int rw4, int tile0,
                                                          int tile1, int fused) {
  __shared__ uint32_t s_bits[PEER_MAX][8];
  __shared__ uint32_t s_union[8];
  if (fused) {  // in-kernel barrier: every rank's rows of this step have arrived in my staging area
    peer_resolve_step(pv);
    peer_wait_flags(pv, PEER_SLOT_PUSHED);
  }
  const uint32_t* stagebits = reinterpret_cast<const uint32_t*>(pv.own + pl.off_stagebits);
  const float4* stage = reinterpret_cast<const float4*>(pv.own + pl.off_stage);
  // my tiles in [tile0, tile1): the first one is tile0 rounded up to rank (mod world)
  const int first = tile0 + ((pv.rank - tile0 % pv.world) + pv.world) % pv.world;
  for (int tile = first + blockIdx.x * pv.world; tile < tile1; tile += gridDim.x * pv.world) {
    const int lt = tile / pv.world;  // index among my tiles
    __syncthreads();
    if (threadIdx.x < 8 * PEER_MAX) {
      const int g = threadIdx.x >> 3, w = threadIdx.x & 7;
      s_bits[g][w] = g < pv.world ? stagebits[((size_t)g * pl.tiles_per_rank + lt) * 8 + w] : 0u;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
      uint32_t u = 0u;
#pragma unroll
      for (int g = 0; g < PEER_MAX; ++g) u |= s_bits[g][threadIdx.x];
      s_union[threadIdx.x] = u;
#pragma unroll
      for (int g = 0; g < PEER_MAX; ++g)
        if (g < pv.world) reinterpret_cast<uint32_t*>(pv.buf[g] + pl.off_rowbits)[tile * 8 + threadIdx.x] = u;
    }
    __syncthreads();
    for (int item = threadIdx.x; item < 256 * rw4; item += 256) {
      const int r = item / rw4, q = item - r * rw4;
      if (!((s_union[r >> 5] >> (r & 31)) & 1u)) continue;
      float4 v[PEER_MAX];
#pragma unroll
      for (int g = 0; g < PEER_MAX; ++g) {
        v[g] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g < pv.world && ((s_bits[g][r >> 5] >> (r & 31)) & 1u))
          v[g] = stage[(((size_t)g * pl.tiles_per_rank + lt) * 256 + r) * rw4 + q];
      }
      float4 acc = v[0];
#pragma unroll
      for (int g = 1; g < PEER_MAX; ++g)
        if (g < pv.world) { acc.x += v[g].x; acc.y += v[g].y; acc.z += v[g].z; acc.w += v[g].w; }
      const size_t at = ((size_t)tile * 256 + r) * rw4 + q;
#pragma unroll
      for (int g = 0; g < PEER_MAX; ++g)
        if (g < pv.world) reinterpret_cast<float4*>(pv.buf[g] + pl.off_rows)[at] = acc;
    }
  }
  if (fused) peer_signal_when_last(pv, PEER_SLOT_SUMMED, 1);  // the sums of my tiles are on their way to every rank
}

int launch_peer_reduce_rows(const gsl_peer_ctx* c, int P, int S, int row_begin, int row_end, cudaStream_t st, bool fused) {
  if (row_end <= row_begin) return 0;
  const int tile0 = row_begin / 256, tile1 = (row_end + 255) / 256;
  const int per_rank = (tile1 - tile0 + c->world - 1) / c->world;
  const int blocks = std::max(1, std::min(per_rank, 148 * 8));
  k_peer_reduce_rows<<<blocks, 256, 0, st>>>(make_view(c, fused), peer_layout((size_t)P, S, c->world), peer_row_width(S) / 4,
                                             tile0, tile1, fused ? 1 : 0);
  return check_cuda(cudaGetLastError(), "k_peer_reduce_rows launch");
}

// ---- summed packed rows -> the dense gradient tensors autograd returns (local) ------------------------------------------
__global__ void __launch_bounds__(256) k_peer_unpack(PeerView pv, int fused, int P, int S, int rw, int prezeroed,
                                                     const float* rows, const uint32_t* bits, float* __restrict__ d_means3D,
                                                     float* __restrict__ d_means2D, float* __restrict__ d_scales,
                                                     float* __restrict__ d_rot, float* __restrict__ d_opacity,
                                                     float* __restrict__ d_features) {
  if (fused) {  // in-kernel barrier: every owner's sums of this step have arrived in my result area
    peer_resolve_step(pv);
    peer_wait_flags(pv, PEER_SLOT_SUMMED);
  }
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 r0 = zero4, r1 = zero4, r2 = zero4, f[3] = {zero4, zero4, zero4};
  bool set = (bits[i >> 5] >> (i & 31)) & 1u;
  const bool failed = peer_error(pv);  // a rank missed a barrier: the sums are undefined -> NaN, never silently wrong
  if (failed) set = true;
  if (!set && prezeroed) return;  // the dense outputs were zero-filled under the backward compositor
  if (failed) {
    const float qnan = __int_as_float(0x7fc00000);
    r0 = r1 = r2 = f[0] = f[1] = f[2] = make_float4(qnan, qnan, qnan, qnan);
  } else if (set) {
    const float4* r = reinterpret_cast<const float4*>(rows + (size_t)i * rw);
    r0 = r[0]; r1 = r[1]; r2 = r[2];
    for (int k = 0; k < rw / 4 - 3; ++k) f[k] = r[3 + k];
  }
  reinterpret_cast<float4*>(d_means2D)[i] = make_float4(r0.x, r0.y, 0.f, 0.f);
  d_scales[3 * (size_t)i] = r0.z; d_scales[3 * (size_t)i + 1] = r0.w; d_scales[3 * (size_t)i + 2] = 0.f;
  reinterpret_cast<float4*>(d_rot)[i] = r1;
  d_means3D[3 * (size_t)i] = r2.x; d_means3D[3 * (size_t)i + 1] = r2.y; d_means3D[3 * (size_t)i + 2] = r2.z;
  d_opacity[i] = r2.w;
  for (int ch = 0; ch < S; ++ch) {
    const float4 v = f[ch >> 2];
    d_features[(size_t)i * S + ch] = (ch & 3) == 0 ? v.x : (ch & 3) == 1 ? v.y : (ch & 3) == 2 ? v.z : v.w;
  }
}

int launch_peer_unpack(const gsl_peer_ctx* c, int P, int S, bool prezeroed, const gsl_bwd_outputs& out, cudaStream_t st,
                       bool fused) {
  if (P == 0) return 0;
  const char* own = (const char*)c->buf[c->rank];
  const PeerLayout pl = peer_layout((size_t)P, S, c->world);
  k_peer_unpack<<<(P + 255) / 256, 256, 0, st>>>(make_view(c, fused), fused ? 1 : 0, P, S, peer_row_width(S), prezeroed ? 1 : 0,
                                                 reinterpret_cast<const float*>(own + pl.off_rows),
                                                 reinterpret_cast<const uint32_t*>(own + pl.off_rowbits), out.dL_dmeans3D,
                                                 out.dL_dmeans2D, out.dL_dscales, out.dL_drotations, out.dL_dopacity,
                                                 out.dL_dfeatures);
  return check_cuda(cudaGetLastError(), "k_peer_unpack launch");
}

}  // namespace gsl
