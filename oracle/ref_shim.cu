// oracle/ref_shim.cu -- TEST INFRASTRUCTURE, not product code.
//
// A thin C-ABI shim around the UNMODIFIED reference rasterizer
// (CudaRasterizer::Rasterizer, /root/reference/diff-gaussian-rasterization-2d/
// cuda_rasterizer/rasterizer.h:20-104).  The reference sources are compiled where
// they lie under /root/reference by oracle/build_ref.sh; nothing of them is copied
// into this repo.  The output (oracle/_ref/libgslidar_ref.so) is git-ignored and
// is used only by tests/, __graft_entry__.smoke() and bench.py (--impl reference /
// the parity gate) as the checker and the GPU baseline.
//
// What the shim adds on top of the reference (all of it host-side plumbing that the
// reference does in its torch binding, rasterize_points.cu:35-247):
//   * grow-only cudaMalloc'd geometry / binning / image chunks handed to the
//     std::function<char*(size_t)> callbacks (rasterize_points.cu:25-33 does this
//     with tensor.resize_);
//   * optional zero fill of outputs / gradient buffers (rasterize_points.cu:77-82,
//     186-197 do this with torch::full / torch::zeros);
//   * a decoder that exposes the internal GeometryState / BinningState / ImageState
//     arrays (rasterizer_impl.h:26-63) so parity tests can compare tile keys, sorted
//     ids and tile ranges bit-for-bit.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <cuda_runtime.h>
#include "cuda_rasterizer/config.h"
#include "cuda_rasterizer/rasterizer.h"
#include "cuda_rasterizer/rasterizer_impl.h"

namespace {
struct RefBuf {
  char* p = nullptr;
  size_t cap = 0;
  char* get(size_t n) {
    if (n > cap) {
      if (p) cudaFree(p);
      size_t c = n + n / 4 + 4096;
      if (cudaMalloc(&p, c) != cudaSuccess) { p = nullptr; cap = 0; return nullptr; }
      cap = c;
    }
    return p;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct RefHandle {
  RefBuf geom, bin, img;
  // scratch gradient buffers the torch binding allocates internally
  RefBuf dtrans, dnormals;
  int P = 0, R = 0, N = 0;
};
}  // namespace

extern "C" {

void* gslref_create() { return new RefHandle(); }

void gslref_destroy(void* h) {
  RefHandle* s = (RefHandle*)h;
  if (!s) return;
  s->geom.release(); s->bin.release(); s->img.release();
  s->dtrans.release(); s->dnormals.release();
  delete s;
}

// Mirrors RasterizeGaussiansCUDA (rasterize_points.cu:35-139).  All pointers are
// device pointers except none; returns num_rendered (>=0) or -1 on failure.
int gslref_forward(void* h, int P, int S, int D, int M, const float* bg, int W, int H,
                   const float* means3D, const float* shs, const float* colors_precomp,
                   const float* features, const float* opacities, const float* scales,
                   float scale_modifier, const float* rotations, const float* cov3D_precomp,
                   const bool* mask, const float* viewmatrix, const float* projmatrix,
                   const float* campos, float tanfovx, float tanfovy, int prefiltered,
                   int* out_contrib, float* out_color, float* out_feature, float* out_depth,
                   float* out_T, int* radii, int debug, float vfov_min, float vfov_max,
                   float hfov_min, float hfov_max, float scale_factor, int zero_fill) {
  RefHandle* s = (RefHandle*)h;
  const size_t N = (size_t)W * H;
  if (zero_fill) {
    cudaMemsetAsync(out_contrib, 0, 2 * N * sizeof(int));
    cudaMemsetAsync(out_color, 0, NUM_CHANNELS * N * sizeof(float));
    cudaMemsetAsync(out_feature, 0, (S + 3) * N * sizeof(float));
    cudaMemsetAsync(out_depth, 0, 4 * N * sizeof(float));
    cudaMemsetAsync(out_T, 0, N * sizeof(float));
    if (P > 0) cudaMemsetAsync(radii, 0, (size_t)P * sizeof(int));
  }
  int rendered = 0;
  if (P != 0) {
    std::function<char*(size_t)> gf = [s](size_t n) { return s->geom.get(n); };
    std::function<char*(size_t)> bf = [s](size_t n) { return s->bin.get(n); };
    std::function<char*(size_t)> imf = [s](size_t n) { return s->img.get(n); };
    try {
      rendered = CudaRasterizer::Rasterizer::forward(
          gf, bf, imf, P, S, D, M, bg, W, H, means3D, shs, colors_precomp, features, opacities,
          scales, scale_modifier, rotations, cov3D_precomp, mask, viewmatrix, projmatrix, campos,
          tanfovx, tanfovy, prefiltered != 0, out_contrib, out_color, out_feature, out_depth, out_T,
          radii, debug != 0, vfov_min, vfov_max, hfov_min, hfov_max, scale_factor);
    } catch (const std::exception& e) {
      fprintf(stderr, "[gslref] forward threw: %s\n", e.what());
      return -1;
    }
  }
  s->P = P; s->R = rendered; s->N = (int)N;
  return rendered;
}

// Mirrors RasterizeGaussiansBackwardCUDA (rasterize_points.cu:141-247).
int gslref_backward(void* h, int P, int S, int D, int M, int R, const float* bg, int W, int H,
                    const float* means3D, const float* shs, const float* colors_precomp,
                    const float* features, const float* scales, float scale_modifier,
                    const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                    const float* projmatrix, const float* campos, float tanfovx, float tanfovy,
                    const int* radii, const int* out_contrib, const float* dL_dpix,
                    const float* dL_depths, const float* dL_masks, const float* dL_dpix_feature,
                    float* dL_dmean2D, float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D,
                    float* dL_dcov3D, float* dL_dsh, float* dL_dfeatures, float* dL_dscale,
                    float* dL_drot, int debug, float vfov_min, float vfov_max, float hfov_min,
                    float hfov_max, float scale_factor, int zero_fill) {
  RefHandle* s = (RefHandle*)h;
  if (P == 0) return 0;
  float* dtrans = (float*)s->dtrans.get((size_t)P * 9 * sizeof(float));
  float* dnorm = (float*)s->dnormals.get((size_t)P * 3 * sizeof(float));
  if (!dtrans || !dnorm) return -1;
  // internal scratch is always zeroed (torch::zeros in the binding)
  cudaMemsetAsync(dtrans, 0, (size_t)P * 9 * sizeof(float));
  cudaMemsetAsync(dnorm, 0, (size_t)P * 3 * sizeof(float));
  if (zero_fill) {
    cudaMemsetAsync(dL_dmean3D, 0, (size_t)P * 3 * sizeof(float));
    cudaMemsetAsync(dL_dmean2D, 0, (size_t)P * 4 * sizeof(float));
    cudaMemsetAsync(dL_dcolor, 0, (size_t)P * NUM_CHANNELS * sizeof(float));
    if (S > 0) cudaMemsetAsync(dL_dfeatures, 0, (size_t)P * S * sizeof(float));
    cudaMemsetAsync(dL_dopacity, 0, (size_t)P * sizeof(float));
    cudaMemsetAsync(dL_dcov3D, 0, (size_t)P * 6 * sizeof(float));
    if (M > 0) cudaMemsetAsync(dL_dsh, 0, (size_t)P * M * NUM_CHANNELS * sizeof(float));
    cudaMemsetAsync(dL_dscale, 0, (size_t)P * 3 * sizeof(float));
    cudaMemsetAsync(dL_drot, 0, (size_t)P * 4 * sizeof(float));
  }
  try {
    CudaRasterizer::Rasterizer::backward(
        P, S, D, M, R, bg, W, H, means3D, shs, colors_precomp, features, scales, scale_modifier,
        rotations, cov3D_precomp, viewmatrix, projmatrix, campos, tanfovx, tanfovy, radii,
        s->geom.p, s->bin.p, s->img.p, out_contrib, dL_dpix, dL_depths, dL_masks, dL_dpix_feature,
        dL_dmean2D, dL_dopacity, dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dsh, dL_dfeatures, dL_dscale,
        dL_drot, dtrans, dnorm, debug != 0, vfov_min, vfov_max, hfov_min, hfov_max, scale_factor);
  } catch (const std::exception& e) {
    fprintf(stderr, "[gslref] backward threw: %s\n", e.what());
    return -1;
  }
  return 0;
}

int gslref_mark_visible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present) {
  if (P > 0) CudaRasterizer::Rasterizer::markVisible(P, means3D, viewmatrix, projmatrix, present);
  return 0;
}

// Decode the chunks of the last forward into raw device pointers
// (layout: rasterizer_impl.cu:159-208).  ptrs[] order:
//  0 depths(f32 P) 1 clamped(bool 4P) 2 internal_radii(i32 P) 3 means2D(f32 2P)
//  4 transMat(f32 9P) 5 normal_opacity(f32 4P) 6 rgb(f32 4P) 7 tiles_touched(u32 P)
//  8 point_offsets(u32 P) 9 point_list(u32 R) 10 point_list_keys(u64 R)
//  11 point_list_unsorted(u32 R) 12 point_list_keys_unsorted(u64 R)
//  13 ranges(uint2 N, first tiles used) 14 accum_alpha(f32 3N) 15 dL_dtransMat(f32 9P)
//  16 dL_dnormals(f32 3P)
int gslref_state(void* h, void** ptrs) {
  RefHandle* s = (RefHandle*)h;
  if (!s->geom.p) return -1;
  char* c = s->geom.p;
  auto g = CudaRasterizer::GeometryState::fromChunk(c, s->P);
  ptrs[0] = g.depths; ptrs[1] = g.clamped; ptrs[2] = g.internal_radii; ptrs[3] = g.means2D;
  ptrs[4] = g.transMat; ptrs[5] = g.normal_opacity; ptrs[6] = g.rgb; ptrs[7] = g.tiles_touched;
  ptrs[8] = g.point_offsets;
  for (int i = 9; i <= 12; ++i) ptrs[i] = nullptr;
  if (s->bin.p) {
    c = s->bin.p;
    auto b = CudaRasterizer::BinningState::fromChunk(c, s->R);
    ptrs[9] = b.point_list; ptrs[10] = b.point_list_keys; ptrs[11] = b.point_list_unsorted;
    ptrs[12] = b.point_list_keys_unsorted;
  }
  c = s->img.p;
  auto im = CudaRasterizer::ImageState::fromChunk(c, s->N);
  ptrs[13] = im.ranges; ptrs[14] = im.accum_alpha;
  ptrs[15] = s->dtrans.p; ptrs[16] = s->dnormals.p;
  return 0;
}

int gslref_num_channels() { return NUM_CHANNELS; }

// device-to-device copy helper so the Python side needs no cudart binding of its own
int gslref_copy(void* dst, const void* src, size_t nbytes) {
  return (int)cudaMemcpy(dst, src, nbytes, cudaMemcpyDeviceToDevice);
}

}  // extern "C"
