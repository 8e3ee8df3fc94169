// gsl_preprocess.cu -- per-surfel stages of the panoramic surfel rasterizer:
//   k_preprocess_fwd  : equirectangular projection, ray-splat transform T, 12-sample AABB, tile
//                       rect, conservative pixel box, SH -> 4-channel colour
//                       (semantics of forward.cu:173-287 incl. helpers :17-171, auxiliary.h:47-55,
//                        182-228, 276-283)
//   k_preprocess_bwd  : VJP of the above from the packed gradient accumulators to the dense
//                       per-parameter gradients (semantics of backward.cu:517-712, :17-134)
//   k_mark_visible    : pinhole frustum test (rasterizer_impl.cu:51-64, auxiliary.h:157-180)
// One thread per surfel, 256-thread blocks; every load of the per-surfel record is a float4.
#include <math.h>
#include "gsl_common.cuh"
#include <cstdlib>
#include <algorithm>
#include "gsl_math.cuh"
#include "gsl_peer.cuh"

namespace gsl {

__device__ __constant__ float kSH_C0 = 0.28209479177387814f;
__device__ __constant__ float kSH_C1 = 0.4886025119029199f;
__device__ __constant__ float kSH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                          -1.0925484305920792f, 0.5462742152960396f};
__device__ __constant__ float kSH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                          0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                          -0.5900435899266435f};

struct PreParams {
  int P, D, M, W, H, gx, gy;
  int wrap;  // GSL_FLAG_WRAP_AZIMUTH: the panorama is periodic in azimuth (opt-in; the reference has no wrap-around)
  float VFOV_min, VFOV_max, HFOV_min, HFOV_max;
  float scale_factor;
  float samp[12];  // (float)(2*MY_PI*i/12), forward.cu:155
};

struct Rot3 {  // columns c0,c1,c2 of the rotation matrix
  float r00, r01, r02, r10, r11, r12, r20, r21, r22;  // rCR : column C, row R
};

// quat_to_rotmat (auxiliary.h:206-228); q = (q0,q1,q2,q3) as stored, w = q0.
__device__ __forceinline__ Rot3 quat_to_rot(float q0, float q1, float q2, float q3) {
  float n = GSL_FF(q2, q2, GSL_FF(q1, q1, GSL_FF(q0, q0, GSL_FM(q3, q3))));
  float s = rsqrtf(n);
  float w = GSL_FM(q0, s), x = GSL_FM(q1, s), y = GSL_FM(q2, s), z = GSL_FM(q3, s);
  float wz = GSL_FM(w, z), wx = GSL_FM(w, x), wy = GSL_FM(w, y);
  float yy = GSL_FM(y, y), zz = GSL_FM(z, z);
  float xy_p = GSL_FF(x, y, wz), xy_m = GSL_FF(x, y, -wz);
  float yz_p = GSL_FF(y, z, wx), yz_m = GSL_FF(y, z, -wx);
  float xz_m = GSL_FF(x, z, -wy), xz_p = GSL_FF(x, z, wy);
  float yyzz = GSL_FA(yy, zz);
  float xxzz = GSL_FF(x, x, zz);
  float xxyy = GSL_FF(x, x, yy);
  Rot3 R;
  R.r00 = GSL_FS(1.f, GSL_FA(yyzz, yyzz));
  R.r01 = GSL_FA(xy_p, xy_p);
  R.r02 = GSL_FA(xz_m, xz_m);
  R.r10 = GSL_FA(xy_m, xy_m);
  R.r11 = GSL_FS(1.f, GSL_FA(xxzz, xxzz));
  R.r12 = GSL_FA(yz_p, yz_p);
  R.r20 = GSL_FA(xz_p, xz_p);
  R.r21 = GSL_FA(yz_m, yz_m);
  R.r22 = GSL_FS(1.f, GSL_FA(xxyy, xxyy));
  return R;
}

// The M float4 SH coefficients of one surfel: either one (P, M, 4) tensor like the reference's `shs`, or GS-LiDAR's two
// parameter tensors _features_dc (P, 1, 4) and _features_rest (P, M - 1, 4) taken as they are (the reference
// concatenates them on every call, scene/gaussian_model.py:167-171: 256 MB written and read again at 1M surfels).
struct ShView {
  const float4* c0;  // coefficient 0
  const float4* cr;  // coefficients 1 .. M-1
  __device__ __forceinline__ float4 at(int k) const { return k == 0 ? c0[0] : cr[k - 1]; }
};
struct ShOut {
  float4* c0;
  float4* cr;
  __device__ __forceinline__ void set(int k, float4 v) const { if (k == 0) c0[0] = v; else cr[k - 1] = v; }
};
__device__ __forceinline__ ShView sh_view(const float* shs, const float* shs_rest, size_t idx, int M) {
  ShView v;
  if (shs_rest) {
    v.c0 = reinterpret_cast<const float4*>(shs) + idx;
    v.cr = reinterpret_cast<const float4*>(shs_rest) + idx * (size_t)(M - 1);
  } else {
    v.c0 = reinterpret_cast<const float4*>(shs) + idx * (size_t)M;
    v.cr = v.c0 + 1;
  }
  return v;
}
__device__ __forceinline__ ShOut sh_out(float* d, float* d_rest, size_t idx, int M) {
  ShOut v;
  if (d_rest) {
    v.c0 = reinterpret_cast<float4*>(d) + idx;
    v.cr = reinterpret_cast<float4*>(d_rest) + idx * (size_t)(M - 1);
  } else {
    v.c0 = reinterpret_cast<float4*>(d) + idx * (size_t)M;
    v.cr = v.c0 + 1;
  }
  return v;
}

// SH -> 4 channels (forward.cu:17-69).
__device__ __forceinline__ float4 eval_sh(int deg, const ShView sh, float x, float y, float z) {
  float4 c0 = sh.at(0);
  float r[4] = {kSH_C0 * c0.x, kSH_C0 * c0.y, kSH_C0 * c0.z, kSH_C0 * c0.w};
#define GSL_ACC(coef, idx, sign)                         \
  {                                                      \
    float4 c = sh.at(idx);                               \
    float k = (coef);                                    \
    r[0] = r[0] sign k * c.x;                            \
    r[1] = r[1] sign k * c.y;                            \
    r[2] = r[2] sign k * c.z;                            \
    r[3] = r[3] sign k * c.w;                            \
  }
  if (deg > 0) {
    GSL_ACC(kSH_C1 * y, 1, -)
    GSL_ACC(kSH_C1 * z, 2, +)
    GSL_ACC(kSH_C1 * x, 3, -)
    if (deg > 1) {
      float xx = x * x, yy = y * y, zz = z * z;
      float xy = x * y, yz = y * z, xz = x * z;
      GSL_ACC(kSH_C2[0] * xy, 4, +)
      GSL_ACC(kSH_C2[1] * yz, 5, +)
      GSL_ACC(kSH_C2[2] * (2.0f * zz - xx - yy), 6, +)
      GSL_ACC(kSH_C2[3] * xz, 7, +)
      GSL_ACC(kSH_C2[4] * (xx - yy), 8, +)
      if (deg > 2) {
        GSL_ACC(kSH_C3[0] * y * (3.0f * xx - yy), 9, +)
        GSL_ACC(kSH_C3[1] * xy * z, 10, +)
        GSL_ACC(kSH_C3[2] * y * (4.0f * zz - xx - yy), 11, +)
        GSL_ACC(kSH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy), 12, +)
        GSL_ACC(kSH_C3[4] * x * (4.0f * zz - xx - yy), 13, +)
        GSL_ACC(kSH_C3[5] * z * (xx - yy), 14, +)
        GSL_ACC(kSH_C3[6] * x * (xx - 3.0f * yy), 15, +)
      }
    }
  }
#undef GSL_ACC
  return make_float4(r[0] + 0.5f, r[1] + 0.5f, r[2] + 0.5f, r[3] + 0.5f);
}

// One sample of the 12-point AABB (forward.cu:139-171): pixel coordinates of the point (vx, vy) of the splat plane,
// with exactly the reference's rounding sequence.
__device__ __forceinline__ float2 aabb_sample_ref(const Splat& s, float vx, float vy, float hfov_min, float vfov_min,
                                                  float Wf, float Hf, float dH, float dV) {
  float X = GSL_FA(s.Tuz, GSL_FF(s.Tux, vx, GSL_FM(s.Tuy, vy)));
  float Y = GSL_FA(s.Tvz, GSL_FF(s.Tvx, vx, GSL_FM(s.Tvy, vy)));
  float Z = GSL_FA(s.Twz, GSL_FF(s.Twx, vx, GSL_FM(s.Twy, vy)));
  float ph = atan2f(X, Z);
  float th = atan2f(sqrtf(GSL_FF(X, X, GSL_FM(Z, Z))), -Y);
  return make_float2(GSL_FD(GSL_FM(GSL_FS(ph, hfov_min), Wf), dH), GSL_FD(GSL_FM(GSL_FS(th, vfov_min), Hf), dV));
}

// atan(t) for 0 <= t <= 0.25 (odd series to t^11: truncation 5e-9 relative)
__device__ __forceinline__ float atan_small(float t) {
  const float q = t * t;
  float p = fmaf(q, -0.09090909f, 0.11111111f);
  p = fmaf(q, p, -0.14285714f);
  p = fmaf(q, p, 0.2f);
  p = fmaf(q, p, -0.33333333f);
  return fmaf(t * q, p, t);
}

// Filtered AABB radius.  Only ceil(rad) and the test rad < 0.3 consume the 24 atan2f of the reference's 12-sample AABB
// (forward.cu:139-171, :243-264), so the radius is first computed in a form that needs two short arctangent series:
// the azimuth / elevation offsets of a sample RELATIVE to the splat centre are atan(c / d) with
//     azimuth:   c = dX tz - dZ tx,          d = rxz^2 + dX tx + dZ tz           (dX, dY, dZ = sample - centre)
//     elevation: c = dY rxz - drho ty,       d = Y ty + rho rxz,                 drho = rho - rxz without cancellation
// atan is odd and monotone, so max(max_i atan(t_i), -min_i atan(t_i)) = atan(max(t_max, -t_min)): ONE series per axis.
// The result differs from the reference's float value by at most `eps` pixels -- the reference's own rounding (atan2f
// 2 ulp of pi, the subtraction of the fov origin, one multiplication and one division at magnitude <= W: 2.4e-4 px at
// W = 1030 for each of the two coordinates it subtracts) plus 2e-4 for this form -- so whenever [rad - eps, rad + eps]
// contains neither an integer nor 0.3 the radius and the culling decision ARE the reference's.  Returns false when that
// cannot be certified (band hit, splat wider than atan(0.25) = 14 degrees, sample behind the centre's meridian plane,
// azimuth range touching the +-pi seam, non-finite values): the caller then evaluates the reference sequence itself.
__device__ __forceinline__ bool aabb_radius_filtered(const Splat& s, float cutoff, const float* s_sin, const float* s_cos,
                                                     float rxz, float phi, float kx, float ky, float eps,
                                                     int& radius_out, bool& small_out) {
  const float tx = s.Tuz, ty = s.Tvz, tz = s.Twz;
  const float rxz2 = rxz * rxz;
  float tmax = -1e30f, tmin = 1e30f, umax = -1e30f, umin = 1e30f;
  float slack = 1e30f;  // min over the samples of (d - 4|c|) of both axes, and of rho^2: all must be > 1e-30
#pragma unroll 4
  for (int i = 0; i < 12; ++i) {
    const float vx = s_sin[i] * cutoff, vy = s_cos[i] * cutoff;
    const float dX = fmaf(s.Tux, vx, s.Tuy * vy), dY = fmaf(s.Tvx, vx, s.Tvy * vy), dZ = fmaf(s.Twx, vx, s.Twy * vy);
    const float X = tx + dX, Y = ty + dY, Z = tz + dZ;
    const float c = fmaf(dX, tz, -dZ * tx);
    const float along = fmaf(dX, tx, dZ * tz);
    const float d = rxz2 + along;
    const float t = c * fast_rcp(d);
    tmax = fmaxf(tmax, t); tmin = fminf(tmin, t);
    const float rho2 = fmaf(X, X, Z * Z);
    const float rho = rho2 * rsqrtf(rho2);
    const float drho = fmaf(2.f, along, fmaf(dX, dX, dZ * dZ)) * fast_rcp(rho + rxz);
    const float c2 = fmaf(dY, rxz, -drho * ty);
    const float d2 = fmaf(Y, ty, rho * rxz);
    const float u = c2 * fast_rcp(d2);
    umax = fmaxf(umax, u); umin = fminf(umin, u);
    slack = fminf(slack, fminf(fmaf(-4.f, fabsf(c), d), fmaf(-4.f, fabsf(c2), d2)));
    slack = fminf(slack, rho2);
  }
  const float ax = atan_small(fmaxf(tmax, -tmin)), ay = atan_small(fmaxf(umax, -umin));
  const float rad = fmaxf(ax * kx, ay * ky);
  const float lo = rad - eps, hi = rad + eps;
  // NaN anywhere makes a comparison false -> not certified
  const bool base = (slack > 1e-30f) && (fabsf(phi) + ax < 3.1414f) && (rad < 1e6f);  // (denominators are normal numbers)
  small_out = hi < 0.2999f;                                         // certainly culled by `radii < 0.3`
  const bool big = (lo > 0.3001f) && (ceilf(lo) == ceilf(hi));     // certainly kept, with a certain ceil
  radius_out = (int)ceilf(hi);
  return base && (small_out || big);
}

#ifdef GSL_PFWD_MINB
#define GSL_PFWD_BOUNDS __launch_bounds__(256, GSL_PFWD_MINB)
#else
#define GSL_PFWD_BOUNDS __launch_bounds__(256)
#endif
__global__ void GSL_PFWD_BOUNDS k_preprocess_fwd(
    PreParams pp, const float* __restrict__ means3D, const float* __restrict__ scales,
    const float* __restrict__ rotations, const float* __restrict__ opacities,
    const float* __restrict__ shs, const float* __restrict__ shs_rest, const float* __restrict__ colors_precomp,
    const uint8_t* __restrict__ mask, const float* __restrict__ viewmatrix,
    const float* __restrict__ campos, int* __restrict__ radii, float4* __restrict__ rec,
    float4* __restrict__ rgb, ushort4* __restrict__ rect, short4* __restrict__ pixbox,
    uint32_t* __restrict__ tiles, uint8_t* __restrict__ clamped) {
  __shared__ float s_sin[12], s_cos[12];
  if (threadIdx.x < 12) {
    s_sin[threadIdx.x] = sinf(pp.samp[threadIdx.x]);
    s_cos[threadIdx.x] = cosf(pp.samp[threadIdx.x]);
  }
  __syncthreads();
  const int gidx = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = gidx < pp.P;          // the whole warp stays: the exact AABB fallback below is warp-cooperative
  const int idx = live ? gidx : pp.P - 1;  // (loads of a dead lane read the last surfel; nothing is stored for it)

  // Invisible unless proven otherwise (forward.cu:214-215).
  int out_radius = 0;
  uint32_t out_tiles = 0;
  ushort4 out_rect = make_ushort4(0, 0, 0, 0);
  short4 out_box = make_short4(0, 1, 0, 0);  // y0 > y1 : empty

  const float vm0 = viewmatrix[0], vm1 = viewmatrix[1], vm2 = viewmatrix[2];
  const float vm4 = viewmatrix[4], vm5 = viewmatrix[5], vm6 = viewmatrix[6];
  const float vm8 = viewmatrix[8], vm9 = viewmatrix[9], vm10 = viewmatrix[10];
  const float vm12 = viewmatrix[12], vm13 = viewmatrix[13], vm14 = viewmatrix[14];

  const float px = means3D[3 * idx], py = means3D[3 * idx + 1], pz = means3D[3 * idx + 2];
  const float opacity = opacities[idx];

  // view-space centre, polar coordinates (forward.cu:116-125, auxiliary.h:77-85)
  const float tx = GSL_FA(vm12, dot3_ref(px, vm0, py, vm4, pz, vm8));
  const float ty = GSL_FA(vm13, dot3_ref(px, vm1, py, vm5, pz, vm9));
  const float tz = GSL_FA(vm14, dot3_ref(px, vm2, py, vm6, pz, vm10));
  const float phi = atan2f(tx, tz);
  const float tx2 = GSL_FM(tx, tx), tz2 = GSL_FM(tz, tz);
  const float rxz = sqrtf(GSL_FA(tx2, tz2));
  const float theta = atan2f(rxz, -ty);
  const float r = sqrtf(GSL_FA(GSL_FF(ty, ty, tx2), tz2));

  bool visible = live && mask[idx] != 0;
  const float dV = GSL_FS(pp.VFOV_max, pp.VFOV_min);
  const float dH = GSL_FS(pp.HFOV_max, pp.HFOV_min);
  if (visible) {
    // in_frustum_panorama (auxiliary.h:182-204); the 1.3 compares are in double.
    float center_v = GSL_FM(GSL_FA(pp.VFOV_max, pp.VFOV_min), 0.5f);
    float ratio_v = fabsf(GSL_FD(GSL_FS(theta, center_v), GSL_FM(dV, 0.5f)));
    float center_h = GSL_FM(GSL_FA(pp.HFOV_min, pp.HFOV_max), 0.5f);
    float ratio_h = fabsf(GSL_FD(GSL_FS(phi, center_h), GSL_FM(dH, 0.5f)));
    float near_ = GSL_FA(pp.scale_factor, pp.scale_factor);
    if (r <= near_ || (double)ratio_v > 1.3 || (double)ratio_h > 1.3) visible = false;
  }

  // state of a visible surfel that survives the (warp-uniform) AABB stage below
  Splat s;
  float nx = 0.f, ny = 0.f, nz = 0.f, cutoff = 0.f, cx = 0.f, cy = 0.f;
  float4* my = rec + 4 * (size_t)idx;
  const float Wf = (float)pp.W, Hf = (float)pp.H;
  s.Tux = s.Tuy = s.Tuz = s.Tvx = s.Tvy = s.Tvz = s.Twx = s.Twy = s.Twz = 0.f;
  if (visible) {
    const float sx = scales[3 * idx], sy = scales[3 * idx + 1];
    const float4 q = reinterpret_cast<const float4*>(rotations)[idx];
    const Rot3 R = quat_to_rot(q.x, q.y, q.z, q.w);
    // L = R * diag(sx, sy, 1)  (scale.z and scale_modifier are ignored: auxiliary.h:276-283)
    const float l0x = GSL_FM(sx, R.r00), l0y = GSL_FM(sx, R.r01), l0z = GSL_FM(sx, R.r02);
    const float l1x = GSL_FM(sy, R.r10), l1y = GSL_FM(sy, R.r11), l1z = GSL_FM(sy, R.r12);
    // T rows (forward.cu:84-105): Tu = x-coefficients, Tv = y, Tw = z over splat coords (u,v,1)
    s.Tux = dot3_ref(l0x, vm0, l0y, vm4, l0z, vm8);
    s.Tuy = dot3_ref(l1x, vm0, l1y, vm4, l1z, vm8);
    s.Tuz = tx;
    s.Tvx = dot3_ref(l0x, vm1, l0y, vm5, l0z, vm9);
    s.Tvy = dot3_ref(l1x, vm1, l1y, vm5, l1z, vm9);
    s.Tvz = ty;
    s.Twx = dot3_ref(l0x, vm2, l0y, vm6, l0z, vm10);
    s.Twy = dot3_ref(l1x, vm2, l1y, vm6, l1z, vm10);
    s.Twz = tz;
    // view-space normal, flipped towards the sensor (forward.cu:106-112)
    nx = dot3_ref(vm0, R.r20, vm4, R.r21, vm8, R.r22);
    ny = dot3_ref(vm1, R.r20, vm5, R.r21, vm9, R.r22);
    nz = dot3_ref(vm2, R.r20, vm6, R.r21, vm10, R.r22);
    float ndot = dot3_ref(nx, tx, ny, ty, nz, tz);
    float mult = ndot < 0.f ? 1.f : -1.f;
    nx = GSL_FM(nx, mult); ny = GSL_FM(ny, mult); nz = GSL_FM(nz, mult);

    // The reference stores T before any further culling (forward.cu:238-241).
    my[0] = make_float4(s.Tux, s.Tuy, s.Tuz, s.Tvx);
    my[1] = make_float4(s.Tvy, s.Tvz, s.Twx, s.Twy);

    cutoff = sqrtf((float)fmax((double)GSL_FF(logf(opacity), 2.f, 9.f), 0.000001));
    cx = GSL_FD(GSL_FM(GSL_FS(phi, pp.HFOV_min), Wf), dH);
    cy = GSL_FD(GSL_FM(GSL_FS(theta, pp.VFOV_min), Hf), dV);
  }

  // ---- AABB radius (forward.cu:139-171, :243-257): my_radius = ceil(rad), culled if rad < 0.3
  int my_radius = 0;
  bool keep = false;
  if (pp.wrap) {
    // wrap-around mode (opt-in, not the reference's semantics): azimuth of a sample RELATIVE to the centre
    if (visible) {
      float minx = INFINITY, miny = INFINITY, maxx = -INFINITY, maxy = -INFINITY;
#pragma unroll 1
      for (int i = 0; i < 12; ++i) {
        const float vx = GSL_FM(s_sin[i], cutoff), vy = GSL_FM(s_cos[i], cutoff);
        const float X = GSL_FA(s.Tuz, GSL_FF(s.Tux, vx, GSL_FM(s.Tuy, vy)));
        const float Z = GSL_FA(s.Twz, GSL_FF(s.Twx, vx, GSL_FM(s.Twy, vy)));
        const float2 pxy = aabb_sample_ref(s, vx, vy, pp.HFOV_min, pp.VFOV_min, Wf, Hf, dH, dV);
        float dphi = atan2f(X, Z) - phi;
        if (dphi > 3.14159265f) dphi -= 6.2831853f;
        else if (dphi < -3.14159265f) dphi += 6.2831853f;
        const float ppx = cx + dphi * Wf / dH;
        minx = fminf(minx, ppx); maxx = fmaxf(maxx, ppx);
        miny = fminf(miny, pxy.y); maxy = fmaxf(maxy, pxy.y);
      }
      const float rad = fmaxf(fmaxf(GSL_FS(maxx, cx), GSL_FS(cx, minx)), fmaxf(GSL_FS(maxy, cy), GSL_FS(cy, miny)));
      keep = !((double)rad < 0.3);
      my_radius = (int)ceilf(rad);
    }
  } else {
    // 1. filtered form (two short arctangent series); certified for all but ~1 % of the surfels
    bool exact = false;
    if (visible) {
      bool small = false;
      const float kx = Wf / dH, ky = Hf / dV;
      const float eps = 3.0f * (7.2e-7f * fmaxf(kx, ky) + 1.2e-7f * fmaxf(Wf, Hf)) + 2e-4f;
      if (aabb_radius_filtered(s, cutoff, s_sin, s_cos, rxz, phi, kx, ky, eps, my_radius, small)) keep = !small;
      else exact = true;
    }
    // 2. the rest (band hits, splats on the +-pi seam, very large or degenerate ones): the reference sequence itself, one
    //    surfel at a time with its 12 samples spread over 12 lanes; min / max are order-independent (fminf / fmaxf drop NaN
    //    operands exactly like the reference's running min / max, seeded with +-infinity)
    uint32_t todo = __ballot_sync(0xffffffffu, exact);
    const int lane = threadIdx.x & 31;
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      Splat q;
      q.Tux = __shfl_sync(0xffffffffu, s.Tux, src); q.Tuy = __shfl_sync(0xffffffffu, s.Tuy, src);
      q.Tuz = __shfl_sync(0xffffffffu, s.Tuz, src); q.Tvx = __shfl_sync(0xffffffffu, s.Tvx, src);
      q.Tvy = __shfl_sync(0xffffffffu, s.Tvy, src); q.Tvz = __shfl_sync(0xffffffffu, s.Tvz, src);
      q.Twx = __shfl_sync(0xffffffffu, s.Twx, src); q.Twy = __shfl_sync(0xffffffffu, s.Twy, src);
      q.Twz = __shfl_sync(0xffffffffu, s.Twz, src);
      const float qcut = __shfl_sync(0xffffffffu, cutoff, src);
      float minx = INFINITY, miny = INFINITY, maxx = -INFINITY, maxy = -INFINITY;
      if (lane < 12) {
        const float2 pxy = aabb_sample_ref(q, GSL_FM(s_sin[lane], qcut), GSL_FM(s_cos[lane], qcut), pp.HFOV_min, pp.VFOV_min,
                                           Wf, Hf, dH, dV);
        minx = fminf(minx, pxy.x); maxx = fmaxf(maxx, pxy.x);
        miny = fminf(miny, pxy.y); maxy = fmaxf(maxy, pxy.y);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {  // lanes 0..15 hold the samples (12..15: the seeds)
        minx = fminf(minx, __shfl_xor_sync(0xffffffffu, minx, o)); maxx = fmaxf(maxx, __shfl_xor_sync(0xffffffffu, maxx, o));
        miny = fminf(miny, __shfl_xor_sync(0xffffffffu, miny, o)); maxy = fmaxf(maxy, __shfl_xor_sync(0xffffffffu, maxy, o));
      }
      minx = __shfl_sync(0xffffffffu, minx, 0); maxx = __shfl_sync(0xffffffffu, maxx, 0);
      miny = __shfl_sync(0xffffffffu, miny, 0); maxy = __shfl_sync(0xffffffffu, maxy, 0);
      if (lane == src) {
        const float rad = fmaxf(fmaxf(GSL_FS(maxx, cx), GSL_FS(cx, minx)), fmaxf(GSL_FS(maxy, cy), GSL_FS(cy, miny)));
        keep = !((double)rad < 0.3);
        my_radius = (int)ceilf(rad);
      }
    }
  }

  if (visible && keep) {
    {
      const float rf = (float)my_radius;
      // getRect (auxiliary.h:47-55)
      int rminx = min(pp.gx, max(0, (int)(GSL_FM(GSL_FS(cx, rf), 0.0625f))));
      int rminy = min(pp.gy, max(0, (int)(GSL_FM(GSL_FS(cy, rf), 0.0625f))));
      int rmaxx = min(pp.gx, max(0, (int)(GSL_FM(GSL_FS(GSL_FA(GSL_FA(cx, rf), 16.f), 1.f), 0.0625f))));
      int rmaxy = min(pp.gy, max(0, (int)(GSL_FM(GSL_FS(GSL_FA(GSL_FA(cy, rf), 16.f), 1.f), 0.0625f))));
      if (pp.wrap) {
        // Tile columns as a modular range [rminx, rmaxx) over gx columns, rmaxx may exceed gx: column = x mod gx.
        // The part of the footprint left of pixel 0 continues at pixel W, the part right of pixel W - 1 at pixel 0
        // (same rounding rules as getRect for each part).
        const int tx0 = (int)floorf((cx - rf) * 0.0625f), tx1 = (int)floorf((cx + rf + 15.f) * 0.0625f);
        const bool left = (cx - rf) < 0.f, right = (cx + rf) >= Wf;
        int start = max(tx0, 0), len = min(tx1, pp.gx) - max(tx0, 0);
        if (left) {
          const int a0 = min(pp.gx, max(0, (int)floorf((Wf + cx - rf) * 0.0625f)));
          start = a0;
          len += pp.gx - a0;
        }
        if (right) len += min(pp.gx, max(0, (int)floorf((cx + rf - Wf + 15.f) * 0.0625f)));
        if ((left && right) || len >= pp.gx) { start = 0; len = pp.gx; }
        rminx = start;
        rmaxx = start + max(len, 0);
      }
      const int area = (rmaxx - rminx) * (rmaxy - rminy);
      if (area != 0) {
        // colour: SH or precomputed (forward.cu:269-279)
        if (colors_precomp == nullptr) {
          float dx = px - campos[0], dy = py - campos[1], dz = pz - campos[2];
          float len = sqrtf(dx * dx + dy * dy + dz * dz);
          dx = dx / len; dy = dy / len; dz = dz / len;
          float4 c = eval_sh(pp.D, sh_view(shs, shs_rest, (size_t)idx, pp.M), dx, dy, dz);
          uint8_t cl = (uint8_t)((c.x < 0.f ? 1 : 0) | (c.y < 0.f ? 2 : 0) | (c.z < 0.f ? 4 : 0) | (c.w < 0.f ? 8 : 0));
          clamped[idx] = cl;
          rgb[idx] = make_float4(fmaxf(c.x, 0.f), fmaxf(c.y, 0.f), fmaxf(c.z, 0.f), fmaxf(c.w, 0.f));
        }
        my[2] = make_float4(s.Twz, cx, cy, opacity);
        my[3] = make_float4(nx, ny, nz, r);
        out_radius = my_radius;
        out_tiles = (uint32_t)area;
        out_rect = make_ushort4((unsigned short)rminx, (unsigned short)rminy, (unsigned short)rmaxx,
                                (unsigned short)rmaxy);

        // ---- conservative pixel box of the pair-level support (this design, not in the reference).
        // A pair can only contribute if alpha >= 1/255  <=>  min(rho3d, rho2d) <= tau := 2 ln(255 o).
        //  * rho2d <= tau : |pixel - means2D| <= sqrt(tau/2) (plain pixel distance, no wrap);
        //  * rho3d <= tau with a positive in-range depth: the pixel's ray meets the splat plane at a point
        //    p = t + u a + v b of the ellipse E = {u^2 + v^2 <= tau} in front of the sensor, so the pixel's
        //    (phi, theta) is the direction of some p in E.  With q = p - t written in the local spherical
        //    frame (r^, phi^, theta^) at t and the extents E_x = sqrt(tau ((a.x^)^2 + (b.x^)^2)) of E
        //    along each frame axis:
        //       dphi   = atan2(q_phi, (r + q_r) sin th + q_th cos th)            (exact)
        //       dtheta = atan2(q_th, r + q_r) + xi,  |xi| <= (q_phi / |(r + q_r, q_th)|)^2 / (2 sin(th + .))
        //    which gives the half-widths below.  Where the local bound degenerates (huge or very close
        //    splats, poles) the spherical-cap bound asin(D / r), D = sqrt(tau lambda_max), is used, and
        //    failing that the whole image.  All bounds are inflated by safety margins.
        int bx0 = 0, bx1 = pp.W - 1, by0 = 0, by1 = pp.H - 1;
        float tau = 2.f * logf(255.f * opacity);
        float tauc = tau + 1e-3f * fabsf(tau) + 1e-3f;
        if (tauc == tauc && fabsf(tauc) < 1e30f) {
          if (tauc <= 0.f) {
            by0 = 1; by1 = 0;  // can never reach alpha >= 1/255
          } else {
            float A = s.Tux * s.Tux + s.Tvx * s.Tvx + s.Twx * s.Twx;
            float B = s.Tuy * s.Tuy + s.Tvy * s.Tvy + s.Twy * s.Twy;
            float C = s.Tux * s.Tuy + s.Tvx * s.Tvy + s.Twx * s.Twy;
            float hd = 0.5f * (A - B);
            float lam = 0.5f * (A + B) + sqrtf(hd * hd + C * C);
            float Dm = sqrtf(tauc * lam) * 1.001f;
            float rd = sqrtf(0.5f * tauc) * 1.001f;
            float sd = Dm / r;
            float sth = rxz / r;  // sin of the polar angle of the centre
            float hx = -1.f, hy = -1.f;  // negative: unbounded
            // spherical-cap bound
            if (sd == sd && sd < 0.9f) {
              float delta = asinf(sd) * 1.001f + 1e-6f;
              hy = delta * Hf / dV;
              if (sd < 0.9f * sth) {
                float dphi = asinf(sd / sth) * 1.001f + 1e-6f;
                hx = dphi * Wf / dH;
              }
            }
            // local-frame bound (tighter for anisotropic / foreshortened splats)
            {
              const float inv_r = 1.f / r, inv_rxz = 1.f / rxz;
              const float cth = -ty * inv_r;
              const float phx = tz * inv_rxz, phz = -tx * inv_rxz;                      // phi^
              const float thx = cth * tx * inv_rxz, thy = sth, thz = cth * tz * inv_rxz;  // theta^
              const float rhx = tx * inv_r, rhy = ty * inv_r, rhz = tz * inv_r;          // r^
              const float ap = s.Tux * phx + s.Twx * phz, bp = s.Tuy * phx + s.Twy * phz;
              const float at = s.Tux * thx + s.Tvx * thy + s.Twx * thz, bt = s.Tuy * thx + s.Tvy * thy + s.Twy * thz;
              const float ar = s.Tux * rhx + s.Tvx * rhy + s.Twx * rhz, br = s.Tuy * rhx + s.Tvy * rhy + s.Twy * rhz;
              const float sq = sqrtf(tauc) * 1.002f;
              const float Eph = sq * sqrtf(ap * ap + bp * bp), Eth = sq * sqrtf(at * at + bt * bt);
              const float Er = sq * sqrtf(ar * ar + br * br);
              const float rr = r - Er;
              if (rr > 0.25f * r) {
                const float rho_min = r * sth - (Er * sth + Eth * fabsf(cth));
                if (rho_min > 0.1f * rxz) {
                  const float t_hx = (atanf(Eph / rho_min) * 1.001f + 1e-6f) * Wf / dH;
                  if (t_hx >= 0.f && (hx < 0.f || t_hx < hx)) hx = t_hx;
                }
                const float d1 = Eth / rr, smin = sth - d1;
                if (smin > 0.1f) {
                  const float ep = Eph / rr;
                  const float t_hy = ((d1 + 0.5f * ep * ep / smin) * 1.001f + 1e-6f) * Hf / dV;
                  if (t_hy >= 0.f && (hy < 0.f || t_hy < hy)) hy = t_hy;
                }
              }
            }
            if (hx >= 0.f) hx = fmaxf(hx, rd) + 0.02f;
            if (hy >= 0.f) hy = fmaxf(hy, rd) + 0.02f;
            // pixel centres sit at integer coordinates: pixel x is reachable iff |x - cx| <= hx
            if (hy >= 0.f && hy < 30000.f) {
              by0 = max(0, (int)ceilf(cy - hy));
              by1 = min(pp.H - 1, (int)floorf(cy + hy));
              if (by0 > by1) { by0 = 1; by1 = 0; }
            }
            if (hx >= 0.f && hx < 30000.f && !(by0 > by1)) {
              // azimuth is periodic with Wp pixels; collect the images that intersect the picture
              float Wp = 6.2831853071795864f * Wf / dH;
              int lo[3], hi[3], n = 0;
#pragma unroll
              for (int k = -1; k <= 1; ++k) {
                float c = cx + (float)k * Wp;
                int a = max(0, (int)ceilf(c - hx)), b = min(pp.W - 1, (int)floorf(c + hx));
                if (a <= b) { lo[n] = a; hi[n] = b; ++n; }
              }
              if (n == 0) { by0 = 1; by1 = 0; }
              else if (n == 1) { bx0 = lo[0]; bx1 = hi[0]; }
              else if (n == 2 && lo[0] == 0 && hi[1] == pp.W - 1 && hi[0] + 1 < lo[1]) {
                bx0 = lo[1]; bx1 = hi[0];  // wrapped: x >= bx0 || x <= bx1
              }
            }
          }
        }
        out_box = make_short4((short)bx0, (short)by0, (short)bx1, (short)by1);
      }
    }
  }
  if (live) {
    radii[idx] = out_radius;
    tiles[idx] = out_tiles;
    rect[idx] = out_rect;
    pixbox[idx] = out_box;
  }
}

// Keys of the surfel depth sort: bits of the view-space range r, computed with exactly the
// instruction sequence of k_preprocess_fwd (forward.cu:116-125) so that the order is the order of the stored
// depths.  Culled surfels get a key too -- they emit no instances, so where they sort does not matter.  Kept
// separate from k_preprocess_fwd so that the (latency-bound) sort of gsl_sort.cu can run on a side stream UNDER the
// (issue-bound) preprocess kernel.
__device__ __forceinline__ void depth_key_of(int idx, const float* __restrict__ means3D, const float* __restrict__ viewmatrix,
                                             uint32_t* __restrict__ skey, uint32_t& key);
__global__ void __launch_bounds__(256) k_depth_keys(int P, const float* __restrict__ means3D,
                                                    const float* __restrict__ viewmatrix, uint32_t* __restrict__ skey,
                                                    uint32_t* __restrict__ ctrl, uint32_t* __restrict__ hist, uint32_t nb) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  for (uint32_t b = (uint32_t)idx; b < nb; b += gridDim.x * blockDim.x) hist[b] = 0u;  // the sort's histogram (gsl_sort.cu)
  uint32_t key = 0;
  const bool live = idx < P;
  if (live) depth_key_of(idx, means3D, viewmatrix, skey, key);
  // key range for the bucket sort: ctrl[8] = max(~key), ctrl[9] = max(key), one pair of atomics per CTA
  __shared__ uint32_t s_max[2];
  if (threadIdx.x < 2) s_max[threadIdx.x] = 0u;
  __syncthreads();
  const uint32_t kmax = __reduce_max_sync(0xffffffffu, live ? key : 0u);
  const uint32_t nmin = __reduce_max_sync(0xffffffffu, live ? ~key : 0u);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&s_max[0], nmin);
    atomicMax(&s_max[1], kmax);
  }
  __syncthreads();
  if (threadIdx.x < 2) atomicMax(&ctrl[8 + threadIdx.x], s_max[threadIdx.x]);
}

__device__ __forceinline__ void depth_key_of(int idx, const float* __restrict__ means3D, const float* __restrict__ viewmatrix,
                                             uint32_t* __restrict__ skey, uint32_t& key) {
  const float vm0 = viewmatrix[0], vm1 = viewmatrix[1], vm2 = viewmatrix[2];
  const float vm4 = viewmatrix[4], vm5 = viewmatrix[5], vm6 = viewmatrix[6];
  const float vm8 = viewmatrix[8], vm9 = viewmatrix[9], vm10 = viewmatrix[10];
  const float vm12 = viewmatrix[12], vm13 = viewmatrix[13], vm14 = viewmatrix[14];
  const float px = means3D[3 * idx], py = means3D[3 * idx + 1], pz = means3D[3 * idx + 2];
  const float tx = GSL_FA(vm12, dot3_ref(px, vm0, py, vm4, pz, vm8));
  const float ty = GSL_FA(vm13, dot3_ref(px, vm1, py, vm5, pz, vm9));
  const float tz = GSL_FA(vm14, dot3_ref(px, vm2, py, vm6, pz, vm10));
  const float tx2 = GSL_FM(tx, tx), tz2 = GSL_FM(tz, tz);
  const float r = sqrtf(GSL_FA(GSL_FF(ty, ty, tx2), tz2));
  key = __float_as_uint(r);
  skey[idx] = key;
}

bool is_depth_keys_kernel(const void* func) { return func == (const void*)k_depth_keys; }

int launch_depth_keys(const gsl_params& p, const gsl_fwd_inputs& in, const GeomView& g, cudaStream_t st) {
  if (p.P == 0) return 0;
  // ctrl[8..9] (the key range, accumulated with atomicMax) were reset by the previous sort's last kernel; a workspace
  // is zero-filled before its first use.  A stale range only widens the bucket domain: the sort stays exact.
  k_depth_keys<<<(p.P + 255) / 256, 256, 0, st>>>(p.P, in.means3D, in.viewmatrix, g.skey_a, g.ctrl, g.sort_buckets,
                                                  sort_num_buckets_host(p.P));
  return check_cuda(cudaGetLastError(), "k_depth_keys launch");
}

int launch_preprocess(const gsl_params& p, const gsl_fwd_inputs& in, gsl_fwd_outputs& out,
                      const GeomView& g, cudaStream_t st) {
  if (p.P == 0) return 0;
  PreParams pp;
  pp.P = p.P; pp.D = p.D; pp.M = p.M; pp.W = p.W; pp.H = p.H;
  pp.gx = (p.W + GSL_BLOCK_X - 1) / GSL_BLOCK_X;
  pp.gy = (p.H + GSL_BLOCK_Y - 1) / GSL_BLOCK_Y;
  Fov f = make_fov(p);
  pp.VFOV_min = f.VFOV_min; pp.VFOV_max = f.VFOV_max; pp.HFOV_min = f.HFOV_min; pp.HFOV_max = f.HFOV_max;
  pp.scale_factor = p.scale_factor;
  pp.wrap = (p.flags & GSL_FLAG_WRAP_AZIMUTH) ? 1 : 0;
  for (int i = 0; i < 12; ++i) pp.samp[i] = (float)(2 * GSL_MY_PI * i / 12);
  int blocks = (p.P + 255) / 256;
  ProfScope prof(GSL_K_PREPROCESS_FWD, st);
  k_preprocess_fwd<<<blocks, 256, 0, st>>>(pp, in.means3D, in.scales, in.rotations, in.opacities, in.shs, in.shs_rest,
                                          in.colors_precomp, in.mask, in.viewmatrix, in.campos, out.radii,
                                          g.rec, g.rgb, g.rect, g.pixbox, g.tiles, g.clamped);
  return check_cuda(cudaGetLastError(), "k_preprocess_fwd launch");
}

// ------------------------------------------------------------------------------------------------
// backward preprocess
// ------------------------------------------------------------------------------------------------

struct PreBwdParams {
  int P, D, M, S, W, H, gstride;
  int factored;  // GSL_FLAG_BWD_SH_FACTORED: dL_dcolors receives the clamp-masked dL_dRGB, dL_dsh is not written
  int prezeroed; // the dense outputs were zero-filled already (side stream, under the backward compositor):
                 // only non-zero values are written here
  int row0, row1; // surfel range of this launch (chunked launches pipeline the peer exchange behind the kernel)
  int rw;         // > 0: GSL_FLAG_BWD_PEER_ROWS -- floats per packed exchange row (peer_row_width(S))
  int fused;      // peer mode: part of a fused step -- step / parity come from the device-side counter (gsl_peer.cuh)
  int factors_done; // peer mode: the SH factors were pushed already (k_peer_factor_extract / _push), rows only here
  // peer mode, gsl_peer_glue: the time-dependent part of render()'s glue VJP is applied to the row before it is pushed
  int glue, g_dynamic;
  float g_ts, g_shift, g_a, g_inv2T_decay;  // as GlueParams (gsl_glue.cu)
  const float* g_vel; const float* g_t0; const float* g_sigt; const float* g_opa;
  float VFOV_min, VFOV_max, HFOV_min, HFOV_max;
};

__device__ __forceinline__ float3 dnormvdv3(float3 v, float3 dv) {  // auxiliary.h:128-139
  float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
  float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
  float3 o;
  o.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
  o.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
  o.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
  return o;
}

__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float4 operator*(float s, float4 a) { return make_float4(s * a.x, s * a.y, s * a.z, s * a.w); }
__device__ __forceinline__ float4 operator*(float4 a, float s) { return make_float4(s * a.x, s * a.y, s * a.z, s * a.w); }
__device__ __forceinline__ float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ void operator+=(float4& a, float4 b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

// SH VJP (backward.cu:17-134).  Writes dL_dsh[0..(D+1)^2) and returns the gradient w.r.t. the
// mean through the view direction.
template <bool WRITE>
__device__ __forceinline__ float3 sh_backward(int deg, int M, const ShView sh, float4 dL_dRGB, float3 dir_orig,
                                              const ShOut dL_dsh) {
  float len = sqrtf(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
  float x = dir_orig.x / len, y = dir_orig.y / len, z = dir_orig.z / len;
  float4 dRGBdx = make_float4(0, 0, 0, 0), dRGBdy = dRGBdx, dRGBdz = dRGBdx;
  if (WRITE) dL_dsh.set(0, kSH_C0 * dL_dRGB);
  if (deg > 0) {
    if (WRITE) dL_dsh.set(1, (-kSH_C1 * y) * dL_dRGB);
    if (WRITE) dL_dsh.set(2, (kSH_C1 * z) * dL_dRGB);
    if (WRITE) dL_dsh.set(3, (-kSH_C1 * x) * dL_dRGB);
    dRGBdx = (-kSH_C1) * sh.at(3);
    dRGBdy = (-kSH_C1) * sh.at(1);
    dRGBdz = kSH_C1 * sh.at(2);
    if (deg > 1) {
      float xx = x * x, yy = y * y, zz = z * z;
      float xy = x * y, yz = y * z, xz = x * z;
      if (WRITE) dL_dsh.set(4, (kSH_C2[0] * xy) * dL_dRGB);
      if (WRITE) dL_dsh.set(5, (kSH_C2[1] * yz) * dL_dRGB);
      if (WRITE) dL_dsh.set(6, (kSH_C2[2] * (2.f * zz - xx - yy)) * dL_dRGB);
      if (WRITE) dL_dsh.set(7, (kSH_C2[3] * xz) * dL_dRGB);
      if (WRITE) dL_dsh.set(8, (kSH_C2[4] * (xx - yy)) * dL_dRGB);
      float4 s4 = sh.at(4), s5 = sh.at(5), s6 = sh.at(6), s7 = sh.at(7), s8 = sh.at(8);
      dRGBdx += (kSH_C2[0] * y) * s4 + (kSH_C2[2] * 2.f * -x) * s6 + (kSH_C2[3] * z) * s7 + (kSH_C2[4] * 2.f * x) * s8;
      dRGBdy += (kSH_C2[0] * x) * s4 + (kSH_C2[1] * z) * s5 + (kSH_C2[2] * 2.f * -y) * s6 + (kSH_C2[4] * 2.f * -y) * s8;
      dRGBdz += (kSH_C2[1] * y) * s5 + (kSH_C2[2] * 2.f * 2.f * z) * s6 + (kSH_C2[3] * x) * s7;
      if (deg > 2) {
#ifdef GSL_PBWD_FENCE
        asm volatile("" ::: "memory");  // the seven degree-3 coefficient loads start after the degree-2 terms are consumed
#endif
        if (WRITE) dL_dsh.set(9, (kSH_C3[0] * y * (3.f * xx - yy)) * dL_dRGB);
        if (WRITE) dL_dsh.set(10, (kSH_C3[1] * xy * z) * dL_dRGB);
        if (WRITE) dL_dsh.set(11, (kSH_C3[2] * y * (4.f * zz - xx - yy)) * dL_dRGB);
        if (WRITE) dL_dsh.set(12, (kSH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy)) * dL_dRGB);
        if (WRITE) dL_dsh.set(13, (kSH_C3[4] * x * (4.f * zz - xx - yy)) * dL_dRGB);
        if (WRITE) dL_dsh.set(14, (kSH_C3[5] * z * (xx - yy)) * dL_dRGB);
        if (WRITE) dL_dsh.set(15, (kSH_C3[6] * x * (xx - 3.f * yy)) * dL_dRGB);
        float4 s9 = sh.at(9), s10 = sh.at(10), s11 = sh.at(11), s12 = sh.at(12), s13 = sh.at(13), s14 = sh.at(14), s15 = sh.at(15);
        dRGBdx += (kSH_C3[0] * 3.f * 2.f * xy) * s9 + (kSH_C3[1] * yz) * s10 + (kSH_C3[2] * -2.f * xy) * s11 +
                  (kSH_C3[3] * -3.f * 2.f * xz) * s12 + (kSH_C3[4] * (-3.f * xx + 4.f * zz - yy)) * s13 +
                  (kSH_C3[5] * 2.f * xz) * s14 + (kSH_C3[6] * 3.f * (xx - yy)) * s15;
        dRGBdy += (kSH_C3[0] * 3.f * (xx - yy)) * s9 + (kSH_C3[1] * xz) * s10 +
                  (kSH_C3[2] * (-3.f * yy + 4.f * zz - xx)) * s11 + (kSH_C3[3] * -3.f * 2.f * yz) * s12 +
                  (kSH_C3[4] * -2.f * xy) * s13 + (kSH_C3[5] * -2.f * yz) * s14 + (kSH_C3[6] * -3.f * 2.f * xy) * s15;
        dRGBdz += (kSH_C3[1] * xy) * s10 + (kSH_C3[2] * 4.f * 2.f * yz) * s11 +
                  (kSH_C3[3] * 3.f * (2.f * zz - xx - yy)) * s12 + (kSH_C3[4] * 4.f * 2.f * xz) * s13 +
                  (kSH_C3[5] * (xx - yy)) * s14;
      }
    }
  }
  float3 dL_ddir = make_float3(dot4(dRGBdx, dL_dRGB), dot4(dRGBdy, dL_dRGB), dot4(dRGBdz, dL_dRGB));
  return dnormvdv3(dir_orig, dL_ddir);
}

// The 16 real SH basis values of a unit direction, as sh_backward scales dL_dRGB with them (backward.cu:17-134).
__device__ __forceinline__ void sh_basis16(int deg, float x, float y, float z, float (&b)[16]) {
#pragma unroll
  for (int k = 0; k < 16; ++k) b[k] = 0.f;
  b[0] = kSH_C0;
  if (deg > 0) {
    b[1] = -kSH_C1 * y; b[2] = kSH_C1 * z; b[3] = -kSH_C1 * x;
    if (deg > 1) {
      const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      b[4] = kSH_C2[0] * xy; b[5] = kSH_C2[1] * yz; b[6] = kSH_C2[2] * (2.f * zz - xx - yy);
      b[7] = kSH_C2[3] * xz; b[8] = kSH_C2[4] * (xx - yy);
      if (deg > 2) {
        b[9] = kSH_C3[0] * y * (3.f * xx - yy); b[10] = kSH_C3[1] * xy * z; b[11] = kSH_C3[2] * y * (4.f * zz - xx - yy);
        b[12] = kSH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy); b[13] = kSH_C3[4] * x * (4.f * zz - xx - yy);
        b[14] = kSH_C3[5] * z * (xx - yy); b[15] = kSH_C3[6] * x * (xx - 3.f * yy);
      }
    }
  }
}

// The frame-dependent part of the VJP of render()'s glue (k_glue_bwd, gsl_glue.cu; gaussian_model.py:151-157,185-186,
// gaussian_renderer/__init__.py:69-79), applied to one surfel's (dL/dmeans3D, dL/dopacity) of THIS frame before the rows
// of the frames are summed: gv = (dL/dvelocity.xyz, dL/dt), gs.x = dL/dscaling_t, gop becomes dL/d sigmoid(opacity).
__device__ __forceinline__ void glue_fold(const PreBwdParams& pp, int i, float3 dmean, float& gop, float4& gv, float4& gs) {
  const float vx = pp.g_vel[3 * (size_t)i], vy = pp.g_vel[3 * (size_t)i + 1], vz = pp.g_vel[3 * (size_t)i + 2];
  const float tt = pp.g_t0[i];
  const float sig = expf(pp.g_sigt[i]);
  const float ph = (pp.g_ts - tt) * pp.g_a;
  float sn, cs;
  sincosf(ph, &sn, &cs);
  float coef = sn / pp.g_a;
  float ev = 0.f;
  if (pp.g_shift != 0.f) { ev = expf(-sig * pp.g_inv2T_decay); coef += ev * pp.g_shift; }
  const float gvv = dmean.x * vx + dmean.y * vy + dmean.z * vz;
  float g_tt = -cs * gvv;
  float g_sig = (pp.g_shift != 0.f) ? gvv * pp.g_shift * ev * (-pp.g_inv2T_decay) : 0.f;
  if (pp.g_dynamic) {
    const float d = tt - pp.g_ts;
    const float mt = expf(-0.5f * d * d / (sig * sig));
    const float os = 1.f / (1.f + expf(-pp.g_opa[i]));
    const float g_mt = gop * os;
    gop = gop * mt;
    g_tt += g_mt * mt * (-d / (sig * sig));
    g_sig += g_mt * mt * d * d / (sig * sig * sig);
  }
  gv = make_float4(dmean.x * coef, dmean.y * coef, dmean.z * coef, g_tt);
  gs = make_float4(g_sig * sig, 0.f, 0.f, 0.f);
}

// Position a rank rasterized surfel `row` at: xyz + velocity * coef(that rank's timestamp) (k_glue_fwd, gsl_glue.cu).
struct GlueRow {
  float x, y, z, vx, vy, vz, t0, ev;
};
__device__ __forceinline__ GlueRow glue_row(const GlueMean& gm, size_t row) {
  GlueRow r;
  r.x = gm.xyz[3 * row]; r.y = gm.xyz[3 * row + 1]; r.z = gm.xyz[3 * row + 2];
  r.vx = gm.vel[3 * row]; r.vy = gm.vel[3 * row + 1]; r.vz = gm.vel[3 * row + 2];
  r.t0 = gm.t0[row];
  r.ev = expf(-expf(gm.sigt[row]) * gm.inv2T_decay);
  return r;
}
__device__ __forceinline__ float3 glue_mean_of(const GlueMean& gm, const GlueRow& r, float4 stamp) {  // stamp = (ts, shift)
  float coef = sinf((stamp.x - r.t0) * gm.a) / gm.a;
  if (stamp.y != 0.f) coef += r.ev * stamp.y;
  return make_float3(r.x + r.vx * coef, r.y + r.vy * coef, r.z + r.vz * coef);
}

// (k_preprocess_bwd is defined after preprocess_vjp_one)
// VJP of k_preprocess_fwd and of the SH evaluation for one surfel with accumulator record (g0, g1, g2, dcol, gn).
// PEER: the result is a packed exchange row (pp.rw floats) instead of entries of the dense tensors.
template <bool PEER>
__device__ __forceinline__ void preprocess_vjp_one(
    const PreBwdParams& pp, int i, float4 g0, float4 g1, float4 g2, float4 dcol, float4 gn,
    const float* __restrict__ means3D, const float* __restrict__ scales, const float* __restrict__ rotations,
    const float* __restrict__ shs, const float* __restrict__ shs_rest, const float* __restrict__ viewmatrix,
    const float* __restrict__ campos, const float4* __restrict__ rec, const uint8_t* __restrict__ clamped,
    float* __restrict__ dL_dmeans3D, float* __restrict__ dL_dmeans2D, float* __restrict__ dL_dsh,
    float* __restrict__ dL_dsh_rest, float* __restrict__ dL_dscales, float* __restrict__ dL_drot,
    float* __restrict__ rows, bool write_sh = true) {
  {
    const float vm0 = viewmatrix[0], vm1 = viewmatrix[1], vm2 = viewmatrix[2];
    const float vm4 = viewmatrix[4], vm5 = viewmatrix[5], vm6 = viewmatrix[6];
    const float vm8 = viewmatrix[8], vm9 = viewmatrix[9], vm10 = viewmatrix[10];
    const float sx = scales[3 * (size_t)i], sy = scales[3 * (size_t)i + 1];
    const float4 q = reinterpret_cast<const float4*>(rotations)[i];
    const Rot3 R = quat_to_rot(q.x, q.y, q.z, q.w);
    const float4 r0 = rec[4 * (size_t)i], r1 = rec[4 * (size_t)i + 1], r2 = rec[4 * (size_t)i + 2];
    // view-space centre = z column of T (backward.cu:584-586, :684-686)
    const float u = r0.z, v = r1.y, w = r2.x;
    // raw accumulated dL_dT rows
    float dTu_x = g0.x, dTu_y = g0.y, dTu_z = g0.z;
    float dTv_x = g0.w, dTv_y = g1.x, dTv_z = g1.y;
    float dTw_x = g1.z, dTw_y = g1.w, dTw_z = g2.x;
    const float raw_du = dTu_z, raw_dv = dTv_z, raw_dw = dTw_z;
    const float gm_x = g2.y, gm_y = g2.z;
    const float dHr = pp.HFOV_max - pp.HFOV_min, dVr = pp.VFOV_max - pp.VFOV_min;
    if (gm_x != 0.f || gm_y != 0.f) {  // backward.cu:579-595
      const float Wrange = pp.W / dHr;
      const float Hrange = pp.H / dVr;
      const float r2_uw = u * u + w * w;
      const float r_uw = sqrtf(u * u + w * w);
      const float r2 = u * u + v * v + w * w;
      dTu_z += gm_x * Wrange * w / r2_uw - gm_y * Hrange * u * v / (r_uw * r2);
      dTv_z += gm_y * Hrange * r_uw / r2;
      dTw_z += -gm_x * Wrange * u / r2_uw - gm_y * Hrange * v * w / (r_uw * r2);
    }
    // dL_dM = P * dL_dT^T, P = rotation part of the view matrix as used in T = M^T P
    // (backward.cu:598): world-space gradients of L0, L1 and the mean.
    const float3 dL0 = make_float3(vm0 * dTu_x + vm1 * dTv_x + vm2 * dTw_x, vm4 * dTu_x + vm5 * dTv_x + vm6 * dTw_x,
                                   vm8 * dTu_x + vm9 * dTv_x + vm10 * dTw_x);
    const float3 dL1 = make_float3(vm0 * dTu_y + vm1 * dTv_y + vm2 * dTw_y, vm4 * dTu_y + vm5 * dTv_y + vm6 * dTw_y,
                                   vm8 * dTu_y + vm9 * dTv_y + vm10 * dTw_y);
    float3 dmean = make_float3(vm0 * dTu_z + vm1 * dTv_z + vm2 * dTw_z, vm4 * dTu_z + vm5 * dTv_z + vm6 * dTw_z,
                               vm8 * dTu_z + vm9 * dTv_z + vm10 * dTw_z);
    // normal gradient back to world, sign by the UNFLIPPED view normal's z (backward.cu:599-603)
    float3 dtn = make_float3(vm0 * gn.x + vm1 * gn.y + vm2 * gn.z, vm4 * gn.x + vm5 * gn.y + vm6 * gn.z,
                             vm8 * gn.x + vm9 * gn.y + vm10 * gn.z);
    const float nz_view = vm2 * R.r20 + vm6 * R.r21 + vm10 * R.r22;
    const float mult = nz_view < 0.f ? 1.f : -1.f;
    dtn.x *= mult; dtn.y *= mult; dtn.z *= mult;
    // dL_dscale, dL_dR (backward.cu:604-619)
    float3 dscale;
    dscale.x = dL0.x * R.r00 + dL0.y * R.r01 + dL0.z * R.r02;
    dscale.y = dL1.x * R.r10 + dL1.y * R.r11 + dL1.z * R.r12;
    dscale.z = 0.f;
    // v_R columns: c0 = dL0*sx, c1 = dL1*sy, c2 = dtn ; vR[c][r]
    const float v00 = dL0.x * sx, v01 = dL0.y * sx, v02 = dL0.z * sx;
    const float v10 = dL1.x * sy, v11 = dL1.y * sy, v12 = dL1.z * sy;
    const float v20 = dtn.x, v21 = dtn.y, v22 = dtn.z;
    float4 drot;
    {  // quat_to_rotmat_vjp (auxiliary.h:230-274)
      const float s = rsqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
      const float qw = q.x * s, qx = q.y * s, qy = q.z * s, qz = q.w * s;
      drot.x = 2.f * (qx * (v12 - v21) + qy * (v20 - v02) + qz * (v01 - v10));
      drot.y = 2.f * (-2.f * qx * (v11 + v22) + qy * (v01 + v10) + qz * (v02 + v20) + qw * (v12 - v21));
      drot.z = 2.f * (qx * (v01 + v10) - 2.f * qy * (v00 + v22) + qz * (v12 + v21) + qw * (v20 - v02));
      drot.w = 2.f * (qx * (v02 + v20) + qy * (v12 + v21) - 2.f * qz * (v00 + v11) + qw * (v01 - v10));
    }
    // SH (backward.cu:676-677)
    if (shs != nullptr) {
      const uint8_t cl = clamped[i];
      const float4 dRGB = make_float4((cl & 1) ? 0.f : dcol.x, (cl & 2) ? 0.f : dcol.y, (cl & 4) ? 0.f : dcol.z,
                                      (cl & 8) ? 0.f : dcol.w);
      // the reference masks dL_dcolors only in a local copy; the returned tensor is discarded on the SH
      // path by the Python wrapper semantics (colors_precomp is empty), so dL_dcolors stays as accumulated.
      const float3 dir = make_float3(means3D[3 * (size_t)i] - campos[0], means3D[3 * (size_t)i + 1] - campos[1],
                                     means3D[3 * (size_t)i + 2] - campos[2]);
      const ShOut out_sh = sh_out(dL_dsh, dL_dsh_rest, (size_t)i, pp.M);
      const ShView sh_i = sh_view(shs, shs_rest, (size_t)i, pp.M);
      const float3 dm = (pp.factored || !write_sh) ? sh_backward<false>(pp.D, pp.M, sh_i, dRGB, dir, out_sh)
                                                   : sh_backward<true>(pp.D, pp.M, sh_i, dRGB, dir, out_sh);
      // coefficients >= (D+1)^2 keep the zeros of the sweep
      dmean.x += dm.x; dmean.y += dm.y; dmean.z += dm.z;
    }
    // densification proxy (backward.cu:684-711), from the RAW accumulated dT z-column
    const float phi = atan2f(u, w);
    float2 dm2;
    dm2.x = (float)((raw_du * w + raw_dw * (-u)) * 0.5 * dHr);
    const float du_dth = -v * sinf(phi), dv_dth = sqrtf(u * u + w * w), dw_dth = -v * cosf(phi);
    dm2.y = (float)((raw_du * du_dth + raw_dv * dv_dth + raw_dw * dw_dth) * 0.5 * dVr * pp.W / pp.H);

    if (PEER) {  // packed exchange row: [means2D.xy scales.xy | rot | means3D opacity | features... | glue]
      float4* r = reinterpret_cast<float4*>(rows + (size_t)i * pp.rw);
      r[0] = make_float4(dm2.x, dm2.y, dscale.x, dscale.y);
      r[1] = drot;
      float gop = g2.w;
      if (pp.glue) {
        float4 gv, gs;
        glue_fold(pp, i, dmean, gop, gv, gs);
        const int q = 3 + (pp.S + 3) / 4;  // first quad behind the feature quads
        r[q] = gv;
        r[q + 1] = gs;
      }
      r[2] = make_float4(dmean.x, dmean.y, dmean.z, gop);
      return;
    }
    dL_dmeans3D[3 * (size_t)i] = dmean.x; dL_dmeans3D[3 * (size_t)i + 1] = dmean.y; dL_dmeans3D[3 * (size_t)i + 2] = dmean.z;
    dL_dscales[3 * (size_t)i] = dscale.x; dL_dscales[3 * (size_t)i + 1] = dscale.y; dL_dscales[3 * (size_t)i + 2] = dscale.z;
    reinterpret_cast<float4*>(dL_drot)[i] = drot;
    reinterpret_cast<float4*>(dL_dmeans2D)[i] = make_float4(dm2.x, dm2.y, 0.f, 0.f);
  }
}

// Backward preprocess, 256 surfels per CTA, two phases.
//  1. thread = surfel, pure streaming: read the packed accumulator record, re-zero it (it must be all-zero
//     when the next backward starts), copy the pass-through gradients (opacity, colour, features), zero-fill
//     every other dense output with coalesced stores and queue the surfels whose record is non-zero in
//     shared memory.  A surfel no pixel contributed to -- culled, hidden behind saturated pixels, outside
//     every box: about half of the 1M-surfel scene -- has all-zero gradients and its parameters / SH
//     coefficients are never read.
//  2. the queued surfels, densely packed over the CTA's threads: VJP of k_preprocess_fwd and of the SH
//     evaluation (preprocess_vjp_one).  Every element of every dense output is written by one of the phases.
// Three CTAs per SM for the plain kernel (80 registers, no spills): the kernel is bound by the latency of its dependent
// gathers, and 768 instead of 512 resident threads took it from 0.150 to 0.123 ms at 1M surfels (four CTAs: the same).
// The peer-exchange variant keeps two (it spills at 80 registers).
#ifndef GSL_PBWD_MINB
#define GSL_PBWD_MINB 3
#endif
#ifndef GSL_PBWD_PEER_MINB
#define GSL_PBWD_PEER_MINB 2
#endif
template <bool PEER>
__global__ void __launch_bounds__(256, PEER ? GSL_PBWD_PEER_MINB : GSL_PBWD_MINB) k_preprocess_bwd(
    PreBwdParams pp, const float* __restrict__ means3D, const float* __restrict__ scales,
    const float* __restrict__ rotations, const float* __restrict__ shs, const float* __restrict__ shs_rest,
    const float* __restrict__ viewmatrix, const float* __restrict__ campos,
    const int* __restrict__ radii, const float4* __restrict__ rec, const uint8_t* __restrict__ clamped,
    uint8_t* __restrict__ touched, float* __restrict__ grad, float* __restrict__ dL_dmeans3D, float* __restrict__ dL_dmeans2D,
    float* __restrict__ dL_dsh, float* __restrict__ dL_dsh_rest, float* __restrict__ dL_dcolors,
    float* __restrict__ dL_dfeatures,
    float* __restrict__ dL_dopacity, float* __restrict__ dL_dscales, float* __restrict__ dL_drot,
    float* __restrict__ dL_dcov3D, PeerView pv, const PeerLayout pl) {
  __shared__ float4 s_g[5][256];  // queued accumulator records (dT, mean2D, opacity, colour, normal)
  __shared__ uint16_t s_who[256];
  __shared__ int s_count;
  const bool have_sh = shs != nullptr;
  const int cta0 = pp.row0 + blockIdx.x * 256;
  const int idx = cta0 + threadIdx.x;
  const int S = pp.S;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  if (PEER) {
    // ---- peer-memory gradient exchange: this CTA's 256 surfels are one tile; its packed rows (only those with a
    // non-zero record, + a row bitmap) are PUSHED into the staging area of the rank that owns the tile, its non-zero SH
    // factors (packed to the front of the tile's segment, + bits / prefix words) into every rank's factor table.  Remote
    // stores only: nothing here waits for NVLink.  Nothing is zero-filled (readers look at the bits first).
    __shared__ int s_warp[8];
    __shared__ uint32_t s_any[8];
    __shared__ float4 s_rows[256 * 6];  // the tile's 64- or 96-byte rows, flushed as contiguous 16-byte pieces at the end
    if (pp.fused) peer_resolve_step(pv);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile = cta0 >> 8;
    const int owner = tile % pv.world;
    char* obuf = nullptr;
#pragma unroll
    for (int g = 0; g < PEER_MAX; ++g)
      if (g == owner) obuf = pv.buf[g];
    const size_t slot0 = ((size_t)pv.rank * pl.tiles_per_rank + tile / pv.world) * 256;  // staging row of the tile's row 0
    float* rows_remote = reinterpret_cast<float*>(obuf + pl.off_stage) + (ptrdiff_t)(slot0 - (size_t)cta0) * pp.rw;
    // 16-byte stores 64 bytes apart make poor NVLink packets: 64-byte rows are staged in shared memory first
    const bool staged = pp.rw == 16 || pp.rw == 24;
    float* rows = staged ? reinterpret_cast<float*>(&s_rows[0]) - (ptrdiff_t)cta0 * pp.rw : rows_remote;
    bool any = false;
    float4 fac = zero4;
    if (idx < pp.row1 && touched[idx]) {  // no pixel composited the others: their accumulators are all-zero, never read
      touched[idx] = 0;
      float4* gq = reinterpret_cast<float4*>(grad + (size_t)idx * pp.gstride);
      const float4 g0 = gq[0], g1 = gq[1], g2 = gq[2], gc = gq[3], gn = gq[4];
      any = (g0.x != 0.f) | (g0.y != 0.f) | (g0.z != 0.f) | (g0.w != 0.f) | (g1.x != 0.f) | (g1.y != 0.f) |
            (g1.z != 0.f) | (g1.w != 0.f) | (g2.x != 0.f) | (g2.y != 0.f) | (g2.z != 0.f) | (g2.w != 0.f) |
            (gc.x != 0.f) | (gc.y != 0.f) | (gc.z != 0.f) | (gc.w != 0.f) | (gn.x != 0.f) | (gn.y != 0.f) |
            (gn.z != 0.f);
      if (!pp.factors_done && (gc.x != 0.f || gc.y != 0.f || gc.z != 0.f || gc.w != 0.f)) {
        const uint8_t cl = clamped[idx];
        fac = make_float4((cl & 1) ? 0.f : gc.x, (cl & 2) ? 0.f : gc.y, (cl & 4) ? 0.f : gc.z, (cl & 8) ? 0.f : gc.w);
      }
      const int nf4 = (S + 3) / 4;
      float4 f[3] = {zero4, zero4, zero4};
      for (int k = 0; k < nf4; ++k) {
        f[k] = gq[5 + k];
        if (f[k].x != 0.f || f[k].y != 0.f || f[k].z != 0.f || f[k].w != 0.f) {
          any = true;
          gq[5 + k] = zero4;
        }
      }
      if (any) {
        float4* r = reinterpret_cast<float4*>(rows + (size_t)idx * pp.rw);
        for (int k = 0; k < pp.rw / 4 - 3; ++k) r[3 + k] = k < nf4 ? f[k] : zero4;
        gq[0] = zero4; gq[1] = zero4; gq[2] = zero4; gq[3] = zero4; gq[4] = zero4;
        if (radii[idx] > 0) {
          const int slot = atomicAdd(&s_count, 1);
          s_who[slot] = (uint16_t)threadIdx.x;
          s_g[0][slot] = g0; s_g[1][slot] = g1; s_g[2][slot] = g2; s_g[3][slot] = gc; s_g[4][slot] = gn;
        } else {
          float gop = g2.w;
          if (pp.glue) {  // (a culled surfel's opacity gradient is zero in practice; kept exact all the same)
            float4 gv, gs;
            glue_fold(pp, idx, make_float3(0.f, 0.f, 0.f), gop, gv, gs);
            r[3 + nf4] = gv;
            r[4 + nf4] = gs;
          }
          r[0] = zero4; r[1] = zero4; r[2] = make_float4(0.f, 0.f, 0.f, gop);
        }
      }
    }
    const bool live = cta0 + warp * 32 < pp.row1;  // this warp's word exists
    const uint32_t bits = __ballot_sync(0xffffffffu, any);
    if (lane == 0 && live) reinterpret_cast<uint32_t*>(obuf + pl.off_stagebits)[(slot0 >> 5) + warp] = bits;
    if (lane == 0) s_any[warp] = bits;
    if (pp.factors_done) {
      __syncthreads();  // s_count, the queued records and s_any are complete
    } else {
      // SH factors: block-local compaction, pushed to every rank
      const bool nz = fac.x != 0.f || fac.y != 0.f || fac.z != 0.f || fac.w != 0.f;
      const uint32_t fbits = __ballot_sync(0xffffffffu, nz);
      if (lane == 0) s_warp[warp] = __popc(fbits);
      __syncthreads();
      int before = 0;
      for (int w = 0; w < warp; ++w) before += s_warp[w];
      const size_t table = (size_t)(pv.parity * pv.world + pv.rank) * pl.tiles + tile;  // [parity][source rank][tile]
      const size_t fpos = table * 256 + before + __popc(fbits & ((1u << lane) - 1u));
      const size_t mpos = table * 8 + warp;
#pragma unroll
      for (int g = 0; g < PEER_MAX; ++g) {
        if (g < pv.world) {
          if (nz) reinterpret_cast<float4*>(pv.buf[g] + pl.off_factor)[fpos] = fac;
          if (lane == 0 && live)
            reinterpret_cast<uint2*>(pv.buf[g] + pl.off_fmeta)[mpos] = make_uint2(fbits, (uint32_t)before);
        }
      }
    }
    const int count = s_count;
    for (int slot = threadIdx.x; slot < count; slot += 256)
      preprocess_vjp_one<true>(pp, cta0 + (int)s_who[slot], s_g[0][slot], s_g[1][slot], s_g[2][slot], s_g[3][slot],
                         s_g[4][slot], means3D, scales, rotations, shs, shs_rest, viewmatrix, campos, rec, clamped,
                         dL_dmeans3D, dL_dmeans2D, dL_dsh, dL_dsh_rest, dL_dscales, dL_drot, rows);
    if (staged) {
      __syncthreads();
      const int rw4 = pp.rw >> 2;  // 4 or 6 sixteen-byte pieces per row; consecutive threads store consecutive pieces
      float4* dst = reinterpret_cast<float4*>(obuf + pl.off_stage) + slot0 * rw4;
      for (int item = threadIdx.x; item < 256 * rw4; item += 256) {
        const int r = rw4 == 4 ? item >> 2 : item / 6;
        if ((s_any[r >> 5] >> (r & 31)) & 1u) dst[item] = s_rows[item];
      }
    }
    return;  // fused step: the "pushed" flag is published by k_peer_signal, the next kernel of the stream
  }
  const bool was_touched = idx < pp.row1 && touched[idx] != 0;
  if (was_touched) touched[idx] = 0;
  if (idx < pp.row1 && (was_touched || !pp.prezeroed)) {
    // a surfel no pixel composited has an all-zero accumulator: it is not read (prezeroed outputs: nothing to do at all)
    float4* gq = reinterpret_cast<float4*>(grad + (size_t)idx * pp.gstride);
    float4 g0 = zero4, g1 = zero4, g2 = zero4, gc = zero4, gn = zero4;
    if (was_touched) { g0 = gq[0]; g1 = gq[1]; g2 = gq[2]; gc = gq[3]; gn = gq[4]; }
    const bool any = (g0.x != 0.f) | (g0.y != 0.f) | (g0.z != 0.f) | (g0.w != 0.f) | (g1.x != 0.f) | (g1.y != 0.f) |
                     (g1.z != 0.f) | (g1.w != 0.f) | (g2.x != 0.f) | (g2.y != 0.f) | (g2.z != 0.f) | (g2.w != 0.f) |
                     (gc.x != 0.f) | (gc.y != 0.f) | (gc.z != 0.f) | (gc.w != 0.f) | (gn.x != 0.f) | (gn.y != 0.f) |
                     (gn.z != 0.f);
    const int nf4 = (S + 3) / 4;
    for (int k = 0; k < nf4; ++k) {
      const float4 f = was_touched ? gq[5 + k] : zero4;
      const float fv[4] = {f.x, f.y, f.z, f.w};
      bool fany = false;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ch = 4 * k + j;
        if (ch < S) fany |= (fv[j] != 0.f);
      }
      if (fany || !pp.prezeroed) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int ch = 4 * k + j;
          if (ch < S) dL_dfeatures[(size_t)idx * S + ch] = fv[j];
        }
      }
      if (fany) gq[5 + k] = zero4;
    }
    if (any || !pp.prezeroed) dL_dopacity[idx] = g2.w;
    float4 gc_out = gc;
    if (pp.factored && any) {  // clamp-masked dL_dRGB: the factor the SH gradient is rebuilt from (k_sh_expand)
      const uint8_t cl = clamped[idx];
      gc_out = make_float4((cl & 1) ? 0.f : gc.x, (cl & 2) ? 0.f : gc.y, (cl & 4) ? 0.f : gc.z, (cl & 8) ? 0.f : gc.w);
    }
    if (any || !pp.prezeroed) reinterpret_cast<float4*>(dL_dcolors)[idx] = gc_out;
    const bool heavy = any && radii[idx] > 0;
    if (any) { gq[0] = zero4; gq[1] = zero4; gq[2] = zero4; gq[3] = zero4; gq[4] = zero4; }
    if (heavy) {
      const int slot = atomicAdd(&s_count, 1);
      s_who[slot] = (uint16_t)threadIdx.x;
      s_g[0][slot] = g0; s_g[1][slot] = g1; s_g[2][slot] = g2; s_g[3][slot] = gc; s_g[4][slot] = gn;
    } else if (!pp.prezeroed) {
      reinterpret_cast<float4*>(dL_drot)[idx] = zero4;
      reinterpret_cast<float4*>(dL_dmeans2D)[idx] = zero4;
    }
  }
  // coalesced zero-fill of the CTA's rows of the strided outputs (queued surfels overwrite theirs later)
  const int nrows = min(256, pp.row1 - cta0);
  if (pp.prezeroed) {
    // nothing to fill
  } else if (have_sh && !pp.factored) {
    const int m0 = dL_dsh_rest ? 1 : pp.M;
    float4* o = reinterpret_cast<float4*>(dL_dsh) + (size_t)cta0 * m0;
    for (int k = threadIdx.x; k < nrows * m0; k += 256) o[k] = zero4;
    if (dL_dsh_rest) {
      o = reinterpret_cast<float4*>(dL_dsh_rest) + (size_t)cta0 * (pp.M - 1);
      for (int k = threadIdx.x; k < nrows * (pp.M - 1); k += 256) o[k] = zero4;
    }
  }
  if (!pp.prezeroed) {
    float* o = dL_dmeans3D + (size_t)cta0 * 3;
    for (int k = threadIdx.x; k < nrows * 3; k += 256) o[k] = 0.f;
    o = dL_dscales + (size_t)cta0 * 3;
    for (int k = threadIdx.x; k < nrows * 3; k += 256) o[k] = 0.f;
    if (dL_dcov3D) {
      o = dL_dcov3D + (size_t)cta0 * 6;
      for (int k = threadIdx.x; k < nrows * 6; k += 256) o[k] = 0.f;
    }
  }
  __syncthreads();
  // ---- phase 2: the queued surfels, densely packed over the CTA's threads (count <= 256: one per thread)
  const int count = s_count;
  const int slot = threadIdx.x;
  const bool valid = slot < count;
  const int row = cta0 + (int)s_who[valid ? slot : 0];
  const float4 q0 = s_g[0][slot], q1 = s_g[1][slot], q2 = s_g[2][slot], q3 = s_g[3][slot], q4 = s_g[4][slot];
  const bool sh_rows = have_sh && !pp.factored;  // dL_dsh is written here, through shared memory (below)
  if (valid)
    preprocess_vjp_one<false>(pp, row, q0, q1, q2, q3, q4, means3D, scales, rotations, shs, shs_rest, viewmatrix, campos, rec,
                       clamped, dL_dmeans3D, dL_dmeans2D, dL_dsh, dL_dsh_rest, dL_dscales, dL_drot, nullptr, !sh_rows);
  if (!sh_rows) return;
  // dL_dsh rows (backward.cu:17-134: basis_k(view direction) x dL_dRGB).  A thread storing its own 16 float4 writes 16-byte
  // pieces 256 bytes apart -- half sectors, read-modify-written in DRAM (ncu: 1.18x the algorithmic traffic).  The rows go
  // through shared memory instead, four coefficients at a time (the queue's storage is free by now), so that a store
  // instruction writes 8 rows x 64 contiguous bytes = whole sectors.
  __syncthreads();  // every thread holds its queue entry in registers: s_g becomes the staging area
  float4 (*stage)[5] = reinterpret_cast<float4 (*)[5]>(&s_g[0][0]) + (threadIdx.x & ~31);  // [32 rows][4 + 1 pad] per warp
  const int lane = threadIdx.x & 31;
  const int wbase = threadIdx.x & ~31;
  if (wbase >= count) return;  // whole warp without queue entries
  float b[16];
  float4 dRGB = zero4;
  {
    const size_t rr = (size_t)row;
    const uint8_t cl = clamped[rr];
    dRGB = make_float4((cl & 1) ? 0.f : q3.x, (cl & 2) ? 0.f : q3.y, (cl & 4) ? 0.f : q3.z, (cl & 8) ? 0.f : q3.w);
    const float dx = means3D[3 * rr] - campos[0], dy = means3D[3 * rr + 1] - campos[1], dz = means3D[3 * rr + 2] - campos[2];
    const float len = sqrtf(dx * dx + dy * dy + dz * dz);
    sh_basis16(pp.D, dx / len, dy / len, dz / len, b);
  }
  const int ncoef = (pp.D + 1) * (pp.D + 1);  // coefficients >= (D+1)^2 keep the zeros of the sweep
#pragma unroll
  for (int quarter = 0; quarter < 4; ++quarter) {
    if (4 * quarter < ncoef) {  // warp-uniform
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 4; ++c) stage[lane][c] = b[4 * quarter + c] * dRGB;
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int r = it * 8 + (lane >> 2), c = lane & 3;  // 8 rows x 4 coefficients per store instruction
        const int k = 4 * quarter + c;
        if (wbase + r < count && k < ncoef) {
          const int rw = cta0 + (int)s_who[wbase + r];
          sh_out(dL_dsh, dL_dsh_rest, (size_t)rw, pp.M).set(k, stage[r][c]);
        }
      }
    }
  }
}


// The SH factor of frame-parallel training, extracted from the packed accumulators right after the backward
// compositor (so that its all-gather can run under k_preprocess_bwd): clamp-masked dL_dRGB per surfel.
__global__ void __launch_bounds__(256) k_extract_sh_factor(int P, int gstride, const float* __restrict__ grad,
                                                           const uint8_t* __restrict__ clamped, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float4 gc = reinterpret_cast<const float4*>(grad + (size_t)i * gstride)[3];
  float4 o = gc;
  if (gc.x != 0.f || gc.y != 0.f || gc.z != 0.f || gc.w != 0.f) {
    const uint8_t cl = clamped[i];
    o = make_float4((cl & 1) ? 0.f : gc.x, (cl & 2) ? 0.f : gc.y, (cl & 4) ? 0.f : gc.z, (cl & 8) ? 0.f : gc.w);
  }
  out[i] = o;
}

int launch_extract_sh_factor(const gsl_params& p, const GeomView& g, float* out, cudaStream_t st) {
  if (p.P == 0) return 0;
  k_extract_sh_factor<<<(p.P + 255) / 256, 256, 0, st>>>(p.P, grad_stride(p.S), g.grad, g.clamped,
                                                        reinterpret_cast<float4*>(out));
  return check_cuda(cudaGetLastError(), "k_extract_sh_factor launch");
}

// ---- fused exchange step: the SH factors leave EARLY ----------------------------------------------------------------------
// A factor (clamp-masked dL_dRGB) is complete when the backward compositor has finished, 0.15 ms before the packed rows
// k_preprocess_bwd produces, and the factor tables are half of the step's NVLink traffic (16 B x touched surfels into every
// rank).  k_peer_factor_extract (main stream, ~10 us: one 32-byte sector per accumulator record) packs this rank's non-zero
// factors tile by tile into its OWN factor table (+ bits / prefix words), before k_preprocess_bwd re-zeroes the accumulators;
// k_peer_factor_push (side stream) copies the packed part of every tile into the other ranks' tables while k_preprocess_bwd
// runs, and the "factors" flag behind it lets every rank start its expansion without waiting for anybody's rows.
__global__ void __launch_bounds__(256) k_peer_factor_extract(PeerView pv, const PeerLayout pl, int P, int gstride,
                                                             const float* __restrict__ grad,
                                                             const uint8_t* __restrict__ clamped,
                                                             const uint8_t* __restrict__ touched) {
  __shared__ int s_warp[8];
  peer_resolve_step(pv);
  const int tile = blockIdx.x;
  const int idx = tile * 256 + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 fac = make_float4(0.f, 0.f, 0.f, 0.f);
  if (idx < P && touched[idx]) {
    const float4 gc = reinterpret_cast<const float4*>(grad + (size_t)idx * gstride)[3];
    if (gc.x != 0.f || gc.y != 0.f || gc.z != 0.f || gc.w != 0.f) {
      const uint8_t cl = clamped[idx];
      fac = make_float4((cl & 1) ? 0.f : gc.x, (cl & 2) ? 0.f : gc.y, (cl & 4) ? 0.f : gc.z, (cl & 8) ? 0.f : gc.w);
    }
  }
  const bool nz = fac.x != 0.f || fac.y != 0.f || fac.z != 0.f || fac.w != 0.f;
  const uint32_t fbits = __ballot_sync(0xffffffffu, nz);
  if (lane == 0) s_warp[warp] = __popc(fbits);
  __syncthreads();
  int before = 0;
  for (int w = 0; w < warp; ++w) before += s_warp[w];
  const size_t table = (size_t)(pv.parity * pv.world + pv.rank) * pl.tiles + tile;  // [parity][source rank][tile]
  if (nz) reinterpret_cast<float4*>(pv.own + pl.off_factor)[table * 256 + before + __popc(fbits & ((1u << lane) - 1u))] = fac;
  if (lane == 0) reinterpret_cast<uint2*>(pv.own + pl.off_fmeta)[table * 8 + warp] = make_uint2(fbits, (uint32_t)before);
}

__global__ void __launch_bounds__(256) k_peer_factor_push(PeerView pv, const PeerLayout pl) {
  peer_resolve_step(pv);
  const size_t slab = (size_t)(pv.parity * pv.world + pv.rank) * pl.tiles;
  for (int tile = blockIdx.x; tile < pl.tiles; tile += gridDim.x) {
    const size_t table = slab + tile;
    const uint2* meta = reinterpret_cast<const uint2*>(pv.own + pl.off_fmeta) + table * 8;
    const uint2 last = meta[7];
    const int count = (int)last.y + __popc(last.x);
    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
    uint2 m = make_uint2(0u, 0u);
    if ((int)threadIdx.x < count) f = reinterpret_cast<const float4*>(pv.own + pl.off_factor)[table * 256 + threadIdx.x];
    if (threadIdx.x < 8) m = meta[threadIdx.x];
#pragma unroll
    for (int g = 0; g < PEER_MAX; ++g) {
      if (g < pv.world && g != pv.rank) {
        if ((int)threadIdx.x < count) reinterpret_cast<float4*>(pv.buf[g] + pl.off_factor)[table * 256 + threadIdx.x] = f;
        if (threadIdx.x < 8) reinterpret_cast<uint2*>(pv.buf[g] + pl.off_fmeta)[table * 8 + threadIdx.x] = m;
      }
    }
  }
}

int launch_peer_factor_extract(const gsl_peer_ctx* c, const gsl_params& p, const GeomView& g, cudaStream_t st) {
  if (p.P == 0) return 0;
  const PeerLayout pl = peer_layout((size_t)p.P, p.S, c->world);
  k_peer_factor_extract<<<pl.tiles, 256, 0, st>>>(make_view(c, true), pl, p.P, grad_stride(p.S), g.grad, g.clamped,
                                                  g.touched);
  return check_cuda(cudaGetLastError(), "k_peer_factor_extract launch");
}

int launch_peer_factor_push(const gsl_peer_ctx* c, const gsl_params& p, cudaStream_t st) {
  if (p.P == 0 || c->world < 2) return 0;
  const PeerLayout pl = peer_layout((size_t)p.P, p.S, c->world);
  k_peer_factor_push<<<std::min(pl.tiles, 148 * 8), 256, 0, st>>>(make_view(c, true), pl);
  return check_cuda(cudaGetLastError(), "k_peer_factor_push launch");
}

// zero-fill of every dense gradient output (enqueued on a side stream under the backward compositor)
int launch_zero_outputs(const gsl_params& p, const gsl_fwd_inputs& in, gsl_bwd_outputs& gout, cudaStream_t st) {
  const size_t P = (size_t)p.P;
  if (P == 0) return 0;
  struct Range { char* ptr; size_t bytes; };
  Range r[10];
  int n = 0;
  auto add = [&](float* ptr, size_t bytes) { if (ptr && bytes) { r[n].ptr = (char*)ptr; r[n].bytes = bytes; ++n; } };
  add(gout.dL_dmeans3D, P * 12);
  add(gout.dL_dmeans2D, P * 16);
  add(gout.dL_dcolors, P * 16);
  add(gout.dL_dopacity, P * 4);
  add(gout.dL_dscales, P * 12);
  add(gout.dL_drotations, P * 16);
  const int S_out = (p.flags & GSL_FLAG_BWD_PEER_ROWS) ? peer_rows_S(p.S, gout.peer) : p.S;
  if (S_out > 0) add(gout.dL_dfeatures, P * 4 * (size_t)S_out);
  if (in.shs && (p.flags & GSL_FLAG_BWD_PEER_ROWS)) {
    add(gout.dL_dsh, P * 16 * (size_t)p.M);  // summed over the ranks by k_peer_sh_expand, non-zero rows only
  } else if (in.shs && !(p.flags & GSL_FLAG_BWD_SH_FACTORED)) {
    add(gout.dL_dsh, P * 16 * (size_t)(in.shs_rest ? 1 : p.M));
    if (in.shs_rest) add(gout.dL_dsh_rest, P * 16 * (size_t)(p.M - 1));
  }
  add(gout.dL_dcov3D, P * 24);
  // exactly adjacent tensors (the Python wrapper packs all gradients into one allocation) become one memset
  for (int i = 1; i < n; ++i)
    for (int j = i; j > 0 && r[j].ptr < r[j - 1].ptr; --j) { Range t = r[j]; r[j] = r[j - 1]; r[j - 1] = t; }
  for (int i = 0; i < n;) {
    char* start = r[i].ptr;
    size_t bytes = r[i].bytes;
    int j = i + 1;
    while (j < n && r[j].ptr == start + bytes) { bytes += r[j].bytes; ++j; }
    cudaMemsetAsync(start, 0, bytes, st);  // (an own 16-byte-store kernel measured 0.02 ms / step slower)
    i = j;
  }
  return check_cuda(cudaGetLastError(), "zero-fill of the gradient outputs");
}

int launch_preprocess_backward(const gsl_params& p, const gsl_fwd_inputs& in, const gsl_fwd_outputs& fwd,
                               gsl_bwd_outputs& gout, const GeomView& g, bool prezeroed, int row0, int row1,
                               cudaStream_t st, bool fused, bool factors_done) {
  if (p.P == 0 || row1 <= row0) return 0;
  PreBwdParams pp;
  pp.P = p.P; pp.D = p.D; pp.M = p.M; pp.S = p.S;
  {
    // The reference re-derives W,H from focal*tan*2 in float (rasterizer_impl.cu:440-441,
    // backward.cu:658-659); reproduce the truncation instead of assuming it round-trips.
    volatile float fy = p.H / (2.0f * p.tanfovy), fx = p.W / (2.0f * p.tanfovx);
    volatile float wf = fx * p.tanfovx, hf = fy * p.tanfovy;
    volatile float w2 = wf * 2, h2 = hf * 2;
    pp.W = (int)w2; pp.H = (int)h2;
    if (!(p.tanfovx == p.tanfovx) || p.tanfovx == 0.f) pp.W = p.W;
    if (!(p.tanfovy == p.tanfovy) || p.tanfovy == 0.f) pp.H = p.H;
  }
  pp.gstride = grad_stride(p.S);
  pp.factored = (p.flags & GSL_FLAG_BWD_SH_FACTORED) ? 1 : 0;
  pp.prezeroed = prezeroed ? 1 : 0;
  pp.row0 = row0; pp.row1 = row1;
  pp.rw = (p.flags & GSL_FLAG_BWD_PEER_ROWS) ? peer_row_width(peer_rows_S(p.S, gout.peer)) : 0;
  PeerView pv = {};
  PeerLayout pl = {};
  pp.fused = (fused && pp.rw > 0) ? 1 : 0;
  pp.factors_done = (factors_done && pp.rw > 0) ? 1 : 0;
  pp.glue = 0; pp.g_dynamic = 0;
  pp.g_ts = pp.g_shift = pp.g_a = pp.g_inv2T_decay = 0.f;
  pp.g_vel = pp.g_t0 = pp.g_sigt = pp.g_opa = nullptr;
  if (pp.rw > 0) {
    pp.factored = 1;
    pv = make_view(gout.peer, fused);
    pl = peer_layout((size_t)p.P, peer_rows_S(p.S, gout.peer), gout.peer->world);
    if (const gsl_peer_glue* gl = gout.peer->glue) {  // same constants as make_glue_params (gsl_glue.cu)
      pp.glue = 1;
      pp.g_dynamic = gl->dynamic;
      pp.g_ts = gl->timestamp - gl->time_shift;
      pp.g_shift = gl->time_shift;
      pp.g_a = (float)(1.0 / (double)gl->cycle * 3.141592653589793 * 2.0);
      pp.g_inv2T_decay = gl->velocity_decay / gl->cycle / 2.f;
      pp.g_vel = gl->velocity; pp.g_t0 = gl->t; pp.g_sigt = gl->scaling_t; pp.g_opa = gl->opacity;
    }
  }
  Fov f = make_fov(p);
  pp.VFOV_min = f.VFOV_min; pp.VFOV_max = f.VFOV_max; pp.HFOV_min = f.HFOV_min; pp.HFOV_max = f.HFOV_max;
  int blocks = (row1 - row0 + 255) / 256;
  ProfScope prof(GSL_K_PREPROCESS_BWD, st);
#define GSL_LAUNCH_PBWD(PEER)                                                                                         \
  k_preprocess_bwd<PEER><<<blocks, 256, 0, st>>>(pp, in.means3D, in.scales, in.rotations, in.shs, in.shs_rest,         \
                                                in.viewmatrix, in.campos, fwd.radii, g.rec, g.clamped, g.touched, g.grad, \
                                                gout.dL_dmeans3D, gout.dL_dmeans2D, gout.dL_dsh,                         \
                                                in.shs_rest ? gout.dL_dsh_rest : nullptr, gout.dL_dcolors,               \
                                                gout.dL_dfeatures, gout.dL_dopacity, gout.dL_dscales, gout.dL_drotations, \
                                                gout.dL_dcov3D, pv, pl)
  if (pp.rw > 0) GSL_LAUNCH_PBWD(true);
  else GSL_LAUNCH_PBWD(false);
#undef GSL_LAUNCH_PBWD
  return check_cuda(cudaGetLastError(), "k_preprocess_bwd launch");
}

// ------------------------------------------------------------------------------------------------
// SH gradient expansion for frame-parallel training (gsl_sh_expand).  The SH gradient a frame contributes to a
// surfel is the outer product basis(dir) x dL_dRGB (backward.cu:17-134), so instead of all-reducing 16*M bytes
// per surfel the ranks all-gather the 16-byte factor dL_dRGB (already clamp-masked) and every rank rebuilds
//     dL_dsh[i] = sum_g basis((mean_i - campos_g) / |.|) x dL_dRGB_g[i]
// -- the same sum an all-reduce of the dense gradients would produce, in a different order.
// ------------------------------------------------------------------------------------------------
// acc[k] += basis_k(normalize(x, y, z)) * d for the coefficients of degree <= D (the dL_dsh rows of backward.cu:17-134)
__device__ __forceinline__ void sh_basis_accumulate(float4 (&acc)[16], int D, float x, float y, float z, const float4 d) {
  const float inv = rsqrtf(x * x + y * y + z * z);  // 2 ulp: the sum is compared with the dense gradients to 1e-4
  x *= inv; y *= inv; z *= inv;
  acc[0] += kSH_C0 * d;
  if (D > 0) {
    acc[1] += (-kSH_C1 * y) * d;
    acc[2] += (kSH_C1 * z) * d;
    acc[3] += (-kSH_C1 * x) * d;
    if (D > 1) {
      const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      acc[4] += (kSH_C2[0] * xy) * d;
      acc[5] += (kSH_C2[1] * yz) * d;
      acc[6] += (kSH_C2[2] * (2.f * zz - xx - yy)) * d;
      acc[7] += (kSH_C2[3] * xz) * d;
      acc[8] += (kSH_C2[4] * (xx - yy)) * d;
      if (D > 2) {
        acc[9] += (kSH_C3[0] * y * (3.f * xx - yy)) * d;
        acc[10] += (kSH_C3[1] * xy * z) * d;
        acc[11] += (kSH_C3[2] * y * (4.f * zz - xx - yy)) * d;
        acc[12] += (kSH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy)) * d;
        acc[13] += (kSH_C3[4] * x * (4.f * zz - xx - yy)) * d;
        acc[14] += (kSH_C3[5] * z * (xx - yy)) * d;
        acc[15] += (kSH_C3[6] * x * (xx - 3.f * yy)) * d;
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_sh_expand(int P, int D, int M, int G, const float* __restrict__ means3D,
                                                   const float* __restrict__ campos_all, const float* __restrict__ drgb_all,
                                                   size_t drgb_stride, float* __restrict__ dL_dsh) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = zero4;
  const float mx = means3D[3 * (size_t)i], my = means3D[3 * (size_t)i + 1], mz = means3D[3 * (size_t)i + 2];
  for (int g = 0; g < G; ++g) {
    const float4 d = reinterpret_cast<const float4*>(drgb_all + (size_t)g * drgb_stride)[i];
    if (d.x == 0.f && d.y == 0.f && d.z == 0.f && d.w == 0.f) continue;
    sh_basis_accumulate(acc, D, mx - campos_all[3 * g], my - campos_all[3 * g + 1], mz - campos_all[3 * g + 2], d);
  }
  float4* out = reinterpret_cast<float4*>(dL_dsh) + (size_t)i * M;
#pragma unroll
  for (int k = 0; k < 16; ++k)
    if (k < M) out[k] = acc[k];
  for (int k = 16; k < M; ++k) out[k] = zero4;
}

int launch_sh_expand(int P, int D, int M, int G, const float* means3D, const float* campos_all, const float* drgb_all,
                     size_t drgb_stride, float* dL_dsh, cudaStream_t st) {
  if (P == 0 || M == 0) return 0;
  k_sh_expand<<<(P + 255) / 256, 256, 0, st>>>(P, D, M, G, means3D, campos_all, drgb_all, drgb_stride, dL_dsh);
  return check_cuda(cudaGetLastError(), "k_sh_expand launch");
}

// ---- SH expansion straight from the peers' factor buffers --------------------------------------------------------
// gsl_sh_expand over the factor tables the ranks pushed into this rank's buffer (local reads only), rows [row0, row1).
__global__ void __launch_bounds__(256) k_peer_sh_expand(const PeerView pv, const PeerLayout pl, int row0, int row1, int D,
                                                        int M, int prezeroed, const float* __restrict__ means3D,
                                                        const GlueMean gm, float* __restrict__ dL_dsh) {
  const int i = row0 + blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int word = i >> 5;  // warp-uniform (row0 is a multiple of 256)
  const uint2* fmeta = reinterpret_cast<const uint2*>(pv.own + pl.off_fmeta);
  const float4* factor = reinterpret_cast<const float4*>(pv.own + pl.off_factor);
  uint2 mine = make_uint2(0u, 0u);
  if (lane < pv.world && word * 32 < row1) mine = fmeta[(size_t)(pv.parity * pv.world + lane) * pl.tiles * 8 + word];
  float4 d[PEER_MAX];
#pragma unroll
  for (int g = 0; g < PEER_MAX; ++g) {
    d[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g < pv.world) {  // warp-uniform
      const uint32_t bits = __shfl_sync(0xffffffffu, mine.x, g);
      const uint32_t before = __shfl_sync(0xffffffffu, mine.y, g);
      if ((bits >> lane) & 1u)
        d[g] = factor[((size_t)(pv.parity * pv.world + g) * pl.tiles + (word >> 3)) * 256 + before +
                      __popc(bits & ((1u << lane) - 1u))];
    }
  }
  float4 acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  bool some = false;  // a rank has a non-zero factor for this surfel
#pragma unroll
  for (int g = 0; g < PEER_MAX; ++g) some |= (d[g].x != 0.f) | (d[g].y != 0.f) | (d[g].z != 0.f) | (d[g].w != 0.f);
  // rows to store: all of them, or -- when dL_dsh was zero-filled under the backward compositor -- the non-zero ones
  const uint32_t rowmask = prezeroed ? __ballot_sync(0xffffffffu, some) : 0xffffffffu;
  if (rowmask == 0u) return;
  const bool in_range = i < row1;
  const size_t ic = in_range ? (size_t)i : 0;
  float mx = 0.f, my = 0.f, mz = 0.f;
  GlueRow grow = {};
  if (gm.xyz) grow = glue_row(gm, ic);
  else { mx = means3D[3 * ic]; my = means3D[3 * ic + 1]; mz = means3D[3 * ic + 2]; }
  const float4* campos_all = reinterpret_cast<const float4*>(pv.own + PEER_CAMPOS_ALL_OFF) + pv.parity * PEER_MAX;
  const float4* stamp_all = reinterpret_cast<const float4*>(pv.own + PEER_GLUE_ALL_OFF) + pv.parity * PEER_MAX;
#pragma unroll
  for (int g = 0; g < PEER_MAX; ++g) {
    const float4 dg = d[g];
    if (g >= pv.world || (dg.x == 0.f && dg.y == 0.f && dg.z == 0.f && dg.w == 0.f)) continue;
    const float4 cpos = campos_all[g];
    if (gm.xyz) {  // rank g's basis is evaluated where rank g rasterized the surfel (its own timestamp)
      const float3 m = glue_mean_of(gm, grow, stamp_all[g]);
      mx = m.x; my = m.y; mz = m.z;
    }
    sh_basis_accumulate(acc, D, mx - cpos.x, my - cpos.y, mz - cpos.z, dg);
  }
  // coalesced stores: a warp's 32 rows x 16 float4 go through shared memory, 8 coefficients at a time, so that one
  // store instruction writes 4 rows x 128 contiguous bytes instead of 32 pieces of 16 bytes 256 bytes apart
  __shared__ float4 s_t[8][32][9];
  const int warp = threadIdx.x >> 5;
  const int row_base = i - lane;  // first row of this warp
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k) s_t[warp][lane][k] = acc[8 * half + k];
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int idx = it * 32 + lane;
      const int r = idx >> 3, q = idx & 7;
      const int row = row_base + r, k = 8 * half + q;
      if (row < row1 && k < M && ((rowmask >> r) & 1u)) reinterpret_cast<float4*>(dL_dsh)[(size_t)row * M + k] = s_t[warp][r][q];
    }
  }
  if (i < row1 && !prezeroed) {
    float4* out = reinterpret_cast<float4*>(dL_dsh) + (size_t)i * M;
    for (int k = 16; k < M; ++k) out[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}


// The same expansion for a fused step (dL_dsh was zero-filled under the backward compositor, so only rows with a factor are
// written).  About half of the surfels have no factor on any rank and most of the others on only some ranks, so the kernel
// above idles half of its lanes through up to 8 x ~150 instructions.  Here a CTA (= one 256-surfel tile) first waits for the
// "pushed" flags of the step (in-kernel barrier), builds the list of rows with a factor on SOME rank (OR of the ranks' bit
// words, prefix over the 8 words) and only ceil(n / 32) dense warps do the work.
__global__ void __launch_bounds__(256, 2) k_peer_sh_expand_tiles(PeerView pv, const PeerLayout pl, int tile0, int row1, int D,
                                                                 int M, int fused, int wait_slot,
                                                                 const float* __restrict__ means3D, const GlueMean gm,
                                                                 float* __restrict__ dL_dsh) {
  __shared__ uint2 s_meta[PEER_MAX][8];
  __shared__ uint32_t s_union[8];
  __shared__ int s_pref[9];
  __shared__ uint16_t s_list[256];
  __shared__ float4 s_t[8][32][9];
  if (fused) {
    peer_resolve_step(pv);
    peer_wait_flags(pv, wait_slot);
  }
  const int tile = tile0 + blockIdx.x;
  const int base_row = tile * 256;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint2* fmeta = reinterpret_cast<const uint2*>(pv.own + pl.off_fmeta);
  const float4* factor = reinterpret_cast<const float4*>(pv.own + pl.off_factor);
  if (threadIdx.x < 8 * PEER_MAX) {
    const int g = threadIdx.x >> 3, w = threadIdx.x & 7;
    uint2 m = make_uint2(0u, 0u);
    if (g < pv.world && base_row + 32 * w < row1) m = fmeta[((size_t)(pv.parity * pv.world + g) * pl.tiles + tile) * 8 + w];
    s_meta[g][w] = m;
  }
  __syncthreads();
  if (threadIdx.x < 32) {  // union of the ranks' bit words + exclusive prefix of their populations (one warp)
    uint32_t u = 0u;
    if (lane < 8) {
#pragma unroll
      for (int g = 0; g < PEER_MAX; ++g) u |= s_meta[g][lane].x;
    }
    const int c = __popc(u);
    int inc = c;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane < 8) { s_union[lane] = u; s_pref[lane] = inc - c; }
    if (lane == 7) s_pref[8] = inc;
  }
  __syncthreads();
  {
    const uint32_t u = s_union[warp];
    if ((u >> lane) & 1u) s_list[s_pref[warp] + __popc(u & ((1u << lane) - 1u))] = (uint16_t)threadIdx.x;
  }
  __syncthreads();
  const int n = s_pref[8];
  if (warp * 32 >= n) return;  // whole warps only: the listed rows fill the first ceil(n / 32) warps
  const bool valid = (int)threadIdx.x < n;
  const int r = valid ? (int)s_list[threadIdx.x] : 0;
  const size_t row = (size_t)base_row + r;
  const bool failed = peer_error(pv);
  float4 d[PEER_MAX];
#pragma unroll
  for (int g = 0; g < PEER_MAX; ++g) {
    d[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g < pv.world) {
      const uint2 m = s_meta[g][r >> 5];
      if (valid && ((m.x >> (r & 31)) & 1u))
        d[g] = factor[((size_t)(pv.parity * pv.world + g) * pl.tiles + tile) * 256 + m.y + __popc(m.x & ((1u << (r & 31)) - 1u))];
    }
  }
  float4 acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  float mx = 0.f, my = 0.f, mz = 0.f;
  GlueRow grow = {};
  if (gm.xyz) grow = glue_row(gm, row);
  else { mx = means3D[3 * row]; my = means3D[3 * row + 1]; mz = means3D[3 * row + 2]; }
  const float4* campos_all = reinterpret_cast<const float4*>(pv.own + PEER_CAMPOS_ALL_OFF) + pv.parity * PEER_MAX;
  const float4* stamp_all = reinterpret_cast<const float4*>(pv.own + PEER_GLUE_ALL_OFF) + pv.parity * PEER_MAX;
#pragma unroll
  for (int g = 0; g < PEER_MAX; ++g) {
    const float4 dg = d[g];
    if (g >= pv.world || (dg.x == 0.f && dg.y == 0.f && dg.z == 0.f && dg.w == 0.f)) continue;
    const float4 cpos = campos_all[g];
    if (gm.xyz) {  // rank g's basis is evaluated where rank g rasterized the surfel (its own timestamp)
      const float3 m = glue_mean_of(gm, grow, stamp_all[g]);
      mx = m.x; my = m.y; mz = m.z;
    }
    sh_basis_accumulate(acc, D, mx - cpos.x, my - cpos.y, mz - cpos.z, dg);
  }
  if (failed) {  // a rank missed a barrier of this step: never return silently wrong gradients
    const float qnan = __int_as_float(0x7fc00000);
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = make_float4(qnan, qnan, qnan, qnan);
  }
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k) s_t[warp][lane][k] = acc[8 * half + k];
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int idx = it * 32 + lane;
      const int rr = idx >> 3, q = idx & 7;
      const int cid = warp * 32 + rr, k = 8 * half + q;
      if (cid < n && k < M)
        reinterpret_cast<float4*>(dL_dsh)[((size_t)base_row + s_list[cid]) * M + k] = s_t[warp][rr][q];
    }
  }
}

int launch_peer_sh_expand(const gsl_peer_ctx* c, int P, int S, int D, int M, int row0, int row1, bool prezeroed,
                          const float* means3D, float* dL_dsh, cudaStream_t st, bool fused, int wait_slot) {
  if (row1 <= row0 || M == 0) return 0;
  if (prezeroed && M <= 16) {
    k_peer_sh_expand_tiles<<<(row1 - row0 + 255) / 256, 256, 0, st>>>(make_view(c, fused), peer_layout((size_t)P, S, c->world),
                                                                     row0 >> 8, row1, D, M, fused ? 1 : 0, wait_slot, means3D, make_glue_mean(c), dL_dsh);
    return check_cuda(cudaGetLastError(), "k_peer_sh_expand_tiles launch");
  }
  if (fused) return set_error(GSL_EINVAL, "peer_sh_expand: the fused step needs zero-filled outputs and M <= 16");
  k_peer_sh_expand<<<(row1 - row0 + 255) / 256, 256, 0, st>>>(make_view(c), peer_layout((size_t)P, S, c->world), row0, row1, D,
                                                             M, prezeroed ? 1 : 0, means3D, make_glue_mean(c), dL_dsh);
  return check_cuda(cudaGetLastError(), "k_peer_sh_expand launch");
}

// ------------------------------------------------------------------------------------------------
// markVisible (pinhole test kept for API parity; never on the training path)
// ------------------------------------------------------------------------------------------------
__global__ void k_mark_visible(int P, const float* __restrict__ pts, const float* __restrict__ vm,
                               const float* __restrict__ pm, uint8_t* __restrict__ present) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  float x = pts[3 * idx], y = pts[3 * idx + 1], z = pts[3 * idx + 2];
  float hx = pm[0] * x + pm[4] * y + pm[8] * z + pm[12];
  float hy = pm[1] * x + pm[5] * y + pm[9] * z + pm[13];
  float hw = pm[3] * x + pm[7] * y + pm[11] * z + pm[15];
  float pw = 1.0f / (hw + 0.0000001f);
  float projx = hx * pw, projy = hy * pw;
  float vz = vm[2] * x + vm[6] * y + vm[10] * z + vm[14];
  bool out = vz <= 0.2f || (projx < -1.3 || projx > 1.3 || projy < -1.3 || projy > 1.3);
  present[idx] = out ? 0 : 1;
}

int launch_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                        uint8_t* present, cudaStream_t st) {
  if (P == 0) return 0;
  k_mark_visible<<<(P + 255) / 256, 256, 0, st>>>(P, means3D, viewmatrix, projmatrix, present);
  return check_cuda(cudaGetLastError(), "k_mark_visible launch");
}

}  // namespace gsl
