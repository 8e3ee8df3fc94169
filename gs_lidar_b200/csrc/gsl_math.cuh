// gsl_math.cuh -- device math shared by the preprocess / compositing kernels.
//
// PARITY NOTE.  Tile keys embed float_bits(depth) and the tile rect depends on ceil()/trunc() of
// float expressions, and the compositors branch on float thresholds (alpha < 1/255, T < 1e-4 ...),
// so "same formula" is not enough: the rounding of every intermediate has to be the one the
// reference binary performs.  Where the order matters the code below therefore spells out each
// rounding with __fmul_rn / __fmaf_rn / __fadd_rn / __fdiv_rn (never contracted by nvcc/ptxas)
// following the dataflow of the reference kernels as compiled by the same nvcc 12.9 for sm_100a
// (read from its PTX/SASS; see DESIGN.md "bit-exactness").  The formulas themselves are the
// published 2DGS ray-splat intersection specialised to an equirectangular camera
// (reference: cuda_rasterizer/forward.cu:73-171,397-447).
#pragma once
#include <cuda_runtime.h>

namespace gsl {

#define GSL_FM(a, b) __fmul_rn((a), (b))
#define GSL_FA(a, b) __fadd_rn((a), (b))
#define GSL_FS(a, b) __fsub_rn((a), (b))
#define GSL_FF(a, b, c) __fmaf_rn((a), (b), (c))
#define GSL_FD(a, b) __fdiv_rn((a), (b))

// 1/x to ~1 ulp in one MUFU instruction (backward pass only; the forward pass keeps IEEE divisions)
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// a*b + c*d + e*f as the reference evaluates 3-term dot products:
// the middle product is rounded on its own, the others are fused.
__device__ __forceinline__ float dot3_ref(float a, float b, float c, float d, float e, float f) {
  return GSL_FF(e, f, GSL_FF(a, b, GSL_FM(c, d)));
}

// Per-pixel constants of the equirectangular camera (forward.cu:404-405).
struct PixelRay {
  float sphi, cphi, sth, cth;
  float sphi_cth;  // sin(phi)*cos(theta)
  float cphi_cth;  // cos(phi)*cos(theta)
  float px, py;    // float pixel coordinates (integer valued)
  float wrapW;     // > 0 in azimuth wrap-around mode: the period of the panorama in pixels
};

__device__ __forceinline__ PixelRay make_pixel_ray(float px, float py, float hfov_min, float hfov_max,
                                                   float vfov_min, float vfov_max, int W, int H) {
  PixelRay r;
  r.px = px;
  r.py = py;
  r.wrapW = 0.f;
  // phi = pixf.x * (HFOV_max - HFOV_min) / W + HFOV_min : mul, div, add (no contraction possible)
  float phi = GSL_FA(GSL_FD(GSL_FM(GSL_FS(hfov_max, hfov_min), px), (float)W), hfov_min);
  float theta = GSL_FA(GSL_FD(GSL_FM(GSL_FS(vfov_max, vfov_min), py), (float)H), vfov_min);
  r.sphi = sinf(phi);
  r.cphi = cosf(phi);
  r.sth = sinf(theta);
  r.cth = cosf(theta);
  r.sphi_cth = GSL_FM(r.sphi, r.cth);
  r.cphi_cth = GSL_FM(r.cphi, r.cth);
  return r;
}

// One surfel record as staged for the compositors (see gsl_common.cuh for the layout).
struct Splat {
  float Tux, Tuy, Tuz, Tvx, Tvy, Tvz, Twx, Twy, Twz;
  float mx, my, opacity;
  float nx, ny, nz, depth;
};

__device__ __forceinline__ Splat load_splat(const float4* __restrict__ rec) {
  float4 a = rec[0], b = rec[1], c = rec[2], d = rec[3];
  Splat s;
  s.Tux = a.x; s.Tuy = a.y; s.Tuz = a.z; s.Tvx = a.w;
  s.Tvy = b.x; s.Tvz = b.y; s.Twx = b.z; s.Twy = b.w;
  s.Twz = c.x; s.mx = c.y; s.my = c.z; s.opacity = c.w;
  s.nx = d.x; s.ny = d.y; s.nz = d.z; s.depth = d.w;
  return s;
}

// Result of evaluating one (pixel, surfel) pair up to the blending weight.
struct PairEval {
  float sx, sy;        // splat-space intersection
  float rho3d, rho2d;
  float dx, dy;        // means2D - pixel
  float pz;            // homogeneous w of the intersection
  float kx, ky, kz, lx, ly, lz;
  float depth;
  float G;             // exp(-rho/2)
  float alpha;
  bool valid;
};

// Ray-splat intersection, low-pass filter, depth and alpha for one pair, with the skip tests of
// forward.cu:410-441 (identical in backward.cu:308-339).  `near`/`far` are 2*sf and 300*sf.
// The rounding sequence is the reference forward kernel's.  Written branch-free (the skip tests only
// feed `valid`) so that the compiler can interleave the evaluation of consecutive candidates; values
// computed past a failed test are never used.
// FAST (backward pass only): the two IEEE divisions and expf are replaced by their approximate forms.  The
// backward never re-decides which pairs contribute (the forward's pair masks do), so a last-ulp difference
// cannot flip a threshold there; it only perturbs gradients at the 1e-7 level (contract: 1e-4).
template <bool KEEP_KL, bool FAST = false>
__device__ __forceinline__ PairEval eval_pair(const Splat& s, const PixelRay& r, float near_, float far_) {
  PairEval e;
  // k = cos(phi)*Tu - sin(phi)*Tw            (x,y,z components over Tu/Tv... columns)
  float kx = GSL_FF(s.Tux, r.cphi, -GSL_FM(s.Twx, r.sphi));
  float ky = GSL_FF(s.Tuy, r.cphi, -GSL_FM(s.Twy, r.sphi));
  float kz = GSL_FF(s.Tuz, r.cphi, -GSL_FM(s.Twz, r.sphi));
  // l = sin(phi)cos(theta)*Tu + sin(theta)*Tv + cos(phi)cos(theta)*Tw
  float lx = GSL_FF(s.Twx, r.cphi_cth, GSL_FF(s.Tux, r.sphi_cth, GSL_FM(s.Tvx, r.sth)));
  float ly = GSL_FF(s.Twy, r.cphi_cth, GSL_FF(s.Tuy, r.sphi_cth, GSL_FM(s.Tvy, r.sth)));
  float lz = GSL_FF(s.Twz, r.cphi_cth, GSL_FF(s.Tuz, r.sphi_cth, GSL_FM(s.Tvz, r.sth)));
  // p = k x l
  float px = GSL_FF(ky, lz, -GSL_FM(kz, ly));
  float py = GSL_FF(kz, lx, -GSL_FM(kx, lz));
  float pz = GSL_FF(kx, ly, -GSL_FM(ky, lx));
  if (KEEP_KL) { e.kx = kx; e.ky = ky; e.kz = kz; e.lx = lx; e.ly = ly; e.lz = lz; }
  e.pz = pz;
  bool ok = pz != 0.0f;
  float sx, sy;
  if (FAST) {
    const float ipz = fast_rcp(pz);
    sx = px * ipz;
    sy = py * ipz;
  } else {
    sx = GSL_FD(px, pz);
    sy = GSL_FD(py, pz);
  }
  float rho3d = GSL_FF(sx, sx, GSL_FM(sy, sy));
  float dx = GSL_FS(s.mx, r.px);
  if (r.wrapW > 0.f) {  // wrap-around mode: distance to the nearest periodic image of the projected centre
    if (dx > 0.5f * r.wrapW) dx -= r.wrapW;
    else if (dx < -0.5f * r.wrapW) dx += r.wrapW;
  }
  float dy = GSL_FS(s.my, r.py);
  float h = GSL_FF(dx, dx, GSL_FM(dy, dy));
  float rho2d = GSL_FA(h, h);  // FilterInvSquare (=2) * |d|^2
  float sTu = GSL_FA(s.Tuz, GSL_FF(s.Tux, sx, GSL_FM(s.Tuy, sy)));
  float sTv = GSL_FA(s.Tvz, GSL_FF(s.Tvx, sx, GSL_FM(s.Tvy, sy)));
  float sTw = GSL_FA(s.Twz, GSL_FF(s.Twx, sx, GSL_FM(s.Twy, sy)));
  // depth_3d = s_Tu*sin(th)*sin(ph) - s_Tv*cos(th) + s_Tw*sin(th)*cos(ph)
  float d3 = GSL_FF(-sTv, r.cth, GSL_FM(GSL_FM(sTu, r.sth), r.sphi));
  d3 = GSL_FF(GSL_FM(sTw, r.sth), r.cphi, d3);
  float depth = (rho3d <= rho2d) ? d3 : s.depth;
  e.sx = sx; e.sy = sy; e.rho3d = rho3d; e.rho2d = rho2d; e.dx = dx; e.dy = dy; e.depth = depth;
  ok = ok && !(depth < near_ || depth > far_);
  float power = GSL_FM(fminf(rho3d, rho2d), -0.5f);
  ok = ok && !(power > 0.0f);
  float G = expf(power);  // exact also in the backward: G scales every gradient term directly
  float alpha = fminf(GSL_FM(s.opacity, G), 0.99f);
  e.G = G;
  e.alpha = alpha;
  e.valid = ok && !(alpha < 1.0f / 255.0f);
  return e;
}

}  // namespace gsl
