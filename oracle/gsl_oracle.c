/*
 * oracle/gsl_oracle.c -- TEST INFRASTRUCTURE ONLY (never imported by the product package).
 *
 * A plain-C, CPU restatement of the reference's panoramic surfel rasterizer, stage by stage.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it, and only as the checker or the timed CPU baseline.
 *
 * Each function cites the reference lines it restates (paths relative to
 * /root/reference/diff-gaussian-rasterization-2d/).  Arithmetic is float32 with the reference's
 * float<->double promotions; libm's sinf/atan2f/expf differ from CUDA's libdevice in the last ulp
 * and x86 does not contract a*b+c, so comparisons against the CUDA path are tolerance-based for
 * floats and exact for the integer stages when fed identical (depth, rect) inputs.
 *
 * PINNING: the oracle is pinned against outputs of the reference's own CUDA kernels
 * (oracle/_ref/libgslidar_ref.so run on a B200) committed under tests/golden/ -- see
 * tests/golden/README.md and tests/test_oracle_golden.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_PI 3.14159265 /* auxiliary.h:17 */
#define BLOCK_X 16        /* config.h:13 */
#define BLOCK_Y 16        /* config.h:14 */

static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

typedef struct {
  int P, S, D, M, W, H;
  float vfov_min, vfov_max, hfov_min, hfov_max; /* degrees */
  float scale_factor;
  float tanfovx, tanfovy;
  int wrap; /* != 0: azimuth wrap-around mode (this repo's opt-in extension, NOT the reference's semantics): the
               reference has getRect_panorama (auxiliary.h:67-75) but never calls it; the mode restated here is the
               periodic panorama of DESIGN.md 5b -- AABB samples measured relative to the centre azimuth, modular
               tile-column ranges, low-pass distance to the nearest periodic image */
} orc_params;

typedef struct { float VFOV_min, VFOV_max, HFOV_min, HFOV_max; } orc_fov;

/* forward.cu:221-226 (double arithmetic on float inputs, rounded to float) */
static orc_fov fov_consts(const orc_params* p) {
  orc_fov f;
  f.VFOV_max = (float)(ORC_PI / 2 - p->vfov_min * ORC_PI / 180);
  f.VFOV_min = (float)(ORC_PI / 2 - p->vfov_max * ORC_PI / 180);
  f.HFOV_max = (float)(p->hfov_max * ORC_PI / 180);
  f.HFOV_min = (float)(p->hfov_min * ORC_PI / 180);
  return f;
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* auxiliary.h:206-228 quat_to_rotmat; R[c][r] column-major like glm */
static void quat_to_rot(const float* q, float R[3][3]) {
  float s = 1.0f / sqrtf(q[3] * q[3] + q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
  float w = q[0] * s, x = q[1] * s, y = q[2] * s, z = q[3] * s;
  R[0][0] = 1.f - 2.f * (y * y + z * z); R[0][1] = 2.f * (x * y + w * z); R[0][2] = 2.f * (x * z - w * y);
  R[1][0] = 2.f * (x * y - w * z); R[1][1] = 1.f - 2.f * (x * x + z * z); R[1][2] = 2.f * (y * z + w * x);
  R[2][0] = 2.f * (x * z + w * y); R[2][1] = 2.f * (y * z - w * x); R[2][2] = 1.f - 2.f * (x * x + y * y);
}

/* auxiliary.h:47-55 getRect */
static void get_rect(float px, float py, int radius, int gx, int gy, int* rect /*minx,miny,maxx,maxy*/) {
  int v;
  v = (int)((px - radius) / BLOCK_X); rect[0] = v < 0 ? 0 : (v > gx ? gx : v);
  v = (int)((py - radius) / BLOCK_Y); rect[1] = v < 0 ? 0 : (v > gy ? gy : v);
  v = (int)((px + radius + BLOCK_X - 1) / BLOCK_X); rect[2] = v < 0 ? 0 : (v > gx ? gx : v);
  v = (int)((py + radius + BLOCK_Y - 1) / BLOCK_Y); rect[3] = v < 0 ? 0 : (v > gy ? gy : v);
}

/* wrap-around mode: tile columns [rect[0], rect[2]) as a modular range over gx columns (rect[2] may exceed gx: column
   = x mod gx); the part of the footprint left of pixel 0 continues at pixel W, the part right of pixel W-1 at pixel 0;
   rows as in get_rect. */
static void get_rect_wrap(float px, float py, int radius, int gx, int gy, int W, int* rect) {
  get_rect(px, py, radius, gx, gy, rect);
  const float rf = (float)radius, Wf = (float)W;
  const int tx0 = (int)floorf((px - rf) * 0.0625f), tx1 = (int)floorf((px + rf + 15.f) * 0.0625f);
  const int left = (px - rf) < 0.f, right = (px + rf) >= Wf;
  int start = tx0 > 0 ? tx0 : 0;
  int len = (tx1 < gx ? tx1 : gx) - start;
  if (left) {
    int a0 = (int)floorf((Wf + px - rf) * 0.0625f);
    a0 = a0 < 0 ? 0 : (a0 > gx ? gx : a0);
    start = a0;
    len += gx - a0;
  }
  if (right) {
    int b1 = (int)floorf((px + rf - Wf + 15.f) * 0.0625f);
    b1 = b1 < 0 ? 0 : (b1 > gx ? gx : b1);
    len += b1;
  }
  if ((left && right) || len >= gx) { start = 0; len = gx; }
  rect[0] = start;
  rect[2] = start + (len > 0 ? len : 0);
}

/* forward.cu:17-69 computeColorFromSH (4 channels) */
static void sh_to_color(int deg, const float* sh /*M x 4*/, const float* pos, const float* campos, float* out,
                        uint8_t* clamped) {
  float dx = pos[0] - campos[0], dy = pos[1] - campos[1], dz = pos[2] - campos[2];
  float len = sqrtf(dx * dx + dy * dy + dz * dz);
  float x = dx / len, y = dy / len, z = dz / len;
  for (int c = 0; c < 4; ++c) {
    float r = SH_C0 * sh[0 * 4 + c];
    if (deg > 0) {
      r = r - SH_C1 * y * sh[1 * 4 + c] + SH_C1 * z * sh[2 * 4 + c] - SH_C1 * x * sh[3 * 4 + c];
      if (deg > 1) {
        float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        r = r + SH_C2[0] * xy * sh[4 * 4 + c] + SH_C2[1] * yz * sh[5 * 4 + c] +
            SH_C2[2] * (2.0f * zz - xx - yy) * sh[6 * 4 + c] + SH_C2[3] * xz * sh[7 * 4 + c] +
            SH_C2[4] * (xx - yy) * sh[8 * 4 + c];
        if (deg > 2) {
          r = r + SH_C3[0] * y * (3.0f * xx - yy) * sh[9 * 4 + c] + SH_C3[1] * xy * z * sh[10 * 4 + c] +
              SH_C3[2] * y * (4.0f * zz - xx - yy) * sh[11 * 4 + c] +
              SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[12 * 4 + c] +
              SH_C3[4] * x * (4.0f * zz - xx - yy) * sh[13 * 4 + c] + SH_C3[5] * z * (xx - yy) * sh[14 * 4 + c] +
              SH_C3[6] * x * (xx - 3.0f * yy) * sh[15 * 4 + c];
        }
      }
    }
    r += 0.5f;
    clamped[c] = r < 0;
    out[c] = r > 0.f ? r : 0.f;
  }
}

/*
 * K1 preprocess, forward.cu:173-287 (+ compute_transmat :73-113, computePanoramaCoordinate :116-125,
 * compute_aabb :129-171, in_frustum_panorama auxiliary.h:182-204).
 * Outputs follow the reference's GeometryState arrays; slots of culled surfels are left untouched
 * except radii / tiles_touched (= 0), like the reference.
 */
void orc_preprocess(const orc_params* p, const float* means3D, const float* scales, const float* rotations,
                    const float* opacities, const float* shs, const float* colors_precomp, const uint8_t* mask,
                    const float* vm, const float* campos, int* radii, float* means2D, float* depths,
                    float* transMat, float* normal_opacity, float* rgb, uint8_t* clamped, uint32_t* tiles_touched) {
  const orc_fov f = fov_consts(p);
  const int gx = (p->W + BLOCK_X - 1) / BLOCK_X, gy = (p->H + BLOCK_Y - 1) / BLOCK_Y;
  const int W = p->W, H = p->H;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < p->P; ++i) {
    radii[i] = 0;
    tiles_touched[i] = 0;
    const float* po = means3D + 3 * i;
    float opacity = opacities[i];
    /* auxiliary.h:77-85 transformPoint4x3 */
    float tx = vm[0] * po[0] + vm[4] * po[1] + vm[8] * po[2] + vm[12];
    float ty = vm[1] * po[0] + vm[5] * po[1] + vm[9] * po[2] + vm[13];
    float tz = vm[2] * po[0] + vm[6] * po[1] + vm[10] * po[2] + vm[14];
    float phi = atan2f(tx, tz);
    float theta = atan2f(sqrtf(tx * tx + tz * tz), -ty);
    float r = sqrtf(tx * tx + ty * ty + tz * tz);
    if (!mask[i]) continue;
    {
      float center_v = (f.VFOV_max + f.VFOV_min) / 2, half_v = (f.VFOV_max - f.VFOV_min) / 2;
      float ratio_v = fabsf((theta - center_v) / half_v);
      float center_h = (f.HFOV_max + f.HFOV_min) / 2, half_h = (f.HFOV_max - f.HFOV_min) / 2;
      float ratio_h = fabsf((phi - center_h) / half_h);
      if (r <= 2.0f * p->scale_factor || (double)ratio_v > 1.3 || (double)ratio_h > 1.3) continue;
    }
    float R[3][3];
    quat_to_rot(rotations + 4 * i, R);
    float sx = scales[3 * i], sy = scales[3 * i + 1];
    float L0[3] = {R[0][0] * sx, R[0][1] * sx, R[0][2] * sx};
    float L1[3] = {R[1][0] * sy, R[1][1] * sy, R[1][2] * sy};
    float L2[3] = {R[2][0], R[2][1], R[2][2]};
    float T[9]; /* Tu(3), Tv(3), Tw(3) = transMat rows, forward.cu:238-241 */
    for (int c = 0; c < 3; ++c) {
      T[3 * c + 0] = L0[0] * vm[c] + L0[1] * vm[4 + c] + L0[2] * vm[8 + c];
      T[3 * c + 1] = L1[0] * vm[c] + L1[1] * vm[4 + c] + L1[2] * vm[8 + c];
    }
    T[2] = tx; T[5] = ty; T[8] = tz;
    float nx = vm[0] * L2[0] + vm[4] * L2[1] + vm[8] * L2[2];
    float ny = vm[1] * L2[0] + vm[5] * L2[1] + vm[9] * L2[2];
    float nz = vm[2] * L2[0] + vm[6] * L2[1] + vm[10] * L2[2];
    float mult = (nx * tx + ny * ty + nz * tz) < 0 ? 1.f : -1.f; /* forward.cu:108-112 */
    nx *= mult; ny *= mult; nz *= mult;
    memcpy(transMat + 9 * (size_t)i, T, sizeof(T));

    float cutoff = sqrtf((float)fmax((double)(9.f + 2.f * logf(opacity)), 0.000001)); /* forward.cu:243 */
    float minx = INFINITY, miny = INFINITY, maxx = -INFINITY, maxy = -INFINITY;
    const float cx0 = (phi - f.HFOV_min) * W / (f.HFOV_max - f.HFOV_min);
    for (int k = 0; k < 12; ++k) { /* forward.cu:153-168 */
      float a = (float)(2 * ORC_PI * k / 12);
      float vx = cutoff * sinf(a), vy = cutoff * cosf(a);
      float X = T[0] * vx + T[1] * vy + T[2];
      float Y = T[3] * vx + T[4] * vy + T[5];
      float Z = T[6] * vx + T[7] * vy + T[8];
      float ph = atan2f(X, Z);
      float th = atan2f(sqrtf(X * X + Z * Z), -Y);
      float ppx = (ph - f.HFOV_min) * W / (f.HFOV_max - f.HFOV_min);
      if (p->wrap) { /* azimuth of the sample relative to the centre: a splat on the seam keeps its true extent */
        float dphi = ph - phi;
        if (dphi > 3.14159265f) dphi -= 6.2831853f;
        else if (dphi < -3.14159265f) dphi += 6.2831853f;
        ppx = cx0 + dphi * (float)W / (f.HFOV_max - f.HFOV_min);
      }
      float ppy = (th - f.VFOV_min) * H / (f.VFOV_max - f.VFOV_min);
      minx = fminf(minx, ppx); maxx = fmaxf(maxx, ppx);
      miny = fminf(miny, ppy); maxy = fmaxf(maxy, ppy);
    }
    float cx = (phi - f.HFOV_min) * W / (f.HFOV_max - f.HFOV_min);
    float cy = (theta - f.VFOV_min) * H / (f.VFOV_max - f.VFOV_min);
    float rad = fmaxf(fmaxf(maxx - cx, cx - minx), fmaxf(maxy - cy, cy - miny));
    if ((double)rad < 0.3) continue;
    int my_radius = (int)ceilf(rad);
    int rect[4];
    if (p->wrap) get_rect_wrap(cx, cy, my_radius, gx, gy, W, rect);
    else get_rect(cx, cy, my_radius, gx, gy, rect);
    int area = (rect[2] - rect[0]) * (rect[3] - rect[1]);
    if (area == 0) continue;
    if (colors_precomp == NULL) sh_to_color(p->D, shs + (size_t)i * p->M * 4, po, campos, rgb + 4 * (size_t)i, clamped + 4 * (size_t)i);
    depths[i] = r;
    radii[i] = my_radius;
    means2D[2 * i] = cx; means2D[2 * i + 1] = cy;
    normal_opacity[4 * i] = nx; normal_opacity[4 * i + 1] = ny; normal_opacity[4 * i + 2] = nz;
    normal_opacity[4 * i + 3] = opacity;
    tiles_touched[i] = (uint32_t)area;
  }
}

/* rasterizer_impl.cu:32-47 */
static uint32_t higher_msb(uint32_t n) {
  uint32_t msb = sizeof(n) * 4, step = msb;
  while (step > 1) {
    step /= 2;
    if (n >> msb) msb += step; else msb -= step;
  }
  if (n >> msb) msb++;
  return msb;
}

/*
 * K2-K7 binning, rasterizer_impl.cu:310-354 (duplicateWithKeys :68-111, identifyTileRanges :116-142).
 * point_offsets: inclusive scan.  keys/vals must hold R = offsets[P-1] entries (call once with
 * keys == NULL to get R).  Stable LSD radix sort over bits [0, 32+higher_msb(tiles)).
 */
int64_t orc_binning(const orc_params* p, const int* radii, const float* means2D, const float* depths,
                    const uint32_t* tiles_touched, uint32_t* point_offsets, uint64_t* keys, uint32_t* vals,
                    uint32_t* ranges /* tiles x 2 */) {
  const int gx = (p->W + BLOCK_X - 1) / BLOCK_X, gy = (p->H + BLOCK_Y - 1) / BLOCK_Y;
  uint32_t acc = 0;
  for (int i = 0; i < p->P; ++i) { acc += tiles_touched[i]; point_offsets[i] = acc; }
  const int64_t R = p->P > 0 ? acc : 0;
  if (keys == NULL) return R;
  for (int i = 0; i < p->P; ++i) {
    if (radii[i] <= 0) continue;
    uint32_t off = i == 0 ? 0 : point_offsets[i - 1];
    int rect[4];
    if (p->wrap) get_rect_wrap(means2D[2 * i], means2D[2 * i + 1], radii[i], gx, gy, p->W, rect);
    else get_rect(means2D[2 * i], means2D[2 * i + 1], radii[i], gx, gy, rect);
    uint32_t dbits;
    memcpy(&dbits, depths + i, 4);
    for (int y = rect[1]; y < rect[3]; ++y)
      for (int x = rect[0]; x < rect[2]; ++x) {
        keys[off] = ((uint64_t)(y * gx + (x >= gx ? x - gx : x)) << 32) | dbits;
        vals[off] = (uint32_t)i;
        off++;
      }
  }
  /* stable LSD radix sort, 8 bits per pass */
  const int end_bit = 32 + (int)higher_msb((uint32_t)(gx * gy));
  uint64_t* k2 = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(R > 0 ? R : 1));
  uint32_t* v2 = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(R > 0 ? R : 1));
  uint64_t *ka = keys, *kb = k2;
  uint32_t *va = vals, *vb = v2;
  for (int shift = 0; shift < end_bit; shift += 8) {
    int bits = end_bit - shift < 8 ? end_bit - shift : 8;
    uint32_t maskb = (1u << bits) - 1;
    size_t cnt[257];
    memset(cnt, 0, sizeof(cnt));
    for (int64_t i = 0; i < R; ++i) cnt[((ka[i] >> shift) & maskb) + 1]++;
    for (int b = 0; b < 256; ++b) cnt[b + 1] += cnt[b];
    for (int64_t i = 0; i < R; ++i) {
      size_t d = cnt[(ka[i] >> shift) & maskb]++;
      kb[d] = ka[i]; vb[d] = va[i];
    }
    uint64_t* tk = ka; ka = kb; kb = tk;
    uint32_t* tv = va; va = vb; vb = tv;
  }
  if (ka != keys) { memcpy(keys, ka, sizeof(uint64_t) * (size_t)R); memcpy(vals, va, sizeof(uint32_t) * (size_t)R); }
  free(k2); free(v2);
  memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)gx * gy);
  for (int64_t i = 0; i < R; ++i) {
    uint32_t cur = (uint32_t)(keys[i] >> 32);
    if (i == 0) ranges[2 * cur] = 0;
    else {
      uint32_t prev = (uint32_t)(keys[i - 1] >> 32);
      if (cur != prev) { ranges[2 * prev + 1] = (uint32_t)i; ranges[2 * cur] = (uint32_t)i; }
    }
    if (i == R - 1) ranges[2 * cur + 1] = (uint32_t)R;
  }
  return R;
}

typedef struct {
  float sx, sy, rho3d, rho2d, dx, dy, pz, depth, G, alpha;
  float k[3], l[3];
  int valid;
} pair_eval;

/* forward.cu:397-441 == backward.cu:296-339: ray-splat intersection, low-pass, depth, alpha + skips */
static void eval_pair(const float* T, const float* xy, float opa, float sdepth, float pxf, float pyf, float cph,
                      float sph, float cth, float sth, float near_, float far_, float wrapW, pair_eval* e) {
  const float* Tu = T; const float* Tv = T + 3; const float* Tw = T + 6;
  e->valid = 0;
  for (int c = 0; c < 3; ++c) {
    e->k[c] = cph * Tu[c] - sph * Tw[c];
    e->l[c] = sph * cth * Tu[c] + sth * Tv[c] + cph * cth * Tw[c];
  }
  float px = e->k[1] * e->l[2] - e->k[2] * e->l[1];
  float py = e->k[2] * e->l[0] - e->k[0] * e->l[2];
  float pz = e->k[0] * e->l[1] - e->k[1] * e->l[0];
  e->pz = pz;
  if (pz == 0.0f) return;
  float sx = px / pz, sy = py / pz;
  float rho3d = sx * sx + sy * sy;
  float dx = xy[0] - pxf, dy = xy[1] - pyf;
  if (wrapW > 0.f) { /* wrap-around mode: distance to the nearest periodic image of the projected centre */
    if (dx > 0.5f * wrapW) dx -= wrapW;
    else if (dx < -0.5f * wrapW) dx += wrapW;
  }
  float rho2d = 2.0f * (dx * dx + dy * dy);
  float rho = fminf(rho3d, rho2d);
  float sTu = sx * Tu[0] + sy * Tu[1] + Tu[2];
  float sTv = sx * Tv[0] + sy * Tv[1] + Tv[2];
  float sTw = sx * Tw[0] + sy * Tw[1] + Tw[2];
  float d3 = sTu * sth * sph - sTv * cth + sTw * sth * cph;
  float depth = (rho3d <= rho2d) ? d3 : sdepth;
  e->sx = sx; e->sy = sy; e->rho3d = rho3d; e->rho2d = rho2d; e->dx = dx; e->dy = dy; e->depth = depth;
  if (depth < near_ || depth > far_) return;
  float power = -0.5f * rho;
  if (power > 0.0f) return;
  float G = expf(power);
  float alpha = fminf(0.99f, opa * G);
  e->G = G; e->alpha = alpha;
  if (alpha < 1.0f / 255.0f) return;
  e->valid = 1;
}

/*
 * K8 forward compositing, forward.cu:292-505.  One pixel at a time (the per-pixel result does not
 * depend on the reference's 256-entry batching).  Outputs: final_T (3N: T, M1, M2), n_contrib (2N),
 * out_color (4N), out_feature ((S+3)N), out_depth (4N).
 */
void orc_render_forward(const orc_params* p, const uint32_t* ranges, const uint32_t* point_list, const float* means2D,
                        const float* colors, const float* features, const float* transMat, const float* depths,
                        const float* normal_opacity, const float* bg, float* final_T, int32_t* n_contrib,
                        float* out_color, float* out_feature, float* out_depth) {
  const orc_fov f = fov_consts(p);
  const int W = p->W, H = p->H, S = p->S, N = W * H;
  const int gx = (W + BLOCK_X - 1) / BLOCK_X;
  const float near_ = 2.0f * p->scale_factor, far_ = 300.0f * p->scale_factor;
#pragma omp parallel for schedule(dynamic, 64)
  for (int pix = 0; pix < N; ++pix) {
    const int x = pix % W, y = pix / W;
    const int tile = (y / BLOCK_Y) * gx + (x / BLOCK_X);
    const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
    const float pxf = (float)x, pyf = (float)y;
    const float phi = pxf * (f.HFOV_max - f.HFOV_min) / W + f.HFOV_min;
    const float theta = pyf * (f.VFOV_max - f.VFOV_min) / H + f.VFOV_min;
    const float sph = sinf(phi), cph = cosf(phi), sth = sinf(theta), cth = cosf(theta);
    float T = 1.0f, C[4] = {0}, F[13] = {0}, D = 0, D2 = 0, M1 = 0, M2 = 0, dist = 0, med = 0;
    int contributor = 0, last = 0, medc = 0;
    for (uint32_t q = r0; q < r1; ++q) {
      contributor++;
      const uint32_t id = point_list[q];
      pair_eval e;
      eval_pair(transMat + 9 * (size_t)id, means2D + 2 * (size_t)id, normal_opacity[4 * (size_t)id + 3], depths[id], pxf,
                pyf, cph, sph, cth, sth, near_, far_, p->wrap ? (float)p->W : 0.f, &e);
      if (!e.valid) continue;
      float alpha = e.alpha, depth = e.depth;
      float test_T = T * (1 - alpha);
      if (test_T < 0.0001f) break; /* done = true */
      float w = alpha * T;
      float A = 1 - T;
      float m = far_ / (far_ - near_) * (1 - near_ / depth);
      dist += (m * m * A + M2 - 2 * m * M1) * w;
      M1 += m * w;
      M2 += m * m * w;
      if (T > 0.5) { med = depth; medc = contributor; }
      for (int ch = 0; ch < 4; ++ch) C[ch] += colors[4 * (size_t)id + ch] * alpha * T;
      for (int ch = 0; ch < S + 3; ++ch) {
        if (ch < S) F[ch] += features[(size_t)id * S + ch] * alpha * T;
        else F[ch] += normal_opacity[4 * (size_t)id + ch - S] * alpha * T;
      }
      D += depth * alpha * T;
      D2 += depth * depth * alpha * T;
      T = test_T;
      last = contributor;
    }
    final_T[pix] = T; final_T[pix + N] = M1; final_T[pix + 2 * N] = M2;
    n_contrib[pix] = last; n_contrib[pix + N] = medc;
    for (int ch = 0; ch < 4; ++ch) out_color[ch * N + pix] = C[ch] + T * bg[ch];
    for (int ch = 0; ch < S + 3; ++ch) out_feature[ch * N + pix] = F[ch];
    out_depth[pix] = D; out_depth[N + pix] = med; out_depth[2 * N + pix] = dist; out_depth[3 * N + pix] = D2;
  }
}

static inline void atomic_addf(float* a, float v) {
#pragma omp atomic
  *a += v;
}

/*
 * K10 backward compositing, backward.cu:137-515.  Accumulates into dL_dtransMat (9P), dL_dmean2D (4P),
 * dL_dopacity (P), dL_dcolors (4P), dL_dfeatures (SP), dL_dnormals (3P); all must be zeroed by the caller.
 */
void orc_render_backward(const orc_params* p, const uint32_t* ranges, const uint32_t* point_list, const float* bg,
                         const float* means2D, const float* normal_opacity, const float* transMat,
                         const float* colors, const float* depths, const float* features, const float* final_T,
                         const int32_t* n_contrib, const float* dL_dpix, const float* dL_ddepth,
                         const float* dL_dmask, const float* dL_dfeat, float* dL_dtransMat, float* dL_dmean2D,
                         float* dL_dopacity, float* dL_dcolors, float* dL_dfeatures, float* dL_dnormals) {
  const orc_fov f = fov_consts(p);
  const int W = p->W, H = p->H, S = p->S, N = W * H;
  const int gx = (W + BLOCK_X - 1) / BLOCK_X;
  const float near_ = 2.0f * p->scale_factor, far_ = 300.0f * p->scale_factor;
#pragma omp parallel for schedule(dynamic, 64)
  for (int pix = 0; pix < N; ++pix) {
    const int x = pix % W, y = pix / W;
    const int tile = (y / BLOCK_Y) * gx + (x / BLOCK_X);
    const uint32_t r0 = ranges[2 * tile];
    const float pxf = (float)x, pyf = (float)y;
    const float phi = pxf * (f.HFOV_max - f.HFOV_min) / W + f.HFOV_min;
    const float theta = pyf * (f.VFOV_max - f.VFOV_min) / H + f.VFOV_min;
    const float sph = sinf(phi), cph = cosf(phi), sth = sinf(theta), cth = cosf(theta);
    const float T_final = final_T[pix];
    float T = T_final;
    const int last_contributor = n_contrib[pix], median_contributor = n_contrib[pix + N];
    const float final_D = final_T[pix + N], final_D2 = final_T[pix + 2 * N], final_A = 1 - T_final;
    float dpix[4], dfe[13];
    for (int i = 0; i < 4; ++i) dpix[i] = dL_dpix[i * N + pix];
    for (int i = 0; i < S + 3; ++i) dfe[i] = dL_dfeat[i * N + pix];
    const float dL_depth = dL_ddepth[pix], dL_dmedian = dL_ddepth[N + pix], dL_ddist = dL_ddepth[2 * N + pix],
                dL_depth_sq = dL_ddepth[3 * N + pix], dL_mask = dL_dmask[pix];
    float accum_rec[4] = {0}, last_color[4] = {0}, accum_frec[16] = {0}, last_feature[13] = {0};
    float accum_depth = 0, last_depth = 0, accum_mask = 0, last_alpha = 0, last_dL_dT = 0;
    for (int contributor = last_contributor - 1; contributor >= 0; --contributor) {
      const uint32_t id = point_list[r0 + contributor];
      const float* Tm = transMat + 9 * (size_t)id;
      pair_eval e;
      eval_pair(Tm, means2D + 2 * (size_t)id, normal_opacity[4 * (size_t)id + 3], depths[id], pxf, pyf, cph, sph, cth, sth,
                near_, far_, p->wrap ? (float)p->W : 0.f, &e);
      if (!e.valid) continue;
      const float alpha = e.alpha, G = e.G, depth = e.depth;
      T = T / (1.f - alpha);
      const float wgt = alpha * T;
      float dL_dalpha = 0.0f;
      for (int ch = 0; ch < 4; ++ch) {
        const float c = colors[4 * (size_t)id + ch];
        accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
        last_color[ch] = c;
        dL_dalpha += (c - accum_rec[ch]) * dpix[ch];
        atomic_addf(dL_dcolors + 4 * (size_t)id + ch, wgt * dpix[ch]);
      }
      float dL_dr = 0.0f;
      dL_dr += alpha * T * dL_depth;
      dL_dr += alpha * T * 2 * depth * dL_depth_sq;
      if (contributor == median_contributor - 1) dL_dr += dL_dmedian;
      const float m_d = far_ / (far_ - near_) * (1 - near_ / depth);
      const float dmd_dd = (far_ * near_) / ((far_ - near_) * depth * depth);
      const float dL_dweight = (final_D2 + m_d * m_d * final_A - 2 * m_d * final_D) * dL_ddist;
      dL_dalpha += dL_dweight - last_dL_dT;
      last_dL_dT = dL_dweight * alpha + (1 - alpha) * last_dL_dT;
      const float dL_dmd = 2.0f * (T * alpha) * (m_d * final_A - final_D) * dL_ddist;
      dL_dr += dL_dmd * dmd_dd;
      for (int ch = 0; ch < S + 3; ++ch) {
        float feat = ch < S ? features[(size_t)id * S + ch] : normal_opacity[4 * (size_t)id + ch - S];
        accum_frec[ch] = last_alpha * last_feature[ch] + (1.f - last_alpha) * accum_frec[ch];
        last_feature[ch] = feat;
        if (ch < S) atomic_addf(dL_dfeatures + (size_t)id * S + ch, wgt * dfe[ch]);
        else {
          dL_dalpha += (feat - accum_frec[ch]) * dfe[ch];
          atomic_addf(dL_dnormals + 3 * (size_t)id + ch - S, wgt * dfe[ch]);
        }
      }
      accum_depth = last_alpha * last_depth + (1.f - last_alpha) * accum_depth;
      last_depth = depth;
      dL_dalpha += (depth - accum_depth) * dL_depth;
      accum_mask = last_alpha + (1.f - last_alpha) * accum_mask;
      dL_dalpha = (float)((double)dL_dalpha + (1.0 - (double)accum_mask) * (double)dL_mask);
      dL_dalpha *= T;
      last_alpha = alpha;
      float bg_dot = 0;
      for (int i = 0; i < 4; ++i) bg_dot += bg[i] * dpix[i];
      dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
      const float dL_dG = normal_opacity[4 * (size_t)id + 3] * dL_dalpha;
      const float* Tu = Tm; const float* Tv = Tm + 3; const float* Tw = Tm + 6;
      float* dT = dL_dtransMat + 9 * (size_t)id;
      if (e.rho3d <= e.rho2d) {
        const float ex = sth * sph, ey = -cth, ez = sth * cph;
        const float dsx = dL_dG * -G * e.sx + dL_dr * (Tu[0] * ex + Tv[0] * ey + Tw[0] * ez);
        const float dsy = dL_dG * -G * e.sy + dL_dr * (Tu[1] * ex + Tv[1] * ey + Tw[1] * ez);
        const float qx = dsx / e.pz, qy = dsy / e.pz;
        const float dp[3] = {qx, qy, -(qx * e.sx + qy * e.sy)};
        const float* k = e.k; const float* l = e.l;
        const float dk[3] = {l[1] * dp[2] - l[2] * dp[1], l[2] * dp[0] - l[0] * dp[2], l[0] * dp[1] - l[1] * dp[0]};
        const float dl[3] = {dp[1] * k[2] - dp[2] * k[1], dp[2] * k[0] - dp[0] * k[2], dp[0] * k[1] - dp[1] * k[0]};
        const float s3[3] = {e.sx, e.sy, 1.f};
        for (int c = 0; c < 3; ++c) {
          atomic_addf(dT + c, cph * dk[c] + sph * cth * dl[c] + dL_dr * ex * s3[c]);
          atomic_addf(dT + 3 + c, sth * dl[c] + dL_dr * ey * s3[c]);
          atomic_addf(dT + 6 + c, -sph * dk[c] + cph * cth * dl[c] + dL_dr * ez * s3[c]);
        }
      } else {
        atomic_addf(dL_dmean2D + 4 * (size_t)id, dL_dG * (-G * 2.0f * e.dx));
        atomic_addf(dL_dmean2D + 4 * (size_t)id + 1, dL_dG * (-G * 2.0f * e.dy));
        atomic_addf(dT + 2, dL_dr * Tu[2] / depth);
        atomic_addf(dT + 5, dL_dr * Tv[2] / depth);
        atomic_addf(dT + 8, dL_dr * Tw[2] / depth);
      }
      atomic_addf(dL_dopacity + id, G * dL_dalpha);
    }
  }
}

/* auxiliary.h:128-139 */
static void dnormvdv(const float* v, const float* dv, float* o) {
  float sum2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  float inv = 1.0f / sqrtf(sum2 * sum2 * sum2);
  o[0] = ((+sum2 - v[0] * v[0]) * dv[0] - v[1] * v[0] * dv[1] - v[2] * v[0] * dv[2]) * inv;
  o[1] = (-v[0] * v[1] * dv[0] + (sum2 - v[1] * v[1]) * dv[1] - v[2] * v[1] * dv[2]) * inv;
  o[2] = (-v[0] * v[2] * dv[0] - v[1] * v[2] * dv[1] + (sum2 - v[2] * v[2]) * dv[2]) * inv;
}

/* backward.cu:17-134 SH VJP: writes dL_dsh[0..(D+1)^2) (4 channels each), adds to dL_dmean */
static void sh_backward(int deg, const float* sh, const uint8_t* clamped, const float* dL_dcolor, const float* pos,
                        const float* campos, float* dL_dsh, float* dL_dmean) {
  float d0[3] = {pos[0] - campos[0], pos[1] - campos[1], pos[2] - campos[2]};
  float len = sqrtf(d0[0] * d0[0] + d0[1] * d0[1] + d0[2] * d0[2]);
  float x = d0[0] / len, y = d0[1] / len, z = d0[2] / len;
  float g[4];
  for (int c = 0; c < 4; ++c) g[c] = dL_dcolor[c] * (clamped[c] ? 0.f : 1.f);
  float dx[4] = {0}, dy[4] = {0}, dz[4] = {0};
#define SHV(i, c) sh[(i)*4 + (c)]
#define OUT(i, coef) for (int c = 0; c < 4; ++c) dL_dsh[(i)*4 + c] = (coef) * g[c];
  OUT(0, SH_C0)
  if (deg > 0) {
    OUT(1, -SH_C1 * y) OUT(2, SH_C1 * z) OUT(3, -SH_C1 * x)
    for (int c = 0; c < 4; ++c) { dx[c] = -SH_C1 * SHV(3, c); dy[c] = -SH_C1 * SHV(1, c); dz[c] = SH_C1 * SHV(2, c); }
    if (deg > 1) {
      float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      OUT(4, SH_C2[0] * xy) OUT(5, SH_C2[1] * yz) OUT(6, SH_C2[2] * (2.f * zz - xx - yy)) OUT(7, SH_C2[3] * xz)
      OUT(8, SH_C2[4] * (xx - yy))
      for (int c = 0; c < 4; ++c) {
        dx[c] += SH_C2[0] * y * SHV(4, c) + SH_C2[2] * 2.f * -x * SHV(6, c) + SH_C2[3] * z * SHV(7, c) + SH_C2[4] * 2.f * x * SHV(8, c);
        dy[c] += SH_C2[0] * x * SHV(4, c) + SH_C2[1] * z * SHV(5, c) + SH_C2[2] * 2.f * -y * SHV(6, c) + SH_C2[4] * 2.f * -y * SHV(8, c);
        dz[c] += SH_C2[1] * y * SHV(5, c) + SH_C2[2] * 2.f * 2.f * z * SHV(6, c) + SH_C2[3] * x * SHV(7, c);
      }
      if (deg > 2) {
        OUT(9, SH_C3[0] * y * (3.f * xx - yy)) OUT(10, SH_C3[1] * xy * z) OUT(11, SH_C3[2] * y * (4.f * zz - xx - yy))
        OUT(12, SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy)) OUT(13, SH_C3[4] * x * (4.f * zz - xx - yy))
        OUT(14, SH_C3[5] * z * (xx - yy)) OUT(15, SH_C3[6] * x * (xx - 3.f * yy))
        for (int c = 0; c < 4; ++c) {
          dx[c] += (SH_C3[0] * SHV(9, c) * 3.f * 2.f * xy + SH_C3[1] * SHV(10, c) * yz + SH_C3[2] * SHV(11, c) * -2.f * xy +
                    SH_C3[3] * SHV(12, c) * -3.f * 2.f * xz + SH_C3[4] * SHV(13, c) * (-3.f * xx + 4.f * zz - yy) +
                    SH_C3[5] * SHV(14, c) * 2.f * xz + SH_C3[6] * SHV(15, c) * 3.f * (xx - yy));
          dy[c] += (SH_C3[0] * SHV(9, c) * 3.f * (xx - yy) + SH_C3[1] * SHV(10, c) * xz +
                    SH_C3[2] * SHV(11, c) * (-3.f * yy + 4.f * zz - xx) + SH_C3[3] * SHV(12, c) * -3.f * 2.f * yz +
                    SH_C3[4] * SHV(13, c) * -2.f * xy + SH_C3[5] * SHV(14, c) * -2.f * yz + SH_C3[6] * SHV(15, c) * -3.f * 2.f * xy);
          dz[c] += (SH_C3[1] * SHV(10, c) * xy + SH_C3[2] * SHV(11, c) * 4.f * 2.f * yz +
                    SH_C3[3] * SHV(12, c) * 3.f * (2.f * zz - xx - yy) + SH_C3[4] * SHV(13, c) * 4.f * 2.f * xz +
                    SH_C3[5] * SHV(14, c) * (xx - yy));
        }
      }
    }
  }
#undef SHV
#undef OUT
  float ddir[3] = {0, 0, 0};
  for (int c = 0; c < 4; ++c) { ddir[0] += dx[c] * g[c]; ddir[1] += dy[c] * g[c]; ddir[2] += dz[c] * g[c]; }
  float dm[3];
  dnormvdv(d0, ddir, dm);
  dL_dmean[0] += dm[0]; dL_dmean[1] += dm[1]; dL_dmean[2] += dm[2];
}

/*
 * K11 backward preprocess, backward.cu:517-712 (compute_transmat_aabb :517-620, preprocessCUDA :622-712,
 * quat_to_rotmat_vjp auxiliary.h:230-274).  dL_dmean2D (4P) carries the low-pass accumulations in and the
 * densification proxy out; all other outputs must be zero on entry (torch::zeros in the reference).
 */
void orc_preprocess_backward(const orc_params* p, const float* means3D, const float* scales, const float* rotations,
                             const float* shs, const uint8_t* clamped, const float* vm, const float* campos,
                             const int* radii, const float* transMat, const float* dL_dtransMat,
                             const float* dL_dnormals, const float* dL_dcolors, float* dL_dmean2D, float* dL_dmeans3D,
                             float* dL_dsh, float* dL_dscales, float* dL_drot) {
  const orc_fov f = fov_consts(p);
  /* backward.cu:658-659 with rasterizer_impl.cu:440-441 */
  const float focal_y = p->H / (2.0f * p->tanfovy), focal_x = p->W / (2.0f * p->tanfovx);
  const int W = (int)(focal_x * p->tanfovx * 2), H = (int)(focal_y * p->tanfovy * 2);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < p->P; ++i) {
    if (!(radii[i] > 0)) continue;
    float R[3][3];
    const float* q = rotations + 4 * i;
    quat_to_rot(q, R);
    const float sx = scales[3 * i], sy = scales[3 * i + 1];
    float dT[9];
    memcpy(dT, dL_dtransMat + 9 * (size_t)i, sizeof(dT));
    const float* Tm = transMat + 9 * (size_t)i;
    const float u = Tm[2], v = Tm[5], w = Tm[8];
    const float gx_ = dL_dmean2D[4 * i], gy_ = dL_dmean2D[4 * i + 1];
    if (gx_ != 0 || gy_ != 0) { /* backward.cu:579-595 */
      const float Wrange = W / (f.HFOV_max - f.HFOV_min), Hrange = H / (f.VFOV_max - f.VFOV_min);
      const float r2_uw = u * u + w * w, r_uw = sqrtf(u * u + w * w), r2 = u * u + v * v + w * w;
      dT[2] += gx_ * Wrange * w / r2_uw - gy_ * Hrange * u * v / (r_uw * r2);
      dT[5] += gy_ * Hrange * r_uw / r2;
      dT[8] += -gx_ * Wrange * u / r2_uw - gy_ * Hrange * v * w / (r_uw * r2);
    }
    /* dL_dM = P * dL_dT^T (backward.cu:598): column j of dL_dM from the j-th components of dTu,dTv,dTw */
    float dM[3][3];
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < 3; ++k) dM[j][k] = vm[4 * k + 0] * dT[j] + vm[4 * k + 1] * dT[3 + j] + vm[4 * k + 2] * dT[6 + j];
    const float* dn = dL_dnormals + 3 * (size_t)i;
    float dtn[3];
    for (int k = 0; k < 3; ++k) dtn[k] = vm[4 * k + 0] * dn[0] + vm[4 * k + 1] * dn[1] + vm[4 * k + 2] * dn[2];
    const float nz = vm[2] * R[2][0] + vm[6] * R[2][1] + vm[10] * R[2][2];
    const float mult = nz < 0 ? 1.f : -1.f; /* backward.cu:600-603 */
    for (int k = 0; k < 3; ++k) dtn[k] *= mult;
    dL_dscales[3 * i] = dM[0][0] * R[0][0] + dM[0][1] * R[0][1] + dM[0][2] * R[0][2];
    dL_dscales[3 * i + 1] = dM[1][0] * R[1][0] + dM[1][1] * R[1][1] + dM[1][2] * R[1][2];
    dL_dscales[3 * i + 2] = 0;
    float vR[3][3]; /* vR[c][r] */
    for (int k = 0; k < 3; ++k) { vR[0][k] = dM[0][k] * sx; vR[1][k] = dM[1][k] * sy; vR[2][k] = dtn[k]; }
    {
      float s = 1.0f / sqrtf(q[3] * q[3] + q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
      float qw = q[0] * s, qx = q[1] * s, qy = q[2] * s, qz = q[3] * s;
      dL_drot[4 * i + 0] = 2.f * (qx * (vR[1][2] - vR[2][1]) + qy * (vR[2][0] - vR[0][2]) + qz * (vR[0][1] - vR[1][0]));
      dL_drot[4 * i + 1] = 2.f * (-2.f * qx * (vR[1][1] + vR[2][2]) + qy * (vR[0][1] + vR[1][0]) + qz * (vR[0][2] + vR[2][0]) + qw * (vR[1][2] - vR[2][1]));
      dL_drot[4 * i + 2] = 2.f * (qx * (vR[0][1] + vR[1][0]) - 2.f * qy * (vR[0][0] + vR[2][2]) + qz * (vR[1][2] + vR[2][1]) + qw * (vR[2][0] - vR[0][2]));
      dL_drot[4 * i + 3] = 2.f * (qx * (vR[0][2] + vR[2][0]) + qy * (vR[1][2] + vR[2][1]) - 2.f * qz * (vR[0][0] + vR[1][1]) + qw * (vR[0][1] - vR[1][0]));
    }
    dL_dmeans3D[3 * i] = dM[2][0]; dL_dmeans3D[3 * i + 1] = dM[2][1]; dL_dmeans3D[3 * i + 2] = dM[2][2];
    if (shs != NULL)
      sh_backward(p->D, shs + (size_t)i * p->M * 4, clamped + 4 * (size_t)i, dL_dcolors + 4 * (size_t)i, means3D + 3 * i,
                  campos, dL_dsh + (size_t)i * p->M * 4, dL_dmeans3D + 3 * i);
    /* densification proxy, backward.cu:684-711 (raw accumulated dT, double promotions) */
    const float dL_du = dL_dtransMat[9 * (size_t)i + 2], dL_dv = dL_dtransMat[9 * (size_t)i + 5], dL_dw = dL_dtransMat[9 * (size_t)i + 8];
    const float phi = atan2f(u, w);
    dL_dmean2D[4 * i] = (float)((dL_du * w + dL_dw * -u) * 0.5 * (f.HFOV_max - f.HFOV_min));
    const float du_dth = -v * sinf(phi), dv_dth = sqrtf(u * u + w * w), dw_dth = -v * cosf(phi);
    dL_dmean2D[4 * i + 1] = (float)((dL_du * du_dth + dL_dv * dv_dth + dL_dw * dw_dth) * 0.5 * (f.VFOV_max - f.VFOV_min) * W / H);
  }
}

/* auxiliary.h:157-180 + rasterizer_impl.cu:51-64 */
void orc_mark_visible(int P, const float* pts, const float* vm, const float* pm, uint8_t* present) {
  for (int i = 0; i < P; ++i) {
    const float* q = pts + 3 * i;
    float hx = pm[0] * q[0] + pm[4] * q[1] + pm[8] * q[2] + pm[12];
    float hy = pm[1] * q[0] + pm[5] * q[1] + pm[9] * q[2] + pm[13];
    float hw = pm[3] * q[0] + pm[7] * q[1] + pm[11] * q[2] + pm[15];
    float pw = 1.0f / (hw + 0.0000001f);
    float projx = hx * pw, projy = hy * pw;
    float vz = vm[2] * q[0] + vm[6] * q[1] + vm[10] * q[2] + vm[14];
    int out = vz <= 0.2f || (projx < -1.3 || projx > 1.3 || projy < -1.3 || projy > 1.3);
    present[i] = out ? 0 : 1;
  }
}
