// gsl_postops.cu -- "next-3" of SURVEY.md section 8(f): the two panorama post-ops the training loop runs on every rendered
// range image (train.py:261-262,306), each ~15 element-wise PyTorch kernels in the reference
// (utils/graphics_utils.py:96-118 pano_to_lidar, :121-149 depth_to_normal), as one pass over the range image:
//   points  (K, 3)     = direction(pixel) * range(pixel) for the pixels with range > 0, row-major order
//   normals (3, H, W)  = normalize(cross(p[y+1,x] - p[y-1,x], p[y,x+1] - p[y,x-1])), zero on the one-pixel border
// with direction(row, col) = normalize(sin th sin ph, -cos th, sin th cos ph),
//   th = (90 - vfov_max + row / H * (vfov_max - vfov_min)) pi / 180,  ph = (hfov_min + col / W * (hfov_max - hfov_min)) pi / 180
// evaluated in float32 in the reference's order of operations, and their backward passes (both ops are on the loss path:
// the Chamfer loss back-propagates through pano_to_lidar, the normal-consistency loss through depth_to_normal).
// One thread per pixel; the image is 68k - 262k pixels, so the forward is launch-latency-bound: the point of the fusion
// is 2 launches instead of ~30.
#include "gsl_common.cuh"

namespace gsl {

struct PanoParams {
  int H, W;
  float a_v, d_v;  // 90 - vfov_max, vfov_max - vfov_min   (degrees)
  float a_h, d_h;  // hfov_min, hfov_max - hfov_min
};

__device__ __forceinline__ float3 pano_direction(const PanoParams& pp, int row, int col) {
  // reference: (a + idx / N * d) * pi / 180 in float32, then sin / cos, then F.normalize (eps 1e-12)
  const float pi = 3.14159265358979323846f;
  const float th = __fdiv_rn(__fmul_rn(__fadd_rn(pp.a_v, __fmul_rn(__fdiv_rn((float)row, (float)pp.H), pp.d_v)), pi), 180.f);
  const float ph = __fdiv_rn(__fmul_rn(__fadd_rn(pp.a_h, __fmul_rn(__fdiv_rn((float)col, (float)pp.W), pp.d_h)), pi), 180.f);
  const float st = sinf(th), ct = cosf(th), sp = sinf(ph), cp = cosf(ph);
  float3 d = make_float3(__fmul_rn(st, sp), -ct, __fmul_rn(st, cp));
  const float n = fmaxf(sqrtf(d.x * d.x + d.y * d.y + d.z * d.z), 1e-12f);
  return make_float3(d.x / n, d.y / n, d.z / n);
}

constexpr int PANO_BLOCK = 256;

// pass 1: valid pixels per block
__global__ void __launch_bounds__(PANO_BLOCK) k_pano_count(int N, const float* __restrict__ range, int32_t* __restrict__ blk) {
  const int i = blockIdx.x * PANO_BLOCK + threadIdx.x;
  const bool v = i < N && range[i] > 0.f;
  const int c = __syncthreads_count(v);
  if (threadIdx.x == 0) blk[blockIdx.x] = c;
}

// pass 2: compacted points (+ the pixel of every point, for the backward pass) and the normal map
__global__ void __launch_bounds__(PANO_BLOCK) k_pano_forward(PanoParams pp, const float* __restrict__ range,
                                                             const int32_t* __restrict__ blk, float* __restrict__ points,
                                                             int32_t* __restrict__ index, int32_t* __restrict__ count,
                                                             float* __restrict__ normals) {
  __shared__ int s_base, s_warp[PANO_BLOCK / 32];
  const int N = pp.H * pp.W;
  const int i = blockIdx.x * PANO_BLOCK + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = i / pp.W, col = i - row * pp.W;
  const float r = i < N ? range[i] : 0.f;
  float3 d = make_float3(0.f, 0.f, 0.f);
  if (i < N) d = pano_direction(pp, row, col);
  if (points) {
    // exclusive prefix of the block counts (a few hundred blocks: one warp sums them)
    if (warp == 0) {
      int s = 0;
      for (int b = lane; b < (int)blockIdx.x; b += 32) s += blk[b];
      s = __reduce_add_sync(0xffffffffu, s);
      if (lane == 0) s_base = s;
      if (blockIdx.x == gridDim.x - 1 && lane == 0 && count) *count = s + blk[blockIdx.x];
    }
    const bool v = r > 0.f;
    const uint32_t bits = __ballot_sync(0xffffffffu, v);
    if (lane == 0) s_warp[warp] = __popc(bits);
    __syncthreads();
    int before = s_base;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    if (v) {
      const int k = before + __popc(bits & ((1u << lane) - 1u));
      points[3 * (size_t)k] = d.x * r;
      points[3 * (size_t)k + 1] = d.y * r;
      points[3 * (size_t)k + 2] = d.z * r;
      if (index) index[k] = i;
    }
  }
  if (normals && i < N) {
    float3 n = make_float3(0.f, 0.f, 0.f);
    if (row > 0 && row < pp.H - 1 && col > 0 && col < pp.W - 1) {
      const float3 du = pano_direction(pp, row - 1, col), dd = pano_direction(pp, row + 1, col);
      const float3 dl = pano_direction(pp, row, col - 1), dr = pano_direction(pp, row, col + 1);
      const float ru = range[i - pp.W], rd = range[i + pp.W], rl = range[i - 1], rr = range[i + 1];
      const float3 a = make_float3(dd.x * rd - du.x * ru, dd.y * rd - du.y * ru, dd.z * rd - du.z * ru);  // "dx": down - up
      const float3 b = make_float3(dr.x * rr - dl.x * rl, dr.y * rr - dl.y * rl, dr.z * rr - dl.z * rl);  // "dy": right - left
      const float3 c = make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
      const float len = fmaxf(sqrtf(c.x * c.x + c.y * c.y + c.z * c.z), 1e-12f);
      n = make_float3(c.x / len, c.y / len, c.z / len);
    }
    normals[i] = n.x;
    normals[N + i] = n.y;
    normals[2 * (size_t)N + i] = n.z;
  }
}

// backward of the points: d range[pixel] = <direction(pixel), d points[k]>; pixels without a point get 0 (g_range is
// zero-filled by the launcher)
__global__ void __launch_bounds__(PANO_BLOCK) k_pano_points_bwd(PanoParams pp, int K, const float* __restrict__ g_points,
                                                                const int32_t* __restrict__ index, float* __restrict__ g_range) {
  const int k = blockIdx.x * PANO_BLOCK + threadIdx.x;
  if (k >= K) return;
  const int i = index[k];
  const int row = i / pp.W, col = i - row * pp.W;
  const float3 d = pano_direction(pp, row, col);
  g_range[i] += d.x * g_points[3 * (size_t)k] + d.y * g_points[3 * (size_t)k + 1] + d.z * g_points[3 * (size_t)k + 2];
}

// backward of the normal map, gather form: pixel q collects from its four neighbours p (as their up / down / left / right
// point): n = c / |c|, c = a x b  =>  dc = (g - n <n, g>) / |c|,  da = b x dc,  db = dc x a,  and point q enters a or b of p
// with sign +-1 scaled by direction(q).
__global__ void __launch_bounds__(PANO_BLOCK) k_pano_normals_bwd(PanoParams pp, const float* __restrict__ range,
                                                                 const float* __restrict__ g_normals, float* __restrict__ g_range,
                                                                 int accumulate) {
  const int N = pp.H * pp.W;
  const int q = blockIdx.x * PANO_BLOCK + threadIdx.x;
  if (q >= N) return;
  const int qr = q / pp.W, qc = q - qr * pp.W;
  float3 gp = make_float3(0.f, 0.f, 0.f);  // gradient w.r.t. the back-projected point of pixel q
  // neighbours p for which q is: down (p = q - W, enters a with +), up (p = q + W, a with -), right (p = q - 1, b +), left (p = q + 1, b -)
  const int drow[4] = {-1, 1, 0, 0}, dcol[4] = {0, 0, -1, 1};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int pr = qr + drow[t], pc = qc + dcol[t];
    if (pr <= 0 || pr >= pp.H - 1 || pc <= 0 || pc >= pp.W - 1) continue;  // p must be an interior pixel
    const int p = pr * pp.W + pc;
    const float3 du = pano_direction(pp, pr - 1, pc), dd = pano_direction(pp, pr + 1, pc);
    const float3 dl = pano_direction(pp, pr, pc - 1), dr = pano_direction(pp, pr, pc + 1);
    const float ru = range[p - pp.W], rd = range[p + pp.W], rl = range[p - 1], rr = range[p + 1];
    const float3 a = make_float3(dd.x * rd - du.x * ru, dd.y * rd - du.y * ru, dd.z * rd - du.z * ru);
    const float3 b = make_float3(dr.x * rr - dl.x * rl, dr.y * rr - dl.y * rl, dr.z * rr - dl.z * rl);
    const float3 c = make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
    const float len = sqrtf(c.x * c.x + c.y * c.y + c.z * c.z);
    const float3 g = make_float3(g_normals[p], g_normals[N + p], g_normals[2 * (size_t)N + p]);
    float3 dc;
    if (len > 1e-12f) {
      const float inv = 1.f / len;
      const float3 n = make_float3(c.x * inv, c.y * inv, c.z * inv);
      const float ng = n.x * g.x + n.y * g.y + n.z * g.z;
      dc = make_float3((g.x - n.x * ng) * inv, (g.y - n.y * ng) * inv, (g.z - n.z * ng) * inv);
    } else {  // F.normalize clamps the norm at eps: n = c / eps
      dc = make_float3(g.x * 1e12f, g.y * 1e12f, g.z * 1e12f);
    }
    float3 dv;  // gradient w.r.t. the vector (a or b) that q's point enters
    if (t < 2) dv = make_float3(b.y * dc.z - b.z * dc.y, b.z * dc.x - b.x * dc.z, b.x * dc.y - b.y * dc.x);  // da = b x dc
    else dv = make_float3(dc.y * a.z - dc.z * a.y, dc.z * a.x - dc.x * a.z, dc.x * a.y - dc.y * a.x);        // db = dc x a
    const float sgn = (t == 0 || t == 2) ? 1.f : -1.f;
    gp.x += sgn * dv.x; gp.y += sgn * dv.y; gp.z += sgn * dv.z;
  }
  const float3 dq = pano_direction(pp, qr, qc);
  const float v = dq.x * gp.x + dq.y * gp.y + dq.z * gp.z;
  g_range[q] = accumulate ? g_range[q] + v : v;
}

static PanoParams make_pano(const gsl_pano_params& p) {
  PanoParams pp;
  pp.H = p.H; pp.W = p.W;
  // the scalar parts are Python doubles in the reference (90 - vfov[1], vfov[1] - vfov[0]) and enter the tensor expression
  // rounded to float32
  pp.a_v = (float)(90.0 - (double)p.vfov_max);
  pp.d_v = (float)((double)p.vfov_max - (double)p.vfov_min);
  pp.a_h = (float)(double)p.hfov_min;
  pp.d_h = (float)((double)p.hfov_max - (double)p.hfov_min);
  return pp;
}

int launch_pano_forward(const gsl_pano_params& p, const float* range, float* points, int32_t* index, int32_t* count,
                        float* normals, void* scratch, cudaStream_t st) {
  const int N = p.H * p.W;
  if (N <= 0) return 0;
  const int blocks = (N + PANO_BLOCK - 1) / PANO_BLOCK;
  const PanoParams pp = make_pano(p);
  if (points) k_pano_count<<<blocks, PANO_BLOCK, 0, st>>>(N, range, (int32_t*)scratch);
  k_pano_forward<<<blocks, PANO_BLOCK, 0, st>>>(pp, range, (const int32_t*)scratch, points, index, count, normals);
  return check_cuda(cudaGetLastError(), "k_pano_forward launch");
}

int launch_pano_backward(const gsl_pano_params& p, const float* range, int K, const float* g_points, const int32_t* index,
                         const float* g_normals, float* g_range, cudaStream_t st) {
  const int N = p.H * p.W;
  if (N <= 0) return 0;
  const PanoParams pp = make_pano(p);
  if (g_normals) {
    k_pano_normals_bwd<<<(N + PANO_BLOCK - 1) / PANO_BLOCK, PANO_BLOCK, 0, st>>>(pp, range, g_normals, g_range, 0);
  } else {
    cudaMemsetAsync(g_range, 0, sizeof(float) * (size_t)N, st);
  }
  if (g_points && K > 0)
    k_pano_points_bwd<<<(K + PANO_BLOCK - 1) / PANO_BLOCK, PANO_BLOCK, 0, st>>>(pp, K, g_points, index, g_range);
  return check_cuda(cudaGetLastError(), "k_pano_backward launch");
}

}  // namespace gsl
