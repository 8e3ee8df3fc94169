// gsl_glue.cu -- "next-1" of SURVEY.md section 8(f): the per-surfel glue the reference's render() runs as ~12 separate
// PyTorch element-wise kernels (and as many autograd nodes) in front of the rasterizer, fused into ONE streaming kernel
// per direction:
//   means3D  = xyz + v sin((t - t0) a) / a  [+ v exp(-sigma_t / T / 2 * decay) * time_shift]
//                                           GaussianModel.get_xyz_SHM / get_inst_velocity  scene/gaussian_model.py:151-157,
//                                           render()  gaussian_renderer/__init__.py:69-75
//   marginal = exp(-0.5 (t0 - t)^2 / sigma_t^2),  sigma_t = exp(_scaling_t)        gaussian_model.py:144-145,185-186
//   opacity  = sigmoid(_opacity) [* marginal when pipe.dynamic]                      gaussian_model.py:174-175, __init__.py:77-79
//   scales   = exp(_scaling);  rotations = _rotation / max(|_rotation|, 1e-12)       gaussian_model.py:140-149
//   mask     = [mask &] opacity > 1/255 [& marginal > 0.05 when dynamic]             __init__.py:112-115
// One thread per surfel; 64 B in, 45 B out per surfel in the forward -- a pure HBM stream.
#include "gsl_common.cuh"

namespace gsl {

struct GlueParams {
  int P;
  float ts;        // timestamp - time_shift (or the timestamp itself)
  float shift;     // time_shift, 0 when absent
  float a;         // 2 pi / T
  float inv2T_decay;  // velocity_decay / (2 T)
  int dynamic;
};

__global__ void __launch_bounds__(256) k_glue_fwd(GlueParams gp, const float* __restrict__ xyz, const float* __restrict__ vel,
                                                  const float* __restrict__ t0, const float* __restrict__ scaling_t,
                                                  const float* __restrict__ opacity_raw, const float* __restrict__ scaling_raw,
                                                  const float4* __restrict__ rotation_raw, const uint8_t* __restrict__ mask_in,
                                                  float* __restrict__ means3D, float* __restrict__ opacity,
                                                  float* __restrict__ scales, float4* __restrict__ rotations,
                                                  float* __restrict__ marginal_t, uint8_t* __restrict__ mask_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= gp.P) return;
  const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
  const float vx = vel[3 * i], vy = vel[3 * i + 1], vz = vel[3 * i + 2];
  const float tt = t0[i];
  const float sig = expf(scaling_t[i]);
  const float ph = (gp.ts - tt) * gp.a;
  float coef = sinf(ph) / gp.a;
  if (gp.shift != 0.f) coef += expf(-sig * gp.inv2T_decay) * gp.shift;
  means3D[3 * i] = x + vx * coef;
  means3D[3 * i + 1] = y + vy * coef;
  means3D[3 * i + 2] = z + vz * coef;
  const float d = tt - gp.ts;
  const float mt = expf(-0.5f * d * d / (sig * sig));
  const float os = 1.f / (1.f + expf(-opacity_raw[i]));
  const float op = gp.dynamic ? os * mt : os;
  opacity[i] = op;
  if (marginal_t) marginal_t[i] = mt;
  scales[3 * i] = expf(scaling_raw[3 * i]);
  scales[3 * i + 1] = expf(scaling_raw[3 * i + 1]);
  scales[3 * i + 2] = expf(scaling_raw[3 * i + 2]);
  const float4 q = rotation_raw[i];
  const float inv = 1.f / fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);
  rotations[i] = make_float4(q.x * inv, q.y * inv, q.z * inv, q.w * inv);
  bool m = op > (1.0f / 255.0f);
  if (mask_in) m = m && mask_in[i] != 0;
  if (gp.dynamic) m = m && mt > 0.05f;
  mask_out[i] = m ? 1 : 0;
}

// VJP of k_glue_fwd.  Gradient pointers of the outputs may be NULL (treated as zero).
__global__ void __launch_bounds__(256) k_glue_bwd(GlueParams gp, const float* __restrict__ vel, const float* __restrict__ t0,
                                                  const float* __restrict__ scaling_t, const float* __restrict__ opacity_raw,
                                                  const float* __restrict__ scaling_raw, const float4* __restrict__ rotation_raw,
                                                  const float* __restrict__ g_means3D, const float* __restrict__ g_opacity,
                                                  const float* __restrict__ g_scales, const float4* __restrict__ g_rotations,
                                                  float* __restrict__ g_xyz, float* __restrict__ g_vel, float* __restrict__ g_t0,
                                                  float* __restrict__ g_scaling_t, float* __restrict__ g_opacity_raw,
                                                  float* __restrict__ g_scaling_raw, float4* __restrict__ g_rotation_raw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= gp.P) return;
  float gx = 0.f, gy = 0.f, gz = 0.f;
  if (g_means3D) { gx = g_means3D[3 * i]; gy = g_means3D[3 * i + 1]; gz = g_means3D[3 * i + 2]; }
  const float vx = vel[3 * i], vy = vel[3 * i + 1], vz = vel[3 * i + 2];
  const float tt = t0[i];
  const float sig = expf(scaling_t[i]);
  const float ph = (gp.ts - tt) * gp.a;
  float sn, cs;
  sincosf(ph, &sn, &cs);
  float coef = sn / gp.a;
  float ev = 0.f;
  if (gp.shift != 0.f) { ev = expf(-sig * gp.inv2T_decay); coef += ev * gp.shift; }
  g_xyz[3 * i] = gx; g_xyz[3 * i + 1] = gy; g_xyz[3 * i + 2] = gz;
  g_vel[3 * i] = gx * coef; g_vel[3 * i + 1] = gy * coef; g_vel[3 * i + 2] = gz * coef;
  const float gv = gx * vx + gy * vy + gz * vz;
  float g_tt = -cs * gv;
  float g_sig = (gp.shift != 0.f) ? gv * gp.shift * ev * (-gp.inv2T_decay) : 0.f;
  const float d = tt - gp.ts;
  const float mt = expf(-0.5f * d * d / (sig * sig));
  const float os = 1.f / (1.f + expf(-opacity_raw[i]));
  const float gop = g_opacity ? g_opacity[i] : 0.f;
  float g_os = gop;
  if (gp.dynamic) {
    const float g_mt = gop * os;
    g_os = gop * mt;
    g_tt += g_mt * mt * (-d / (sig * sig));
    g_sig += g_mt * mt * d * d / (sig * sig * sig);
  }
  g_t0[i] = g_tt;
  g_scaling_t[i] = g_sig * sig;
  g_opacity_raw[i] = g_os * os * (1.f - os);
#pragma unroll
  for (int k = 0; k < 3; ++k) g_scaling_raw[3 * i + k] = (g_scales ? g_scales[3 * i + k] : 0.f) * expf(scaling_raw[3 * i + k]);
  const float4 q = rotation_raw[i];
  const float4 gq = g_rotations ? g_rotations[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float nrm = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  float4 out;
  if (nrm > 1e-12f) {
    const float inv = 1.f / nrm;
    const float nx = q.x * inv, ny = q.y * inv, nz = q.z * inv, nw = q.w * inv;
    const float dt = nx * gq.x + ny * gq.y + nz * gq.z + nw * gq.w;
    out = make_float4((gq.x - nx * dt) * inv, (gq.y - ny * dt) * inv, (gq.z - nz * dt) * inv, (gq.w - nw * dt) * inv);
  } else {  // clamped denominator: the output is q * 1e12
    out = make_float4(gq.x * 1e12f, gq.y * 1e12f, gq.z * 1e12f, gq.w * 1e12f);
  }
  g_rotation_raw[i] = out;
}

static GlueParams make_glue_params(const gsl_glue_params& p) {
  GlueParams g;
  g.P = p.P;
  g.ts = p.timestamp - p.time_shift;
  g.shift = p.time_shift;
  g.a = (float)(1.0 / (double)p.cycle * 3.141592653589793 * 2.0);  // a = 1 / T * np.pi * 2 (gaussian_model.py:152)
  g.inv2T_decay = p.velocity_decay / p.cycle / 2.f;
  g.dynamic = p.dynamic;
  return g;
}

int launch_glue_forward(const gsl_glue_params& p, const gsl_glue_inputs& in, const gsl_glue_outputs& out, cudaStream_t st) {
  if (p.P == 0) return 0;
  ProfScope prof(GSL_K_GLUE_FWD, st);
  k_glue_fwd<<<(p.P + 255) / 256, 256, 0, st>>>(make_glue_params(p), in.xyz, in.velocity, in.t, in.scaling_t, in.opacity,
                                               in.scaling, reinterpret_cast<const float4*>(in.rotation), in.mask,
                                               out.means3D, out.opacity, out.scales, reinterpret_cast<float4*>(out.rotations),
                                               out.marginal_t, out.mask);
  return check_cuda(cudaGetLastError(), "k_glue_fwd launch");
}

int launch_glue_backward(const gsl_glue_params& p, const gsl_glue_inputs& in, const gsl_glue_outputs& gout,
                         const gsl_glue_inputs_grad& gin, cudaStream_t st) {
  if (p.P == 0) return 0;
  ProfScope prof(GSL_K_GLUE_BWD, st);
  k_glue_bwd<<<(p.P + 255) / 256, 256, 0, st>>>(make_glue_params(p), in.velocity, in.t, in.scaling_t, in.opacity, in.scaling,
                                               reinterpret_cast<const float4*>(in.rotation), gout.means3D, gout.opacity,
                                               gout.scales, reinterpret_cast<const float4*>(gout.rotations), gin.xyz,
                                               gin.velocity, gin.t, gin.scaling_t, gin.opacity, gin.scaling,
                                               reinterpret_cast<float4*>(gin.rotation));
  return check_cuda(cudaGetLastError(), "k_glue_bwd launch");
}

}  // namespace gsl
