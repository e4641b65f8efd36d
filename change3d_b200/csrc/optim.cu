// Fused Adam over a flat fp32 parameter buffer — torch.optim.Adam(lr, betas=(0.9,0.99), eps=1e-8,
// weight_decay=1e-4) as configured at scripts/train_BCD.py:284-290 (L2 added to the gradient, not
// decoupled), one launch for every parameter of the model.  grad_clip > 0 clamps every (scaled) gradient element to
// [-grad_clip, grad_clip] first: `clip_gradient` of the captioning loop (model/utils.py:481-491, called at
// scripts/train_CC.py:141-144 before both optimizers step).
#include "c3d_common.cuh"
#include "../../include/change3d_b200.h"

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, long long n,
                                                   float step_size, float inv_sqrt_bc2, float beta1, float beta2,
                                                   float eps, float wd, float grad_scale, float clip) {
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pv = *reinterpret_cast<float4*>(p + 4 * i), mv = *reinterpret_cast<float4*>(m + 4 * i),
           vv = *reinterpret_cast<float4*>(v + 4 * i);
    const float4 gv = ldg4(g + 4 * i);
    float pa[4] = {pv.x, pv.y, pv.z, pv.w}, ma[4] = {mv.x, mv.y, mv.z, mv.w}, va[4] = {vv.x, vv.y, vv.z, vv.w},
          ga[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gg = fmaf(wd, pa[j], fminf(fmaxf(ga[j] * grad_scale, -clip), clip));
      ma[j] = fmaf(beta1, ma[j], (1.f - beta1) * gg);
      va[j] = fmaf(beta2, va[j], (1.f - beta2) * gg * gg);
      const float denom = sqrtf(va[j]) * inv_sqrt_bc2 + eps;
      pa[j] -= step_size * (ma[j] / denom);
    }
    st4(p + 4 * i, make_float4(pa[0], pa[1], pa[2], pa[3]));
    st4(m + 4 * i, make_float4(ma[0], ma[1], ma[2], ma[3]));
    st4(v + 4 * i, make_float4(va[0], va[1], va[2], va[3]));
  }
  // tail
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    const float gg = fmaf(wd, p[i], fminf(fmaxf(g[i] * grad_scale, -clip), clip));
    m[i] = fmaf(beta1, m[i], (1.f - beta1) * gg);
    v[i] = fmaf(beta2, v[i], (1.f - beta2) * gg * gg);
    p[i] -= step_size * (m[i] / (sqrtf(v[i]) * inv_sqrt_bc2 + eps));
  }
}

extern "C" int c3d_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                             float beta2, float eps, float weight_decay, int step, float grad_scale, float grad_clip,
                             void* stream_) {
  if (!p || !g || !m || !v || n <= 0 || step <= 0) return C3D_ERR_ARG;
  if ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) != 0) return C3D_ERR_ARG;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1), inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  long long blocks = ((n >> 2) + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  adam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(p, g, m, v, n, step_size, inv_sqrt_bc2, beta1, beta2,
                                                                   eps, weight_decay, grad_scale, grad_clip > 0.f ? grad_clip : HUGE_VALF);
  return c3d_check_last(cudaGetLastError());
}
