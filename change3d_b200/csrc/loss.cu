// On-device losses and confusion matrices of the train step, sm_100a.
//
//   BCEDiceLoss          model/utils.py:154-169   (+ the `output > 0.5` mask and 2x2 confusion matrix of
//                                                  scripts/train_BCD.py:203-225 in the same pass)
//   CrossEntropyLoss2d   model/utils.py:171-178   (log_softmax + NLLLoss(ignore_index), + torch.argmax map,
//                                                  scripts/train_SCD.py:243-244)
//   ChangeSimilarity     model/utils.py:180-203   (softmax, CosineEmbeddingLoss(margin 0, mean))
//   get_confuse_matrix   utils/metric_tool.py:111-128
//
// Every forward is ONE pass over its inputs: per-thread fp32 partials -> per-block fp64 partials -> fp64
// atomics into a small workspace; the last block to finish (ticket counter) turns the sums into the loss and
// the backward coefficients and re-zeroes the workspace, so a step needs no memset and is graph-capturable.
// All kernels are byte-bound streaming reductions (grid = 148 SMs x 8 CTAs of 256 threads, 16-byte loads).
#include "c3d_common.cuh"
#include "../../include/change3d_b200.h"

namespace {

struct LossWs {                 // C3D_LOSS_WS_BYTES; zeroed once by the caller, self-cleaning afterwards
  double acc[4];
  unsigned long long cnt[4];
  unsigned int ticket;
  unsigned int pad;
};
static_assert(sizeof(LossWs) <= C3D_LOSS_WS_BYTES, "workspace size");

constexpr int kThreads = 256;
constexpr int kMaxBlocks = 148 * 8;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ unsigned warp_sum(unsigned v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-reduces NA float partials and NC counters into the workspace; returns true in every thread of the
// block that finished last (all other blocks' atomics are then visible to it).
template <int NA, int NC>
__device__ __forceinline__ bool block_commit(LossWs* ws, const float (&a)[NA], const unsigned (&c)[NC]) {
  __shared__ double s_acc[kThreads / 32][NA];
  __shared__ unsigned s_cnt[kThreads / 32][NC];
  __shared__ unsigned s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    const double v = warp_sum((double)a[i]);
    if (lane == 0) s_acc[warp][i] = v;
  }
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const unsigned v = warp_sum(c[i]);
    if (lane == 0) s_cnt[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      double v = 0.0;
      for (int w = 0; w < kThreads / 32; ++w) v += s_acc[w][i];
      atomicAdd(&ws->acc[i], v);
    }
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      unsigned long long v = 0;
      for (int w = 0; w < kThreads / 32; ++w) v += s_cnt[w][i];
      if (v) atomicAdd(&ws->cnt[i], v);
    }
    __threadfence();
    s_last = (atomicAdd(&ws->ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  return s_last != 0;
}

__device__ __forceinline__ double ws_take(double* p) {            // read-and-clear through L2
  return __longlong_as_double((long long)atomicExch((unsigned long long*)p, 0ull));
}

// ------------------------------------------------------------------------------------------------ BCE + Dice
// F.binary_cross_entropy clamps both logarithms at -100 (torch/aten Loss.cu).
__device__ __forceinline__ void bce_dice_elem(float p, float t, float (&a)[4], unsigned (&c)[4]) {
  const float lp = fmaxf(logf(p), -100.f), lq = fmaxf(log1pf(-p), -100.f);
  a[0] -= fmaf(t, lp, (1.f - t) * lq);
  a[1] = fmaf(p, t, a[1]);
  a[2] += p;
  a[3] += t;
  if (t >= 0.f && t < 2.f) c[2 * (int)t + (p > 0.5f ? 1 : 0)] += 1u;   // mask (gt >= 0) & (gt < n_class), astype(int)
}

__global__ void __launch_bounds__(kThreads) bce_dice_fwd_kernel(const float* __restrict__ pred,
                                                                const float* __restrict__ target, long long n,
                                                                int vec, LossWs* ws, float* __restrict__ out,
                                                                long long* __restrict__ cm) {
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  unsigned c[4] = {0u, 0u, 0u, 0u};
  const long long tid = blockIdx.x * (long long)kThreads + threadIdx.x, nth = (long long)gridDim.x * kThreads;
  if (vec) {
    const long long n4 = n >> 2;
    for (long long i = tid; i < n4; i += nth) {
      const float4 p = ldg4(pred + 4 * i), t = ldg4(target + 4 * i);
      bce_dice_elem(p.x, t.x, a, c);
      bce_dice_elem(p.y, t.y, a, c);
      bce_dice_elem(p.z, t.z, a, c);
      bce_dice_elem(p.w, t.w, a, c);
    }
    for (long long i = (n4 << 2) + tid; i < n; i += nth) bce_dice_elem(pred[i], target[i], a, c);
  } else {
    for (long long i = tid; i < n; i += nth) bce_dice_elem(pred[i], target[i], a, c);
  }
  if (block_commit<4, 4>(ws, a, c) && threadIdx.x == 0) {
    const double bce = ws_take(&ws->acc[0]) / (double)n, inter = ws_take(&ws->acc[1]);
    const double S = ws_take(&ws->acc[2]) + ws_take(&ws->acc[3]) + 1e-5;
    const double dice = (2.0 * inter + 1e-5) / S;
    out[0] = (float)(bce + 1.0 - dice);                            // the loss
    out[1] = (float)(1.0 / (double)n);                             // d bce / d(sum of terms)
    out[2] = (float)(2.0 / S);                                     // d dice / d inter
    out[3] = (float)(dice / S);                                    // -d dice / d(sum p)
    out[4] = (float)bce;
    out[5] = (float)dice;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const unsigned long long v = atomicExch(&ws->cnt[i], 0ull);
      if (cm) cm[i] += (long long)v;
    }
    ws->ticket = 0u;
  }
}

// d loss / d p = g * [ (p - t) / max(p (1 - p), 1e-12) / n  -  t * 2/S  +  dice/S ]
__device__ __forceinline__ float bce_dice_grad(float p, float t, float g, float inv_n, float c_i, float c_s) {
  const float d = fmaxf((1.f - p) * p, 1e-12f);
  return g * (fmaf((p - t) / d, inv_n, c_s) - t * c_i);
}

__global__ void __launch_bounds__(kThreads) bce_dice_bwd_kernel(const float* __restrict__ pred,
                                                                const float* __restrict__ target,
                                                                const float* __restrict__ coef,
                                                                const float* __restrict__ gout, float gscale,
                                                                float* __restrict__ dpred, long long n, int vec) {
  const float g = (gout ? __ldg(gout) : 1.f) * gscale, inv_n = __ldg(coef + 1), c_i = __ldg(coef + 2),
              c_s = __ldg(coef + 3);
  const long long tid = blockIdx.x * (long long)kThreads + threadIdx.x, nth = (long long)gridDim.x * kThreads;
  if (vec) {
    const long long n4 = n >> 2;
    for (long long i = tid; i < n4; i += nth) {
      const float4 p = ldg4(pred + 4 * i), t = ldg4(target + 4 * i);
      st4(dpred + 4 * i, make_float4(bce_dice_grad(p.x, t.x, g, inv_n, c_i, c_s), bce_dice_grad(p.y, t.y, g, inv_n, c_i, c_s),
                                     bce_dice_grad(p.z, t.z, g, inv_n, c_i, c_s), bce_dice_grad(p.w, t.w, g, inv_n, c_i, c_s)));
    }
    for (long long i = (n4 << 2) + tid; i < n; i += nth) dpred[i] = bce_dice_grad(pred[i], target[i], g, inv_n, c_i, c_s);
  } else {
    for (long long i = tid; i < n; i += nth) dpred[i] = bce_dice_grad(pred[i], target[i], g, inv_n, c_i, c_s);
  }
}

// ------------------------------------------------------------------------------------------- cross entropy 2d
constexpr int kMaxClasses = 16;

// One thread per pixel; the C logits of a pixel are HW apart (NCHW), so a warp reads C coalesced rows.
template <int MAXC>
__global__ void __launch_bounds__(kThreads) ce2d_fwd_kernel(const float* __restrict__ logits,
                                                            const long long* __restrict__ target, long long npix,
                                                            int C, long long HW, long long batch_stride,
                                                            long long ignore_index, LossWs* ws,
                                                            float* __restrict__ out, long long* __restrict__ argmax_out,
                                                            long long* __restrict__ cm) {
  __shared__ unsigned s_cm[kMaxClasses * kMaxClasses];
  if (cm) {
    for (int i = threadIdx.x; i < C * C; i += kThreads) s_cm[i] = 0u;
    __syncthreads();
  }
  float a[1] = {0.f};
  unsigned c[1] = {0u};
  const long long nth = (long long)gridDim.x * kThreads;
  for (long long pix = blockIdx.x * (long long)kThreads + threadIdx.x; pix < npix; pix += nth) {
    const long long b = pix / HW, hw = pix - b * HW;
    const float* x = logits + b * batch_stride + hw;
    float v[MAXC];
    float mx = -INFINITY;
    int am = 0;
#pragma unroll
    for (int k = 0; k < MAXC; ++k) {
      v[k] = (k < C) ? __ldg(x + k * HW) : -INFINITY;
      if (v[k] > mx) { mx = v[k]; am = k; }                         // first maximum, like torch.argmax
    }
    float se = 0.f;
#pragma unroll
    for (int k = 0; k < MAXC; ++k) se += (k < C) ? expf(v[k] - mx) : 0.f;
    const long long t = target[pix];
    if (t != ignore_index && t >= 0 && t < C) {
      float vt = 0.f;
#pragma unroll
      for (int k = 0; k < MAXC; ++k) vt = (k == (int)t) ? v[k] : vt;
      a[0] += (mx - vt) + logf(se);                                 // -log_softmax(x)[t]
      c[0] += 1u;
    }
    if (argmax_out) argmax_out[pix] = am;
    if (cm && t >= 0 && t < C) atomicAdd(&s_cm[(int)t * C + am], 1u);
  }
  if (cm) {
    __syncthreads();
    for (int i = threadIdx.x; i < C * C; i += kThreads)
      if (s_cm[i]) atomicAdd((unsigned long long*)&cm[i], (unsigned long long)s_cm[i]);
  }
  if (block_commit<1, 1>(ws, a, c) && threadIdx.x == 0) {
    const double s = ws_take(&ws->acc[0]);
    const unsigned long long cnt = atomicExch(&ws->cnt[0], 0ull);
    out[0] = (float)(s / (double)cnt);                              // NLLLoss 'mean': 0/0 = nan when all ignored
    out[1] = cnt ? (float)(1.0 / (double)cnt) : 0.f;
    ws->ticket = 0u;
  }
}

template <int MAXC>
__global__ void __launch_bounds__(kThreads) ce2d_bwd_kernel(const float* __restrict__ logits,
                                                            const long long* __restrict__ target, long long npix,
                                                            int C, long long HW, long long batch_stride,
                                                            long long ignore_index, const float* __restrict__ coef,
                                                            const float* __restrict__ gout, float gscale,
                                                            float* __restrict__ dlogits) {
  const float g = (gout ? __ldg(gout) : 1.f) * gscale * __ldg(coef + 1);
  const long long nth = (long long)gridDim.x * kThreads;
  for (long long pix = blockIdx.x * (long long)kThreads + threadIdx.x; pix < npix; pix += nth) {
    const long long b = pix / HW, hw = pix - b * HW;
    float* d = dlogits + b * (long long)C * HW + hw;
    const long long t = target[pix];
    if (t == ignore_index || t < 0 || t >= C) {
#pragma unroll
      for (int k = 0; k < MAXC; ++k)
        if (k < C) d[k * HW] = 0.f;
      continue;
    }
    const float* x = logits + b * batch_stride + hw;
    float v[MAXC];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < MAXC; ++k) {
      v[k] = (k < C) ? __ldg(x + k * HW) : -INFINITY;
      mx = fmaxf(mx, v[k]);
    }
    float se = 0.f;
#pragma unroll
    for (int k = 0; k < MAXC; ++k) {
      v[k] = (k < C) ? expf(v[k] - mx) : 0.f;
      se += v[k];
    }
    const float r = g / se;
#pragma unroll
    for (int k = 0; k < MAXC; ++k)
      if (k < C) d[k * HW] = fmaf(v[k], r, (k == (int)t) ? -g : 0.f);
  }
}

// ------------------------------------------------------------------------------------------ change similarity
// cosine_embedding_loss (aten Loss.cpp): cos = <a,b> / sqrt((|a|^2 + 1e-12)(|b|^2 + 1e-12));
// target +1 (unchanged): 1 - cos;  target -1 (changed): max(0, cos - margin), margin = 0.
template <int MAXC, bool BWD>
__global__ void __launch_bounds__(kThreads) sim_kernel(const float* __restrict__ x1, const float* __restrict__ x2,
                                                       const long long* __restrict__ label_change, long long npix,
                                                       int C, long long HW, long long bs1, long long bs2, LossWs* ws,
                                                       float* __restrict__ out, const float* __restrict__ gout,
                                                       float gscale, float* __restrict__ dx1, float* __restrict__ dx2) {
  float acc[1] = {0.f};
  unsigned cn[1] = {0u};
  const float g = BWD ? (gout ? __ldg(gout) : 1.f) * gscale / (float)npix : 0.f;
  const long long nth = (long long)gridDim.x * kThreads;
  for (long long pix = blockIdx.x * (long long)kThreads + threadIdx.x; pix < npix; pix += nth) {
    const long long b = pix / HW, hw = pix - b * HW;
    const float* p1 = x1 + b * bs1 + hw;
    const float* p2 = x2 + b * bs2 + hw;
    float a[MAXC], q[MAXC];
    float m1 = -INFINITY, m2 = -INFINITY;
#pragma unroll
    for (int k = 0; k < MAXC; ++k) {
      a[k] = (k < C) ? __ldg(p1 + k * HW) : -INFINITY;
      q[k] = (k < C) ? __ldg(p2 + k * HW) : -INFINITY;
      m1 = fmaxf(m1, a[k]);
      m2 = fmaxf(m2, q[k]);
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < MAXC; ++k) {
      a[k] = (k < C) ? expf(a[k] - m1) : 0.f;
      q[k] = (k < C) ? expf(q[k] - m2) : 0.f;
      s1 += a[k];
      s2 += q[k];
    }
    const float r1 = 1.f / s1, r2 = 1.f / s2;
    float dot = 0.f, n1 = 0.f, n2 = 0.f;
#pragma unroll
    for (int k = 0; k < MAXC; ++k) {
      a[k] *= r1;
      q[k] *= r2;
      dot = fmaf(a[k], q[k], dot);
      n1 = fmaf(a[k], a[k], n1);
      n2 = fmaf(q[k], q[k], n2);
    }
    n1 += 1e-12f;
    n2 += 1e-12f;
    const float rden = rsqrtf(n1 * n2);
    const float cs = dot * rden;
    const bool changed = label_change[pix] != 0;
    if (!BWD) {
      acc[0] += changed ? fmaxf(cs, 0.f) : 1.f - cs;
    } else {
      // d loss_i / d cos
      const float dc = changed ? (cs > 0.f ? g : 0.f) : -g;
      // d cos / d a_k = q_k * rden - cos * a_k / n1 ; softmax backward: dx_k = a_k (ga_k - sum_j ga_j a_j)
      float ga[MAXC], gq[MAXC];
      float da = 0.f, dq = 0.f;
      const float c1 = cs / n1, c2 = cs / n2;
#pragma unroll
      for (int k = 0; k < MAXC; ++k) {
        ga[k] = dc * fmaf(q[k], rden, -c1 * a[k]);
        gq[k] = dc * fmaf(a[k], rden, -c2 * q[k]);
        da = fmaf(ga[k], a[k], da);
        dq = fmaf(gq[k], q[k], dq);
      }
      float* d1 = dx1 + b * (long long)C * HW + hw;
      float* d2 = dx2 + b * (long long)C * HW + hw;
#pragma unroll
      for (int k = 0; k < MAXC; ++k)
        if (k < C) {
          d1[k * HW] = a[k] * (ga[k] - da);
          d2[k * HW] = q[k] * (gq[k] - dq);
        }
    }
  }
  if (!BWD) {
    if (block_commit<1, 1>(ws, acc, cn) && threadIdx.x == 0) {
      out[0] = (float)(ws_take(&ws->acc[0]) / (double)npix);
      ws->ticket = 0u;
    }
  }
}

// ------------------------------------------------------------------------------------------ confusion matrix
template <typename TG>
__global__ void __launch_bounds__(kThreads) confusion_kernel(const TG* __restrict__ gt,
                                                             const long long* __restrict__ pred, long long n, int C,
                                                             long long* __restrict__ cm) {
  __shared__ unsigned s_cm[kMaxClasses * kMaxClasses];
  for (int i = threadIdx.x; i < C * C; i += kThreads) s_cm[i] = 0u;
  __syncthreads();
  const long long nth = (long long)gridDim.x * kThreads;
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < n; i += nth) {
    const TG g = gt[i];                                            // mask on the stored value, then astype(int)
    const long long p = pred[i];
    if (g >= (TG)0 && g < (TG)C && p >= 0 && p < C) atomicAdd(&s_cm[(int)g * C + (int)p], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * C; i += kThreads)
    if (s_cm[i]) atomicAdd((unsigned long long*)&cm[i], (unsigned long long)s_cm[i]);
}

inline unsigned grid_for(long long work_items) {
  long long blocks = (work_items + kThreads - 1) / kThreads;
  if (blocks > kMaxBlocks) blocks = kMaxBlocks;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}
inline bool aligned16(const void* a, const void* b, const void* c = nullptr) {
  return (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) == 0;
}

}  // namespace

extern "C" int c3d_bce_dice_fwd(const float* pred, const float* target, long long n, void* ws, float* out,
                                long long* cm, void* stream_) {
  if (!pred || !target || !ws || !out || n <= 0) return C3D_ERR_ARG;
  const int vec = aligned16(pred, target) ? 1 : 0;
  bce_dice_fwd_kernel<<<grid_for(vec ? (n + 3) / 4 : n), kThreads, 0, (cudaStream_t)stream_>>>(
      pred, target, n, vec, (LossWs*)ws, out, cm);
  return c3d_check_last(cudaGetLastError());
}

extern "C" int c3d_bce_dice_bwd(const float* pred, const float* target, const float* coef, const float* gout,
                                float gscale, float* dpred, long long n, void* stream_) {
  if (!pred || !target || !coef || !dpred || n <= 0) return C3D_ERR_ARG;
  const int vec = aligned16(pred, target, dpred) ? 1 : 0;
  bce_dice_bwd_kernel<<<grid_for(vec ? (n + 3) / 4 : n), kThreads, 0, (cudaStream_t)stream_>>>(
      pred, target, coef, gout, gscale, dpred, n, vec);
  return c3d_check_last(cudaGetLastError());
}

extern "C" int c3d_ce2d_fwd(const float* logits, const long long* target, int B, int C, long long HW,
                            long long batch_stride, long long ignore_index, void* ws, float* out,
                            long long* argmax_out, long long* cm, void* stream_) {
  if (!logits || !target || !ws || !out || B <= 0 || C <= 0 || C > kMaxClasses || HW <= 0) return C3D_ERR_ARG;
  if (batch_stride == 0) batch_stride = (long long)C * HW;
  const long long npix = (long long)B * HW;
  cudaStream_t st = (cudaStream_t)stream_;
  if (C <= 8)
    ce2d_fwd_kernel<8><<<grid_for(npix), kThreads, 0, st>>>(logits, target, npix, C, HW, batch_stride, ignore_index,
                                                            (LossWs*)ws, out, argmax_out, cm);
  else
    ce2d_fwd_kernel<16><<<grid_for(npix), kThreads, 0, st>>>(logits, target, npix, C, HW, batch_stride, ignore_index,
                                                             (LossWs*)ws, out, argmax_out, cm);
  return c3d_check_last(cudaGetLastError());
}

extern "C" int c3d_ce2d_bwd(const float* logits, const long long* target, int B, int C, long long HW,
                            long long batch_stride, long long ignore_index, const float* coef, const float* gout,
                            float gscale, float* dlogits, void* stream_) {
  if (!logits || !target || !coef || !dlogits || B <= 0 || C <= 0 || C > kMaxClasses || HW <= 0) return C3D_ERR_ARG;
  if (batch_stride == 0) batch_stride = (long long)C * HW;
  const long long npix = (long long)B * HW;
  cudaStream_t st = (cudaStream_t)stream_;
  if (C <= 8)
    ce2d_bwd_kernel<8><<<grid_for(npix), kThreads, 0, st>>>(logits, target, npix, C, HW, batch_stride, ignore_index,
                                                            coef, gout, gscale, dlogits);
  else
    ce2d_bwd_kernel<16><<<grid_for(npix), kThreads, 0, st>>>(logits, target, npix, C, HW, batch_stride, ignore_index,
                                                             coef, gout, gscale, dlogits);
  return c3d_check_last(cudaGetLastError());
}

extern "C" int c3d_change_similarity_fwd(const float* x1, const float* x2, const long long* label_change, int B, int C,
                                         long long HW, long long batch_stride1, long long batch_stride2, void* ws,
                                         float* out, void* stream_) {
  if (!x1 || !x2 || !label_change || !ws || !out || B <= 0 || C <= 0 || C > kMaxClasses || HW <= 0) return C3D_ERR_ARG;
  if (batch_stride1 == 0) batch_stride1 = (long long)C * HW;
  if (batch_stride2 == 0) batch_stride2 = (long long)C * HW;
  const long long npix = (long long)B * HW;
  cudaStream_t st = (cudaStream_t)stream_;
  if (C <= 8)
    sim_kernel<8, false><<<grid_for(npix), kThreads, 0, st>>>(x1, x2, label_change, npix, C, HW, batch_stride1,
                                                              batch_stride2, (LossWs*)ws, out, nullptr, 1.f, nullptr,
                                                              nullptr);
  else
    sim_kernel<16, false><<<grid_for(npix), kThreads, 0, st>>>(x1, x2, label_change, npix, C, HW, batch_stride1,
                                                               batch_stride2, (LossWs*)ws, out, nullptr, 1.f, nullptr,
                                                               nullptr);
  return c3d_check_last(cudaGetLastError());
}

extern "C" int c3d_change_similarity_bwd(const float* x1, const float* x2, const long long* label_change, int B, int C,
                                         long long HW, long long batch_stride1, long long batch_stride2,
                                         const float* gout, float gscale, float* dx1, float* dx2, void* stream_) {
  if (!x1 || !x2 || !label_change || !dx1 || !dx2 || B <= 0 || C <= 0 || C > kMaxClasses || HW <= 0)
    return C3D_ERR_ARG;
  if (batch_stride1 == 0) batch_stride1 = (long long)C * HW;
  if (batch_stride2 == 0) batch_stride2 = (long long)C * HW;
  const long long npix = (long long)B * HW;
  cudaStream_t st = (cudaStream_t)stream_;
  if (C <= 8)
    sim_kernel<8, true><<<grid_for(npix), kThreads, 0, st>>>(x1, x2, label_change, npix, C, HW, batch_stride1,
                                                             batch_stride2, nullptr, nullptr, gout, gscale, dx1, dx2);
  else
    sim_kernel<16, true><<<grid_for(npix), kThreads, 0, st>>>(x1, x2, label_change, npix, C, HW, batch_stride1,
                                                              batch_stride2, nullptr, nullptr, gout, gscale, dx1, dx2);
  return c3d_check_last(cudaGetLastError());
}

extern "C" int c3d_confusion_matrix(const void* gt, int gt_is_float, const long long* pred, long long n,
                                    int num_classes, long long* cm, void* stream_) {
  if (!gt || !pred || !cm || n <= 0 || num_classes <= 0 || num_classes > kMaxClasses) return C3D_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream_;
  if (gt_is_float)
    confusion_kernel<float><<<grid_for(n), kThreads, 0, st>>>((const float*)gt, pred, n, num_classes, cm);
  else
    confusion_kernel<long long><<<grid_for(n), kThreads, 0, st>>>((const long long*)gt, pred, n, num_classes, cm);
  return c3d_check_last(cudaGetLastError());
}
