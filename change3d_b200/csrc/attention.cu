// Attention core of the change-captioning head: the scaled-dot-product part of the
// nn.MultiheadAttention calls in Mesh_TransformerDecoderLayer.forward (model/caption_decoder.py:393-423: causal
// self-attention over <= 52 tokens, cross-attention to the 256-token memory; 8 heads of 24 channels), forward and
// backward, and the same kernel with one query row per sequence for the cached decode step of the batched greedy / beam
// search (scripts/train_CC.py:209-322 re-runs the whole decoder over all 52 positions for every generated token).
//
// One CTA per (batch element, head): Q, K, V and the Lq x Lk probability tile live in shared memory (<= 64 x 256 x 4 B),
// fp32 FFMA throughout — the head is ~1 % of a CC step's FLOPs and these tiles are far below tensor-core tile sizes;
// what this removes is the six eager kernels per attention call and their HBM round trips of the score matrix.
//   S = scale * Q K^T (+ causal mask) -> P = softmax_rows(S) -> [dropout: P * keep / (1 - p)] -> O = P V
// Operand layout: element (position l, batch b, head h, channel d) at  ptr + l * ls + b * bs + h * hd + d,  which covers
// nn.MultiheadAttention's (L, B, E) projections and slices of a packed (L, B, 3E) in-projection alike.
#include <stdint.h>

#include "c3d_common.cuh"
#include "../../include/change3d_b200.h"

namespace {

struct AttnArgs {
  const float* q; const float* k; const float* v;
  long long q_ls, q_bs, k_ls, k_bs, v_ls, v_bs;
  float* o; long long o_ls, o_bs;
  float* P;                  // (B, nh, Lq, Lk) softmax probabilities before dropout, or null
  const uint8_t* keep;       // (B, nh, Lq, Lk) dropout keep mask, or null
  float keep_scale;          // 1 / (1 - p)
  int B, nh, hd, Lq, Lk, causal;
  float scale;
};

constexpr int ATT_THREADS = 256;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(ATT_THREADS) mha_fwd_kernel(const AttnArgs a) {
  extern __shared__ float sm[];
  const int hd = a.hd, hp = hd + 1, Lq = a.Lq, Lk = a.Lk, lp = Lk + 1;
  float* Qs = sm;                       // [Lq][hd]
  float* Ks = Qs + Lq * hd;             // [Lk][hd + 1]
  float* Vs = Ks + Lk * hp;             // [Lk][hd + 1]
  float* Ps = Vs + Lk * hp;             // [Lq][Lk + 1]
  const int b = blockIdx.x / a.nh, h = blockIdx.x - b * a.nh;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long hoff = (long long)h * hd;
  for (int i = tid; i < Lq * hd; i += ATT_THREADS) {
    const int l = i / hd, d = i - l * hd;
    Qs[i] = a.q[l * a.q_ls + b * a.q_bs + hoff + d];
  }
  for (int i = tid; i < Lk * hd; i += ATT_THREADS) {
    const int l = i / hd, d = i - l * hd;
    Ks[l * hp + d] = a.k[l * a.k_ls + b * a.k_bs + hoff + d];
    Vs[l * hp + d] = a.v[l * a.v_ls + b * a.v_bs + hoff + d];
  }
  __syncthreads();
  const int shift = Lk - Lq;            // causal: key j visible to query i iff j <= i + shift
  for (int idx = tid; idx < Lq * Lk; idx += ATT_THREADS) {
    const int i = idx / Lk, j = idx - i * Lk;
    float s = 0.f;
    for (int d = 0; d < hd; ++d) s = fmaf(Qs[i * hd + d], Ks[j * hp + d], s);
    s *= a.scale;
    if (a.causal && j > i + shift) s = -INFINITY;
    Ps[i * lp + j] = s;
  }
  __syncthreads();
  const long long pbase = ((long long)(b * a.nh + h) * Lq) * Lk;
  for (int i = warp; i < Lq; i += ATT_THREADS / 32) {
    float m = -INFINITY;
    for (int j = lane; j < Lk; j += 32) m = fmaxf(m, Ps[i * lp + j]);
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < Lk; j += 32) {
      const float e = __expf(Ps[i * lp + j] - m);
      Ps[i * lp + j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j = lane; j < Lk; j += 32) {
      float p = Ps[i * lp + j] * inv;
      if (a.P) a.P[pbase + (long long)i * Lk + j] = p;
      if (a.keep) p = a.keep[pbase + (long long)i * Lk + j] ? p * a.keep_scale : 0.f;
      Ps[i * lp + j] = p;
    }
  }
  __syncthreads();
  for (int idx = tid; idx < Lq * hd; idx += ATT_THREADS) {
    const int i = idx / hd, d = idx - i * hd;
    float o0 = 0.f, o1 = 0.f;
    int j = 0;
    for (; j + 1 < Lk; j += 2) {
      o0 = fmaf(Ps[i * lp + j], Vs[j * hp + d], o0);
      o1 = fmaf(Ps[i * lp + j + 1], Vs[(j + 1) * hp + d], o1);
    }
    if (j < Lk) o0 = fmaf(Ps[i * lp + j], Vs[j * hp + d], o0);
    a.o[i * a.o_ls + b * a.o_bs + hoff + d] = o0 + o1;
  }
}

struct AttnBwdArgs {
  AttnArgs f;                // q, k, v, P (required), keep, geometry; `o` unused
  const float* dO; long long do_ls, do_bs;
  float* dq; float* dk; float* dv;      // same layouts as q / k / v (ls / bs of the forward operands)
};

__global__ void __launch_bounds__(ATT_THREADS) mha_bwd_kernel(const AttnBwdArgs g) {
  extern __shared__ float sm[];
  const AttnArgs& a = g.f;
  const int hd = a.hd, hp = hd + 1, Lq = a.Lq, Lk = a.Lk, lp = Lk + 1;
  float* Qs = sm;                       // [Lq][hd + 1]
  float* dOs = Qs + Lq * hp;            // [Lq][hd + 1]
  float* Ks = dOs + Lq * hp;            // [Lk][hd + 1]
  float* Vs = Ks + Lk * hp;             // [Lk][hd + 1]
  float* Ps = Vs + Lk * hp;             // [Lq][Lk + 1]  probabilities, later dS
  float* Ds = Ps + Lq * lp;             // [Lq][Lk + 1]  dropped probabilities (for dV), later dP
  float* delta = Ds + Lq * lp;          // [Lq]
  const int b = blockIdx.x / a.nh, h = blockIdx.x - b * a.nh;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long hoff = (long long)h * hd;
  const long long pbase = ((long long)(b * a.nh + h) * Lq) * Lk;
  for (int i = tid; i < Lq * hd; i += ATT_THREADS) {
    const int l = i / hd, d = i - l * hd;
    Qs[l * hp + d] = a.q[l * a.q_ls + b * a.q_bs + hoff + d];
    dOs[l * hp + d] = g.dO[l * g.do_ls + b * g.do_bs + hoff + d];
  }
  for (int i = tid; i < Lk * hd; i += ATT_THREADS) {
    const int l = i / hd, d = i - l * hd;
    Ks[l * hp + d] = a.k[l * a.k_ls + b * a.k_bs + hoff + d];
    Vs[l * hp + d] = a.v[l * a.v_ls + b * a.v_bs + hoff + d];
  }
  for (int idx = tid; idx < Lq * Lk; idx += ATT_THREADS) {
    const int i = idx / Lk, j = idx - i * Lk;
    const float p = a.P[pbase + idx];
    const float m = a.keep ? (a.keep[pbase + idx] ? a.keep_scale : 0.f) : 1.f;
    Ps[i * lp + j] = p;
    Ds[i * lp + j] = p * m;
  }
  __syncthreads();
  // dV[j][d] = sum_i Pd[i][j] dO[i][d]
  for (int idx = tid; idx < Lk * hd; idx += ATT_THREADS) {
    const int j = idx / hd, d = idx - j * hd;
    float s = 0.f;
    for (int i = 0; i < Lq; ++i) s = fmaf(Ds[i * lp + j], dOs[i * hp + d], s);
    g.dv[j * a.v_ls + b * a.v_bs + hoff + d] = s;
  }
  __syncthreads();
  // dP[i][j] = (dO[i] . V[j]) * keep;  delta_i = sum_j dP[i][j] P[i][j]
  for (int i = warp; i < Lq; i += ATT_THREADS / 32) {
    float dl = 0.f;
    for (int j = lane; j < Lk; j += 32) {
      float s = 0.f;
      for (int d = 0; d < hd; ++d) s = fmaf(dOs[i * hp + d], Vs[j * hp + d], s);
      if (a.keep) s = a.keep[pbase + (long long)i * Lk + j] ? s * a.keep_scale : 0.f;
      Ds[i * lp + j] = s;
      dl = fmaf(s, Ps[i * lp + j], dl);
    }
    dl = warp_sum(dl);
    if (lane == 0) delta[i] = dl;
  }
  __syncthreads();
  // dS = P * (dP - delta), scaled by the logit scale (masked entries have P = 0)
  for (int idx = tid; idx < Lq * Lk; idx += ATT_THREADS) {
    const int i = idx / Lk, j = idx - i * Lk;
    Ps[i * lp + j] = Ps[i * lp + j] * (Ds[i * lp + j] - delta[i]) * a.scale;
  }
  __syncthreads();
  for (int idx = tid; idx < Lq * hd; idx += ATT_THREADS) {      // dQ[i][d] = sum_j dS[i][j] K[j][d]
    const int i = idx / hd, d = idx - i * hd;
    float s = 0.f;
    for (int j = 0; j < Lk; ++j) s = fmaf(Ps[i * lp + j], Ks[j * hp + d], s);
    g.dq[i * a.q_ls + b * a.q_bs + hoff + d] = s;
  }
  for (int idx = tid; idx < Lk * hd; idx += ATT_THREADS) {      // dK[j][d] = sum_i dS[i][j] Q[i][d]
    const int j = idx / hd, d = idx - j * hd;
    float s = 0.f;
    for (int i = 0; i < Lq; ++i) s = fmaf(Ps[i * lp + j], Qs[i * hp + d], s);
    g.dk[j * a.k_ls + b * a.k_bs + hoff + d] = s;
  }
}

int check(const c3d_attn_desc* d) {
  if (!d || !d->q || !d->k || !d->v || d->B <= 0 || d->nh <= 0 || d->hd <= 0 || d->Lq <= 0 || d->Lk <= 0) return C3D_ERR_ARG;
  if (d->Lq > 64 || d->Lk > 256 || d->hd > 64) return C3D_ERR_ARG;
  return C3D_OK;
}

AttnArgs fill(const c3d_attn_desc* d) {
  AttnArgs a;
  a.q = d->q; a.k = d->k; a.v = d->v;
  a.q_ls = d->q_ls; a.q_bs = d->q_bs; a.k_ls = d->k_ls; a.k_bs = d->k_bs; a.v_ls = d->v_ls; a.v_bs = d->v_bs;
  a.o = d->o; a.o_ls = d->o_ls; a.o_bs = d->o_bs;
  a.P = d->P; a.keep = d->keep; a.keep_scale = d->keep_scale;
  a.B = d->B; a.nh = d->nh; a.hd = d->hd; a.Lq = d->Lq; a.Lk = d->Lk; a.causal = d->causal;
  a.scale = d->scale;
  return a;
}

}  // namespace

extern "C" int c3d_attention_fwd(const c3d_attn_desc* d, void* stream_) {
  if (int e = check(d)) return e;
  if (!d->o) return C3D_ERR_ARG;
  const AttnArgs a = fill(d);
  const size_t smem = ((size_t)a.Lq * a.hd + 2 * (size_t)a.Lk * (a.hd + 1) + (size_t)a.Lq * (a.Lk + 1)) * sizeof(float);
  if (cudaFuncSetAttribute(mha_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return C3D_ERR_SMEM;
  mha_fwd_kernel<<<(unsigned)(a.B * a.nh), ATT_THREADS, smem, (cudaStream_t)stream_>>>(a);
  return c3d_check_last(cudaGetLastError());
}

extern "C" int c3d_attention_bwd(const c3d_attn_desc* d, const float* dO, long long do_ls, long long do_bs, float* dq,
                                 float* dk, float* dv, void* stream_) {
  if (int e = check(d)) return e;
  if (!d->P || !dO || !dq || !dk || !dv) return C3D_ERR_ARG;
  AttnBwdArgs g;
  g.f = fill(d);
  g.dO = dO; g.do_ls = do_ls; g.do_bs = do_bs; g.dq = dq; g.dk = dk; g.dv = dv;
  const AttnArgs& a = g.f;
  const size_t smem = (2 * (size_t)a.Lq * (a.hd + 1) + 2 * (size_t)a.Lk * (a.hd + 1) + 2 * (size_t)a.Lq * (a.Lk + 1) + a.Lq) * sizeof(float);
  if (cudaFuncSetAttribute(mha_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return C3D_ERR_SMEM;
  mha_bwd_kernel<<<(unsigned)(a.B * a.nh), ATT_THREADS, smem, (cudaStream_t)stream_>>>(g);
  return c3d_check_last(cudaGetLastError());
}
