// Depthwise 3x3x3 Conv3d (X3D conv_b, model/x3d.py:184-193): groups = C, pad 1, stride (1,s,s).
// NDHWC, V (= 2 or 4) channels per thread; the temporal extent (T <= 5) is kept whole in registers; BN_a + ReLU
// are applied to the input on the fly; the epilogue accumulates the per-sample per-channel sum / sum-of-squares
// that BN_b and the SE pool need.  V = 2 halves the register footprint (the backward keeps 27 weight-gradient
// accumulators per channel group in registers) at the price of 8-byte instead of 16-byte accesses.
#include <stdlib.h>
#include "c3d_common.cuh"
#include "../../include/change3d_b200.h"

template <int V> struct VF { float v[V]; };

template <int V> __device__ __forceinline__ VF<V> vzero() {
  VF<V> r;
#pragma unroll
  for (int i = 0; i < V; ++i) r.v[i] = 0.f;
  return r;
}
template <int V> __device__ __forceinline__ VF<V> vldg(const float* p) {
  VF<V> r;
  if (V == 4) { const float4 t = __ldg(reinterpret_cast<const float4*>(p)); r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; }
  else { const float2 t = __ldg(reinterpret_cast<const float2*>(p)); r.v[0] = t.x; r.v[1] = t.y; }
  return r;
}
template <int V> __device__ __forceinline__ VF<V> vlds(const float* p) {
  VF<V> r;
  if (V == 4) { const float4 t = *reinterpret_cast<const float4*>(p); r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; }
  else { const float2 t = *reinterpret_cast<const float2*>(p); r.v[0] = t.x; r.v[1] = t.y; }
  return r;
}
template <int V> __device__ __forceinline__ void vst(float* p, const VF<V>& a) {
  if (V == 4) *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
  else *reinterpret_cast<float2*>(p) = make_float2(a.v[0], a.v[1]);
}
template <int V> __device__ __forceinline__ void vfma(VF<V>& acc, const VF<V>& a, const VF<V>& b) {
#pragma unroll
  for (int i = 0; i < V; ++i) acc.v[i] = fmaf(a.v[i], b.v[i], acc.v[i]);
}
// relu((x - mean) * scale + beta)
template <int V> __device__ __forceinline__ VF<V> vbnrelu(const VF<V>& x, const VF<V>& mean, const VF<V>& scale, const VF<V>& beta) {
  VF<V> r;
#pragma unroll
  for (int i = 0; i < V; ++i) r.v[i] = fmaxf(fmaf(x.v[i] - mean.v[i], scale.v[i], beta.v[i]), 0.f);
  return r;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int T, int OWT, int S, int V, int MINB>
__global__ void __launch_bounds__(256, MINB) dw_fwd_kernel(const float* __restrict__ X, const float* __restrict__ bnp,
                                                           const float* __restrict__ w, float* __restrict__ Y,
                                                           double* __restrict__ stats, int IH, int IW, int OH, int OW,
                                                           int C, int Cs) {
  constexpr int NIN = (OWT - 1) * S + 3;
  extern __shared__ __align__(16) float sm[];
  float* ws = sm;                 // [27][Cs]
  float* s_sum = ws + 27 * Cs;    // [Cs]
  float* s_sq = s_sum + Cs;       // [Cs]
  const int n = blockIdx.x / OH, oh = blockIdx.x - n * OH;
  for (int i = threadIdx.x; i < 27 * Cs; i += 256) {
    int tap = i / Cs, c = i - tap * Cs;
    ws[i] = c < C ? __ldg(w + c * 27 + tap) : 0.f;
  }
  for (int i = threadIdx.x; i < 2 * Cs; i += 256) s_sum[i] = 0.f;
  __syncthreads();

  const int nq = Cs / V;
  const int nowg = (OW + OWT - 1) / OWT;
  for (int item = threadIdx.x; item < nowg * nq; item += 256) {
    const int owg = item / nq, q = item - owg * nq, c = V * q;
    const int ow0 = owg * OWT;
    const VF<V> mean = vldg<V>(bnp + c), scale = vldg<V>(bnp + 2 * Cs + c), beta = vldg<V>(bnp + 3 * Cs + c);
    VF<V> acc[T][OWT];
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
      for (int j = 0; j < OWT; ++j) acc[t][j] = vzero<V>();

#pragma unroll
    for (int ti = 0; ti < T; ++ti) {
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int ih = oh * S - 1 + kh;
        if (ih < 0 || ih >= IH) continue;
        const float* xrow = X + (((long long)(n * T + ti) * IH + ih) * IW) * Cs + c;
        VF<V> in[NIN];
#pragma unroll
        for (int x = 0; x < NIN; ++x) {
          const int iw = ow0 * S - 1 + x;
          in[x] = (iw >= 0 && iw < IW) ? vbnrelu<V>(vldg<V>(xrow + (long long)iw * Cs), mean, scale, beta) : vzero<V>();
        }
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
          const int to = ti - kt + 1;
          if (to < 0 || to >= T) continue;
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const VF<V> wv = vlds<V>(ws + (kt * 9 + kh * 3 + kw) * Cs + c);
#pragma unroll
            for (int j = 0; j < OWT; ++j) vfma<V>(acc[to][j], wv, in[j * S + kw]);
          }
        }
      }
    }
    VF<V> s = vzero<V>(), sq = vzero<V>();
#pragma unroll
    for (int t = 0; t < T; ++t) {
      float* yrow = Y + (((long long)(n * T + t) * OH + oh) * OW) * Cs + c;
#pragma unroll
      for (int j = 0; j < OWT; ++j) {
        if (ow0 + j < OW) {
          vst<V>(yrow + (long long)(ow0 + j) * Cs, acc[t][j]);
#pragma unroll
          for (int e = 0; e < V; ++e) { s.v[e] += acc[t][j].v[e]; sq.v[e] = fmaf(acc[t][j].v[e], acc[t][j].v[e], sq.v[e]); }
        }
      }
    }
    if (stats) {
#pragma unroll
      for (int e = 0; e < V; ++e) { atomicAdd(&s_sum[c + e], s.v[e]); atomicAdd(&s_sq[c + e], sq.v[e]); }
    }
  }
  if (stats) {
    __syncthreads();
    for (int c = threadIdx.x; c < Cs; c += 256) {
      atomicAdd(stats + ((long long)n * 2 + 0) * Cs + c, (double)s_sum[c]);
      atomicAdd(stats + ((long long)n * 2 + 1) * Cs + c, (double)s_sq[c]);
    }
  }
}

static int dw_env(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

template <int T, int OWT, int V, int MINB>
static int launch_dw_fwd(const float* X, const float* bnp, const float* w, float* Y, double* stats, int N, int IH,
                         int IW, int C, int Cs, int stride, cudaStream_t st) {
  const int OH = (IH - 1) / stride + 1, OW = (IW - 1) / stride + 1;   // k=3, pad=1
  const size_t smem = (size_t)(29 * Cs) * sizeof(float);
  dim3 grid((unsigned)(N * OH));
  if (stride == 1) {
    cudaFuncSetAttribute(dw_fwd_kernel<T, OWT, 1, V, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dw_fwd_kernel<T, OWT, 1, V, MINB><<<grid, 256, smem, st>>>(X, bnp, w, Y, stats, IH, IW, OH, OW, C, Cs);
  } else {
    cudaFuncSetAttribute(dw_fwd_kernel<T, OWT, 2, V, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dw_fwd_kernel<T, OWT, 2, V, MINB><<<grid, 256, smem, st>>>(X, bnp, w, Y, stats, IH, IW, OH, OW, C, Cs);
  }
  return c3d_check_last(cudaGetLastError());
}

int c3d_launch_dw_fwd_ring(const float* X, const float* bnp, const float* w, float* Y, double* stats, int N, int T, int IH,
                           int IW, int C, int Cs, cudaStream_t st);      // dw_ring.cu

int c3d_launch_dw_fwd_ring2(const float* X, const float* bnp, const float* w, float* Y, double* stats, int N, int T, int IH,
                            int IW, int C, int Cs, cudaStream_t st);     // dw_ring.cu (stride 2)
int c3d_launch_dw_bwd_ring2(const float* dy, const float* ya, const float* bnp_a, const float* w, float* dr, float* dW,
                            double* stats_a, int N, int T, int IH, int IW, int C, int Cs, cudaStream_t st);

extern "C" int c3d_dw_conv_fwd(const float* X, const float* bnp_a, const float* w, float* Y, double* stats, int N,
                               int T, int IH, int IW, int C, int Cs, int stride, void* stream_) {
  if (!X || !bnp_a || !w || !Y || N <= 0 || IH <= 0 || IW <= 0 || C <= 0 || Cs < C || (Cs & 3)) return C3D_ERR_ARG;
  if (stride != 1 && stride != 2) return C3D_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream_;
  if (stride == 1 && T >= 3 && T <= 5) {          // row-streaming kernel (each input element read from HBM once)
    const int r = c3d_launch_dw_fwd_ring(X, bnp_a, w, Y, stats, N, T, IH, IW, C, Cs, st);
    if (r >= 0) return r;
  }
  if (stride == 2 && T >= 3 && T <= 5) {
    const int r = c3d_launch_dw_fwd_ring2(X, bnp_a, w, Y, stats, N, T, IH, IW, C, Cs, st);
    if (r >= 0) return r;
  }
  const int variant = dw_env("C3D_DW_FWD", 0);     // tuning switch: channels/thread x outputs/thread
  switch (T) {
    case 3:
      // measured on B200 (scratch/bench_dw.py): 2 channels/thread wins; 4 outputs/thread at stride 1
      // (1.1-1.3 TB/s), 2 outputs/thread + 3 CTAs/SM at stride 2 (2.6 TB/s)
      if (variant == 1) return launch_dw_fwd<3, 4, 4, 1>(X, bnp_a, w, Y, stats, N, IH, IW, C, Cs, stride, st);
      if (variant == 4) return launch_dw_fwd<3, 2, 4, 2>(X, bnp_a, w, Y, stats, N, IH, IW, C, Cs, stride, st);
      if (variant == 2 || (variant == 0 && stride == 2))
        return launch_dw_fwd<3, 2, 2, 3>(X, bnp_a, w, Y, stats, N, IH, IW, C, Cs, stride, st);
      return launch_dw_fwd<3, 4, 2, 2>(X, bnp_a, w, Y, stats, N, IH, IW, C, Cs, stride, st);
    case 4: return launch_dw_fwd<4, 2, 4, 2>(X, bnp_a, w, Y, stats, N, IH, IW, C, Cs, stride, st);
    case 5: return launch_dw_fwd<5, 2, 2, 2>(X, bnp_a, w, Y, stats, N, IH, IW, C, Cs, stride, st);
    default: return C3D_ERR_ARG;
  }
}

// ------------------------------------------------------------------------------------------------
// Backward of conv_b fused with the BN_b / SE backward on its output side and the ReLU + BN_a
// statistics on its input side:
//   dy_b = scale_b * (du*gate + dpool - c1 - zhat*c2)            (computed on the fly from du, y_b)
//   da   = conv_transpose(dy_b, w);  dW[tap] += a * dy_b          (a = relu(bn_a(y_a)) recomputed)
//   dr   = da * (a > 0)  -> written;  stats_a += (sum dr, sum dr * yhat_a)
// A thread owns one group of V channels for the whole launch (persistent over (n, ih) rows), so the 27
// weight-gradient accumulators live in registers and are flushed once.
// ------------------------------------------------------------------------------------------------
// dy_b = scale_b * (du*gate + dpool - c1 - zhat*c2), elementwise and in place over du: the stencil then reads every
// dy value once per tap instead of recomputing it from two tensors for each of the ~6 threads that need it.
__global__ void __launch_bounds__(256) dw_dy_kernel(float* __restrict__ du, const float* __restrict__ yb,
                                                    const float* __restrict__ bnp_b, const float* __restrict__ gate,
                                                    const float* __restrict__ dpool, const float* __restrict__ coef_b,
                                                    long long total4, int Cs, long long rows_per_sample) {
  const int q4 = Cs >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / q4;
    const int c = (int)(i - row * q4) * 4;
    const float4 mean = ldg4(bnp_b + c), rstd = ldg4(bnp_b + Cs + c), scale = ldg4(bnp_b + 2 * Cs + c);
    const float4 c1 = ldg4(coef_b + c), c2 = ldg4(coef_b + Cs + c);
    float4 g = make_float4(1.f, 1.f, 1.f, 1.f), dp = f4zero();
    if (gate) {
      const long long n = row / rows_per_sample;
      g = ldg4(gate + n * Cs + c);
      dp = ldg4(dpool + n * Cs + c);
    }
    const float4 d = *reinterpret_cast<const float4*>(du + i * 4);
    const float4 y = ldg4(yb + i * 4);
    float4 o;
    o.x = scale.x * (fmaf(d.x, g.x, dp.x) - c1.x - (y.x - mean.x) * rstd.x * c2.x);
    o.y = scale.y * (fmaf(d.y, g.y, dp.y) - c1.y - (y.y - mean.y) * rstd.y * c2.y);
    o.z = scale.z * (fmaf(d.z, g.z, dp.z) - c1.z - (y.z - mean.z) * rstd.z * c2.z);
    o.w = scale.w * (fmaf(d.w, g.w, dp.w) - c1.w - (y.w - mean.w) * rstd.w * c2.w);
    st4(du + i * 4, o);
  }
}

template <int T, int IWT, int S, int V, bool PRE>
__global__ void __launch_bounds__(128) dw_bwd_kernel(
    const float* __restrict__ du, const float* __restrict__ yb, const float* __restrict__ bnp_b,
    const float* __restrict__ gate, const float* __restrict__ dpool, const float* __restrict__ coef_b,
    const float* __restrict__ ya, const float* __restrict__ bnp_a, const float* __restrict__ w,
    float* __restrict__ dr, float* __restrict__ dW, double* __restrict__ stats_a, int N, int IH, int IW, int OH,
    int OW, int C, int Cs) {
  constexpr int NSEG = (S == 1) ? IWT + 2 : IWT / 2 + 1;
  static_assert(S == 1 || (IWT % 2) == 0, "stride-2 path pairs even/odd input columns");
  extern __shared__ __align__(16) float sm[];
  float* ws = sm;                  // [27][Cs]
  float* s_dw = ws + 27 * Cs;      // [27][Cs]
  float* s_st = s_dw + 27 * Cs;    // [2][Cs]
  const int nq = Cs / V;
  const int nslots = blockDim.x / nq;
  const int tid = threadIdx.x;
  for (int i = tid; i < 27 * Cs; i += blockDim.x) {
    int tap = i / Cs, c = i - tap * Cs;
    ws[i] = c < C ? __ldg(w + c * 27 + tap) : 0.f;
    s_dw[i] = 0.f;
  }
  for (int i = tid; i < 2 * Cs; i += blockDim.x) s_st[i] = 0.f;
  __syncthreads();
  const bool active = tid < nq * nslots;
  const int q = tid % nq, slot = tid / nq, c = V * q;

  VF<V> dwacc[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) dwacc[i] = vzero<V>();
  VF<V> st_s = vzero<V>(), st_t = vzero<V>();

  if (active) {
    const VF<V> mean_b = vldg<V>(bnp_b + c), rstd_b = vldg<V>(bnp_b + Cs + c), scale_b = vldg<V>(bnp_b + 2 * Cs + c);
    const VF<V> c1 = vldg<V>(coef_b + c), c2 = vldg<V>(coef_b + Cs + c);
    // dy = A1*du + A0 - A2*(yb - mean_b)
    VF<V> A2;
#pragma unroll
    for (int e = 0; e < V; ++e) A2.v[e] = scale_b.v[e] * c2.v[e] * rstd_b.v[e];
    const int niwg = (IW + IWT - 1) / IWT;
    for (int row = blockIdx.x; row < N * IH; row += gridDim.x) {
      const int n = row / IH, ih = row - n * IH;
      VF<V> A1 = scale_b, A0;
      if (gate) {
        const VF<V> g4 = vldg<V>(gate + (long long)n * Cs + c), dp4 = vldg<V>(dpool + (long long)n * Cs + c);
#pragma unroll
        for (int e = 0; e < V; ++e) { A1.v[e] = scale_b.v[e] * g4.v[e]; A0.v[e] = scale_b.v[e] * (dp4.v[e] - c1.v[e]); }
      } else {
#pragma unroll
        for (int e = 0; e < V; ++e) A0.v[e] = -scale_b.v[e] * c1.v[e];
      }
      for (int iwg = slot; iwg < niwg; iwg += nslots) {
        const int iw0 = iwg * IWT;
        VF<V> a[T][IWT], da[T][IWT];
        {
          const VF<V> mean_a = vldg<V>(bnp_a + c), scale_a = vldg<V>(bnp_a + 2 * Cs + c), beta_a = vldg<V>(bnp_a + 3 * Cs + c);
#pragma unroll
          for (int t = 0; t < T; ++t)
#pragma unroll
            for (int j = 0; j < IWT; ++j) {
              da[t][j] = vzero<V>();
              a[t][j] = (iw0 + j < IW)
                            ? vbnrelu<V>(vldg<V>(ya + ((((long long)n * T + t) * IH + ih) * IW + iw0 + j) * Cs + c), mean_a, scale_a, beta_a)
                            : vzero<V>();
            }
        }
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const int num_h = ih + 1 - kh;
          if (S == 2 && (num_h & 1)) continue;
          const int oh = num_h / S;
          if (num_h < 0 || oh >= OH) continue;
          const int owlo = (S == 1) ? iw0 - 1 : iw0 / 2;
#pragma unroll
          for (int to = 0; to < T; ++to) {
            VF<V> dy[NSEG];
            const long long obase = ((((long long)n * T + to) * OH + oh) * OW) * Cs + c;
#pragma unroll
            for (int x = 0; x < NSEG; ++x) {
              const int ow = owlo + x;
              if (ow >= 0 && ow < OW) {
                const VF<V> d = vldg<V>(du + obase + (long long)ow * Cs);
                if (PRE) {
                  dy[x] = d;                                  // du already holds dy_b (dw_dy_kernel)
                } else {
                  const VF<V> y = vldg<V>(yb + obase + (long long)ow * Cs);
#pragma unroll
                  for (int e = 0; e < V; ++e) dy[x].v[e] = fmaf(A1.v[e], d.v[e], A0.v[e]) - A2.v[e] * (y.v[e] - mean_b.v[e]);
                }
              } else {
                dy[x] = vzero<V>();
              }
            }
#pragma unroll
            for (int kt = 0; kt < 3; ++kt) {
              const int ti = to + kt - 1;
              if (ti < 0 || ti >= T) continue;
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) {
                const int tap = kt * 9 + kh * 3 + kw;
                const VF<V> wv = vlds<V>(ws + tap * Cs + c);
#pragma unroll
                for (int j = 0; j < IWT; ++j) {
                  const int num_w = j + 1 - kw;           // relative to iw0 (iw0 is even when S == 2)
                  if (S == 2 && (num_w & 1)) continue;
                  const int x = (S == 1) ? num_w + 1 : num_w / 2;
                  if (x < 0 || x >= NSEG) continue;
                  vfma<V>(da[ti][j], wv, dy[x]);
                  vfma<V>(dwacc[tap], a[ti][j], dy[x]);
                }
              }
            }
          }
        }
        // epilogue: ReLU mask, store, BN_a backward statistics
        const VF<V> mean_a = vldg<V>(bnp_a + c), rstd_a = vldg<V>(bnp_a + Cs + c);
#pragma unroll
        for (int t = 0; t < T; ++t)
#pragma unroll
          for (int j = 0; j < IWT; ++j) {
            if (iw0 + j >= IW) continue;
            const long long off = ((((long long)n * T + t) * IH + ih) * IW + iw0 + j) * Cs + c;
            VF<V> v = da[t][j];
#pragma unroll
            for (int e = 0; e < V; ++e) v.v[e] = a[t][j].v[e] > 0.f ? v.v[e] : 0.f;
            vst<V>(dr + off, v);
            const VF<V> y = vldg<V>(ya + off);
#pragma unroll
            for (int e = 0; e < V; ++e) {
              st_s.v[e] += v.v[e];
              st_t.v[e] = fmaf(v.v[e], (y.v[e] - mean_a.v[e]) * rstd_a.v[e], st_t.v[e]);
            }
          }
      }
    }
    // flush register accumulators
#pragma unroll
    for (int tap = 0; tap < 27; ++tap)
#pragma unroll
      for (int e = 0; e < V; ++e) atomicAdd(&s_dw[tap * Cs + c + e], dwacc[tap].v[e]);
#pragma unroll
    for (int e = 0; e < V; ++e) { atomicAdd(&s_st[c + e], st_s.v[e]); atomicAdd(&s_st[Cs + c + e], st_t.v[e]); }
  }
  __syncthreads();
  for (int i = tid; i < 27 * Cs; i += blockDim.x) {
    int tap = i / Cs, cc = i - tap * Cs;
    if (cc < C) atomicAdd(dW + cc * 27 + tap, s_dw[i]);
  }
  for (int i = tid; i < 2 * Cs; i += blockDim.x) atomicAdd(stats_a + i, (double)s_st[i]);
}

template <int T, int IWT1, int IWT2, int V, bool PRE>
static int launch_dw_bwd(const float* du, const float* yb, const float* bnp_b, const float* gate, const float* dpool,
                         const float* coef_b, const float* ya, const float* bnp_a, const float* w, float* dr, float* dW,
                         double* stats_a, int N, int IH, int IW, int C, int Cs, int stride, cudaStream_t st) {
  const int OH = (IH - 1) / stride + 1, OW = (IW - 1) / stride + 1;
  const int nq = Cs / V;
  if (nq > 128) return -1;
  const int threads = ((128 / nq) * nq + 31) / 32 * 32;
  const size_t smem = (size_t)(56 * Cs) * sizeof(float);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int per_sm = dw_env("C3D_DW_BWD_CTAS", 2);
  int grid = N * IH < per_sm * sms ? N * IH : per_sm * sms;
  if (stride == 1) {
    cudaFuncSetAttribute(dw_bwd_kernel<T, IWT1, 1, V, PRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dw_bwd_kernel<T, IWT1, 1, V, PRE><<<grid, threads, smem, st>>>(du, yb, bnp_b, gate, dpool, coef_b, ya, bnp_a, w, dr,
                                                                   dW, stats_a, N, IH, IW, OH, OW, C, Cs);
  } else {
    cudaFuncSetAttribute(dw_bwd_kernel<T, IWT2, 2, V, PRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dw_bwd_kernel<T, IWT2, 2, V, PRE><<<grid, threads, smem, st>>>(du, yb, bnp_b, gate, dpool, coef_b, ya, bnp_a, w, dr,
                                                                   dW, stats_a, N, IH, IW, OH, OW, C, Cs);
  }
  return c3d_check_last(cudaGetLastError());
}

int c3d_launch_dw_bwd_ring(const float* dy, const float* ya, const float* bnp_a, const float* w, float* dr, float* dW,
                           double* stats_a, int N, int T, int IH, int IW, int C, int Cs, cudaStream_t st, const float* yb,
                           const float* bnp_b, const float* gate, const float* dpool, const float* coef_b);   // dw_ring.cu

extern "C" int c3d_dw_conv_bwd(float* du, const float* y_b, const float* bnp_b, const float* gate,
                               const float* dpool, const float* coef_b, const float* y_a, const float* bnp_a,
                               const float* w, float* dr, float* dW, double* stats_a, int N, int T, int IH, int IW,
                               int C, int Cs, int stride, void* stream_) {
  if (!du || !y_b || !bnp_b || !coef_b || !y_a || !bnp_a || !w || !dr || !dW || !stats_a) return C3D_ERR_ARG;
  if ((gate == nullptr) != (dpool == nullptr)) return C3D_ERR_ARG;
  if (N <= 0 || IH <= 0 || IW <= 0 || C <= 0 || Cs < C || (Cs & 3) || (stride != 1 && stride != 2)) return C3D_ERR_ARG;
  if (stride == 2 && ((IH | IW) & 1)) return C3D_ERR_ARG;   // stride-2 path pairs even/odd input columns
  if (T < 3 || T > 5) return C3D_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream_;
  // tuning switch: 0 default = dy pre-pass + stencil (2 channels, 4 columns per thread); 1 / 2 = fused single kernel
  // with 4 / 2 channels per thread (kept for comparison, scratch/bench_dw.py)
  const int variant = dw_env("C3D_DW_BWD", 0);
  int r = -1;
  if (variant == 0 && stride == 1 && dw_env("C3D_DW_FUSE", 1)) {
    // row-streaming kernel with the BN_b / SE backward transform of du fused into its ring fill (du is left untouched)
    const int rr = c3d_launch_dw_bwd_ring(du, y_a, bnp_a, w, dr, dW, stats_a, N, T, IH, IW, C, Cs, st, y_b, bnp_b, gate, dpool, coef_b);
    if (rr >= 0) return rr;
  }
  if ((variant == 0 || variant == 3) && Cs / 2 <= 128) {
    const int OH = (IH - 1) / stride + 1, OW = (IW - 1) / stride + 1;
    const long long total4 = (long long)N * T * OH * OW * (Cs >> 2);
    long long blocks = (total4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    dw_dy_kernel<<<(unsigned)blocks, 256, 0, st>>>(du, y_b, bnp_b, gate, dpool, coef_b, total4, Cs, (long long)T * OH * OW);
    if (cudaGetLastError() != cudaSuccess) return C3D_ERR_CUDA;
    if (stride == 2) {
      const int rr = c3d_launch_dw_bwd_ring2(du, y_a, bnp_a, w, dr, dW, stats_a, N, T, IH, IW, C, Cs, st);
      if (rr >= 0) return rr;
    }
    if (stride == 1) {          // row-streaming kernel over the pre-computed dy
      const int rr = c3d_launch_dw_bwd_ring(du, y_a, bnp_a, w, dr, dW, stats_a, N, T, IH, IW, C, Cs, st, nullptr, nullptr, nullptr, nullptr, nullptr);
      if (rr >= 0) return rr;
    }
    if (variant == 3 && T == 3)
      return launch_dw_bwd<3, 2, 2, 2, true>(du, y_b, bnp_b, gate, dpool, coef_b, y_a, bnp_a, w, dr, dW, stats_a, N, IH, IW, C, Cs, stride, st);
    switch (T) {
      case 3: r = launch_dw_bwd<3, 4, 4, 2, true>(du, y_b, bnp_b, gate, dpool, coef_b, y_a, bnp_a, w, dr, dW, stats_a, N, IH, IW, C, Cs, stride, st); break;
      case 4: r = launch_dw_bwd<4, 2, 2, 2, true>(du, y_b, bnp_b, gate, dpool, coef_b, y_a, bnp_a, w, dr, dW, stats_a, N, IH, IW, C, Cs, stride, st); break;
      default: r = launch_dw_bwd<5, 2, 2, 2, true>(du, y_b, bnp_b, gate, dpool, coef_b, y_a, bnp_a, w, dr, dW, stats_a, N, IH, IW, C, Cs, stride, st); break;
    }
    return r < 0 ? C3D_ERR_ARG : r;
  }
  if (variant == 2 && Cs / 2 <= 128) {
    switch (T) {
      case 3: r = launch_dw_bwd<3, 2, 2, 2, false>(du, y_b, bnp_b, gate, dpool, coef_b, y_a, bnp_a, w, dr, dW, stats_a, N, IH, IW, C, Cs, stride, st); break;
      case 4: r = launch_dw_bwd<4, 1, 2, 2, false>(du, y_b, bnp_b, gate, dpool, coef_b, y_a, bnp_a, w, dr, dW, stats_a, N, IH, IW, C, Cs, stride, st); break;
      default: r = launch_dw_bwd<5, 1, 2, 2, false>(du, y_b, bnp_b, gate, dpool, coef_b, y_a, bnp_a, w, dr, dW, stats_a, N, IH, IW, C, Cs, stride, st); break;
    }
    return r < 0 ? C3D_ERR_ARG : r;
  }
  switch (T) {
    case 3: r = launch_dw_bwd<3, 1, 2, 4, false>(du, y_b, bnp_b, gate, dpool, coef_b, y_a, bnp_a, w, dr, dW, stats_a, N, IH, IW, C, Cs, stride, st); break;
    case 4: r = launch_dw_bwd<4, 1, 2, 4, false>(du, y_b, bnp_b, gate, dpool, coef_b, y_a, bnp_a, w, dr, dW, stats_a, N, IH, IW, C, Cs, stride, st); break;
    default: r = launch_dw_bwd<5, 1, 2, 4, false>(du, y_b, bnp_b, gate, dpool, coef_b, y_a, bnp_a, w, dr, dW, stats_a, N, IH, IW, C, Cs, stride, st); break;
  }
  return r < 0 ? C3D_ERR_ARG : r;
}
