// Depthwise 3x3x3 Conv3d (X3D conv_b, model/x3d.py:184-193): groups = C, pad 1, stride (1,s,s).
// NDHWC, one float4 = 4 channels per thread; the temporal extent (T <= 5) is kept whole in
// registers; BN_a + ReLU are applied to the input on the fly; the epilogue accumulates the
// per-sample per-channel sum / sum-of-squares that BN_b and the SE pool need.
#include "c3d_common.cuh"
#include "../../include/change3d_b200.h"

template <int T, int OWT, int S>
__global__ void __launch_bounds__(256) dw_fwd_kernel(const float* __restrict__ X, const float* __restrict__ bnp,
                                                     const float* __restrict__ w, float* __restrict__ Y,
                                                     double* __restrict__ stats, int IH, int IW, int OH, int OW, int C,
                                                     int Cs) {
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

  const int nq = Cs >> 2;
  const int nowg = (OW + OWT - 1) / OWT;
  for (int item = threadIdx.x; item < nowg * nq; item += 256) {
    const int owg = item / nq, q = item - owg * nq, c = 4 * q;
    const int ow0 = owg * OWT;
    const float4 mean = ldg4(bnp + c), scale = ldg4(bnp + 2 * Cs + c), beta = ldg4(bnp + 3 * Cs + c);
    float4 acc[T][OWT];
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
      for (int j = 0; j < OWT; ++j) acc[t][j] = f4zero();

#pragma unroll
    for (int ti = 0; ti < T; ++ti) {
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int ih = oh * S - 1 + kh;
        if (ih < 0 || ih >= IH) continue;
        const float* xrow = X + (((long long)(n * T + ti) * IH + ih) * IW) * Cs + c;
        float4 in[NIN];
#pragma unroll
        for (int x = 0; x < NIN; ++x) {
          const int iw = ow0 * S - 1 + x;
          in[x] = (iw >= 0 && iw < IW) ? f4relu(f4bn(ldg4(xrow + (long long)iw * Cs), mean, scale, beta)) : f4zero();
        }
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
          const int to = ti - kt + 1;
          if (to < 0 || to >= T) continue;
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const float4 wv = *reinterpret_cast<const float4*>(ws + (kt * 9 + kh * 3 + kw) * Cs + c);
#pragma unroll
            for (int j = 0; j < OWT; ++j) acc[to][j] = f4fma(wv, in[j * S + kw], acc[to][j]);
          }
        }
      }
    }
    float4 s = f4zero(), sq = f4zero();
#pragma unroll
    for (int t = 0; t < T; ++t) {
      float* yrow = Y + (((long long)(n * T + t) * OH + oh) * OW) * Cs + c;
#pragma unroll
      for (int j = 0; j < OWT; ++j) {
        if (ow0 + j < OW) {
          st4(yrow + (long long)(ow0 + j) * Cs, acc[t][j]);
          s = f4add(s, acc[t][j]);
          sq = f4fma(acc[t][j], acc[t][j], sq);
        }
      }
    }
    if (stats) {
      atomicAdd(&s_sum[c], s.x); atomicAdd(&s_sum[c + 1], s.y); atomicAdd(&s_sum[c + 2], s.z); atomicAdd(&s_sum[c + 3], s.w);
      atomicAdd(&s_sq[c], sq.x); atomicAdd(&s_sq[c + 1], sq.y); atomicAdd(&s_sq[c + 2], sq.z); atomicAdd(&s_sq[c + 3], sq.w);
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

template <int T, int OWT>
static int launch_dw_fwd(const float* X, const float* bnp, const float* w, float* Y, double* stats, int N, int IH,
                         int IW, int C, int Cs, int stride, cudaStream_t st) {
  const int OH = (IH - 1) / stride + 1, OW = (IW - 1) / stride + 1;   // k=3, pad=1
  const size_t smem = (size_t)(29 * Cs) * sizeof(float);
  dim3 grid((unsigned)(N * OH));
  if (stride == 1) {
    cudaFuncSetAttribute(dw_fwd_kernel<T, OWT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dw_fwd_kernel<T, OWT, 1><<<grid, 256, smem, st>>>(X, bnp, w, Y, stats, IH, IW, OH, OW, C, Cs);
  } else {
    cudaFuncSetAttribute(dw_fwd_kernel<T, OWT, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dw_fwd_kernel<T, OWT, 2><<<grid, 256, smem, st>>>(X, bnp, w, Y, stats, IH, IW, OH, OW, C, Cs);
  }
  return c3d_check_last(cudaGetLastError());
}

extern "C" int c3d_dw_conv_fwd(const float* X, const float* bnp_a, const float* w, float* Y, double* stats, int N,
                               int T, int IH, int IW, int C, int Cs, int stride, void* stream_) {
  if (!X || !bnp_a || !w || !Y || N <= 0 || IH <= 0 || IW <= 0 || C <= 0 || Cs < C || (Cs & 3)) return C3D_ERR_ARG;
  if (stride != 1 && stride != 2) return C3D_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream_;
  switch (T) {
    case 3: return launch_dw_fwd<3, 4>(X, bnp_a, w, Y, stats, N, IH, IW, C, Cs, stride, st);
    case 4: return launch_dw_fwd<4, 2>(X, bnp_a, w, Y, stats, N, IH, IW, C, Cs, stride, st);
    case 5: return launch_dw_fwd<5, 2>(X, bnp_a, w, Y, stats, N, IH, IW, C, Cs, stride, st);
    default: return C3D_ERR_ARG;
  }
}

// ------------------------------------------------------------------------------------------------
// Backward of conv_b fused with the BN_b / SE backward on its output side and the ReLU + BN_a
// statistics on its input side:
//   dy_b = scale_b * (du*gate + dpool - c1 - zhat*c2)            (computed on the fly from du, y_b)
//   da   = conv_transpose(dy_b, w);  dW[tap] += a * dy_b          (a = relu(bn_a(y_a)) recomputed)
//   dr   = da * (a > 0)  -> written;  stats_a += (sum dr, sum dr * yhat_a)
// A thread owns one channel quad for the whole launch (persistent over (n, ih) rows), so the 27
// weight-gradient accumulators live in registers and are flushed once.
// ------------------------------------------------------------------------------------------------
template <int T, int IWT, int S>
__global__ void __launch_bounds__(128) dw_bwd_kernel(
    const float* __restrict__ du, const float* __restrict__ yb, const float* __restrict__ bnp_b,
    const float* __restrict__ gate, const float* __restrict__ dpool, const float* __restrict__ coef_b,
    const float* __restrict__ ya, const float* __restrict__ bnp_a, const float* __restrict__ w,
    float* __restrict__ dr, float* __restrict__ dW, double* __restrict__ stats_a, int N, int IH, int IW, int OH,
    int OW, int C, int Cs) {
  constexpr int NSEG = (S == 1) ? IWT + 2 : 2;
  static_assert(S == 1 || IWT == 2, "stride-2 path assumes two input columns per thread");
  extern __shared__ __align__(16) float sm[];
  float* ws = sm;                  // [27][Cs]
  float* s_dw = ws + 27 * Cs;      // [27][Cs]
  float* s_st = s_dw + 27 * Cs;    // [2][Cs]
  const int nq = Cs >> 2;
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
  const int q = tid % nq, slot = tid / nq, c = 4 * q;

  float4 dwacc[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) dwacc[i] = f4zero();
  float4 st_s = f4zero(), st_t = f4zero();

  if (active) {
    const float4 mean_b = ldg4(bnp_b + c), rstd_b = ldg4(bnp_b + Cs + c), scale_b = ldg4(bnp_b + 2 * Cs + c);
    const float4 c1 = ldg4(coef_b + c), c2 = ldg4(coef_b + Cs + c);
    const float4 mean_a = ldg4(bnp_a + c), rstd_a = ldg4(bnp_a + Cs + c), scale_a = ldg4(bnp_a + 2 * Cs + c),
                 beta_a = ldg4(bnp_a + 3 * Cs + c);
    // dy = A1*du + A0 - A2*(yb - mean_b)
    const float4 A2 = f4mul(scale_b, f4mul(c2, rstd_b));
    const int niwg = (IW + IWT - 1) / IWT;
    for (int row = blockIdx.x; row < N * IH; row += gridDim.x) {
      const int n = row / IH, ih = row - n * IH;
      float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f), dp4 = f4zero();
      if (gate) { g4 = ldg4(gate + (long long)n * Cs + c); dp4 = ldg4(dpool + (long long)n * Cs + c); }
      const float4 A1 = f4mul(scale_b, g4);
      const float4 A0 = make_float4(scale_b.x * (dp4.x - c1.x), scale_b.y * (dp4.y - c1.y), scale_b.z * (dp4.z - c1.z),
                                    scale_b.w * (dp4.w - c1.w));
      for (int iwg = slot; iwg < niwg; iwg += nslots) {
        const int iw0 = iwg * IWT;
        float4 a[T][IWT], da[T][IWT];
#pragma unroll
        for (int t = 0; t < T; ++t)
#pragma unroll
          for (int j = 0; j < IWT; ++j) {
            da[t][j] = f4zero();
            a[t][j] = (iw0 + j < IW)
                          ? f4relu(f4bn(ldg4(ya + ((((long long)n * T + t) * IH + ih) * IW + iw0 + j) * Cs + c), mean_a,
                                        scale_a, beta_a))
                          : f4zero();
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
            float4 dy[NSEG];
            const long long obase = ((((long long)n * T + to) * OH + oh) * OW) * Cs + c;
#pragma unroll
            for (int x = 0; x < NSEG; ++x) {
              const int ow = owlo + x;
              if (ow >= 0 && ow < OW) {
                const float4 d = ldg4(du + obase + (long long)ow * Cs);
                const float4 y = ldg4(yb + obase + (long long)ow * Cs);
                dy[x].x = fmaf(A1.x, d.x, A0.x) - A2.x * (y.x - mean_b.x);
                dy[x].y = fmaf(A1.y, d.y, A0.y) - A2.y * (y.y - mean_b.y);
                dy[x].z = fmaf(A1.z, d.z, A0.z) - A2.z * (y.z - mean_b.z);
                dy[x].w = fmaf(A1.w, d.w, A0.w) - A2.w * (y.w - mean_b.w);
              } else {
                dy[x] = f4zero();
              }
            }
#pragma unroll
            for (int kt = 0; kt < 3; ++kt) {
              const int ti = to + kt - 1;
              if (ti < 0 || ti >= T) continue;
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) {
                const int tap = kt * 9 + kh * 3 + kw;
                const float4 wv = *reinterpret_cast<const float4*>(ws + tap * Cs + c);
#pragma unroll
                for (int j = 0; j < IWT; ++j) {
                  const int num_w = j + 1 - kw;           // relative to iw0 (iw0 is even when S == 2)
                  if (S == 2 && (num_w & 1)) continue;
                  const int x = (S == 1) ? num_w + 1 : num_w / 2;
                  if (x < 0 || x >= NSEG) continue;
                  da[ti][j] = f4fma(wv, dy[x], da[ti][j]);
                  dwacc[tap] = f4fma(a[ti][j], dy[x], dwacc[tap]);
                }
              }
            }
          }
        }
        // epilogue: ReLU mask, store, BN_a backward statistics
#pragma unroll
        for (int t = 0; t < T; ++t)
#pragma unroll
          for (int j = 0; j < IWT; ++j) {
            if (iw0 + j >= IW) continue;
            const long long off = ((((long long)n * T + t) * IH + ih) * IW + iw0 + j) * Cs + c;
            float4 v = da[t][j];
            v.x = a[t][j].x > 0.f ? v.x : 0.f; v.y = a[t][j].y > 0.f ? v.y : 0.f;
            v.z = a[t][j].z > 0.f ? v.z : 0.f; v.w = a[t][j].w > 0.f ? v.w : 0.f;
            st4(dr + off, v);
            const float4 y = ldg4(ya + off);
            st_s = f4add(st_s, v);
            st_t.x = fmaf(v.x, (y.x - mean_a.x) * rstd_a.x, st_t.x); st_t.y = fmaf(v.y, (y.y - mean_a.y) * rstd_a.y, st_t.y);
            st_t.z = fmaf(v.z, (y.z - mean_a.z) * rstd_a.z, st_t.z); st_t.w = fmaf(v.w, (y.w - mean_a.w) * rstd_a.w, st_t.w);
          }
      }
    }
    // flush register accumulators
#pragma unroll
    for (int tap = 0; tap < 27; ++tap) {
      atomicAdd(&s_dw[tap * Cs + c], dwacc[tap].x); atomicAdd(&s_dw[tap * Cs + c + 1], dwacc[tap].y);
      atomicAdd(&s_dw[tap * Cs + c + 2], dwacc[tap].z); atomicAdd(&s_dw[tap * Cs + c + 3], dwacc[tap].w);
    }
    atomicAdd(&s_st[c], st_s.x); atomicAdd(&s_st[c + 1], st_s.y); atomicAdd(&s_st[c + 2], st_s.z); atomicAdd(&s_st[c + 3], st_s.w);
    atomicAdd(&s_st[Cs + c], st_t.x); atomicAdd(&s_st[Cs + c + 1], st_t.y);
    atomicAdd(&s_st[Cs + c + 2], st_t.z); atomicAdd(&s_st[Cs + c + 3], st_t.w);
  }
  __syncthreads();
  for (int i = tid; i < 27 * Cs; i += blockDim.x) {
    int tap = i / Cs, cc = i - tap * Cs;
    if (cc < C) atomicAdd(dW + cc * 27 + tap, s_dw[i]);
  }
  for (int i = tid; i < 2 * Cs; i += blockDim.x) atomicAdd(stats_a + i, (double)s_st[i]);
}

template <int T, int IWT1>
static int launch_dw_bwd(const float* du, const float* yb, const float* bnp_b, const float* gate, const float* dpool,
                         const float* coef_b, const float* ya, const float* bnp_a, const float* w, float* dr, float* dW,
                         double* stats_a, int N, int IH, int IW, int C, int Cs, int stride, cudaStream_t st) {
  const int OH = (IH - 1) / stride + 1, OW = (IW - 1) / stride + 1;
  const int nq = Cs >> 2;
  if (nq > 128) return C3D_ERR_ARG;
  const int threads = ((128 / nq) * nq + 31) / 32 * 32;
  const size_t smem = (size_t)(56 * Cs) * sizeof(float);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = N * IH < 2 * sms ? N * IH : 2 * sms;
  if (stride == 1) {
    cudaFuncSetAttribute(dw_bwd_kernel<T, IWT1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dw_bwd_kernel<T, IWT1, 1><<<grid, threads, smem, st>>>(du, yb, bnp_b, gate, dpool, coef_b, ya, bnp_a, w, dr, dW,
                                                           stats_a, N, IH, IW, OH, OW, C, Cs);
  } else {
    cudaFuncSetAttribute(dw_bwd_kernel<T, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dw_bwd_kernel<T, 2, 2><<<grid, threads, smem, st>>>(du, yb, bnp_b, gate, dpool, coef_b, ya, bnp_a, w, dr, dW,
                                                        stats_a, N, IH, IW, OH, OW, C, Cs);
  }
  return c3d_check_last(cudaGetLastError());
}

extern "C" int c3d_dw_conv_bwd(const float* du, const float* y_b, const float* bnp_b, const float* gate,
                               const float* dpool, const float* coef_b, const float* y_a, const float* bnp_a,
                               const float* w, float* dr, float* dW, double* stats_a, int N, int T, int IH, int IW,
                               int C, int Cs, int stride, void* stream_) {
  if (!du || !y_b || !bnp_b || !coef_b || !y_a || !bnp_a || !w || !dr || !dW || !stats_a) return C3D_ERR_ARG;
  if ((gate == nullptr) != (dpool == nullptr)) return C3D_ERR_ARG;
  if (N <= 0 || IH <= 0 || IW <= 0 || C <= 0 || Cs < C || (Cs & 3) || (stride != 1 && stride != 2)) return C3D_ERR_ARG;
  if (stride == 2 && ((IH | IW) & 1)) return C3D_ERR_ARG;   // stride-2 path pairs even/odd columns
  cudaStream_t st = (cudaStream_t)stream_;
  switch (T) {
    case 3: return launch_dw_bwd<3, 1>(du, y_b, bnp_b, gate, dpool, coef_b, y_a, bnp_a, w, dr, dW, stats_a, N, IH, IW, C, Cs, stride, st);
    case 4: return launch_dw_bwd<4, 1>(du, y_b, bnp_b, gate, dpool, coef_b, y_a, bnp_a, w, dr, dW, stats_a, N, IH, IW, C, Cs, stride, st);
    case 5: return launch_dw_bwd<5, 1>(du, y_b, bnp_b, gate, dpool, coef_b, y_a, bnp_a, w, dr, dW, stats_a, N, IH, IW, C, Cs, stride, st);
    default: return C3D_ERR_ARG;
  }
}
