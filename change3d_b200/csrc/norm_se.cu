// BatchNorm3d finalisation (statistics -> parameter block, running-stat update), the SqueezeExcitation
// gate, the fused residual join, and their backward counterparts.
// Reference: nn.BatchNorm3d sites model/x3d.py:94-98,176-180,203-208,217-221,296-298; fvcore
// SqueezeExcitation call site model/x3d.py:194-202; ResBlock fusion model/x3d.py:326-327.
#include <stdio.h>
#include <stdlib.h>
#include "c3d_common.cuh"
#include "../../include/change3d_b200.h"

__device__ __forceinline__ void bn_params_for_channel(const double* stats, int groups, long long count,
                                                      const float* gamma, const float* beta, float* rm, float* rv,
                                                      int c, int Cs, float momentum, float eps, int training,
                                                      bool write_running, float& mean_f, float& rstd_f) {
  double mean, var;
  if (training) {
    // four independent partial sums per moment: the loads of the batch loop are issued together instead of one
    // L2 round trip per sample (the loop is the latency of this single-wave kernel)
    double s = 0.0, q = 0.0, s1 = 0.0, q1 = 0.0, s2 = 0.0, q2 = 0.0, s3 = 0.0, q3 = 0.0;
    int g = 0;
    for (; g + 8 <= groups; g += 8) {      // 16 loads in flight
      const double* p = stats + ((long long)g * 2) * Cs + c;
      double a[8], b[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { a[u] = p[(2 * u) * Cs]; b[u] = p[(2 * u + 1) * Cs]; }
      s += a[0] + a[4]; q += b[0] + b[4]; s1 += a[1] + a[5]; q1 += b[1] + b[5];
      s2 += a[2] + a[6]; q2 += b[2] + b[6]; s3 += a[3] + a[7]; q3 += b[3] + b[7];
    }
    for (; g + 4 <= groups; g += 4) {
      const double* p = stats + ((long long)g * 2) * Cs + c;
      const double a0 = p[0], b0 = p[Cs], a1 = p[2 * Cs], b1 = p[3 * Cs], a2 = p[4 * Cs], b2 = p[5 * Cs], a3 = p[6 * Cs], b3 = p[7 * Cs];
      s += a0; q += b0; s1 += a1; q1 += b1; s2 += a2; q2 += b2; s3 += a3; q3 += b3;
    }
    for (; g < groups; ++g) {
      s += stats[((long long)g * 2 + 0) * Cs + c];
      q += stats[((long long)g * 2 + 1) * Cs + c];
    }
    s += s1 + s2 + s3;
    q += q1 + q2 + q3;
    mean = s / (double)count;
    var = q / (double)count - mean * mean;
    if (var < 0.0) var = 0.0;
    if (write_running && rm && rv) {
      double unbiased = count > 1 ? var * (double)count / (double)(count - 1) : var;
      rm[c] = (float)((1.0 - (double)momentum) * (double)rm[c] + (double)momentum * mean);
      rv[c] = (float)((1.0 - (double)momentum) * (double)rv[c] + (double)momentum * unbiased);
    }
  } else {
    mean = (double)rm[c];
    var = (double)rv[c];
  }
  mean_f = (float)mean;
  rstd_f = (float)(1.0 / sqrt(var + (double)eps));
}

__global__ void bn_finalize_kernel(const double* stats, int groups, long long count, const float* gamma,
                                   const float* beta, float* rm, float* rv, int C, int Cs, float momentum, float eps,
                                   int training, float* bnp) {
  pdl_trigger();   // programmatic dependent launch: let the next kernel start its setup,
  pdl_wait();      // then wait for the earlier kernels whose results this one reads

  for (int c = threadIdx.x; c < Cs; c += blockDim.x) {
    if (c >= C) {
      bnp[c] = 0.f; bnp[Cs + c] = 0.f; bnp[2 * Cs + c] = 0.f; bnp[3 * Cs + c] = 0.f;
      continue;
    }
    float mean, rstd;
    bn_params_for_channel(stats, groups, count, gamma, beta, rm, rv, c, Cs, momentum, eps, training, true, mean, rstd);
    bnp[c] = mean; bnp[Cs + c] = rstd; bnp[2 * Cs + c] = gamma[c] * rstd; bnp[3 * Cs + c] = beta[c];
  }
}

extern "C" int c3d_bn_finalize(const double* stats, int groups, long long count, const float* gamma, const float* beta,
                               float* running_mean, float* running_var, int C, int Cs, float momentum, float eps,
                               int training, float* bnp, void* stream_) {
  if (!gamma || !beta || !bnp || C <= 0 || Cs < C) return C3D_ERR_ARG;
  if (training ? (!stats || groups <= 0 || count <= 0) : (!running_mean || !running_var)) return C3D_ERR_ARG;
  c3d_launch_pdl_small(bn_finalize_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream_, stats, groups, count, gamma, beta, running_mean,
                                                            running_var, C, Cs, momentum, eps, training, bnp);
  return c3d_check_last(cudaGetLastError());
}

// one CTA per batch sample; every CTA derives the BN parameters (cheap), CTA 0 publishes them
__global__ void __launch_bounds__(256) bn_se_finalize_kernel(
    const double* stats, int N, long long cnt, const float* gamma, const float* beta, float* rm, float* rv, int C,
    int Cs, float momentum, float eps, int training, const float* w1, const float* b1, const float* w2,
    const float* b2, int R, float* bnp, float* zhat_mean, float* hidden, float* gate) {
  pdl_trigger();   // programmatic dependent launch: let the next kernel start its setup,
  pdl_wait();      // then wait for the earlier kernels whose results this one reads

  extern __shared__ float sm[];
  float* pooled = sm;          // [Cs]
  float* hid = sm + Cs;        // [R]
  const int n = blockIdx.x;
  const long long count = cnt * N;
  for (int c = threadIdx.x; c < Cs; c += blockDim.x) {
    float zm = 0.f, pl = 0.f;
    if (c < C) {
      float mean, rstd;
      bn_params_for_channel(stats, N, count, gamma, beta, rm, rv, c, Cs, momentum, eps, training, n == 0, mean, rstd);
      const float scale = gamma[c] * rstd, bt = beta[c];
      if (n == 0) { bnp[c] = mean; bnp[Cs + c] = rstd; bnp[2 * Cs + c] = scale; bnp[3 * Cs + c] = bt; }
      const float m_n = (float)(stats[((long long)n * 2) * Cs + c] / (double)cnt);
      zm = (m_n - mean) * rstd;
      pl = fmaf(m_n - mean, scale, bt);
    } else if (n == 0) {
      bnp[c] = 0.f; bnp[Cs + c] = 0.f; bnp[2 * Cs + c] = 0.f; bnp[3 * Cs + c] = 0.f;
    }
    pooled[c] = pl;
    if (zhat_mean) zhat_mean[(long long)n * Cs + c] = zm;
  }
  if (!w1) return;   // BN only
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // fc1: one warp per hidden unit; the weight loads of a row are issued in batches of 8 (the loop is otherwise one L2
  // round trip per 32 channels), and the rows a warp handles are independent
  for (int r = warp; r < R; r += 8) {
    float s = 0.f;
    for (int c0 = 0; c0 < C; c0 += 256) {
      float wv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { const int c = c0 + u * 32 + lane; wv[u] = c < C ? __ldg(w1 + r * C + c) : 0.f; }
#pragma unroll
      for (int u = 0; u < 8; ++u) { const int c = c0 + u * 32 + lane; if (c < C) s = fmaf(wv[u], pooled[c], s); }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      s = fmaxf(s + b1[r], 0.f);
      hid[r] = s;
      hidden[(long long)n * R + r] = s;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < Cs; c += blockDim.x) {
    float g = 0.f;
    if (c < C) {
      float s = b2[c];
      if ((R & 3) == 0 && R <= 32 && (reinterpret_cast<uintptr_t>(w2) & 15) == 0) {      // a channel's fc2 row is R contiguous floats: vector loads, all in flight
        float4 wv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) wv[u] = 4 * u < R ? ldg4(w2 + c * R + 4 * u) : f4zero();
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (4 * u < R) {
            s = fmaf(wv[u].x, hid[4 * u], s); s = fmaf(wv[u].y, hid[4 * u + 1], s);
            s = fmaf(wv[u].z, hid[4 * u + 2], s); s = fmaf(wv[u].w, hid[4 * u + 3], s);
          }
        }
      } else {
        for (int r = 0; r < R; ++r) s = fmaf(w2[c * R + r], hid[r], s);
      }
      g = sigmoidf_(s);
    }
    gate[(long long)n * Cs + c] = g;
  }
}

extern "C" int c3d_bn_se_finalize(const double* stats, int N, long long count_per_sample, const float* gamma,
                                  const float* beta, float* running_mean, float* running_var, int C, int Cs,
                                  float momentum, float eps, int training, const float* w1, const float* b1,
                                  const float* w2, const float* b2, int R, float* bnp, float* zhat_mean, float* hidden,
                                  float* gate, void* stream_) {
  if (!stats || !gamma || !beta || !bnp || N <= 0 || count_per_sample <= 0 || C <= 0 || Cs < C) return C3D_ERR_ARG;
  if (!training && (!running_mean || !running_var)) return C3D_ERR_ARG;
  if (w1 && (!b1 || !w2 || !b2 || R <= 0 || !hidden || !gate)) return C3D_ERR_ARG;
  size_t smem = (size_t)(Cs + (R > 0 ? R : 0)) * sizeof(float);
  c3d_launch_pdl_small(bn_se_finalize_kernel, dim3(N), dim3(256), smem, (cudaStream_t)stream_, stats, N, count_per_sample, gamma, beta, running_mean,
                                                                 running_var, C, Cs, momentum, eps, training, w1, b1,
                                                                 w2, b2, R, bnp, zhat_mean, hidden, gate);
  return c3d_check_last(cudaGetLastError());
}

template <int U>
__global__ void __launch_bounds__(256) bn_add_relu_kernel(const float* __restrict__ A, const float* __restrict__ bnpA,
                                                          const float* __restrict__ B, const float* __restrict__ bnpB,
                                                          float* __restrict__ Y, long long total4, int Cs) {
  pdl_trigger();   // programmatic dependent launch: let the next kernel start its setup,
  pdl_wait();      // then wait for the earlier kernels whose results this one reads

  const int q4 = Cs >> 2;
  // channel quad of element i tracked incrementally (a 64-bit modulo per element costs more than the arithmetic)
  const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const unsigned stride = gridDim.x * blockDim.x;
  const int cstep = (int)(stride % (unsigned)q4) * 4;
  int c = (int)(i0 % q4) * 4;
  auto one = [&](long long i, int cc, float4 a, float4 b) {
    float4 v = f4bn(a, ldg4(bnpA + cc), ldg4(bnpA + 2 * Cs + cc), ldg4(bnpA + 3 * Cs + cc));
    if (B) {
      if (bnpB) b = f4bn(b, ldg4(bnpB + cc), ldg4(bnpB + 2 * Cs + cc), ldg4(bnpB + 3 * Cs + cc));
      v = f4add(v, b);
    }
    st4(Y + i * 4, f4relu(v));
  };
  long long i = i0;
  // U elements per iteration, their streaming loads issued before the first is consumed
  for (; i + (long long)(U - 1) * stride < total4; i += (long long)U * stride) {
    float4 a[U], b[U];
    int cc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long iu = i + (long long)u * stride;
      a[u] = ldg4(A + iu * 4);
      b[u] = B ? ldg4(B + iu * 4) : f4zero();
      cc[u] = c;
      c += cstep;
      if (c >= Cs) c -= Cs;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) one(i + (long long)u * stride, cc[u], a[u], b[u]);
  }
  for (; i < total4; i += stride) {
    one(i, c, ldg4(A + i * 4), B ? ldg4(B + i * 4) : f4zero());
    c += cstep;
    if (c >= Cs) c -= Cs;
  }
}

// CTAs of `kernel` that fit the device at once: the streaming kernels below loop with a grid stride, so a grid of more
// than one resident wave only adds a partially filled last wave
template <typename K>
static long long resident_ctas(K kernel, int threads, size_t smem) {
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  return (long long)sms * per_sm;
}

static int ew_unroll() {             // C3D_EW_UNROLL: elements / rows per loop iteration of the streaming elementwise kernels
  static int u = -1;
  if (u < 0) { const char* v = getenv("C3D_EW_UNROLL"); u = v ? atoi(v) : 4; }
  return u;
}

extern "C" int c3d_bn_add_relu(const float* A, const float* bnpA, const float* B, const float* bnpB, float* Y,
                               long long M, int Cs, void* stream_) {
  if (!A || !bnpA || !Y || M <= 0 || Cs <= 0 || (Cs & 3)) return C3D_ERR_ARG;
  const long long total4 = M * (Cs >> 2);
  long long blocks = (total4 + 255) / 256;
  const int u = ew_unroll();
#define BAR_LAUNCH(U_) do { const long long cap = resident_ctas(bn_add_relu_kernel<U_>, 256, 0); if (blocks > cap) blocks = cap; \
    c3d_launch_pdl(bn_add_relu_kernel<U_>, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream_, A, bnpA, B, bnpB, Y, total4, Cs); } while (0)
  if (u >= 4) BAR_LAUNCH(4);
  else if (u >= 2) BAR_LAUNCH(2);
  else BAR_LAUNCH(1);
#undef BAR_LAUNCH
  return c3d_check_last(cudaGetLastError());
}

extern "C" int c3d_version(void) { return C3D_ABI_VERSION; }

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------

// d_pre = dOut * (out > 0)   [ResBlock / stem ReLU backward]
// stats_c += (sum d_pre, sum d_pre * yhat_c)  and, for a normalised shortcut, stats_1 likewise.
// Threads keep a fixed channel quad so the column sums stay in registers.
// HAS1: normalised shortcut (y1 / bnp_1 / stats_1); MASKY: out == NULL, the mask is recomputed from y_c.  Both are template
// parameters so that the common case (neither) keeps 12 registers per row in flight and three CTAs per SM at U = 4.
template <int U, bool HAS1, bool MASKY>
__global__ void __launch_bounds__(256, (U >= 4 && (HAS1 || MASKY)) ? 2 : U >= 2 ? 3 : 4) relu_bwd_stats_kernel(
    const float* __restrict__ dOut, const float* __restrict__ out, const float* __restrict__ yc,
    const float* __restrict__ bnp_c, const float* __restrict__ y1, const float* __restrict__ bnp_1,
    float* __restrict__ d_pre, double* __restrict__ stats_c, double* __restrict__ stats_1, long long M, int Cs) {
  pdl_trigger();   // programmatic dependent launch: let the next kernel start its setup,
  pdl_wait();      // then wait for the earlier kernels whose results this one reads

  extern __shared__ float sm[];   // [3][Cs]
  for (int i = threadIdx.x; i < 3 * Cs; i += 256) sm[i] = 0.f;
  __syncthreads();
  const int q4 = Cs >> 2;
  const int rpb = 256 / q4;
  const int tid = threadIdx.x;
  if (tid < rpb * q4) {
    const int q = tid % q4, rl = tid / q4, c = 4 * q;
    const float4 mc = ldg4(bnp_c + c), rc = ldg4(bnp_c + Cs + c);
    float4 sc = f4zero(), bc = f4zero();
    if (MASKY) { sc = ldg4(bnp_c + 2 * Cs + c); bc = ldg4(bnp_c + 3 * Cs + c); }
    float4 m1 = f4zero(), r1 = f4zero();
    if (HAS1) { m1 = ldg4(bnp_1 + c); r1 = ldg4(bnp_1 + Cs + c); }
    float4 s = f4zero(), tc = f4zero(), t1 = f4zero();
    auto consume = [&](long long off, float4 d, const float4 y, const float4 o, const float4 z) {
      d.x = o.x > 0.f ? d.x : 0.f; d.y = o.y > 0.f ? d.y : 0.f; d.z = o.z > 0.f ? d.z : 0.f; d.w = o.w > 0.f ? d.w : 0.f;
      if (d_pre) st4(d_pre + off, d);
      s = f4add(s, d);
      tc.x = fmaf(d.x, (y.x - mc.x) * rc.x, tc.x); tc.y = fmaf(d.y, (y.y - mc.y) * rc.y, tc.y);
      tc.z = fmaf(d.z, (y.z - mc.z) * rc.z, tc.z); tc.w = fmaf(d.w, (y.w - mc.w) * rc.w, tc.w);
      if (HAS1) {
        t1.x = fmaf(d.x, (z.x - m1.x) * r1.x, t1.x); t1.y = fmaf(d.y, (z.y - m1.y) * r1.y, t1.y);
        t1.z = fmaf(d.z, (z.z - m1.z) * r1.z, t1.z); t1.w = fmaf(d.w, (z.w - m1.w) * r1.w, t1.w);
      }
    };
    // U rows per iteration: all of their loads are issued before the first is consumed (3-4 loads per row and warp in
    // flight do not cover the DRAM latency at 64 warps per SM: 4.4 TB/s with U = 1)
    const long long rstep = (long long)gridDim.x * rpb;
    long long row = (long long)blockIdx.x * rpb + rl;
    for (; row + (U - 1) * rstep < M; row += U * rstep) {
      float4 d[U], y[U], o[MASKY ? 1 : U], z[HAS1 ? U : 1];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long off = (row + u * rstep) * Cs + c;
        d[u] = ldg4(dOut + off);
        y[u] = ldg4(yc + off);
        if (!MASKY) o[u] = ldg4(out + off);
        if (HAS1) z[u] = ldg4(y1 + off);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        // MASKY: out == relu(bn_c(y_c)) (no shortcut), the mask comes from the forward's expression on y_c
        consume((row + u * rstep) * Cs + c, d[u], y[u], MASKY ? f4bn(y[u], mc, sc, bc) : o[u], HAS1 ? z[u] : f4zero());
      }
    }
    for (; row < M; row += rstep) {
      const long long off = row * Cs + c;
      const float4 y = ldg4(yc + off);
      consume(off, ldg4(dOut + off), y, MASKY ? f4bn(y, mc, sc, bc) : ldg4(out + off), HAS1 ? ldg4(y1 + off) : f4zero());
    }
    atomicAdd(&sm[c], s.x); atomicAdd(&sm[c + 1], s.y); atomicAdd(&sm[c + 2], s.z); atomicAdd(&sm[c + 3], s.w);
    atomicAdd(&sm[Cs + c], tc.x); atomicAdd(&sm[Cs + c + 1], tc.y); atomicAdd(&sm[Cs + c + 2], tc.z); atomicAdd(&sm[Cs + c + 3], tc.w);
    if (HAS1) {
      atomicAdd(&sm[2 * Cs + c], t1.x); atomicAdd(&sm[2 * Cs + c + 1], t1.y);
      atomicAdd(&sm[2 * Cs + c + 2], t1.z); atomicAdd(&sm[2 * Cs + c + 3], t1.w);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Cs; i += 256) {
    atomicAdd(stats_c + i, (double)sm[i]);
    atomicAdd(stats_c + Cs + i, (double)sm[Cs + i]);
    if (HAS1) {
      atomicAdd(stats_1 + i, (double)sm[i]);
      atomicAdd(stats_1 + Cs + i, (double)sm[2 * Cs + i]);
    }
  }
}

extern "C" int c3d_relu_bwd_stats(const float* dOut, const float* out, const float* y_c, const float* bnp_c,
                                  const float* y_1, const float* bnp_1, float* d_pre, double* stats_c, double* stats_1,
                                  long long M, int Cs, void* stream_) {
  if (!dOut || !y_c || !bnp_c || !stats_c || M <= 0 || Cs <= 0 || (Cs & 3) || Cs > 1024) return C3D_ERR_ARG;
  if (y_1 && (!bnp_1 || !stats_1)) return C3D_ERR_ARG;
  if (!out && y_1) return C3D_ERR_ARG;      // the recomputed mask is only valid without a shortcut
  const int rpb = 256 / (Cs >> 2);
  long long blocks = (M + rpb - 1) / rpb;
  if (blocks > 148 * 8) blocks = 148 * 8;
  const int unroll = ew_unroll();
  const dim3 blk(256);
  const size_t smem = 3 * Cs * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream_;
#define RBS_LAUNCH(U_, H_, K_) do { const long long cap = resident_ctas(relu_bwd_stats_kernel<U_, H_, K_>, 256, smem);            \
    if (blocks > cap) blocks = cap;                                                                                            \
    c3d_launch_pdl(relu_bwd_stats_kernel<U_, H_, K_>, dim3((unsigned)blocks), blk, smem, st, dOut, out, y_c, bnp_c, y_1,         \
                   bnp_1, d_pre, stats_c, stats_1, M, Cs); } while (0)
#define RBS_PICK(U_) do { if (y_1) RBS_LAUNCH(U_, true, false); else if (!out) RBS_LAUNCH(U_, false, true); \
                          else RBS_LAUNCH(U_, false, false); } while (0)
  if (unroll >= 4) RBS_PICK(4);
  else if (unroll >= 2) RBS_PICK(2);
  else RBS_PICK(1);
#undef RBS_PICK
#undef RBS_LAUNCH
  return c3d_check_last(cudaGetLastError());
}

// stats = double[groups][2][Cs] (sum d, sum d*yhat) -> coef[2][Cs] = (mean d, mean d*yhat); dgamma, dbeta
__global__ void bn_bwd_finalize_kernel(const double* stats, int groups, long long count, int C, int Cs, float* coef,
                                       float* dgamma, float* dbeta) {
  pdl_trigger();   // programmatic dependent launch: let the next kernel start its setup,
  pdl_wait();      // then wait for the earlier kernels whose results this one reads

  for (int c = threadIdx.x; c < Cs; c += blockDim.x) {
    double s = 0.0, t = 0.0;
    if (c < C)
      for (int g = 0; g < groups; ++g) { s += stats[((long long)g * 2) * Cs + c]; t += stats[((long long)g * 2 + 1) * Cs + c]; }
    coef[c] = (float)(s / (double)count);
    coef[Cs + c] = (float)(t / (double)count);
    if (c < C) { dgamma[c] = (float)t; dbeta[c] = (float)s; }
  }
}

extern "C" int c3d_bn_bwd_finalize(const double* stats, int groups, long long count, int C, int Cs, float* coef,
                                   float* dgamma, float* dbeta, void* stream_) {
  if (!stats || !coef || !dgamma || !dbeta || groups <= 0 || count <= 0 || C <= 0 || Cs < C) return C3D_ERR_ARG;
  c3d_launch_pdl_small(bn_bwd_finalize_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream_, stats, groups, count, C, Cs, coef, dgamma, dbeta);
  return c3d_check_last(cudaGetLastError());
}

// SE backward + BN_b backward coefficients.  stats = double[N][2][Cs]: per-sample (sum du, sum du*zhat).
// One CTA (the work is ~0.5 MFLOP); every phase is laid out so that no thread walks the batch with dependent global
// loads: the per-sample inputs are staged in shared memory once, sums over the batch run on (channel, quarter of the
// batch) threads and are folded through shared memory.
__global__ void __launch_bounds__(1024) se_bn_bwd_finalize_kernel(
    const double* __restrict__ stats, int N, long long cnt, const float* __restrict__ bnp, const float* __restrict__ gamma,
    const float* __restrict__ beta, const float* __restrict__ gate, const float* __restrict__ hidden,
    const float* __restrict__ zhat_mean, const float* __restrict__ w1, const float* __restrict__ w2, int C, int Cs, int R,
    float* __restrict__ coef, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dpool,
    float* __restrict__ dw1, float* __restrict__ db1, float* __restrict__ dw2, float* __restrict__ db2,
    long long* __restrict__ dbg) {
  pdl_trigger();   // programmatic dependent launch: let the next kernel start its setup,
  pdl_wait();      // then wait for the earlier kernels whose results this one reads
#define FIN_T(k) do { if (dbg && threadIdx.x == 0) dbg[k] = clock64(); } while (0)
  FIN_T(0);

  extern __shared__ double smd[];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const bool se = (gate != nullptr);
  const double Mtot = (double)cnt * (double)N;
  // shared: part[4][2][C] doubles | dps[N][C] | dpr[N][R] | hid[N][R] | pin[N][C] (SE only)
  double* part = smd;                                   // [4][2][C]
  float* dps = reinterpret_cast<float*>(part + 8 * C);  // [N][C]   grad wrt pre-sigmoid
  float* dpr = dps + (size_t)N * C;                     // [N][R]   grad wrt pre-relu
  float* hid = dpr + (size_t)N * R;                     // [N][R]   relu output of fc1 (forward)
  float* pin = hid + (size_t)N * R;                     // [N][C]   fc1 input: gamma * zhat_mean + beta
  if (se) {
    for (int i = tid; i < N * C; i += nthr) {
      const int n = i / C, c = i - n * C;
      const double A = stats[((long long)n * 2) * Cs + c], Bz = stats[((long long)n * 2 + 1) * Cs + c];
      const float g = gate[(long long)n * Cs + c];
      const float dg = (float)((double)gamma[c] * Bz + (double)beta[c] * A);   // sum du * z
      dps[i] = dg * g * (1.f - g);
      pin[i] = fmaf(zhat_mean[(long long)n * Cs + c], gamma[c], beta[c]);
    }
    for (int i = tid; i < N * R; i += nthr) hid[i] = hidden[i];
    __syncthreads();
    FIN_T(1);
    for (int c = tid; c < C; c += nthr) {
      float s = 0.f;
      for (int n = 0; n < N; ++n) s += dps[n * C + c];
      db2[c] = s;
    }
    for (int i = tid; i < C * R; i += nthr) {
      const int c = i / R, r = i - c * R;
      float s = 0.f;
      for (int n = 0; n < N; ++n) s = fmaf(dps[n * C + c], hid[n * R + r], s);
      dw2[i] = s;
    }
    FIN_T(2);
    const int warp = tid >> 5, lane = tid & 31;
    for (int i = warp; i < N * R; i += (nthr >> 5)) {
      const int n = i / R, r = i - n * R;
      float s = 0.f;
      for (int c = lane; c < C; c += 32) s = fmaf(w2[c * R + r], dps[n * C + c], s);
#pragma unroll
      for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) dpr[i] = hid[i] > 0.f ? s : 0.f;
    }
    __syncthreads();
    FIN_T(3);
    for (int r = tid; r < R; r += nthr) {
      float s = 0.f;
      for (int n = 0; n < N; ++n) s += dpr[n * R + r];
      db1[r] = s;
    }
    for (int i = tid; i < R * C; i += nthr) {
      const int r = i / C, c = i - r * C;
      float s = 0.f;
      for (int n = 0; n < N; ++n) s = fmaf(dpr[n * R + r], pin[n * C + c], s);
      dw1[i] = s;
    }
    FIN_T(4);
  }
  // batch sums S1 = sum_n (g A + dp), S2 = sum_n (g Bz + dp zm): thread = (channel, quarter of the batch)
  for (int i = tid; i < 4 * C; i += nthr) {
    const int qn = i / C, c = i - qn * C;
    const int n0 = qn * ((N + 3) >> 2), n1 = n0 + ((N + 3) >> 2) < N ? n0 + ((N + 3) >> 2) : N;
    double S1 = 0.0, S2 = 0.0;
    for (int n = n0; n < n1; ++n) {
      const double A = stats[((long long)n * 2) * Cs + c], Bz = stats[((long long)n * 2 + 1) * Cs + c];
      if (se) {
        float dp = 0.f;
        for (int r = 0; r < R; ++r) dp = fmaf(w1[r * C + c], dpr[n * R + r], dp);
        const double g = (double)gate[(long long)n * Cs + c];
        S1 += g * A + (double)dp;
        S2 += g * Bz + (double)dp * (double)zhat_mean[(long long)n * Cs + c];
        dpool[(long long)n * Cs + c] = (float)((double)dp / (double)cnt);
      } else {
        S1 += A; S2 += Bz;
      }
    }
    part[(qn * 2 + 0) * C + c] = S1;
    part[(qn * 2 + 1) * C + c] = S2;
  }
  __syncthreads();
  FIN_T(5);
  for (int c = tid; c < Cs; c += nthr) {
    double S1 = 0.0, S2 = 0.0;
    if (c < C) {
      S1 = part[c] + part[2 * C + c] + part[4 * C + c] + part[6 * C + c];
      S2 = part[C + c] + part[3 * C + c] + part[5 * C + c] + part[7 * C + c];
      dgamma[c] = (float)S2;
      dbeta[c] = (float)S1;
    } else if (se) {
      for (int n = 0; n < N; ++n) dpool[(long long)n * Cs + c] = 0.f;
    }
    coef[c] = (float)(S1 / Mtot);
    coef[Cs + c] = (float)(S2 / Mtot);
  }
  FIN_T(6);
#undef FIN_T
}

// Latency-lean version of the SE case (the one above spends 60 us, almost all of it in loops whose every iteration is a
// dependent global load: 66 K cycles in the fc2-transpose phase alone, profiles/r02_summary.md).  Same arithmetic, but
//   * thread (channel c, batch group j) owns the samples n = j, j + G, ...: their per-sample sums (fp64), gate and zhat
//     mean are loaded ONCE, all loads in flight together, and stay in registers for the final batch sums;
//   * fc1 / fc2 weights and the hidden activations are staged in shared memory by coalesced loads in the same phase;
//   * every later phase reads shared memory only; the fc2-transpose runs on (sample, r, H-way split of the channels)
//     threads with a shuffle fold.
// IPT = samples per thread (compile-time bound of the register arrays).
template <int IPT>
__global__ void __launch_bounds__(1024) se_bn_bwd_fast_kernel(
    const double* __restrict__ stats, int N, long long cnt, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ gate, const float* __restrict__ hidden, const float* __restrict__ zhat_mean,
    const float* __restrict__ w1, const float* __restrict__ w2, int C, int Cs, int R, int G, int H,
    float* __restrict__ coef, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dpool,
    float* __restrict__ dw1, float* __restrict__ db1, float* __restrict__ dw2, float* __restrict__ db2,
    long long* __restrict__ dbg) {
  pdl_trigger();
  pdl_wait();
#define FIN_T(k) do { if (dbg && threadIdx.x == 0) dbg[k] = clock64(); } while (0)
  FIN_T(0);
  extern __shared__ double smd[];
  const int tid = threadIdx.x, nthr = blockDim.x;
  double* part = smd;                                       // [G][2][C]
  float* dps = reinterpret_cast<float*>(part + (size_t)G * 2 * C);   // [N][C] grad wrt the pre-sigmoid fc2 output
  float* pin = dps + (size_t)N * C;                         // [N][C] fc1 input: gamma * zhat_mean + beta
  float* hid = pin + (size_t)N * C;                         // [N][R] relu output of fc1 (forward)
  float* dpr = hid + (size_t)N * R;                         // [N][R] grad wrt the pre-relu fc1 output
  float* w1s = dpr + (size_t)N * R;                         // [R][C]
  float* w2s = w1s + (size_t)R * C;                         // [C][R]
  const int j = tid / C, c = tid - j * C;
  const bool active = j < G;
  double A[IPT], Bz[IPT];
  float gt[IPT], zm[IPT];
  float gam = 0.f, bet = 0.f;
  if (active) {
    gam = gamma[c]; bet = beta[c];
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
      const int n = j + k * G;
      A[k] = 0.0; Bz[k] = 0.0; gt[k] = 0.f; zm[k] = 0.f;
      if (n < N) {
        A[k] = stats[((long long)n * 2) * Cs + c];
        Bz[k] = stats[((long long)n * 2 + 1) * Cs + c];
        gt[k] = gate[(long long)n * Cs + c];
        zm[k] = zhat_mean[(long long)n * Cs + c];
      }
    }
  }
  for (int i = tid; i < N * R; i += nthr) hid[i] = hidden[i];
  for (int i = tid; i < R * C; i += nthr) { w1s[i] = w1[i]; w2s[i] = w2[i]; }
  if (active) {
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
      const int n = j + k * G;
      if (n < N) {
        const float dg = (float)((double)gam * Bz[k] + (double)bet * A[k]);   // sum du * z
        dps[n * C + c] = dg * gt[k] * (1.f - gt[k]);
        pin[n * C + c] = fmaf(zm[k], gam, bet);
      }
    }
  }
  __syncthreads();
  FIN_T(1);
  // ---- fc2 backward: bias / weight gradients and the gradient of the hidden activations ----
  for (int cc = tid; cc < C; cc += nthr) {
    float s0 = 0.f, s1 = 0.f;
    int n = 0;
    for (; n + 1 < N; n += 2) { s0 += dps[n * C + cc]; s1 += dps[(n + 1) * C + cc]; }
    if (n < N) s0 += dps[n * C + cc];
    db2[cc] = s0 + s1;
  }
  // register tiling: one thread per (channel, 4 consecutive hidden units) -- one dps load and one float4 of the
  // hidden activations feed four FMAs (the single SM runs out of shared-memory load slots long before FMA slots)
  if ((R & 3) == 0) {
    const int R4 = R >> 2;
    for (int i = tid; i < C * R4; i += nthr) {
      const int cc = i / R4, r4 = (i - cc * R4) * 4;
      float4 a = f4zero();
      for (int n = 0; n < N; ++n) {
        const float d = dps[n * C + cc];
        const float4 hv = *reinterpret_cast<const float4*>(hid + n * R + r4);
        a.x = fmaf(d, hv.x, a.x); a.y = fmaf(d, hv.y, a.y); a.z = fmaf(d, hv.z, a.z); a.w = fmaf(d, hv.w, a.w);
      }
      dw2[cc * R + r4] = a.x; dw2[cc * R + r4 + 1] = a.y; dw2[cc * R + r4 + 2] = a.z; dw2[cc * R + r4 + 3] = a.w;
    }
  } else {
    for (int i = tid; i < C * R; i += nthr) {
      const int cc = i / R, r = i - cc * R;
      float s0 = 0.f;
      for (int n = 0; n < N; ++n) s0 = fmaf(dps[n * C + cc], hid[n * R + r], s0);
      dw2[i] = s0;
    }
  }
  FIN_T(2);
  {
    const int h = tid & (H - 1);
    for (int base = 0; base < N * R; base += nthr / H) {      // block-uniform trip count: every lane reaches the shuffles
      const int o = base + tid / H;
      const bool ok = o < N * R;
      const int n = ok ? o / R : 0, r = ok ? o - n * R : 0;
      float s0 = 0.f, s1 = 0.f;
      int cc = h;
      for (; cc + H < C; cc += 2 * H) {
        s0 = fmaf(w2s[cc * R + r], dps[n * C + cc], s0);
        s1 = fmaf(w2s[(cc + H) * R + r], dps[n * C + cc + H], s1);
      }
      if (cc < C) s0 = fmaf(w2s[cc * R + r], dps[n * C + cc], s0);
      float sres = s0 + s1;
      for (int of = H >> 1; of; of >>= 1) sres += __shfl_xor_sync(0xffffffffu, sres, of);
      if (ok && h == 0) dpr[o] = hid[o] > 0.f ? sres : 0.f;
    }
  }
  __syncthreads();
  FIN_T(3);
  // ---- fc1 backward + batch sums of this thread's own samples ----
  for (int r = tid; r < R; r += nthr) {
    float sres = 0.f;
    for (int n = 0; n < N; ++n) sres += dpr[n * R + r];
    db1[r] = sres;
  }
  if ((R & 3) == 0) {
    const int R4 = R >> 2;
    for (int i = tid; i < R4 * C; i += nthr) {
      const int r4 = (i / C) * 4, cc = i - (i / C) * C;
      float4 a = f4zero();
      for (int n = 0; n < N; ++n) {
        const float pv = pin[n * C + cc];
        const float4 dv = *reinterpret_cast<const float4*>(dpr + n * R + r4);
        a.x = fmaf(dv.x, pv, a.x); a.y = fmaf(dv.y, pv, a.y); a.z = fmaf(dv.z, pv, a.z); a.w = fmaf(dv.w, pv, a.w);
      }
      dw1[r4 * C + cc] = a.x; dw1[(r4 + 1) * C + cc] = a.y; dw1[(r4 + 2) * C + cc] = a.z; dw1[(r4 + 3) * C + cc] = a.w;
    }
  } else {
    for (int i = tid; i < R * C; i += nthr) {
      const int r = i / C, cc = i - r * C;
      float s0 = 0.f;
      for (int n = 0; n < N; ++n) s0 = fmaf(dpr[n * R + r], pin[n * C + cc], s0);
      dw1[i] = s0;
    }
  }
  FIN_T(4);
  if (active) {
    // dp[k] = sum_r w1[r][c] * dpr[n_k][r]: hidden units outermost, so a weight is loaded once for all of the thread's
    // samples and the per-sample factors come as float4
    float dpv[IPT];
#pragma unroll
    for (int k = 0; k < IPT; ++k) dpv[k] = 0.f;
    if ((R & 3) == 0) {
      for (int r4 = 0; r4 < R; r4 += 4) {
        const float w0 = w1s[r4 * C + c], w1v = w1s[(r4 + 1) * C + c], w2v = w1s[(r4 + 2) * C + c], w3 = w1s[(r4 + 3) * C + c];
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
          const int n = j + k * G;
          if (n < N) {
            const float4 dv = *reinterpret_cast<const float4*>(dpr + n * R + r4);
            dpv[k] = fmaf(w0, dv.x, fmaf(w1v, dv.y, fmaf(w2v, dv.z, fmaf(w3, dv.w, dpv[k]))));
          }
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < IPT; ++k) {
        const int n = j + k * G;
        if (n < N) for (int r = 0; r < R; ++r) dpv[k] = fmaf(w1s[r * C + c], dpr[n * R + r], dpv[k]);
      }
    }
    double S1 = 0.0, S2 = 0.0;
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
      const int n = j + k * G;
      if (n < N) {
        const float dp = dpv[k];
        S1 += (double)gt[k] * A[k] + (double)dp;
        S2 += (double)gt[k] * Bz[k] + (double)dp * (double)zm[k];
        dpool[(long long)n * Cs + c] = (float)((double)dp / (double)cnt);
      }
    }
    part[(j * 2 + 0) * C + c] = S1;
    part[(j * 2 + 1) * C + c] = S2;
  }
  __syncthreads();
  FIN_T(5);
  const double Mtot = (double)cnt * (double)N;
  for (int cc = tid; cc < Cs; cc += nthr) {
    double S1 = 0.0, S2 = 0.0;
    if (cc < C) {
      for (int g = 0; g < G; ++g) { S1 += part[(g * 2 + 0) * C + cc]; S2 += part[(g * 2 + 1) * C + cc]; }
      dgamma[cc] = (float)S2;
      dbeta[cc] = (float)S1;
    } else {
      for (int n = 0; n < N; ++n) dpool[(long long)n * Cs + cc] = 0.f;
    }
    coef[cc] = (float)(S1 / Mtot);
    coef[Cs + cc] = (float)(S2 / Mtot);
  }
  FIN_T(6);
#undef FIN_T
}

// Cluster version of the SE case: the single-CTA kernels above are bound by ONE SM's load bandwidth and shared-memory
// slots (33-60 us for ~0.5 MFLOP; phase timers in profiles/r02_summary.md).  Eight CTAs of a thread-block cluster split
// the work: CTA b owns a slice of the channels for everything that is per channel (dps, the fc2 / fc1 weight-gradient
// columns, dp, the batch sums and BN coefficients) and a slice of the samples for the one reduction over ALL channels
// (the gradient of the hidden activations).  The two transposes between the slicings go through a small global scratch
// (dps: N x C floats, dpr: N x R floats) with a cluster barrier after each; scratch reads bypass L1.
namespace {
constexpr int SE_CL = 8;          // CTAs per cluster
constexpr int SE_NT = 256;
struct SeClArgs {
  const double* stats; const float* gamma; const float* beta; const float* gate; const float* hidden; const float* zhat_mean;
  const float* w1; const float* w2;
  float* scratch;                 // [N * C] dps, then [N * R] dpr
  float* coef; float* dgamma; float* dbeta; float* dpool; float* dw1; float* db1; float* dw2; float* db2;
  long long cnt;
  int N, C, Cs, R, Cb, Nb, H;
};

__device__ __forceinline__ void cluster_sync_all() {
  __threadfence();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
}  // namespace

__global__ void __cluster_dims__(SE_CL, 1, 1) __launch_bounds__(SE_NT) se_bn_bwd_cluster_kernel(const SeClArgs a) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ double smd[];
  const int tid = threadIdx.x, rank = blockIdx.x;
  const int N = a.N, C = a.C, Cs = a.Cs, R = a.R, Cb = a.Cb, Nb = a.Nb;
  const int c_lo = rank * Cb, c_n = max(0, min(C, c_lo + Cb) - c_lo);
  const int n_lo = rank * Nb, n_n = max(0, min(N, n_lo + Nb) - n_lo);
  double* sA = smd;                                   // [N][Cb]
  double* sB = sA + (size_t)N * Cb;                   // [N][Cb]
  double* part = sB + (size_t)N * Cb;                 // [Q][2][Cb]   Q = SE_NT / Cb groups of samples
  const int Q = max(1, min(N, SE_NT / max(Cb, 1)));
  float* sg = reinterpret_cast<float*>(part + (size_t)Q * 2 * Cb);   // [N][Cb]
  float* szm = sg + (size_t)N * Cb;
  float* spin = szm + (size_t)N * Cb;
  float* sdps = spin + (size_t)N * Cb;
  float* shid = sdps + (size_t)N * Cb;                // [N][R]
  float* sdpr = shid + (size_t)N * R;                 // [N][R]
  float* sw2 = sdpr + (size_t)N * R;                  // [C][R]
  float* sw1 = sw2 + (size_t)C * R;                   // [R][Cb]
  float* srow = sw1 + (size_t)R * Cb;                 // [Nb][C]  dps rows of this CTA's samples
  float* dps_g = a.scratch;
  float* dpr_g = a.scratch + (size_t)N * C;

  // ---- phase 0: per (sample, own channel): loads, dps, pin ----
  for (int i = tid; i < N * Cb; i += SE_NT) {
    const int n = i / Cb, cc = i - n * Cb;
    double A = 0.0, B = 0.0;
    float g = 0.f, zm = 0.f, dps = 0.f, pin = 0.f;
    if (cc < c_n) {
      const int c = c_lo + cc;
      A = a.stats[((long long)n * 2) * Cs + c];
      B = a.stats[((long long)n * 2 + 1) * Cs + c];
      g = a.gate[(long long)n * Cs + c];
      zm = a.zhat_mean[(long long)n * Cs + c];
      const float gam = a.gamma[c], bet = a.beta[c];
      const float dg = (float)((double)gam * B + (double)bet * A);      // sum du * z
      dps = dg * g * (1.f - g);
      pin = fmaf(zm, gam, bet);
      dps_g[n * C + c] = dps;
    }
    sA[i] = A; sB[i] = B; sg[i] = g; szm[i] = zm; sdps[i] = dps; spin[i] = pin;
  }
  for (int i = tid; i < N * R; i += SE_NT) shid[i] = a.hidden[i];
  for (int i = tid; i < C * R; i += SE_NT) sw2[i] = a.w2[i];
  for (int i = tid; i < R * Cb; i += SE_NT) {
    const int r = i / Cb, cc = i - r * Cb;
    sw1[i] = cc < c_n ? a.w1[r * C + c_lo + cc] : 0.f;
  }
  cluster_sync_all();

  // ---- phase 1: gradient of the hidden activations for own samples (all channels); fc2 gradients for own channels ----
  for (int i = tid; i < n_n * C; i += SE_NT) srow[i] = __ldcg(dps_g + (size_t)n_lo * C + i);
  __syncthreads();
  {
    const int H = a.H, h = tid & (H - 1);
    for (int base = 0; base < Nb * R; base += SE_NT / H) {            // block-uniform trip count
      const int o = base + tid / H;
      const bool ok = o < n_n * R;
      const int nn = ok ? o / R : 0, r = ok ? o - nn * R : 0;
      float s0 = 0.f, s1 = 0.f;
      int c = h;
      for (; c + H < C; c += 2 * H) {
        s0 = fmaf(sw2[c * R + r], srow[nn * C + c], s0);
        s1 = fmaf(sw2[(c + H) * R + r], srow[nn * C + c + H], s1);
      }
      if (c < C) s0 = fmaf(sw2[c * R + r], srow[nn * C + c], s0);
      float sres = s0 + s1;
      for (int of = H >> 1; of; of >>= 1) sres += __shfl_xor_sync(0xffffffffu, sres, of);
      if (ok && h == 0) {
        const int n = n_lo + nn;
        dpr_g[n * R + r] = shid[n * R + r] > 0.f ? sres : 0.f;
      }
    }
  }
  for (int cc = tid; cc < c_n; cc += SE_NT) {
    float s0 = 0.f;
    for (int n = 0; n < N; ++n) s0 += sdps[n * Cb + cc];
    a.db2[c_lo + cc] = s0;
  }
  for (int i = tid; i < c_n * R; i += SE_NT) {
    const int cc = i / R, r = i - cc * R;
    float s0 = 0.f, s1 = 0.f;
    int n = 0;
    for (; n + 1 < N; n += 2) {
      s0 = fmaf(sdps[n * Cb + cc], shid[n * R + r], s0);
      s1 = fmaf(sdps[(n + 1) * Cb + cc], shid[(n + 1) * R + r], s1);
    }
    if (n < N) s0 = fmaf(sdps[n * Cb + cc], shid[n * R + r], s0);
    a.dw2[(c_lo + cc) * R + r] = s0 + s1;
  }
  cluster_sync_all();

  // ---- phase 2: fc1 gradients and the batch sums / BN coefficients for own channels ----
  for (int i = tid; i < N * R; i += SE_NT) sdpr[i] = __ldcg(dpr_g + i);
  __syncthreads();
  if (rank == 0)
    for (int r = tid; r < R; r += SE_NT) {
      float s0 = 0.f;
      for (int n = 0; n < N; ++n) s0 += sdpr[n * R + r];
      a.db1[r] = s0;
    }
  for (int i = tid; i < R * c_n; i += SE_NT) {
    const int r = i / c_n, cc = i - r * c_n;
    float s0 = 0.f, s1 = 0.f;
    int n = 0;
    for (; n + 1 < N; n += 2) {
      s0 = fmaf(sdpr[n * R + r], spin[n * Cb + cc], s0);
      s1 = fmaf(sdpr[(n + 1) * R + r], spin[(n + 1) * Cb + cc], s1);
    }
    if (n < N) s0 = fmaf(sdpr[n * R + r], spin[n * Cb + cc], s0);
    a.dw1[r * C + c_lo + cc] = s0 + s1;
  }
  {
    const int q = tid / max(Cb, 1), cc = tid - q * max(Cb, 1);
    if (q < Q && cc < c_n) {
      double S1 = 0.0, S2 = 0.0;
      for (int n = q; n < N; n += Q) {
        float dp = 0.f;
        for (int r = 0; r < R; ++r) dp = fmaf(sw1[r * Cb + cc], sdpr[n * R + r], dp);
        const int i = n * Cb + cc;
        S1 += (double)sg[i] * sA[i] + (double)dp;
        S2 += (double)sg[i] * sB[i] + (double)dp * (double)szm[i];
        a.dpool[(long long)n * Cs + c_lo + cc] = (float)((double)dp / (double)a.cnt);
      }
      part[(q * 2 + 0) * Cb + cc] = S1;
      part[(q * 2 + 1) * Cb + cc] = S2;
    }
  }
  __syncthreads();
  const double Mtot = (double)a.cnt * (double)N;
  for (int cc = tid; cc < c_n; cc += SE_NT) {
    double S1 = 0.0, S2 = 0.0;
    for (int q = 0; q < Q; ++q) { S1 += part[(q * 2 + 0) * Cb + cc]; S2 += part[(q * 2 + 1) * Cb + cc]; }
    const int c = c_lo + cc;
    a.dgamma[c] = (float)S2;
    a.dbeta[c] = (float)S1;
    a.coef[c] = (float)(S1 / Mtot);
    a.coef[Cs + c] = (float)(S2 / Mtot);
  }
  if (rank == SE_CL - 1)                               // pad channels: zero coefficients and pooled gradients
    for (int c = C + tid; c < Cs; c += SE_NT) {
      a.coef[c] = 0.f; a.coef[Cs + c] = 0.f;
      for (int n = 0; n < N; ++n) a.dpool[(long long)n * Cs + c] = 0.f;
    }
}

extern "C" int c3d_se_bn_bwd_finalize(const double* stats, int N, long long count_per_sample, const float* bnp,
                                      const float* gamma, const float* beta, const float* gate, const float* hidden,
                                      const float* zhat_mean, const float* w1, const float* w2, int C, int Cs, int R,
                                      float* coef, float* dgamma, float* dbeta, float* dpool, float* dw1, float* db1,
                                      float* dw2, float* db2, void* stream_) {
  if (!stats || !bnp || !gamma || !beta || !coef || !dgamma || !dbeta || N <= 0 || count_per_sample <= 0 || C <= 0 || Cs < C)
    return C3D_ERR_ARG;
  if (gate && (!hidden || !zhat_mean || !w1 || !w2 || !dpool || !dw1 || !db1 || !dw2 || !db2 || R <= 0)) return C3D_ERR_ARG;
  static const bool dbg_on = getenv("C3D_FIN_DBG") && atoi(getenv("C3D_FIN_DBG")) != 0;
  static long long* dbuf = nullptr;
  if (dbg_on && !dbuf) cudaMalloc(&dbuf, 8 * sizeof(long long));
  auto report = [&](const char* which) {      // bring-up aid (synchronising): cycles per phase of the single CTA
    long long h[8];
    cudaMemcpyAsync(h, dbuf, sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream_);
    cudaStreamSynchronize((cudaStream_t)stream_);
    fprintf(stderr, "[findbg] %s se=%d N=%d C=%d R=%d p0=%lld p1=%lld p2=%lld p3=%lld p4=%lld p5=%lld\n", which, gate != nullptr, N, C, R,
            h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4], h[6] - h[5]);
  };
  // C3D_FIN_FAST: 1 (default) latency-lean single CTA, 2 cluster of 8 CTAs, 0 the round-1 kernel.  The cluster kernel is
  // the faster one alone (21 vs 39 us) and in a step without the side stream (54.74 vs 55.13 ms), but with the weight
  // gradients running beside the main stream (the default) an 8-CTA cluster has to wait for eight free SMs of one GPC and
  // the step gets slower (55.25 vs 54.51 ms, same box): profiles/r02_summary.md section 2.3.
  static const int fast_mode = getenv("C3D_FIN_FAST") ? atoi(getenv("C3D_FIN_FAST")) : 1;
  static const bool fast_on = fast_mode != 0;
  if (gate && fast_mode >= 2 && !dbg_on && C <= 1024 && R <= 64 && N <= 64) {
    // scratch for the two transposes: one buffer per device, allocated on first use (the first call of a process is
    // never inside a CUDA-graph capture: train_step warms up eagerly before it captures)
    static float* scratch[16] = {nullptr};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 16) {
      if (!scratch[dev] && cudaMalloc(&scratch[dev], (size_t)(64 * 1024 + 64 * 64) * sizeof(float)) != cudaSuccess) scratch[dev] = nullptr;
      if (scratch[dev]) {
        SeClArgs a;
        a.stats = stats; a.gamma = gamma; a.beta = beta; a.gate = gate; a.hidden = hidden; a.zhat_mean = zhat_mean;
        a.w1 = w1; a.w2 = w2; a.scratch = scratch[dev];
        a.coef = coef; a.dgamma = dgamma; a.dbeta = dbeta; a.dpool = dpool; a.dw1 = dw1; a.db1 = db1; a.dw2 = dw2; a.db2 = db2;
        a.cnt = count_per_sample; a.N = N; a.C = C; a.Cs = Cs; a.R = R;
        a.Cb = (C + SE_CL - 1) / SE_CL; a.Nb = (N + SE_CL - 1) / SE_CL;
        int H = 1;
        while (2 * H * a.Nb * R <= SE_NT && 2 * H <= 32) H *= 2;
        a.H = H;
        int Q = SE_NT / a.Cb;
        if (Q > N) Q = N;
        if (Q < 1) Q = 1;
        const size_t csmem = ((size_t)2 * N * a.Cb + (size_t)Q * 2 * a.Cb) * sizeof(double) +
                             ((size_t)4 * N * a.Cb + (size_t)2 * N * R + (size_t)C * R + (size_t)R * a.Cb + (size_t)a.Nb * C) * sizeof(float);
        if (a.Cb <= SE_NT && csmem <= 200 * 1024 &&
            cudaFuncSetAttribute(se_bn_bwd_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem) == cudaSuccess) {
          se_bn_bwd_cluster_kernel<<<SE_CL, SE_NT, csmem, (cudaStream_t)stream_>>>(a);
          return c3d_check_last(cudaGetLastError());
        }
      }
    }
  }
  if (gate && fast_on && C <= 1024 && R <= 64) {
    int G = 1024 / C;
    if (G > N) G = N;
    const int ipt = (N + G - 1) / G;
    int H = 1;
    while (2 * H * N * R <= 1024 && 2 * H <= 32) H *= 2;
    const size_t fsmem = (size_t)G * 2 * C * sizeof(double) + ((size_t)2 * N * C + (size_t)2 * N * R + (size_t)2 * R * C) * sizeof(float);
    if (ipt <= 8 && fsmem <= 220 * 1024) {
      auto kern = ipt <= 2 ? se_bn_bwd_fast_kernel<2> : ipt <= 4 ? se_bn_bwd_fast_kernel<4> : se_bn_bwd_fast_kernel<8>;
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem) != cudaSuccess) return C3D_ERR_SMEM;
      c3d_launch_pdl_small(kern, dim3(1), dim3(1024), fsmem, (cudaStream_t)stream_, stats, N, count_per_sample, gamma, beta, gate, hidden,
                           zhat_mean, w1, w2, C, Cs, R, G, H, coef, dgamma, dbeta, dpool, dw1, db1, dw2, db2,
                           dbg_on ? dbuf : (long long*)nullptr);
      if (dbg_on) report("fast");
      return c3d_check_last(cudaGetLastError());
    }
  }
  size_t smem = (size_t)8 * C * sizeof(double) + (gate ? ((size_t)2 * N * C + (size_t)2 * N * R) * sizeof(float) : 0);
  if (smem > 200 * 1024) return C3D_ERR_SMEM;
  cudaFuncSetAttribute(se_bn_bwd_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  c3d_launch_pdl_small(se_bn_bwd_finalize_kernel, dim3(1), dim3(1024), smem, (cudaStream_t)stream_, stats, N, count_per_sample, bnp, gamma, beta, gate,
                                                                    hidden, zhat_mean, w1, w2, C, Cs, R, coef, dgamma,
                                                                    dbeta, dpool, dw1, db1, dw2, db2, dbg_on ? dbuf : (long long*)nullptr);
  if (dbg_on) report("slow");
  return c3d_check_last(cudaGetLastError());
}

// out[c] += sum over rows of X[row][c]  (ConvTranspose2d bias gradient)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, long long M, int Cs, float* __restrict__ out) {
  extern __shared__ float sm[];
  for (int i = threadIdx.x; i < Cs; i += 256) sm[i] = 0.f;
  __syncthreads();
  const int q4 = Cs >> 2, rpb = 256 / q4, tid = threadIdx.x;
  if (tid < rpb * q4) {
    const int q = tid % q4, rl = tid / q4;
    float4 s = f4zero();
    for (long long row = (long long)blockIdx.x * rpb + rl; row < M; row += (long long)gridDim.x * rpb)
      s = f4add(s, ldg4(X + row * Cs + 4 * q));
    atomicAdd(&sm[4 * q], s.x); atomicAdd(&sm[4 * q + 1], s.y); atomicAdd(&sm[4 * q + 2], s.z); atomicAdd(&sm[4 * q + 3], s.w);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Cs; i += 256) atomicAdd(out + i, sm[i]);
}

extern "C" int c3d_colsum(const float* X, long long M, int Cs, float* out, void* stream_) {
  if (!X || !out || M <= 0 || Cs <= 0 || (Cs & 3) || Cs > 1024) return C3D_ERR_ARG;
  const int rpb = 256 / (Cs >> 2);
  long long blocks = (M + rpb - 1) / rpb;
  if (blocks > 148 * 4) blocks = 148 * 4;
  colsum_kernel<<<(unsigned)blocks, 256, Cs * sizeof(float), (cudaStream_t)stream_>>>(X, M, Cs, out);
  return c3d_check_last(cudaGetLastError());
}
