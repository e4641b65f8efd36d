// ChangeDecoder head (model/change_decoder.py:53-55,76-79): Conv2d 3x3 (C -> ncls, pad 1, no bias)
// [+ sigmoid], NHWC in, NCHW out (the layout the reference returns), and its backward.
// The 1x1 Conv2d of the up-blocks runs through the pointwise-GEMM family.  ConvTranspose2d(k4, s2, p1)
// (model/change_decoder.py:30-45) is split into a dense GEMM over the input grid that produces all 16 kernel
// positions per input pixel, U[j][i][ky][kx][co] = sum_ci t[j][i][ci] W[ci][co][ky][kx], and the col2im gather
// below (output pixel (y, x) sums the 4 positions with 2j-1+ky = y, 2i-1+kx = x, plus bias and the skip tensor);
// the backward uses the mirrored im2col so that d t and d W are dense GEMMs as well (tcgen05 kernels).
#include <stdlib.h>

#include "c3d_common.cuh"
#include "../../include/change3d_b200.h"

#define HEAD_MAX_CLS 8

__global__ void __launch_bounds__(256) dec_head_fwd_kernel(const float* __restrict__ X, const float* __restrict__ w,
                                                           float* __restrict__ Y, int H, int W, int C, int ncls,
                                                           int apply_sigmoid) {
  extern __shared__ __align__(16) float ws[];   // [9][ncls][C]
  for (int i = threadIdx.x; i < 9 * ncls * C; i += 256) {
    int c = i % C, k = (i / C) % ncls, tap = i / (C * ncls);
    ws[i] = __ldg(w + (k * C + c) * 9 + tap);
  }
  __syncthreads();
  const int n = blockIdx.z;
  const int ow = blockIdx.x * 32 + (threadIdx.x & 31), oh = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (ow >= W || oh >= H) return;
  float acc[HEAD_MAX_CLS];
#pragma unroll
  for (int k = 0; k < HEAD_MAX_CLS; ++k) acc[k] = 0.f;
  for (int kh = 0; kh < 3; ++kh) {
    const int ih = oh - 1 + kh;
    if (ih < 0 || ih >= H) continue;
    for (int kw = 0; kw < 3; ++kw) {
      const int iw = ow - 1 + kw;
      if (iw < 0 || iw >= W) continue;
      const float* xp = X + (((long long)n * H + ih) * W + iw) * C;
      const float* wp = ws + (kh * 3 + kw) * ncls * C;
      for (int c = 0; c < C; c += 4) {
        const float4 x = ldg4(xp + c);
#pragma unroll
        for (int k = 0; k < HEAD_MAX_CLS; ++k) {
          if (k < ncls) {
            const float4 wv = *reinterpret_cast<const float4*>(wp + k * C + c);
            acc[k] = fmaf(x.x, wv.x, fmaf(x.y, wv.y, fmaf(x.z, wv.z, fmaf(x.w, wv.w, acc[k]))));
          }
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < HEAD_MAX_CLS; ++k) {
    if (k < ncls) {
      float v = acc[k];
      if (apply_sigmoid) v = 1.0f / (1.0f + expf(-v));
      Y[(((long long)n * ncls + k) * H + oh) * W + ow] = v;
    }
  }
}


// C = 24 fast path (the Change3D decoders): tile + halo staged in shared memory with whole-sector loads (the kernel
// above reads its 9 x 24 inputs per pixel straight from global memory at a 96-byte lane stride: 32 sectors per
// request, and its class loop is predicated on a runtime bound -- ncu: 87 % issue-active, 246 M warp instructions
// for 0.45 G FMA).  NCLS is a template parameter; one thread per output pixel.
#define HEAD_XLD 28
template <int NCLS>
__global__ void __launch_bounds__(256) dec_head_fwd24_kernel(const float* __restrict__ X, const float* __restrict__ w,
                                                             float* __restrict__ Y, int H, int W, int apply_sigmoid) {
  constexpr int C = 24, NQ = 6, PW = 34, PH = 10, NPIX = PW * PH;
  extern __shared__ __align__(16) float sm[];
  float* ws = sm;                         // [9][NCLS][24]
  float* xs = ws + 9 * NCLS * C;          // [NPIX][HEAD_XLD]
  const int tid = threadIdx.x;
  const int n = blockIdx.z, h0 = blockIdx.y * 8, w0 = blockIdx.x * 32;
  for (int i = tid; i < 9 * NCLS * C; i += 256) {
    const int c = i % C, k = (i / C) % NCLS, tap = i / (C * NCLS);
    ws[i] = __ldg(w + (k * C + c) * 9 + tap);
  }
  // tile staging in batches of 4 quads per thread: loads (clamped, always valid addresses) first, then the stores
  constexpr int XS_IT = (NPIX * NQ + 255) / 256;
#pragma unroll
  for (int k0 = 0; k0 < XS_IT; k0 += 4) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = tid + (k0 + u) * 256;
      const int ic = (k0 + u < XS_IT && i < NPIX * NQ) ? i : 0;
      const int p = ic / NQ, q = ic - p * NQ, y = p / PW, x = p - y * PW;
      const int h = h0 - 1 + y, ww = w0 - 1 + x;
      const int hc = h < 0 ? 0 : h >= H ? H - 1 : h, wc = ww < 0 ? 0 : ww >= W ? W - 1 : ww;
      const float4 t = ldg4(X + (((long long)n * H + hc) * W + wc) * C + 4 * q);
      v[u] = (hc == h && wc == ww) ? t : f4zero();
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = tid + (k0 + u) * 256;
      if (k0 + u < XS_IT && i < NPIX * NQ) {
        const int p = i / NQ, q = i - p * NQ;
        *reinterpret_cast<float4*>(xs + p * HEAD_XLD + 4 * q) = v[u];
      }
    }
  }
  __syncthreads();
  const int lx = tid & 31, ly = tid >> 5;
  const int ow = w0 + lx, oh = h0 + ly;
  if (ow >= W || oh >= H) return;
  float acc[NCLS];
#pragma unroll
  for (int k = 0; k < NCLS; ++k) acc[k] = 0.f;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const float* xp = xs + ((ly + kh) * PW + lx + kw) * HEAD_XLD;
      const float* wp = ws + (kh * 3 + kw) * NCLS * C;
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const float4 x = *reinterpret_cast<const float4*>(xp + 4 * q);
#pragma unroll
        for (int k = 0; k < NCLS; ++k) {
          const float4 wv = *reinterpret_cast<const float4*>(wp + k * C + 4 * q);
          acc[k] = fmaf(x.x, wv.x, fmaf(x.y, wv.y, fmaf(x.z, wv.z, fmaf(x.w, wv.w, acc[k]))));
        }
      }
    }
#pragma unroll
  for (int k = 0; k < NCLS; ++k) {
    float v = acc[k];
    if (apply_sigmoid) v = 1.0f / (1.0f + expf(-v));
    Y[(((long long)n * NCLS + k) * H + oh) * W + ow] = v;
  }
}

template <int NCLS>
static int launch_head_fwd24(const float* X, const float* w, float* Y, int B, int H, int W, int apply_sigmoid, cudaStream_t st) {
  dim3 grid((W + 31) / 32, (H + 7) / 8, B);
  const size_t smem = (size_t)(9 * NCLS * 24 + 340 * HEAD_XLD) * sizeof(float);
  if (cudaFuncSetAttribute(dec_head_fwd24_kernel<NCLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return C3D_ERR_SMEM;
  dec_head_fwd24_kernel<NCLS><<<grid, 256, smem, st>>>(X, w, Y, H, W, apply_sigmoid);
  return c3d_check_last(cudaGetLastError());
}

static bool head_fast_path() {
  const char* v = getenv("C3D_HEAD_FAST");      // 0: the generic kernels (any C, runtime class count)
  return !v || atoi(v) != 0;
}

extern "C" int c3d_dec_head_fwd(const float* X, const float* w, float* Y, int B, int H, int W, int C, int ncls,
                                int apply_sigmoid, void* stream_) {
  if (!X || !w || !Y || B <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 3) || ncls <= 0 || ncls > HEAD_MAX_CLS)
    return C3D_ERR_ARG;
  if (C == 24 && head_fast_path()) {
    cudaStream_t st = (cudaStream_t)stream_;
    switch (ncls) {
      case 1: return launch_head_fwd24<1>(X, w, Y, B, H, W, apply_sigmoid, st);
      case 2: return launch_head_fwd24<2>(X, w, Y, B, H, W, apply_sigmoid, st);
      case 3: return launch_head_fwd24<3>(X, w, Y, B, H, W, apply_sigmoid, st);
      case 4: return launch_head_fwd24<4>(X, w, Y, B, H, W, apply_sigmoid, st);
      case 5: return launch_head_fwd24<5>(X, w, Y, B, H, W, apply_sigmoid, st);
      case 6: return launch_head_fwd24<6>(X, w, Y, B, H, W, apply_sigmoid, st);
      case 7: return launch_head_fwd24<7>(X, w, Y, B, H, W, apply_sigmoid, st);
      default: return launch_head_fwd24<8>(X, w, Y, B, H, W, apply_sigmoid, st);
    }
  }
  dim3 grid((W + 31) / 32, (H + 7) / 8, B);
  size_t smem = (size_t)9 * ncls * C * sizeof(float);
  dec_head_fwd_kernel<<<grid, 256, smem, (cudaStream_t)stream_>>>(X, w, Y, H, W, C, ncls, apply_sigmoid);
  return c3d_check_last(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// Head backward: dlogit = dpred * p * (1 - p) (sigmoid head) or dpred;
//   dX[i][ci]      = sum_{k,tap} w[k][ci][tap] * dlogit[i - off(tap)][k]
//   dW[k][ci][tap] += sum_o dlogit[o][k] * X[o + off(tap)][ci]
// Persistent CTAs over 32x8 tiles; dlogit and X tiles (with a 1-pixel halo) staged in shared memory.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dec_head_bwd_kernel(const float* __restrict__ dpred, const float* __restrict__ pred,
                                                           const float* __restrict__ X, const float* __restrict__ w,
                                                           float* __restrict__ dX, float* __restrict__ dW, int B, int H,
                                                           int W, int C, int ncls, int is_sigmoid) {
  constexpr int PW = 34, PH = 10, NPIX = PW * PH;
  extern __shared__ __align__(16) float sm[];
  float* ws = sm;                        // [9][ncls][C]
  float* dl = ws + 9 * ncls * C;         // [ncls][NPIX]
  float* xs = dl + ncls * NPIX;          // [NPIX][HEAD_XLD]
  const int tid = threadIdx.x;
  for (int i = tid; i < 9 * ncls * C; i += 256) {
    int c = i % C, k = (i / C) % ncls, tap = i / (C * ncls);
    ws[i] = __ldg(w + (k * C + c) * 9 + tap);
  }
  const int nq = C >> 2;
  const int nitems = ncls * nq * 9;      // wgrad owners: (k, quad, tap)
  float4 wacc[2] = {f4zero(), f4zero()};
  const int tiles_x = (W + 31) / 32, tiles_y = (H + 7) / 8, ntiles = tiles_x * tiles_y * B;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int n = tile / (tiles_x * tiles_y);
    const int trem = tile - n * tiles_x * tiles_y;
    const int h0 = (trem / tiles_x) * 8, w0 = (trem % tiles_x) * 32;
    __syncthreads();
    for (int i = tid; i < ncls * NPIX; i += 256) {
      int k = i / NPIX, p = i - k * NPIX, y = p / PW, x = p - y * PW;
      int h = h0 - 1 + y, ww = w0 - 1 + x;
      float v = 0.f;
      if (h >= 0 && h < H && ww >= 0 && ww < W) {
        const long long off = (((long long)n * ncls + k) * H + h) * W + ww;
        v = __ldg(dpred + off);
        if (is_sigmoid) { const float p_ = __ldg(pred + off); v *= p_ * (1.f - p_); }
      }
      dl[i] = v;
    }
    for (int i = tid; i < NPIX * nq; i += 256) {
      int p = i / nq, q = i - p * nq, y = p / PW, x = p - y * PW;
      int h = h0 - 1 + y, ww = w0 - 1 + x;
      float4 v = f4zero();
      if (h >= 0 && h < H && ww >= 0 && ww < W) v = ldg4(X + (((long long)n * H + h) * W + ww) * C + 4 * q);
      *reinterpret_cast<float4*>(xs + p * HEAD_XLD + 4 * q) = v;
    }
    __syncthreads();
    // dgrad: one thread per interior pixel
    {
      const int lx = tid & 31, ly = tid >> 5;
      const int h = h0 + ly, ww = w0 + lx;
      if (h < H && ww < W) {
        for (int q = 0; q < nq; ++q) {
          float4 acc = f4zero();
          for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) {
              // out pixel o = i - off(tap) with off = (kh-1, kw-1): tile coords (ly+1-(kh-1), lx+1-(kw-1))
              const int p = (ly + 2 - kh) * PW + (lx + 2 - kw);
              for (int k = 0; k < ncls; ++k) {
                const float d = dl[k * NPIX + p];
                const float4 wv = *reinterpret_cast<const float4*>(ws + ((kh * 3 + kw) * ncls + k) * C + 4 * q);
                acc.x = fmaf(d, wv.x, acc.x); acc.y = fmaf(d, wv.y, acc.y); acc.z = fmaf(d, wv.z, acc.z); acc.w = fmaf(d, wv.w, acc.w);
              }
            }
          st4(dX + (((long long)n * H + h) * W + ww) * C + 4 * q, acc);
        }
      }
    }
    // wgrad owners
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int item = tid + it * 256;
      if (item >= nitems) break;
      const int tap = item % 9, q = (item / 9) % nq, k = item / (9 * nq);
      const int kh = tap / 3, kw = tap - kh * 3;
      float4 a = wacc[it];
      for (int y = 1; y <= 8; ++y) {
        const float* dp = dl + k * NPIX + y * PW;
        const float* xp = xs + ((y - 1 + kh) * PW + kw) * HEAD_XLD + 4 * q;
#pragma unroll 4
        for (int x = 1; x <= 32; ++x) {
          const float d = dp[x];
          const float4 xv = *reinterpret_cast<const float4*>(xp + (x - 1) * HEAD_XLD);
          a.x = fmaf(d, xv.x, a.x); a.y = fmaf(d, xv.y, a.y); a.z = fmaf(d, xv.z, a.z); a.w = fmaf(d, xv.w, a.w);
        }
      }
      wacc[it] = a;
    }
  }
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int item = tid + it * 256;
    if (item >= nitems) break;
    const int tap = item % 9, q = (item / 9) % nq, k = item / (9 * nq);
    atomicAdd(dW + (k * C + 4 * q + 0) * 9 + tap, wacc[it].x); atomicAdd(dW + (k * C + 4 * q + 1) * 9 + tap, wacc[it].y);
    atomicAdd(dW + (k * C + 4 * q + 2) * 9 + tap, wacc[it].z); atomicAdd(dW + (k * C + 4 * q + 3) * 9 + tap, wacc[it].w);
  }
}


// C = 24 fast path of the head backward.  ncu on the kernel above: 240 M warp instructions, barrier + wait stalls --
// its loops run over runtime bounds, the weight gradient of a 1-class head is owned by 54 of the 256 threads, and
// the dgrad pass issues two LDS per 4 FMA.  Here: the class count is a template parameter; the dgrad pass loads one
// dlogit per (tap, class) and sweeps the 6 channel quads (7 LDS per 24 FMA); a weight-gradient thread owns one
// (class, channel quad) with all 9 taps in registers and walks 8-pixel row segments with a sliding 3 x 3 window of
// input quads (3 LDS.128 + 1 LDS per 36 FMA); the segments of a tile are spread over 256 / (6 NCLS) threads per owner.
template <int NCLS>
__global__ void __launch_bounds__(256, 2) dec_head_bwd24_kernel(const float* __restrict__ dpred, const float* __restrict__ pred,
                                                                const float* __restrict__ X, const float* __restrict__ w,
                                                                float* __restrict__ dX, float* __restrict__ dW, int B, int H,
                                                                int W, int is_sigmoid) {
  constexpr int C = 24, NQ = 6, PW = 34, PH = 10, NPIX = PW * PH;
  constexpr int XPW = 35;   // odd row stride of the input tile: the 8 rows of a quarter warp fall into 8 different bank groups
  constexpr int NI = NCLS * NQ;                                // (class, quad) owners
  constexpr int TP = (256 / NI) > 32 ? 32 : (256 / NI);        // threads per owner; a tile has 32 segments of 8 pixels
  extern __shared__ __align__(16) float sm[];
  float* ws = sm;                        // [9][NCLS][24]
  float* red = ws + 9 * NCLS * C;        // [NCLS][24][9] weight-gradient reduction (end of the launch)
  float* dl = red + 9 * NCLS * C;        // [NCLS][NPIX]
  float* xs = dl + ((NCLS * NPIX + 3) & ~3);   // [PH][XPW][HEAD_XLD]
  const int tid = threadIdx.x;
  for (int i = tid; i < 9 * NCLS * C; i += 256) {
    const int c = i % C, k = (i / C) % NCLS, tap = i / (C * NCLS);
    ws[i] = __ldg(w + (k * C + c) * 9 + tap);
    red[i] = 0.f;
  }
  const int owner = tid / TP, sub = tid - owner * TP;
  const bool wg_active = owner < NI;
  const int wk = owner / NQ, wq = owner - wk * NQ;
  float4 wacc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) wacc[t] = f4zero();
  const int tiles_x = (W + 31) / 32, tiles_y = (H + 7) / 8, ntiles = tiles_x * tiles_y * B;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int n = tile / (tiles_x * tiles_y);
    const int trem = tile - n * tiles_x * tiles_y;
    const int h0 = (trem / tiles_x) * 8, w0 = (trem % tiles_x) * 32;
    __syncthreads();
    for (int i = tid; i < NCLS * NPIX; i += 256) {
      const int k = i / NPIX, p = i - k * NPIX, y = p / PW, x = p - y * PW;
      const int h = h0 - 1 + y, ww = w0 - 1 + x;
      float v = 0.f;
      if (h >= 0 && h < H && ww >= 0 && ww < W) {
        const long long off = (((long long)n * NCLS + k) * H + h) * W + ww;
        v = __ldg(dpred + off);
        if (is_sigmoid) { const float p_ = __ldg(pred + off); v *= p_ * (1.f - p_); }
      }
      dl[i] = v;
    }
    // input tile in batches of 4 quads per thread: loads (clamped, always valid addresses) first, then the stores
    constexpr int XS_IT = (NPIX * NQ + 255) / 256;
#pragma unroll
    for (int k0 = 0; k0 < XS_IT; k0 += 4) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = tid + (k0 + u) * 256;
        const int ic = (k0 + u < XS_IT && i < NPIX * NQ) ? i : 0;
        const int p = ic / NQ, q = ic - p * NQ, y = p / PW, x = p - y * PW;
        const int h = h0 - 1 + y, ww = w0 - 1 + x;
        const int hc = h < 0 ? 0 : h >= H ? H - 1 : h, wc = ww < 0 ? 0 : ww >= W ? W - 1 : ww;
        const float4 t = ldg4(X + (((long long)n * H + hc) * W + wc) * C + 4 * q);
        v[u] = (hc == h && wc == ww) ? t : f4zero();
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = tid + (k0 + u) * 256;
        if (k0 + u < XS_IT && i < NPIX * NQ) {
          const int p = i / NQ, q = i - p * NQ, y = p / PW, x = p - y * PW;
          *reinterpret_cast<float4*>(xs + (y * XPW + x) * HEAD_XLD + 4 * q) = v[u];
        }
      }
    }
    __syncthreads();
    {  // dgrad: one thread per interior pixel; dX[i] = sum_{k,tap} w[k][.][tap] * dlogit[i - off(tap)][k]
      const int lx = tid & 31, ly = tid >> 5;
      const int h = h0 + ly, ww = w0 + lx;
      if (h < H && ww < W) {
        float4 acc[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) acc[q] = f4zero();
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const int p = (ly + 2 - kh) * PW + (lx + 2 - kw);
#pragma unroll
            for (int k = 0; k < NCLS; ++k) {
              const float d = dl[k * NPIX + p];
              const float* wp = ws + ((kh * 3 + kw) * NCLS + k) * C;
#pragma unroll
              for (int q = 0; q < NQ; ++q) {
                const float4 wv = *reinterpret_cast<const float4*>(wp + 4 * q);
                acc[q].x = fmaf(d, wv.x, acc[q].x); acc[q].y = fmaf(d, wv.y, acc[q].y);
                acc[q].z = fmaf(d, wv.z, acc[q].z); acc[q].w = fmaf(d, wv.w, acc[q].w);
              }
            }
          }
        float* o = dX + (((long long)n * H + h) * W + ww) * C;
#pragma unroll
        for (int q = 0; q < NQ; ++q) st4(o + 4 * q, acc[q]);
      }
    }
    if (wg_active) {  // dW[k][ci][tap] += sum_o dlogit[o][k] * X[o + off(tap)][ci]
      for (int seg = sub; seg < 32; seg += TP) {
        const int row = seg & 7, xs0 = (seg >> 3) * 8;              // lanes 0..7: rows 0..7 of one column block
        const float* xb = xs + (row * XPW + xs0) * HEAD_XLD + 4 * wq;       // halo (row + kh, xs0 + i + kw)
        const float* db = dl + wk * NPIX + (row + 1) * PW + xs0 + 1;
        float4 win[3][3];                                                     // [column mod 3][kh]
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) win[j][kh] = *reinterpret_cast<const float4*>(xb + (kh * XPW + j) * HEAD_XLD);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
            win[(i + 2) % 3][kh] = *reinterpret_cast<const float4*>(xb + (kh * XPW + i + 2) * HEAD_XLD);
          const float d = db[i];
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const float4 xv = win[(i + kw) % 3][kh];
              float4& a = wacc[kh * 3 + kw];
              a.x = fmaf(d, xv.x, a.x); a.y = fmaf(d, xv.y, a.y); a.z = fmaf(d, xv.z, a.z); a.w = fmaf(d, xv.w, a.w);
            }
        }
      }
    }
  }
  if (wg_active) {
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float* r = red + ((wk * C + 4 * wq) * 9) + t;
      atomicAdd(r, wacc[t].x); atomicAdd(r + 9, wacc[t].y); atomicAdd(r + 18, wacc[t].z); atomicAdd(r + 27, wacc[t].w);
    }
  }
  __syncthreads();
  for (int i = tid; i < 9 * NCLS * C; i += 256) atomicAdd(dW + i, red[i]);
}

template <int NCLS>
static int launch_head_bwd24(const float* dpred, const float* pred, const float* X, const float* w, float* dX, float* dW,
                             int B, int H, int W, int is_sigmoid, cudaStream_t st) {
  const size_t smem = (size_t)(2 * 9 * NCLS * 24 + ((NCLS * 340 + 3) & ~3) + 350 * HEAD_XLD) * sizeof(float);
  if (cudaFuncSetAttribute(dec_head_bwd24_kernel<NCLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return C3D_ERR_SMEM;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int ntiles = ((W + 31) / 32) * ((H + 7) / 8) * B;
  const int grid = ntiles < 2 * sms ? ntiles : 2 * sms;
  dec_head_bwd24_kernel<NCLS><<<grid, 256, smem, st>>>(dpred, pred, X, w, dX, dW, B, H, W, is_sigmoid);
  return c3d_check_last(cudaGetLastError());
}

extern "C" int c3d_dec_head_bwd(const float* dpred, const float* pred, const float* X, const float* w, float* dX,
                                float* dW, int B, int H, int W, int C, int ncls, int is_sigmoid, void* stream_) {
  if (!dpred || !X || !w || !dX || !dW || B <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 3) || C > 24 || ncls <= 0 ||
      ncls > HEAD_MAX_CLS)
    return C3D_ERR_ARG;
  if (is_sigmoid && !pred) return C3D_ERR_ARG;
  if (ncls * (C >> 2) * 9 > 512) return C3D_ERR_ARG;
  if (C == 24 && head_fast_path()) {
    cudaStream_t st = (cudaStream_t)stream_;
    switch (ncls) {
      case 1: return launch_head_bwd24<1>(dpred, pred, X, w, dX, dW, B, H, W, is_sigmoid, st);
      case 2: return launch_head_bwd24<2>(dpred, pred, X, w, dX, dW, B, H, W, is_sigmoid, st);
      case 3: return launch_head_bwd24<3>(dpred, pred, X, w, dX, dW, B, H, W, is_sigmoid, st);
      case 4: return launch_head_bwd24<4>(dpred, pred, X, w, dX, dW, B, H, W, is_sigmoid, st);
      case 5: return launch_head_bwd24<5>(dpred, pred, X, w, dX, dW, B, H, W, is_sigmoid, st);
      case 6: return launch_head_bwd24<6>(dpred, pred, X, w, dX, dW, B, H, W, is_sigmoid, st);
      case 7: return launch_head_bwd24<7>(dpred, pred, X, w, dX, dW, B, H, W, is_sigmoid, st);
      default: return launch_head_bwd24<8>(dpred, pred, X, w, dX, dW, B, H, W, is_sigmoid, st);
    }
  }
  const size_t smem = (size_t)(9 * ncls * C + ncls * 340 + 340 * HEAD_XLD) * sizeof(float);
  cudaFuncSetAttribute(dec_head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int ntiles = ((W + 31) / 32) * ((H + 7) / 8) * B;
  int grid = ntiles < 2 * sms ? ntiles : 2 * sms;
  dec_head_bwd_kernel<<<grid, 256, smem, (cudaStream_t)stream_>>>(dpred, pred, X, w, dX, dW, B, H, W, C, ncls, is_sigmoid);
  return c3d_check_last(cudaGetLastError());
}


// ---------------- ConvTranspose2d(k4, s2, p1) as dense GEMM + col2im / im2col ----------------
// out[b][y][x][co] = bias[co] + skip[b][y][x][co] + sum_{ky == (y+1) mod 2 (+2)} sum_{kx} U[b][(y+1-ky)/2][(x+1-kx)/2][ky][kx][co]
__global__ void __launch_bounds__(256) convt_col2im_kernel(const float* __restrict__ U, const float* __restrict__ skip,
                                                           long long skip_img_stride, const float* __restrict__ bias,
                                                           float* __restrict__ out, int h, int w, int cout, long long total4) {
  const int c4n = cout >> 2, OW = 2 * w, OH = 2 * h;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total4; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % c4n) * 4;
    long long pix = idx / c4n;
    const int x = (int)(pix % OW);
    pix /= OW;
    const int y = (int)(pix % OH);
    const int b = (int)(pix / OH);
    float4 acc = ldg4(bias + c);
    if (skip) acc = f4add(acc, ldg4(skip + (long long)b * skip_img_stride + ((long long)y * OW + x) * cout + c));
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int ky = ((y + 1) & 1) + 2 * a, j = (y + 1 - ky) >> 1;
      if (y + 1 - ky < 0 || j >= h) continue;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int kx = ((x + 1) & 1) + 2 * e, i = (x + 1 - kx) >> 1;
        if (x + 1 - kx < 0 || i >= w) continue;
        acc = f4add(acc, ldg4(U + ((((long long)b * h + j) * w + i) * 16 + ky * 4 + kx) * cout + c));
      }
    }
    st4(out + (((long long)b * OH + y) * OW + x) * cout + c, acc);
  }
}

// V[b][j][i][ky][kx][co] = d_out[b][2j-1+ky][2i-1+kx][co]  (zero outside the image)
__global__ void __launch_bounds__(256) convt_im2col_kernel(const float* __restrict__ dout, float* __restrict__ V, int h, int w,
                                                           int cout, long long total4) {
  const int c4n = cout >> 2, OW = 2 * w, OH = 2 * h;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total4; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % c4n) * 4;
    long long r = idx / c4n;
    const int tap = (int)(r & 15);
    r >>= 4;
    const int i = (int)(r % w);
    r /= w;
    const int j = (int)(r % h);
    const int b = (int)(r / h);
    const int y = 2 * j - 1 + (tap >> 2), x = 2 * i - 1 + (tap & 3);
    float4 v = f4zero();
    if (y >= 0 && y < OH && x >= 0 && x < OW) v = ldg4(dout + (((long long)b * OH + y) * OW + x) * cout + c);
    st4(V + idx * 4, v);
  }
}

// Same gather with 32-bit index arithmetic (total4 < 2^31), the quad count as a template parameter (division by a
// constant) and U elements per iteration with their loads issued first: the generic kernel above spends its time on
// three 64-bit divisions per 16 bytes (ncu: 66 % issue-active at 3.0 TB/s).
template <int C4N, int U>
__global__ void __launch_bounds__(256) convt_im2col32_kernel(const float* __restrict__ dout, float* __restrict__ V, unsigned h,
                                                             unsigned w, unsigned total4) {
  const unsigned OW = 2 * w, OH = 2 * h, cout = 4 * C4N;
  const unsigned stride = gridDim.x * blockDim.x;
  unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  auto src_of = [&](unsigned id, bool& ok) -> const float* {
    const unsigned c = (id % C4N) * 4;
    unsigned r = id / C4N;
    const unsigned tap = r & 15u;
    r >>= 4;
    const unsigned i = r % w;
    r /= w;
    const unsigned j = r % h, b = r / h;
    const int y = 2 * (int)j - 1 + (int)(tap >> 2), x = 2 * (int)i - 1 + (int)(tap & 3);
    ok = y >= 0 && y < (int)OH && x >= 0 && x < (int)OW;
    return dout + ((size_t)(b * OH + (unsigned)y) * OW + (unsigned)x) * cout + c;
  };
  for (; (unsigned long long)idx + (unsigned long long)(U - 1) * stride < total4; idx += U * stride) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      bool ok;
      const float* p = src_of(idx + u * stride, ok);
      v[u] = ok ? ldg4(p) : f4zero();
    }
#pragma unroll
    for (int u = 0; u < U; ++u) st4(V + (size_t)(idx + u * stride) * 4, v[u]);
  }
  for (; idx < total4; idx += stride) {
    bool ok;
    const float* p = src_of(idx, ok);
    st4(V + (size_t)idx * 4, ok ? ldg4(p) : f4zero());
  }
}

extern "C" int c3d_convt_col2im(const float* U, const float* skip, long long skip_img_stride, const float* bias, float* out,
                                int B, int h, int w, int cout, void* stream_) {
  if (!U || !bias || !out || B <= 0 || h <= 0 || w <= 0 || cout <= 0 || (cout & 3)) return C3D_ERR_ARG;
  const long long total4 = (long long)B * 4 * h * w * (cout >> 2);
  long long blocks = (total4 + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  convt_col2im_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(U, skip, skip_img_stride, bias, out, h, w, cout, total4);
  return c3d_check_last(cudaGetLastError());
}

extern "C" int c3d_convt_im2col(const float* dout, float* V, int B, int h, int w, int cout, void* stream_) {
  if (!dout || !V || B <= 0 || h <= 0 || w <= 0 || cout <= 0 || (cout & 3)) return C3D_ERR_ARG;
  const long long total4 = (long long)B * h * w * 16 * (cout >> 2);
  long long blocks = (total4 + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  // all index products below 2^31: 32-bit kernel (the Change3D decoders: cout = 24 / 48 / 96 ... any multiple of 4 up to 96)
  if (total4 < (1ll << 31) && (long long)B * 4 * h * w * cout < (1ll << 31) && head_fast_path()) {
    const unsigned ub = (unsigned)blocks;
    cudaStream_t st = (cudaStream_t)stream_;
    switch (cout >> 2) {
      case 6: convt_im2col32_kernel<6, 4><<<ub, 256, 0, st>>>(dout, V, (unsigned)h, (unsigned)w, (unsigned)total4); return c3d_check_last(cudaGetLastError());
      case 12: convt_im2col32_kernel<12, 4><<<ub, 256, 0, st>>>(dout, V, (unsigned)h, (unsigned)w, (unsigned)total4); return c3d_check_last(cudaGetLastError());
      case 24: convt_im2col32_kernel<24, 4><<<ub, 256, 0, st>>>(dout, V, (unsigned)h, (unsigned)w, (unsigned)total4); return c3d_check_last(cudaGetLastError());
      default: break;
    }
  }
  convt_im2col_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(dout, V, h, w, cout, total4);
  return c3d_check_last(cudaGetLastError());
}
