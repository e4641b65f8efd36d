// X3D stem (model/x3d.py:23-106) fused with Encoder.forward's frame assembly
// (model/trainer.py:154-162): [pre, P perception frames, post] -> Conv3d 1x3x3 (3->24, pad 0,1,1)
// -> depthwise Conv3d 5x1x1 (pad 2,0,0) -> raw NDHWC output + BatchNorm statistics.
// pytorchvideo's Conv2plus1d runs the module stored as conv_t (the spatial conv) first.
#include "c3d_common.cuh"
#include "../../include/change3d_b200.h"

#define STEM_C 24
#define STEM_TW 32
#define STEM_TH 8

// Frame f of sample n, channel ci is the H*W plane at p[f] + n*sn[f] + ci*sc[f].  The fused assembly
// passes pre / perception[0,:,f-1] (sn = 0: broadcast over the batch) / post; a generic
// (B,3,T,H,W) tensor passes x + f*H*W with sn = 3*T*H*W, sc = T*H*W.
struct StemFrames {
  const float* p[5];
  long long sn[5];
  long long sc[5];
};

template <int T>
__global__ void __launch_bounds__(256) stem_fwd_kernel(const StemFrames fr, const float* __restrict__ wxy,
                                                       const float* __restrict__ wt, float* __restrict__ Y,
                                                       double* __restrict__ stats, int H, int W) {
  constexpr int PW = STEM_TW + 2, PH = STEM_TH + 2;
  __shared__ float patch[T][3][PH][PW];
  __shared__ __align__(16) float s_wxy[27][STEM_C];   // [ci*9+kh*3+kw][c]
  __shared__ float s_wt[5][STEM_C];                   // [tap][c]
  __shared__ float s_stat[2][STEM_C];
  const int n = blockIdx.z, h0 = blockIdx.y * STEM_TH, w0 = blockIdx.x * STEM_TW;
  const int tid = threadIdx.x;
  for (int i = tid; i < 27 * STEM_C; i += 256) { int k = i / STEM_C, c = i - k * STEM_C; s_wxy[k][c] = __ldg(wxy + c * 27 + k); }
  for (int i = tid; i < 5 * STEM_C; i += 256) { int k = i / STEM_C, c = i - k * STEM_C; s_wt[k][c] = __ldg(wt + c * 5 + k); }
  if (tid < 2 * STEM_C) (&s_stat[0][0])[tid] = 0.f;
  for (int i = tid; i < T * 3 * PH * PW; i += 256) {
    int x = i % PW, y = (i / PW) % PH, ci = (i / (PW * PH)) % 3, f = i / (PW * PH * 3);
    int h = h0 - 1 + y, w = w0 - 1 + x;
    float v = 0.f;
    if (h >= 0 && h < H && w >= 0 && w < W) v = __ldg(fr.p[f] + n * fr.sn[f] + ci * fr.sc[f] + (long long)h * W + w);
    patch[f][ci][y][x] = v;
  }
  __syncthreads();
  const int lx = tid & (STEM_TW - 1), ly = tid / STEM_TW;
  const int h = h0 + ly, w = w0 + lx;
  const bool valid = (h < H && w < W);
  for (int cg = 0; cg < STEM_C / 4; ++cg) {
    float4 s[T];
#pragma unroll
    for (int f = 0; f < T; ++f) s[f] = f4zero();
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float4 wv = *reinterpret_cast<const float4*>(&s_wxy[ci * 9 + kh * 3 + kw][cg * 4]);
#pragma unroll
          for (int f = 0; f < T; ++f) {
            const float x = patch[f][ci][ly + kh][lx + kw];
            s[f].x = fmaf(x, wv.x, s[f].x); s[f].y = fmaf(x, wv.y, s[f].y);
            s[f].z = fmaf(x, wv.z, s[f].z); s[f].w = fmaf(x, wv.w, s[f].w);
          }
        }
    float4 sum = f4zero(), sq = f4zero();
#pragma unroll
    for (int to = 0; to < T; ++to) {
      float4 o = f4zero();
#pragma unroll
      for (int f = 0; f < T; ++f) {
        const int tap = f - to + 2;
        if (tap < 0 || tap > 4) continue;
        const float4 wv = *reinterpret_cast<const float4*>(&s_wt[tap][cg * 4]);
        o = f4fma(wv, s[f], o);
      }
      if (valid) {
        st4(Y + ((((long long)n * T + to) * H + h) * W + w) * STEM_C + cg * 4, o);
        sum = f4add(sum, o);
        sq = f4fma(o, o, sq);
      }
    }
    if (stats) {
      float v[8] = {sum.x, sum.y, sum.z, sum.w, sq.x, sq.y, sq.z, sq.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int o = 16; o; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
      }
      if ((tid & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { atomicAdd(&s_stat[0][cg * 4 + k], v[k]); atomicAdd(&s_stat[1][cg * 4 + k], v[4 + k]); }
      }
    }
  }
  if (stats) {
    __syncthreads();
    if (tid < 2 * STEM_C) atomicAdd(stats + tid, (double)(&s_stat[0][0])[tid]);
  }
}

static int stem_frames(StemFrames& fr, const float* const* frame_ptr, const long long* stride_n,
                       const long long* stride_c, int T) {
  if (!frame_ptr || !stride_n || !stride_c || T < 3 || T > 5) return C3D_ERR_ARG;
  for (int f = 0; f < 5; ++f) {
    fr.p[f] = f < T ? frame_ptr[f] : nullptr;
    fr.sn[f] = f < T ? stride_n[f] : 0;
    fr.sc[f] = f < T ? stride_c[f] : 0;
    if (f < T && !fr.p[f]) return C3D_ERR_ARG;
  }
  return C3D_OK;
}

extern "C" int c3d_stem_fwd(const float* const* frame_ptr, const long long* stride_n, const long long* stride_c,
                            const float* w_xy, const float* w_t, float* Y, double* stats, int B, int T, int H, int W,
                            void* stream_) {
  if (!w_xy || !w_t || !Y || B <= 0 || H <= 0 || W <= 0) return C3D_ERR_ARG;
  StemFrames fr;
  if (int e = stem_frames(fr, frame_ptr, stride_n, stride_c, T)) return e;
  dim3 grid((W + STEM_TW - 1) / STEM_TW, (H + STEM_TH - 1) / STEM_TH, B);
  cudaStream_t st = (cudaStream_t)stream_;
  switch (T) {
    case 3: stem_fwd_kernel<3><<<grid, 256, 0, st>>>(fr, w_xy, w_t, Y, stats, H, W); break;
    case 4: stem_fwd_kernel<4><<<grid, 256, 0, st>>>(fr, w_xy, w_t, Y, stats, H, W); break;
    case 5: stem_fwd_kernel<5><<<grid, 256, 0, st>>>(fr, w_xy, w_t, Y, stats, H, W); break;
    default: return C3D_ERR_ARG;
  }
  return c3d_check_last(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// Stem backward (autograd of model/x3d.py:70-99 + Encoder.forward's frame assembly):
//   dy = BN backward of d_pre (on the fly);  ds[f] = sum_t wt[f-t+2] * dy[t]  (temporal conv^T)
//   dwt[c][tap]        += dy[t][c] * s[t+tap-2][c]            (s = spatial conv output, recomputed)
//   dwxy[c][ci,kh,kw]  += ds[f][h,w,c] * in[f][ci][h+kh-1][w+kw-1]
//   dperc[ci][f-1][h][w] += sum_{c,kh,kw} wxy[c][ci,kh,kw] * ds[f][h-kh+1][w-kw+1][c]   (frames 1..P only)
// Persistent CTAs walk BTW x BTH pixel tiles; ds for tile+halo lives in shared memory.
// ------------------------------------------------------------------------------------------------
#define STEM_DSLD 28   // padded pixel stride of the ds tile (bank-conflict-free float4 reads)

template <int T, int BTW, int BTH>
__global__ void __launch_bounds__(256) stem_bwd_kernel(const StemFrames fr, const float* __restrict__ dpre,
                                                       const float* __restrict__ ys, const float* __restrict__ bnp,
                                                       const float* __restrict__ coef, const float* __restrict__ wxy,
                                                       const float* __restrict__ wt, float* __restrict__ dwxy,
                                                       float* __restrict__ dwt, float* __restrict__ dperc, int B, int H,
                                                       int W) {
  constexpr int P = T - 2;
  constexpr int PW = BTW + 2, PH = BTH + 2, NPIX = PW * PH;
  extern __shared__ __align__(16) float sm[];
  float* patch = sm;                                   // [T][3][PH][PW]
  float* ds = patch + T * 3 * NPIX;                    // [T][NPIX][STEM_DSLD]
  float* s_wxy = ds + T * NPIX * STEM_DSLD;            // [27][24]
  float* s_wt = s_wxy + 27 * STEM_C;                   // [5][24]
  float* s_dwt = s_wt + 5 * STEM_C;                    // [5][24]
  float* s_dwxy = s_dwt + 5 * STEM_C;                  // [27][24]
  const int tid = threadIdx.x;
  for (int i = tid; i < 27 * STEM_C; i += 256) { int k = i / STEM_C, c = i - k * STEM_C; s_wxy[i] = __ldg(wxy + c * 27 + k); s_dwxy[i] = 0.f; }
  for (int i = tid; i < 5 * STEM_C; i += 256) { int k = i / STEM_C, c = i - k * STEM_C; s_wt[i] = __ldg(wt + c * 5 + k); s_dwt[i] = 0.f; }

  // role A (ds + dwt): thread owns channel quad qa, walks pixel slots
  const int qa = tid % 6, slot = tid / 6;              // 42 slots active (tid < 252)
  const int ca = 4 * qa;
  const float4 mean = ldg4(bnp + ca), rstd = ldg4(bnp + STEM_C + ca), scale = ldg4(bnp + 2 * STEM_C + ca);
  const float4 c1 = ldg4(coef + ca), c2 = ldg4(coef + STEM_C + ca);
  float4 dwt_acc[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) dwt_acc[k] = f4zero();
  // role B (dwxy): thread owns (channel quad qb, tap tb), 162 threads
  const int qb = tid % 6, tb = tid / 6;                // tb < 27
  const int b_ci = tb / 9, b_kh = (tb % 9) / 3, b_kw = tb % 3;
  float4 dwxy_acc = f4zero();

  const int tiles_x = (W + BTW - 1) / BTW, tiles_y = (H + BTH - 1) / BTH;
  const int ntiles = tiles_x * tiles_y * B;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int n = tile / (tiles_x * tiles_y);
    const int trem = tile - n * tiles_x * tiles_y;
    const int h0 = (trem / tiles_x) * BTH, w0 = (trem % tiles_x) * BTW;
    __syncthreads();
    for (int i = tid; i < T * 3 * NPIX; i += 256) {
      int x = i % PW, y = (i / PW) % PH, ci = (i / NPIX) % 3, f = i / (NPIX * 3);
      int h = h0 - 1 + y, w = w0 - 1 + x;
      float v = 0.f;
      if (h >= 0 && h < H && w >= 0 && w < W) v = __ldg(fr.p[f] + n * fr.sn[f] + ci * fr.sc[f] + (long long)h * W + w);
      patch[i] = v;
    }
    __syncthreads();
    if (tid < 252) {
      for (int p = slot; p < NPIX; p += 42) {
        const int y = p / PW, x = p - y * PW;
        const int h = h0 - 1 + y, w = w0 - 1 + x;
        const bool inimg = (h >= 0 && h < H && w >= 0 && w < W);
        float4 dy[T];
        if (inimg) {
#pragma unroll
          for (int t = 0; t < T; ++t) {
            const long long off = ((((long long)n * T + t) * H + h) * W + w) * STEM_C + ca;
            const float4 d = ldg4(dpre + off), yv = ldg4(ys + off);
            dy[t].x = scale.x * (d.x - c1.x - (yv.x - mean.x) * rstd.x * c2.x);
            dy[t].y = scale.y * (d.y - c1.y - (yv.y - mean.y) * rstd.y * c2.y);
            dy[t].z = scale.z * (d.z - c1.z - (yv.z - mean.z) * rstd.z * c2.z);
            dy[t].w = scale.w * (d.w - c1.w - (yv.w - mean.w) * rstd.w * c2.w);
          }
        } else {
#pragma unroll
          for (int t = 0; t < T; ++t) dy[t] = f4zero();
        }
#pragma unroll
        for (int f = 0; f < T; ++f) {
          float4 o = f4zero();
#pragma unroll
          for (int t = 0; t < T; ++t) {
            const int tap = f - t + 2;
            if (tap < 0 || tap > 4) continue;
            o = f4fma(*reinterpret_cast<const float4*>(s_wt + tap * STEM_C + ca), dy[t], o);
          }
          *reinterpret_cast<float4*>(ds + (f * NPIX + p) * STEM_DSLD + ca) = o;
        }
        // temporal weight gradient needs s = spatial conv output at interior (non-halo, in-image) pixels
        const bool interior = inimg && y >= 1 && y <= BTH && x >= 1 && x <= BTW;
        if (interior) {
          float4 s[T];
#pragma unroll
          for (int f = 0; f < T; ++f) s[f] = f4zero();
#pragma unroll
          for (int ci = 0; ci < 3; ++ci)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) {
                const float4 wv = *reinterpret_cast<const float4*>(s_wxy + (ci * 9 + kh * 3 + kw) * STEM_C + ca);
#pragma unroll
                for (int f = 0; f < T; ++f) {
                  const float xv = patch[((f * 3 + ci) * PH + (y - 1 + kh)) * PW + (x - 1 + kw)];
                  s[f].x = fmaf(xv, wv.x, s[f].x); s[f].y = fmaf(xv, wv.y, s[f].y);
                  s[f].z = fmaf(xv, wv.z, s[f].z); s[f].w = fmaf(xv, wv.w, s[f].w);
                }
              }
#pragma unroll
          for (int t = 0; t < T; ++t)
#pragma unroll
            for (int f = 0; f < T; ++f) {
              const int tap = f - t + 2;
              if (tap < 0 || tap > 4) continue;
              dwt_acc[tap] = f4fma(dy[t], s[f], dwt_acc[tap]);
            }
        }
      }
    }
    __syncthreads();
    // role B: dwxy over the interior pixels of the tile
    if (tid < 162) {
      for (int f = 0; f < T; ++f) {
        const float* pp = patch + (f * 3 + b_ci) * NPIX;
        const float* dp = ds + f * NPIX * STEM_DSLD + 4 * qb;
        for (int y = 1; y <= BTH; ++y) {
#pragma unroll 4
          for (int x = 1; x <= BTW; ++x) {
            const float xv = pp[(y - 1 + b_kh) * PW + (x - 1 + b_kw)];
            const float4 d = *reinterpret_cast<const float4*>(dp + (y * PW + x) * STEM_DSLD);
            dwxy_acc.x = fmaf(xv, d.x, dwxy_acc.x); dwxy_acc.y = fmaf(xv, d.y, dwxy_acc.y);
            dwxy_acc.z = fmaf(xv, d.z, dwxy_acc.z); dwxy_acc.w = fmaf(xv, d.w, dwxy_acc.w);
          }
        }
      }
    }
    // role C: gradient of the perception frames, one thread per interior pixel
    if (dperc) {
      const int lx = tid & (BTW - 1), ly = tid / BTW;
      const int h = h0 + ly, w = w0 + lx;
      if (ly < BTH && h < H && w < W) {
#pragma unroll
        for (int f = 1; f <= P; ++f) {
          float g0 = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              // ds at (h - kh + 1, w - kw + 1) -> tile coords (ly + 1 - kh + 1, lx + 1 - kw + 1)
              const float* dp = ds + (f * NPIX + (ly + 2 - kh) * PW + (lx + 2 - kw)) * STEM_DSLD;
#pragma unroll
              for (int q = 0; q < 6; ++q) {
                const float4 d = *reinterpret_cast<const float4*>(dp + 4 * q);
                const float4 w0v = *reinterpret_cast<const float4*>(s_wxy + (0 * 9 + kh * 3 + kw) * STEM_C + 4 * q);
                const float4 w1v = *reinterpret_cast<const float4*>(s_wxy + (1 * 9 + kh * 3 + kw) * STEM_C + 4 * q);
                const float4 w2v = *reinterpret_cast<const float4*>(s_wxy + (2 * 9 + kh * 3 + kw) * STEM_C + 4 * q);
                g0 = fmaf(d.x, w0v.x, fmaf(d.y, w0v.y, fmaf(d.z, w0v.z, fmaf(d.w, w0v.w, g0))));
                g1 = fmaf(d.x, w1v.x, fmaf(d.y, w1v.y, fmaf(d.z, w1v.z, fmaf(d.w, w1v.w, g1))));
                g2 = fmaf(d.x, w2v.x, fmaf(d.y, w2v.y, fmaf(d.z, w2v.z, fmaf(d.w, w2v.w, g2))));
              }
            }
          const long long HW = (long long)H * W, pix = (long long)h * W + w;
          atomicAdd(dperc + (0 * P + (f - 1)) * HW + pix, g0);
          atomicAdd(dperc + (1 * P + (f - 1)) * HW + pix, g1);
          atomicAdd(dperc + (2 * P + (f - 1)) * HW + pix, g2);
        }
      }
    }
  }
  // flush
  if (tid < 252) {
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      atomicAdd(&s_dwt[k * STEM_C + ca], dwt_acc[k].x); atomicAdd(&s_dwt[k * STEM_C + ca + 1], dwt_acc[k].y);
      atomicAdd(&s_dwt[k * STEM_C + ca + 2], dwt_acc[k].z); atomicAdd(&s_dwt[k * STEM_C + ca + 3], dwt_acc[k].w);
    }
  }
  if (tid < 162) {
    s_dwxy[tb * STEM_C + 4 * qb] = dwxy_acc.x; s_dwxy[tb * STEM_C + 4 * qb + 1] = dwxy_acc.y;
    s_dwxy[tb * STEM_C + 4 * qb + 2] = dwxy_acc.z; s_dwxy[tb * STEM_C + 4 * qb + 3] = dwxy_acc.w;
  }
  __syncthreads();
  for (int i = tid; i < 27 * STEM_C; i += 256) { int k = i / STEM_C, c = i - k * STEM_C; atomicAdd(dwxy + c * 27 + k, s_dwxy[i]); }
  for (int i = tid; i < 5 * STEM_C; i += 256) { int k = i / STEM_C, c = i - k * STEM_C; atomicAdd(dwt + c * 5 + k, s_dwt[i]); }
}

template <int T, int BTW, int BTH>
static int launch_stem_bwd_tile(const StemFrames& fr, const float* dpre, const float* ys, const float* bnp, const float* coef,
                                const float* wxy, const float* wt, float* dwxy, float* dwt, float* dperc, int B, int H, int W,
                                cudaStream_t st) {
  static_assert(BTW * BTH <= 256 && (BTW & (BTW - 1)) == 0, "one thread per interior pixel in the perception-frame pass");
  constexpr int NPIX = (BTW + 2) * (BTH + 2);
  const size_t smem = (size_t)(T * 3 * NPIX + T * NPIX * STEM_DSLD + 27 * STEM_C * 2 + 5 * STEM_C * 2) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(stem_bwd_kernel<T, BTW, BTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return C3D_ERR_SMEM;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int ntiles = ((W + BTW - 1) / BTW) * ((H + BTH - 1) / BTH) * B;
  int per_sm = (int)((220 * 1024) / (smem + 1024));          // co-resident CTAs hide each other's phase barriers
  per_sm = per_sm < 1 ? 1 : per_sm > 2 ? 2 : per_sm;           // 127 registers x 256 threads: two CTAs per SM
  int grid = ntiles < sms * per_sm ? ntiles : sms * per_sm;
  stem_bwd_kernel<T, BTW, BTH><<<grid, 256, smem, st>>>(fr, dpre, ys, bnp, coef, wxy, wt, dwxy, dwt, dperc, B, H, W);
  return c3d_check_last(cudaGetLastError());
}

// Tile = 16 x 8 pixels (73 KB of shared memory at T = 3: two CTAs per SM); C3D_STEM_BWD_TILE=32 selects the
// 32 x 8 tile (133 KB, one CTA per SM) the kernel was first written for.
template <int T>
static int launch_stem_bwd(const StemFrames& fr, const float* dpre, const float* ys, const float* bnp, const float* coef,
                           const float* wxy, const float* wt, float* dwxy, float* dwt, float* dperc, int B, int H, int W,
                           cudaStream_t st) {
  const char* v = getenv("C3D_STEM_BWD_TILE");
  if (v && atoi(v) == 32)
    return launch_stem_bwd_tile<T, 32, 8>(fr, dpre, ys, bnp, coef, wxy, wt, dwxy, dwt, dperc, B, H, W, st);
  return launch_stem_bwd_tile<T, 16, 8>(fr, dpre, ys, bnp, coef, wxy, wt, dwxy, dwt, dperc, B, H, W, st);
}

extern "C" int c3d_stem_bwd(const float* const* frame_ptr, const long long* stride_n, const long long* stride_c,
                            const float* d_pre, const float* y_raw, const float* bnp, const float* coef,
                            const float* w_xy, const float* w_t, float* dw_xy, float* dw_t, float* dperception, int B,
                            int T, int H, int W, void* stream_) {
  if (!d_pre || !y_raw || !bnp || !coef || !w_xy || !w_t || !dw_xy || !dw_t || B <= 0 || H <= 0 || W <= 0) return C3D_ERR_ARG;
  StemFrames fr;
  if (int e = stem_frames(fr, frame_ptr, stride_n, stride_c, T)) return e;
  cudaStream_t st = (cudaStream_t)stream_;
  switch (T) {
    case 3: return launch_stem_bwd<3>(fr, d_pre, y_raw, bnp, coef, w_xy, w_t, dw_xy, dw_t, dperception, B, H, W, st);
    case 4: return launch_stem_bwd<4>(fr, d_pre, y_raw, bnp, coef, w_xy, w_t, dw_xy, dw_t, dperception, B, H, W, st);
    case 5: return launch_stem_bwd<5>(fr, d_pre, y_raw, bnp, coef, w_xy, w_t, dw_xy, dw_t, dperception, B, H, W, st);
    default: return C3D_ERR_ARG;
  }
}
