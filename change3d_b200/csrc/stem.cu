// X3D stem (model/x3d.py:23-106) fused with Encoder.forward's frame assembly
// (model/trainer.py:154-162): [pre, P perception frames, post] -> Conv3d 1x3x3 (3->24, pad 0,1,1)
// -> depthwise Conv3d 5x1x1 (pad 2,0,0) -> raw NDHWC output + BatchNorm statistics.
// pytorchvideo's Conv2plus1d runs the module stored as conv_t (the spatial conv) first.
#include <stdlib.h>

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

// Stem forward, second form (default; C3D_STEM_FWD=0 selects the kernel above).  ncu on the quad-per-thread kernel:
// 332 M warp instructions, shared-memory pipe 56 %, 32 sectors per store request, 40 shuffles per thread and channel
// quad for the statistics.  Here a thread owns ONE CHANNEL (27 + 5 weights in registers) and a strip of four
// horizontally adjacent pixels: the six input columns of a (frame, ci, kh) row arrive as three LDS.64 and feed 12 FMA;
// 24 consecutive lanes store the 24 channels of a pixel (whole sectors); the BatchNorm sums stay in per-thread
// registers (fp32 within a tile, fp64 across the tiles of a persistent CTA) and are reduced once per launch.
template <int T, int NSLOT>
__global__ void __launch_bounds__(NSLOT * STEM_C, T <= 4 ? 4 : 3) stem_fwd_pc_kernel(const StemFrames fr, const float* __restrict__ wxy,
                                                                      const float* __restrict__ wt, float* __restrict__ Y,
                                                                      double* __restrict__ stats, int B, int H, int W) {
  constexpr int NT = NSLOT * STEM_C;
  constexpr int PW = STEM_TW + 2, PH = STEM_TH + 2, NPIX = PW * PH, PWP = STEM_TW + 4;
  constexpr int NSTRIP = (STEM_TW / 4) * STEM_TH;
  __shared__ __align__(16) float patch[T * 3 * PH * PWP];
  __shared__ double s_red[2 * STEM_C];
  const int tid = threadIdx.x;
  const int c = tid % STEM_C, slot = tid / STEM_C;
  float wr[27], wtr[5];
#pragma unroll
  for (int j = 0; j < 27; ++j) wr[j] = __ldg(wxy + c * 27 + j);
#pragma unroll
  for (int k = 0; k < 5; ++k) wtr[k] = __ldg(wt + c * 5 + k);
  if (tid < 2 * STEM_C) s_red[tid] = 0.0;
  double dsum = 0.0, dsq = 0.0;
  const long long fstride = (long long)H * W * STEM_C;
  const int tiles_x = (W + STEM_TW - 1) / STEM_TW, tiles_y = (H + STEM_TH - 1) / STEM_TH;
  const int ntiles = tiles_x * tiles_y * B;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int n = tile / (tiles_x * tiles_y);
    const int trem = tile - n * tiles_x * tiles_y;
    const int h0 = (trem / tiles_x) * STEM_TH, w0 = (trem % tiles_x) * STEM_TW;
    __syncthreads();
    // all loads of a frame's patch are issued (clamped, always valid addresses) before the first store
    constexpr int FILL_IT = (3 * NPIX + NT - 1) / NT;
#pragma unroll
    for (int f = 0; f < T; ++f) {
      const float* fp = fr.p[f] + n * fr.sn[f];
      const long long fsc = fr.sc[f];
      float v[FILL_IT];
#pragma unroll
      for (int k = 0; k < FILL_IT; ++k) {
        const int i = tid + k * NT;
        const int ic = i < 3 * NPIX ? i : 0;
        const int x = ic % PW, y = (ic / PW) % PH, ci = ic / NPIX;
        const int h = h0 - 1 + y, w = w0 - 1 + x;
        const int hc = h < 0 ? 0 : h >= H ? H - 1 : h, wc = w < 0 ? 0 : w >= W ? W - 1 : w;
        const float t = __ldg(fp + ci * fsc + (long long)hc * W + wc);
        v[k] = (hc == h && wc == w) ? t : 0.f;
      }
#pragma unroll
      for (int k = 0; k < FILL_IT; ++k) {
        const int i = tid + k * NT;
        if (i < 3 * NPIX) {
          const int x = i % PW, y = (i / PW) % PH, ci = i / NPIX;
          patch[((f * 3 + ci) * PH + y) * PWP + x] = v[k];
        }
      }
      asm volatile("" ::: "memory");      // one frame's batch at a time: keeps FILL_IT values live, not T * FILL_IT
    }
    __syncthreads();
    float psum = 0.f, psq = 0.f;
    float* ybase = Y + (long long)n * T * fstride + c;
    for (int sp = slot; sp < NSTRIP; sp += NSLOT) {
      const int ly = sp / (STEM_TW / 4), x0 = 4 * (sp - ly * (STEM_TW / 4));
      const int h = h0 + ly, w = w0 + x0;
      if (h >= H || w >= W) continue;
      float sacc[T][4];
#pragma unroll
      for (int f = 0; f < T; ++f)
#pragma unroll
        for (int i = 0; i < 4; ++i) sacc[f][i] = 0.f;
      const float* prow = patch + ly * PWP + x0;
#pragma unroll
      for (int f = 0; f < T; ++f)
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const float* r = prow + ((f * 3 + ci) * PH + kh) * PWP;
            const float2 a = *reinterpret_cast<const float2*>(r), b = *reinterpret_cast<const float2*>(r + 2),
                         e = *reinterpret_cast<const float2*>(r + 4);
            const float in[6] = {a.x, a.y, b.x, b.y, e.x, e.y};
            const int j = ci * 9 + kh * 3;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
#pragma unroll
              for (int i = 0; i < 4; ++i) sacc[f][i] = fmaf(in[i + kw], wr[j + kw], sacc[f][i]);
          }
      float* yp = ybase + ((long long)h * W + w) * STEM_C;
#pragma unroll
      for (int to = 0; to < T; ++to) {
        float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int f = 0; f < T; ++f) {
          const int tap = f - to + 2;
          if (tap < 0 || tap > 4) continue;
#pragma unroll
          for (int i = 0; i < 4; ++i) o[i] = fmaf(wtr[tap], sacc[f][i], o[i]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (w + i < W) {
            yp[to * fstride + i * STEM_C] = o[i];
            psum += o[i];
            psq = fmaf(o[i], o[i], psq);
          }
        }
      }
    }
    dsum += (double)psum;
    dsq += (double)psq;
  }
  if (stats) {
    atomicAdd(&s_red[c], dsum);
    atomicAdd(&s_red[STEM_C + c], dsq);
    __syncthreads();
    if (tid < 2 * STEM_C) atomicAdd(stats + tid, s_red[tid]);
  }
}

template <int T>
static int launch_stem_fwd_pc(const StemFrames& fr, const float* wxy, const float* wt, float* Y, double* stats, int B, int H,
                              int W, cudaStream_t st) {
  constexpr int NSLOT = 8;
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stem_fwd_pc_kernel<T, NSLOT>, NSLOT * STEM_C, 0);
  if (per_sm < 1) per_sm = 1;
  const int ntiles = ((W + STEM_TW - 1) / STEM_TW) * ((H + STEM_TH - 1) / STEM_TH) * B;
  const int grid = ntiles < sms * per_sm ? ntiles : sms * per_sm;
  stem_fwd_pc_kernel<T, NSLOT><<<grid, NSLOT * STEM_C, 0, st>>>(fr, wxy, wt, Y, stats, B, H, W);
  return c3d_check_last(cudaGetLastError());
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
  const char* m = getenv("C3D_STEM_FWD");
  if (!m || atoi(m) != 0) {
    switch (T) {
      case 3: return launch_stem_fwd_pc<3>(fr, w_xy, w_t, Y, stats, B, H, W, st);
      case 4: return launch_stem_fwd_pc<4>(fr, w_xy, w_t, Y, stats, B, H, W, st);
      case 5: return launch_stem_fwd_pc<5>(fr, w_xy, w_t, Y, stats, B, H, W, st);
      default: return C3D_ERR_ARG;
    }
  }
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
                                                       int W, int relu_mask) {
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
            float4 d = ldg4(dpre + off);
            const float4 yv = ldg4(ys + off);
            if (relu_mask) {   // mask out > 0 recomputed from y (f4bn, the forward's expression)
              const float4 o = f4bn(yv, mean, scale, ldg4(bnp + 3 * STEM_C + ca));
              d.x = o.x > 0.f ? d.x : 0.f; d.y = o.y > 0.f ? d.y : 0.f; d.z = o.z > 0.f ? d.z : 0.f; d.w = o.w > 0.f ? d.w : 0.f;
            }
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
                                int relu_mask, cudaStream_t st) {
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
  stem_bwd_kernel<T, BTW, BTH><<<grid, 256, smem, st>>>(fr, dpre, ys, bnp, coef, wxy, wt, dwxy, dwt, dperc, B, H, W, relu_mask);
  return c3d_check_last(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// Stem backward, second form (default).  ncu on the kernel above (profiles/r02_summary.md section 2.5): shared-memory
// pipe at 81 % of its peak, short-scoreboard + barrier stalls, 754 M warp instructions -- one LDS per 2-3 FMA in the
// spatial-conv recompute and in the dwxy pass, three roles run one after the other on 252 / 162 / 128 of 256 threads.
// Here a thread owns ONE CHANNEL and walks pairs of horizontally adjacent pixels:
//   * its 27 spatial weights, 27 dwxy accumulators, 5 temporal weights and 5 dwt accumulators live in registers for
//     the whole launch (no weight loads, no weight-gradient traffic inside the loop);
//   * the 4 input columns a pixel pair needs per (frame, ci, kh) arrive as two LDS.64 and feed 12 FMA: the recompute
//     of s AND the dwxy accumulation use the same values (54 LDS.64 per 324 FMA instead of 135 LDS per 324);
//   * d_pre / y are read as 24 consecutive floats per pixel by 24 consecutive lanes (whole sectors);
//   * ds is written to shared memory for the perception frames only, tile + halo; the halo ring is computed by light
//     work items (BN backward + temporal transpose, no spatial work); the perception-frame gradient is the last
//     phase, one thread per pixel pair (4 ds columns x 9 weight quads per (kh, q): 13 LDS.128 per 72 FMA).
// ------------------------------------------------------------------------------------------------
template <int T, int BTW, int BTH, int NSLOT>
__global__ void __launch_bounds__(NSLOT * STEM_C, 2)
stem_bwd_pc_kernel(const StemFrames fr, const float* __restrict__ dpre, const float* __restrict__ ys,
                   const float* __restrict__ bnp, const float* __restrict__ coef, const float* __restrict__ wxy,
                   const float* __restrict__ wt, float* __restrict__ dwxy, float* __restrict__ dwt,
                   float* __restrict__ dperc, int B, int H, int W, int relu_mask, int l2pf) {
  constexpr int P = T - 2;
  constexpr int NT = NSLOT * STEM_C;
  constexpr int PW = BTW + 2, PH = BTH + 2, NPIX = PW * PH;
  constexpr int PWP = BTW + 4;                       // even patch row stride: LDS.64 at even columns
  constexpr int NPP = (BTW / 2) * BTH;               // interior pixel pairs
  constexpr int NHALO = 2 * PW + 2 * BTH;            // halo ring pixels
  static_assert(NPP <= NT, "one thread per pixel pair in the perception-frame pass");
  extern __shared__ __align__(16) float sm[];
  float* patch = sm;                                 // [T][3][PH][PWP]
  float* dsp = patch + T * 3 * PH * PWP;             // [P][NPIX][STEM_DSLD]
  float* s_wxy = dsp + P * NPIX * STEM_DSLD;         // [27][24]
  float* s_red = s_wxy + 27 * STEM_C;                // [32][24]: dwxy rows 0..26, dwt rows 27..31
  const int tid = threadIdx.x;
  const int c = tid % STEM_C, slot = tid / STEM_C;
  for (int i = tid; i < 27 * STEM_C; i += NT) { int k = i / STEM_C, cc = i - k * STEM_C; s_wxy[i] = __ldg(wxy + cc * 27 + k); }
  for (int i = tid; i < 32 * STEM_C; i += NT) s_red[i] = 0.f;
  float wr[27], acc[27], wtr[5], dwt_acc[5];
#pragma unroll
  for (int j = 0; j < 27; ++j) { wr[j] = __ldg(wxy + c * 27 + j); acc[j] = 0.f; }
#pragma unroll
  for (int k = 0; k < 5; ++k) { wtr[k] = __ldg(wt + c * 5 + k); dwt_acc[k] = 0.f; }
  const float mean = __ldg(bnp + c), rstd = __ldg(bnp + STEM_C + c), scale = __ldg(bnp + 2 * STEM_C + c);
  const float c1 = __ldg(coef + c), c2 = __ldg(coef + STEM_C + c);
  // relu_mask: `dpre` is the gradient w.r.t. the stem's ReLU output; the mask out > 0 is recomputed from y with the
  // forward's own expression (f4bn in c3d_common.cuh), so the masked gradient is never written to memory.
  const float beta = __ldg(bnp + 3 * STEM_C + c);
  const float bT = -scale * rstd * c2, bC = -scale * c1;
  const long long fstride = (long long)H * W * STEM_C;          // one frame of d_pre / y

  const int tiles_x = (W + BTW - 1) / BTW, tiles_y = (H + BTH - 1) / BTH;
  const int ntiles = tiles_x * tiles_y * B;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int n = tile / (tiles_x * tiles_y);
    const int trem = tile - n * tiles_x * tiles_y;
    const int h0 = (trem / tiles_x) * BTH, w0 = (trem % tiles_x) * BTW;
    const float* dbase = dpre + (long long)n * T * fstride + c;
    const float* ybase = ys + (long long)n * T * fstride + c;
    __syncthreads();                                            // previous tile's perception pass is done
    // BN backward of one element (+ the ReLU mask recomputed with the forward's expression when relu_mask):
    // scale * (d - c1 - (y - mean) * rstd * c2) = bA * d + bT * (y - mean) + bC
    auto bn_bwd = [&](float d, float yv) -> float {
      const float t = yv - mean;
      if (relu_mask && !(fmaf(t, scale, beta) > 0.f)) d = 0.f;
      return fmaf(scale, d, fmaf(bT, t, bC));
    };
    float nd[T][2], ny[T][2];                                   // d_pre / y of the NEXT pixel pair (software prefetch)
    auto fetch = [&](int pp) {
      const int ly = pp / (BTW / 2), x0 = 2 * (pp - ly * (BTW / 2));
      const int h = h0 + ly, w = w0 + x0;
      // unconditional loads from clamped (always valid) addresses: pixels outside the image are zeroed through
      // v0 / v1 where dy is formed, so the loop carries no divergent branches
      const int hc = h < H ? h : H - 1, wc = w < W ? w : W - 1;
      const int o1 = (w + 1 < W) ? STEM_C : 0;
      const long long off = ((long long)hc * W + wc) * STEM_C;
#pragma unroll
      for (int t = 0; t < T; ++t) {
        nd[t][0] = __ldg(dbase + t * fstride + off);
        ny[t][0] = __ldg(ybase + t * fstride + off);
        nd[t][1] = __ldg(dbase + t * fstride + off + o1);
        ny[t][1] = __ldg(ybase + t * fstride + off + o1);
      }
    };
    constexpr bool PREF = (T <= 4);                             // T = 5: the 20 prefetch registers would spill
    if (PREF && slot < NPP) fetch(slot);                        // in flight across the patch fill and the halo ring
    // Patch fill and halo ring: the loads of a whole batch are issued (from clamped, always valid addresses) before the
    // first value is used -- with one dependent load -> store chain per loop iteration this phase took 41 % of the
    // kernel's time for 17 % of its instructions (ncu source view, profiles/r02_summary.md section 2.5).
    constexpr int FILL_IT = (3 * NPIX + NT - 1) / NT;
#pragma unroll
    for (int f = 0; f < T; ++f) {                               // static f: the frame table stays in the parameter bank
      const float* fp = fr.p[f] + n * fr.sn[f];
      const long long fsc = fr.sc[f];
      float v[FILL_IT];
#pragma unroll
      for (int k = 0; k < FILL_IT; ++k) {
        const int i = tid + k * NT;
        const int ic = i < 3 * NPIX ? i : 0;
        const int x = ic % PW, y = (ic / PW) % PH, ci = ic / NPIX;
        const int h = h0 - 1 + y, w = w0 - 1 + x;
        const int hc = h < 0 ? 0 : h >= H ? H - 1 : h, wc = w < 0 ? 0 : w >= W ? W - 1 : w;
        const float t = __ldg(fp + ci * fsc + (long long)hc * W + wc);
        v[k] = (hc == h && wc == w) ? t : 0.f;
      }
#pragma unroll
      for (int k = 0; k < FILL_IT; ++k) {
        const int i = tid + k * NT;
        if (i < 3 * NPIX) {
          const int x = i % PW, y = (i / PW) % PH, ci = i / NPIX;
          patch[((f * 3 + ci) * PH + y) * PWP + x] = v[k];
        }
      }
      asm volatile("" ::: "memory");      // one frame's batch at a time: keeps FILL_IT values live, not T * FILL_IT
    }
    if (dperc) {
      // halo ring: ds of the perception frames only, HB pixels per batch
      constexpr int HB = 3;
      constexpr int HALO_IT = (NHALO + NSLOT - 1) / NSLOT;
#pragma unroll 1
      for (int k0 = 0; k0 < HALO_IT; k0 += HB) {
        float hd[HB][T], hy[HB][T];
        int hpos[HB];
        bool hin[HB];
#pragma unroll
        for (int u = 0; u < HB; ++u) {
          const int hp = slot + (k0 + u) * NSLOT;
          const int hq = hp < NHALO ? hp : 0;
          int y, x;
          if (hq < PW) { y = 0; x = hq; }
          else if (hq < 2 * PW) { y = PH - 1; x = hq - PW; }
          else { const int r = hq - 2 * PW; y = 1 + (r >> 1); x = (r & 1) ? PW - 1 : 0; }
          const int h = h0 - 1 + y, w = w0 - 1 + x;
          const int hc = h < 0 ? 0 : h >= H ? H - 1 : h, wc = w < 0 ? 0 : w >= W ? W - 1 : w;
          hin[u] = (hc == h && wc == w);
          hpos[u] = (k0 + u < HALO_IT && hp < NHALO) ? y * PW + x : -1;
          const long long off = ((long long)hc * W + wc) * STEM_C;
#pragma unroll
          for (int t = 0; t < T; ++t) {
            hd[u][t] = __ldg(dbase + t * fstride + off);
            hy[u][t] = __ldg(ybase + t * fstride + off);
          }
        }
#pragma unroll
        for (int u = 0; u < HB; ++u) {
          if (hpos[u] < 0) continue;
          float dy[T];
#pragma unroll
          for (int t = 0; t < T; ++t) dy[t] = hin[u] ? bn_bwd(hd[u][t], hy[u][t]) : 0.f;
#pragma unroll
          for (int f = 1; f <= P; ++f) {
            float o = 0.f;
#pragma unroll
            for (int t = 0; t < T; ++t) {
              const int tap = f - t + 2;
              if (tap < 0 || tap > 4) continue;
              o = fmaf(wtr[tap], dy[t], o);
            }
            dsp[((f - 1) * NPIX + hpos[u]) * STEM_DSLD + c] = o;
          }
        }
      }
    }
    __syncthreads();                                            // patch complete
    if (l2pf && tile + (int)gridDim.x < ntiles) {
      // L2 prefetch of the NEXT tile's d_pre / y rows (tile + halo): its halo ring and its first pixel pairs then pay an
      // L2 round trip instead of a DRAM one
      const int nt = tile + (int)gridDim.x;
      const int nn = nt / (tiles_x * tiles_y);
      const int nrem = nt - nn * tiles_x * tiles_y;
      const int nh0 = (nrem / tiles_x) * BTH, nw0 = (nrem % tiles_x) * BTW;
      constexpr int LPR = (PW * STEM_C * 4 + 127) / 128 + 1;    // 128-byte lines per tile row (unaligned start)
      const int ws = nw0 > 0 ? nw0 - 1 : 0;
      const long long row_end = (long long)W * STEM_C;          // floats per image row
      for (int i = tid; i < 2 * T * PH * LPR; i += NT) {
        const int line = i % LPR, r = i / LPR, y = r % PH, tt = (r / PH) % T, which = r / (PH * T);
        const int h = nh0 - 1 + y;
        const long long col = (long long)ws * STEM_C + line * 32;
        if (h >= 0 && h < H && col < row_end) {
          const float* pbase = which ? ys : dpre;
          const float* q = pbase + (((long long)nn * T + tt) * H + h) * row_end + col;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
        }
      }
    }
    for (int pp = slot; pp < NPP; pp += NSLOT) {
      const int ly = pp / (BTW / 2), x0 = 2 * (pp - ly * (BTW / 2));
      const int h = h0 + ly, w = w0 + x0;
      const bool v0 = (h < H && w < W), v1 = (h < H && w + 1 < W);
      float dy[T][2], ds[T][2], s[T][2];
      if (!PREF) fetch(pp);
#pragma unroll
      for (int t = 0; t < T; ++t) {
        dy[t][0] = v0 ? bn_bwd(nd[t][0], ny[t][0]) : 0.f;
        dy[t][1] = v1 ? bn_bwd(nd[t][1], ny[t][1]) : 0.f;
      }
      if (PREF && pp + NSLOT < NPP) fetch(pp + NSLOT);          // next pair's loads fly during this pair's 360 FMA
#pragma unroll
      for (int f = 0; f < T; ++f) {
        float o0 = 0.f, o1 = 0.f;
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const int tap = f - t + 2;
          if (tap < 0 || tap > 4) continue;
          o0 = fmaf(wtr[tap], dy[t][0], o0);
          o1 = fmaf(wtr[tap], dy[t][1], o1);
        }
        ds[f][0] = o0; ds[f][1] = o1;
        s[f][0] = 0.f; s[f][1] = 0.f;
        if (f >= 1 && f <= P) {
          float* q = dsp + ((f - 1) * NPIX + (ly + 1) * PW + x0 + 1) * STEM_DSLD + c;
          q[0] = o0; q[STEM_DSLD] = o1;
        }
      }
      const float* prow = patch + ly * PWP + x0;
#pragma unroll
      for (int f = 0; f < T; ++f)
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const float* r = prow + ((f * 3 + ci) * PH + kh) * PWP;
            const float2 a = *reinterpret_cast<const float2*>(r), b = *reinterpret_cast<const float2*>(r + 2);
            const int j = ci * 9 + kh * 3;
            s[f][0] = fmaf(a.x, wr[j], s[f][0]);     s[f][1] = fmaf(a.y, wr[j], s[f][1]);
            s[f][0] = fmaf(a.y, wr[j + 1], s[f][0]); s[f][1] = fmaf(b.x, wr[j + 1], s[f][1]);
            s[f][0] = fmaf(b.x, wr[j + 2], s[f][0]); s[f][1] = fmaf(b.y, wr[j + 2], s[f][1]);
            acc[j] = fmaf(ds[f][0], a.x, acc[j]);         acc[j] = fmaf(ds[f][1], a.y, acc[j]);
            acc[j + 1] = fmaf(ds[f][0], a.y, acc[j + 1]); acc[j + 1] = fmaf(ds[f][1], b.x, acc[j + 1]);
            acc[j + 2] = fmaf(ds[f][0], b.x, acc[j + 2]); acc[j + 2] = fmaf(ds[f][1], b.y, acc[j + 2]);
          }
#pragma unroll
      for (int t = 0; t < T; ++t)
#pragma unroll
        for (int f = 0; f < T; ++f) {
          const int tap = f - t + 2;
          if (tap < 0 || tap > 4) continue;
          dwt_acc[tap] = fmaf(dy[t][0], s[f][0], dwt_acc[tap]);
          dwt_acc[tap] = fmaf(dy[t][1], s[f][1], dwt_acc[tap]);
        }
    }
    if (dperc) {
      __syncthreads();                                          // ds of tile + halo complete
      if (tid < NPP) {
        const int ly = tid / (BTW / 2), x0 = 2 * (tid - ly * (BTW / 2));
        const int h = h0 + ly, w = w0 + x0;
        if (h < H && w < W) {
          const long long HW = (long long)H * W, pix = (long long)h * W + w;
#pragma unroll 1
          for (int f = 1; f <= P; ++f) {
            float g[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
              // source pixel = output pixel - (kh - 1, kw - 1): halo row ly + 2 - kh, halo columns x0 + 2 - kw (+1)
              const float* dp = dsp + ((f - 1) * NPIX + (ly + 2 - kh) * PW + x0) * STEM_DSLD;
#pragma unroll 2
              for (int q = 0; q < STEM_C / 4; ++q) {
                float4 d[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) d[k] = *reinterpret_cast<const float4*>(dp + k * STEM_DSLD + 4 * q);
#pragma unroll
                for (int kw = 0; kw < 3; ++kw)
#pragma unroll
                  for (int ci = 0; ci < 3; ++ci) {
                    const float4 wv = *reinterpret_cast<const float4*>(s_wxy + (ci * 9 + kh * 3 + kw) * STEM_C + 4 * q);
                    const float4 da = d[2 - kw], db = d[3 - kw];
                    g[0][ci] = fmaf(da.x, wv.x, fmaf(da.y, wv.y, fmaf(da.z, wv.z, fmaf(da.w, wv.w, g[0][ci]))));
                    g[1][ci] = fmaf(db.x, wv.x, fmaf(db.y, wv.y, fmaf(db.z, wv.z, fmaf(db.w, wv.w, g[1][ci]))));
                  }
              }
            }
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
              atomicAdd(dperc + (ci * P + (f - 1)) * HW + pix, g[0][ci]);
              if (w + 1 < W) atomicAdd(dperc + (ci * P + (f - 1)) * HW + pix + 1, g[1][ci]);
            }
          }
        }
      }
    }
  }
  // flush: slots -> shared memory -> global
#pragma unroll
  for (int j = 0; j < 27; ++j) atomicAdd(&s_red[j * STEM_C + c], acc[j]);
#pragma unroll
  for (int k = 0; k < 5; ++k) atomicAdd(&s_red[(27 + k) * STEM_C + c], dwt_acc[k]);
  __syncthreads();
  for (int i = tid; i < 27 * STEM_C; i += NT) { int k = i / STEM_C, cc = i - k * STEM_C; atomicAdd(dwxy + cc * 27 + k, s_red[i]); }
  for (int i = tid; i < 5 * STEM_C; i += NT) { int k = i / STEM_C, cc = i - k * STEM_C; atomicAdd(dwt + cc * 5 + k, s_red[(27 + k) * STEM_C + cc]); }
}

template <int T, int BTW, int BTH, int NSLOT>
static int launch_stem_bwd_pc(const StemFrames& fr, const float* dpre, const float* ys, const float* bnp, const float* coef,
                              const float* wxy, const float* wt, float* dwxy, float* dwt, float* dperc, int B, int H, int W,
                              int relu_mask, cudaStream_t st) {
  constexpr int P = T - 2, NPIX = (BTW + 2) * (BTH + 2);
  const size_t smem = (size_t)(T * 3 * (BTH + 2) * (BTW + 4) + P * NPIX * STEM_DSLD + 27 * STEM_C + 32 * STEM_C) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(stem_bwd_pc_kernel<T, BTW, BTH, NSLOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return C3D_ERR_SMEM;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int ntiles = ((W + BTW - 1) / BTW) * ((H + BTH - 1) / BTH) * B;
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  per_sm = per_sm < 1 ? 1 : per_sm > 2 ? 2 : per_sm;
  const int grid = ntiles < sms * per_sm ? ntiles : sms * per_sm;
  static int l2pf = -1;             // C3D_STEM_PF=1: L2 prefetch of the next tile's rows (default off: measured 4 % slower)
  if (l2pf < 0) { const char* v = getenv("C3D_STEM_PF"); l2pf = v ? atoi(v) : 0; }
  stem_bwd_pc_kernel<T, BTW, BTH, NSLOT><<<grid, NSLOT * STEM_C, smem, st>>>(fr, dpre, ys, bnp, coef, wxy, wt, dwxy, dwt, dperc,
                                                                              B, H, W, relu_mask, l2pf);
  return c3d_check_last(cudaGetLastError());
}

// Tile = 16 x 8 pixels (73 KB of shared memory at T = 3: two CTAs per SM); C3D_STEM_BWD_TILE=32 selects the
// 32 x 8 tile (133 KB, one CTA per SM) the kernel was first written for.
template <int T>
static int launch_stem_bwd(const StemFrames& fr, const float* dpre, const float* ys, const float* bnp, const float* coef,
                           const float* wxy, const float* wt, float* dwxy, float* dwt, float* dperc, int B, int H, int W,
                           int relu_mask, cudaStream_t st) {
  // C3D_STEM_BWD=0: the quad-per-thread kernel above (round 1; kept as a second implementation for the tests)
  const char* m = getenv("C3D_STEM_BWD");
  if (!m || atoi(m) != 0) {
    if constexpr (T <= 4) return launch_stem_bwd_pc<T, 32, 8, 10>(fr, dpre, ys, bnp, coef, wxy, wt, dwxy, dwt, dperc, B, H, W, relu_mask, st);
    else return launch_stem_bwd_pc<T, 16, 8, 10>(fr, dpre, ys, bnp, coef, wxy, wt, dwxy, dwt, dperc, B, H, W, relu_mask, st);
  }
  const char* v = getenv("C3D_STEM_BWD_TILE");
  if (v && atoi(v) == 32)
    return launch_stem_bwd_tile<T, 32, 8>(fr, dpre, ys, bnp, coef, wxy, wt, dwxy, dwt, dperc, B, H, W, relu_mask, st);
  return launch_stem_bwd_tile<T, 16, 8>(fr, dpre, ys, bnp, coef, wxy, wt, dwxy, dwt, dperc, B, H, W, relu_mask, st);
}

extern "C" int c3d_stem_bwd(const float* const* frame_ptr, const long long* stride_n, const long long* stride_c,
                            const float* d_pre, const float* y_raw, const float* bnp, const float* coef,
                            const float* w_xy, const float* w_t, float* dw_xy, float* dw_t, float* dperception, int B,
                            int T, int H, int W, int relu_mask, void* stream_) {
  if (!d_pre || !y_raw || !bnp || !coef || !w_xy || !w_t || !dw_xy || !dw_t || B <= 0 || H <= 0 || W <= 0) return C3D_ERR_ARG;
  StemFrames fr;
  if (int e = stem_frames(fr, frame_ptr, stride_n, stride_c, T)) return e;
  cudaStream_t st = (cudaStream_t)stream_;
  switch (T) {
    case 3: return launch_stem_bwd<3>(fr, d_pre, y_raw, bnp, coef, w_xy, w_t, dw_xy, dw_t, dperception, B, H, W, relu_mask, st);
    case 4: return launch_stem_bwd<4>(fr, d_pre, y_raw, bnp, coef, w_xy, w_t, dw_xy, dw_t, dperception, B, H, W, relu_mask, st);
    case 5: return launch_stem_bwd<5>(fr, d_pre, y_raw, bnp, coef, w_xy, w_t, dw_xy, dw_t, dperception, B, H, W, relu_mask, st);
    default: return C3D_ERR_ARG;
  }
}
