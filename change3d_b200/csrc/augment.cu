// Input pipeline on the GPU: the reference's per-sample CPU transform chain
//   normalize -> scale -> random_crop_resize -> random_flip -> random_exchange -> to_tensor
// (data/transforms.py:82-154 and 166-206 for BCD; the SCD / BDA classes :210-612 differ only in the label handling) on
// raw uint8 HWC pairs, as ONE gather kernel per batch that writes the float NCHW `pre` / `post` tensors and the label
// tensor the training step consumes.  The random decisions stay on the host (input_pipeline.draw_params draws them from
// Python's `random` in the reference's order, so a seeded run makes the same choices); the kernel is deterministic.
//
// Geometry restated from OpenCV (cv2.resize as called by the reference):
//   INTER_LINEAR, float32 image: fx = (float)((dx + 0.5) * (double)(src_w / dst_w) - 0.5); sx = floor(fx); fx -= sx;
//                                sx < 0 -> (0, 0);  sx >= src_w - 1 -> (src_w - 1, 0);  horizontal pass, then vertical
//   INTER_NEAREST (labels):      sx = min(floor(dx * (1.0 / ((double)dst_w / src_w))), src_w - 1)
//   cv2.flip(img, 0) reverses rows, cv2.flip(img, 1) reverses columns.
// The reference normalises BEFORE resizing, so the four taps are normalised and then interpolated, in that order.
#include <stdint.h>

#include "c3d_common.cuh"
#include "../../include/change3d_b200.h"

namespace {

struct AugGeo {
  int B, Hs, Ws, H, W, L;      // source size, output size, label channels
  int label_mode;              // 0: BCD, label -> ceil(l / 255) as float; 1: class ids as int64
  float std, mean;             // (v / 255 - mean) / std
};

__device__ __forceinline__ void lin_coord(int d, int src, int dst, int off, int& s0, int& s1, float& f) {
  // coordinate inside a crop of `src` pixels starting at `off`
  const double scale = 1.0 / ((double)dst / (double)src);      // cv2: scale_x = 1. / inv_scale_x
  float fx = (float)(((double)d + 0.5) * scale - 0.5);
  int sx = (int)floorf(fx);
  fx -= (float)sx;
  if (sx < 0) { sx = 0; fx = 0.f; }
  if (sx >= src - 1) { sx = src - 1; fx = 0.f; }
  s0 = off + sx;
  s1 = off + (sx + 1 < src ? sx + 1 : sx);
  f = fx;
}

__device__ __forceinline__ int nn_coord(int d, int src, int dst, int off) {
  const double ifx = 1.0 / ((double)dst / (double)src);
  int sx = (int)floor((double)d * ifx);
  if (sx > src - 1) sx = src - 1;
  return off + sx;
}

// params[b] = {do_crop, x1, y1, flip_rows, flip_cols, exchange_images, exchange_labels01, unused}
template <typename SrcT>
__global__ void __launch_bounds__(256) augment_kernel(const SrcT* __restrict__ img, const uint8_t* __restrict__ label,
                                                      const int* __restrict__ params, AugGeo G, float* __restrict__ pre,
                                                      float* __restrict__ post, void* __restrict__ label_out) {
  const long long n = (long long)G.B * G.H * G.W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % G.W);
    const int y = (int)((i / G.W) % G.H);
    const int b = (int)(i / ((long long)G.W * G.H));
    const int* p = params + b * 8;
    const bool crop = p[0] != 0;
    const int x1 = crop ? p[1] : 0, y1 = crop ? p[2] : 0;
    const int yf = p[3] ? G.H - 1 - y : y, xf = p[4] ? G.W - 1 - x : x;
    const bool resample = crop || G.Hs != G.H || G.Ws != G.W;
    const int cw = G.Ws - 2 * x1, ch = G.Hs - 2 * y1;
    const SrcT* src = img + (long long)b * G.Hs * G.Ws * 6;
    float v[6];
    if (!resample) {
      const SrcT* q = src + ((long long)yf * G.Ws + xf) * 6;
#pragma unroll
      for (int c = 0; c < 6; ++c)
        v[c] = sizeof(SrcT) == 1 ? ((float)q[c] / 255.0f - G.mean) / G.std : (float)q[c];
    } else {
      int sx0, sx1, sy0, sy1;
      float fx, fy;
      lin_coord(xf, cw, G.W, x1, sx0, sx1, fx);
      lin_coord(yf, ch, G.H, y1, sy0, sy1, fy);
      const SrcT* q00 = src + ((long long)sy0 * G.Ws + sx0) * 6;
      const SrcT* q01 = src + ((long long)sy0 * G.Ws + sx1) * 6;
      const SrcT* q10 = src + ((long long)sy1 * G.Ws + sx0) * 6;
      const SrcT* q11 = src + ((long long)sy1 * G.Ws + sx1) * 6;
      const float ax0 = 1.f - fx, ay0 = 1.f - fy;
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        float a, bq, cq, d;
        if (sizeof(SrcT) == 1) {
          a = ((float)q00[c] / 255.0f - G.mean) / G.std; bq = ((float)q01[c] / 255.0f - G.mean) / G.std;
          cq = ((float)q10[c] / 255.0f - G.mean) / G.std; d = ((float)q11[c] / 255.0f - G.mean) / G.std;
        } else {
          a = (float)q00[c]; bq = (float)q01[c]; cq = (float)q10[c]; d = (float)q11[c];
        }
        // separate multiplies and adds, like the two passes of the reference's resize (no contraction into FMAs)
        const float r0 = __fadd_rn(__fmul_rn(a, ax0), __fmul_rn(bq, fx));
        const float r1 = __fadd_rn(__fmul_rn(cq, ax0), __fmul_rn(d, fx));
        v[c] = __fadd_rn(__fmul_rn(r0, ay0), __fmul_rn(r1, fy));
      }
    }
    const int sw = p[5] ? 3 : 0;       // exchange: pre <- channels 3..5, post <- channels 0..2
    const long long plane = (long long)G.H * G.W;
    float* po = pre + (long long)b * 3 * plane + (long long)y * G.W + x;
    float* qo = post + (long long)b * 3 * plane + (long long)y * G.W + x;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      po[c * plane] = v[sw + c];
      qo[c * plane] = v[3 - sw + c];
    }
    if (label) {
      const int lx = resample ? nn_coord(xf, cw, G.W, x1) : xf;
      const int ly = resample ? nn_coord(yf, ch, G.H, y1) : yf;
      const uint8_t* lp = label + (((long long)b * G.Hs + ly) * G.Ws + lx) * G.L;
      for (int c = 0; c < G.L; ++c) {
        int cs = c;
        if (p[6] && c < 2) cs = 1 - c;                         // SCD exchange swaps the two class maps
        const uint8_t l = lp[cs];
        const long long o = ((long long)b * G.L + c) * plane + (long long)y * G.W + x;
        if (G.label_mode == 0) reinterpret_cast<float*>(label_out)[o] = l ? 1.f : 0.f;      // ceil(l / 255)
        else reinterpret_cast<long long*>(label_out)[o] = (long long)l;
      }
    }
  }
}

}  // namespace

extern "C" int c3d_augment_pairs(const void* img, int img_is_float, const unsigned char* label, const int* params, int B,
                                 int Hs, int Ws, int H, int W, int L, int label_mode, float mean, float std, float* pre,
                                 float* post, void* label_out, void* stream_) {
  if (!img || !params || !pre || !post || B <= 0 || Hs <= 0 || Ws <= 0 || H <= 0 || W <= 0 || std == 0.f)
    return C3D_ERR_ARG;
  if (label && (!label_out || L <= 0 || L > 4 || label_mode < 0 || label_mode > 1)) return C3D_ERR_ARG;
  AugGeo G;
  G.B = B; G.Hs = Hs; G.Ws = Ws; G.H = H; G.W = W; G.L = label ? L : 0; G.label_mode = label_mode;
  G.std = std; G.mean = mean;
  const long long n = (long long)B * H * W;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaStream_t st = (cudaStream_t)stream_;
  if (img_is_float)
    augment_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float*>(img), label, params, G, pre, post, label_out);
  else
    augment_kernel<uint8_t><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const uint8_t*>(img), label, params, G, pre, post, label_out);
  return c3d_check_last(cudaGetLastError());
}
