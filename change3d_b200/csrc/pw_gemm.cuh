// Operand staging shared by the pointwise-GEMM family (forward / dgrad GEMM and wgrad).
// A GEMM row is one NDHWC pixel; staging applies the fused prologue (BN apply, SE gate, Swish,
// BN-backward transform, |a-b|) while the tile moves global -> shared.
#pragma once
#include "c3d_common.cuh"

struct GemmArgs {
  TileSrc a;
  const float* W; long long w_sr, w_so; int Kred;
  int N, Ns; long long M;
  float* Y; long long out_img_stride;
  int epi;
  double* stats;
  const float* E1; long long e1_img_stride;
  const float* E2;
  const float* ebnp; const float* egate;
  float* Y2;
  long long rows_per_sample;
  int NB;        // columns handled per CTA (multiple of 64)
  int nsplit;    // grid.y
  int w_const;   // C3D_GEMM_W_CONSTANT: W may be read before pdl_wait()
};

struct RowMeta {       // per tile row, in shared memory
  long long off;       // element offset of the row in A (dense / sub2 maps)
  long long off2;      // ... in A2
  int img;             // -1 when the row is past M
  int oh, ow;
  int samp;
};

__device__ __forceinline__ void tile_row_meta(const TileSrc& s, long long row0, long long M, int BM, RowMeta* meta) {
  for (int r = threadIdx.x; r < BM; r += blockDim.x) {
    long long row = row0 + r;
    RowMeta m;
    if (row >= M) {
      m.img = -1; m.oh = m.ow = m.samp = 0; m.off = m.off2 = 0;
    } else {
      long long img = row / s.OHW;
      int rem = (int)(row - img * s.OHW);
      m.img = (int)img;
      m.oh = rem / s.OW;
      m.ow = rem - m.oh * s.OW;
      m.samp = (int)(img / s.frames_per_sample);
      int mul = (s.map == MAP_SUB2) ? 2 : 1;
      long long pix = (long long)(m.oh * mul) * s.IW + m.ow * mul;
      m.off = img * s.img_stride + pix * s.ld;
      m.off2 = img * s.img_stride2 + pix * s.ld;
    }
    meta[r] = m;
  }
}

// value of the staged tile at (row r, columns kq..kq+3)
__device__ __forceinline__ float4 tile_fetch(const TileSrc& s, const RowMeta& m, int kq) {
  if (m.img < 0) return f4zero();
  const int c = kq;
  const long long off = m.off + kq, off2 = m.off2 + kq;
  float4 v = ldg4(s.A + off);
  switch (s.mode) {
    case PRO_NONE: break;
    case PRO_BN_RELU: {
      v = f4relu(f4bn(v, ldg4(BNP_MEAN(s.bnp, s.ld) + c), ldg4(BNP_SCALE(s.bnp, s.ld) + c), ldg4(BNP_BETA(s.bnp, s.ld) + c)));
    } break;
    case PRO_BN_GATE_SWISH: {
      v = f4bn(v, ldg4(BNP_MEAN(s.bnp, s.ld) + c), ldg4(BNP_SCALE(s.bnp, s.ld) + c), ldg4(BNP_BETA(s.bnp, s.ld) + c));
      if (s.gate) v = f4mul(v, ldg4(s.gate + (long long)m.samp * s.ld + c));
      v = make_float4(swishf_(v.x), swishf_(v.y), swishf_(v.z), swishf_(v.w));
    } break;
    case PRO_BNBWD: {
      float4 y = ldg4(s.A2 + off2);
      float4 mean = ldg4(BNP_MEAN(s.bnp, s.ld) + c), rstd = ldg4(BNP_RSTD(s.bnp, s.ld) + c);
      float4 scale = ldg4(BNP_SCALE(s.bnp, s.ld) + c);
      float4 c1 = ldg4(s.coef + c), c2 = ldg4(s.coef + s.ld + c);
      v.x = scale.x * (v.x - c1.x - (y.x - mean.x) * rstd.x * c2.x);
      v.y = scale.y * (v.y - c1.y - (y.y - mean.y) * rstd.y * c2.y);
      v.z = scale.z * (v.z - c1.z - (y.z - mean.z) * rstd.z * c2.z);
      v.w = scale.w * (v.w - c1.w - (y.w - mean.w) * rstd.w * c2.w);
    } break;
    case PRO_ABSDIFF: {
      float4 y = ldg4(s.A2 + off2);
      v = make_float4(fabsf(v.x - y.x), fabsf(v.y - y.y), fabsf(v.z - y.z), fabsf(v.w - y.w));
    } break;
    case PRO_MASK_POS: {   // A where A2 > 0 (ReLU backward with the pre-activation in A2)
      float4 y = ldg4(s.A2 + off2);
      v.x = y.x > 0.f ? v.x : 0.f; v.y = y.y > 0.f ? v.y : 0.f; v.z = y.z > 0.f ? v.z : 0.f; v.w = y.w > 0.f ? v.w : 0.f;
    } break;
  }
  return v;
}

// row-major: dst[r * ldS + k]
__device__ __forceinline__ void stage_tile_rowmajor(const TileSrc& s, const RowMeta* meta, int BM, float* dst, int ldS) {
  const int kq4 = s.K >> 2;
  for (int idx = threadIdx.x; idx < BM * kq4; idx += blockDim.x) {
    int r = idx / kq4, q = idx - r * kq4;
    st4(dst + r * ldS + 4 * q, tile_fetch(s, meta[r], 4 * q));
  }
}

// transposed: dst[k * ldT + r]
__device__ __forceinline__ void stage_tile_transposed(const TileSrc& s, const RowMeta* meta, int BM, float* dst, int ldT) {
  const int kq4 = s.K >> 2;
  for (int idx = threadIdx.x; idx < BM * kq4; idx += blockDim.x) {
    int r = idx / kq4, q = idx - r * kq4;
    float4 v = tile_fetch(s, meta[r], 4 * q);
    float* d = dst + (4 * q) * ldT + r;
    d[0] = v.x; d[ldT] = v.y; d[2 * ldT] = v.z; d[3 * ldT] = v.w;
  }
}
