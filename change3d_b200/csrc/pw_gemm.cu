// Pointwise (1x1x1 / 1x1 / ConvTranspose-as-gather) convolution family as implicit GEMMs over
// NDHWC pixels.  v1 inner product: fp32 FFMA register tiling with the weights resident in
// shared memory per persistent CTA; prologues/epilogues fuse BN apply, SE gate, Swish,
// BN-backward, residual/shortcut adds and the per-channel batch statistics.
//
// Replaces (reference): nn.Conv3d 1x1x1 conv_a/conv_c/branch1_conv (model/x3d.py:173-175,214-216,
// 301-311), Encoder.enhance's Conv2d (model/trainer.py:57-69,88-108), ChangeDecoder's Conv2d 1x1 and
// ConvTranspose2d (model/change_decoder.py:30-45) and their autograd backward.
#include <stdlib.h>
#include "pw_gemm.cuh"
#include "../../include/change3d_b200.h"


template <int BM>
__global__ void __launch_bounds__(256) pw_gemm_kernel(const GemmArgs g) {
  constexpr int RI = BM / 16;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSrc a = g.a;
  const int n0 = blockIdx.y * g.NB;
  const int NB = g.NB, NBw = NB + 4;
  const int K = a.K, ldS = K + 4;

  float* Wt = reinterpret_cast<float*>(smem_raw);                  // [K][NBw]
  float* Xs = Wt + (size_t)K * NBw;                                // [BM][ldS]
  float* s_stat = Xs + (size_t)BM * ldS;                           // [2][NB]
  RowMeta* meta = reinterpret_cast<RowMeta*>(s_stat + 2 * NB);     // [BM]

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

  // resident weights: Wt[red][out] (zero outside the logical matrix)
  {
    const float* Wg = g.W;
    const bool red_fast = (g.w_sr == 1);
    const int total = K * NB;
    for (int idx = tid; idx < total; idx += 256) {
      int red, out;
      if (red_fast) { out = idx / K; red = idx - out * K; } else { red = idx / NB; out = idx - red * NB; }
      float v = 0.f;
      if (red < g.Kred && n0 + out < g.N) v = __ldg(Wg + (long long)red * g.w_sr + (long long)(n0 + out) * g.w_so);
      Wt[red * NBw + out] = v;
    }
    for (int i = tid; i < 2 * NB; i += 256) s_stat[i] = 0.f;
  }

  const long long ntiles = (g.M + BM - 1) / BM;
  const long long tpc = (ntiles + gridDim.x - 1) / gridDim.x;
  const long long t_begin = (long long)blockIdx.x * tpc;
  const long long t_end = t_begin + tpc < ntiles ? t_begin + tpc : ntiles;
  const bool has_stats = (g.stats != nullptr);
  long long cur_samp = -1;
  const int OW = a.OW;

  for (long long tile = t_begin; tile < t_end; ++tile) {
    const long long row0 = tile * BM;
    long long last = row0 + BM - 1 < g.M - 1 ? row0 + BM - 1 : g.M - 1;
    const long long samp0 = row0 / g.rows_per_sample;
    const bool straddle = (last / g.rows_per_sample) != samp0;
    __syncthreads();   // previous tile fully consumed (Xs, meta), s_stat updates visible
    if (has_stats && samp0 != cur_samp) {
      if (cur_samp >= 0) {
        for (int c = tid; c < NB; c += 256) {
          if (n0 + c < g.Ns) {
            atomicAdd(g.stats + (cur_samp * 2 + 0) * g.Ns + n0 + c, (double)s_stat[c]);
            atomicAdd(g.stats + (cur_samp * 2 + 1) * g.Ns + n0 + c, (double)s_stat[NB + c]);
          }
          s_stat[c] = 0.f; s_stat[NB + c] = 0.f;
        }
      }
      cur_samp = samp0;
    }
    tile_row_meta(a, row0, g.M, BM, meta);
    __syncthreads();
    stage_tile_rowmajor(a, meta, BM, Xs, ldS);
    __syncthreads();

    for (int nc = 0; nc < NB / 64; ++nc) {
      float acc[RI][4];
#pragma unroll
      for (int i = 0; i < RI; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
      const float* wp = Wt + nc * 64 + 4 * tx;
      for (int k4 = 0; k4 < K; k4 += 4) {
        float4 av[RI];
#pragma unroll
        for (int i = 0; i < RI; ++i) av[i] = *reinterpret_cast<const float4*>(Xs + (ty + 16 * i) * ldS + k4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          float4 b = *reinterpret_cast<const float4*>(wp + (k4 + kk) * NBw);
#pragma unroll
          for (int i = 0; i < RI; ++i) {
            float s = kk == 0 ? av[i].x : kk == 1 ? av[i].y : kk == 2 ? av[i].z : av[i].w;
            acc[i][0] = fmaf(s, b.x, acc[i][0]); acc[i][1] = fmaf(s, b.y, acc[i][1]);
            acc[i][2] = fmaf(s, b.z, acc[i][2]); acc[i][3] = fmaf(s, b.w, acc[i][3]);
          }
        }
      }
      // ---------------- epilogue ----------------
      const int col = n0 + nc * 64 + 4 * tx;
      const bool col_ok = col < g.Ns;
      float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < RI; ++i) {
        const int r = ty + 16 * i;
        const RowMeta m = meta[r];
        if (m.img < 0 || !col_ok) continue;
        const long long row = row0 + r;
        float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        if (g.epi == EPI_STORE) {
          long long addr = (long long)m.img * g.out_img_stride + (long long)(m.oh * OW + m.ow) * g.Ns + col;
          st4(g.Y + addr, v);
          s1[0] += v.x; s1[1] += v.y; s1[2] += v.z; s1[3] += v.w;
          s2[0] += v.x * v.x; s2[1] += v.y * v.y; s2[2] += v.z * v.z; s2[3] += v.w * v.w;
        } else if (g.epi == EPI_RELU_ADD) {
          long long addr = (long long)m.img * g.out_img_stride + (long long)(m.oh * OW + m.ow) * g.Ns + col;
          float4 old = *reinterpret_cast<const float4*>(g.Y + addr);
          if (g.Y2) st4(g.Y2 + row * g.Ns + col, old);
          st4(g.Y + addr, f4add(old, f4relu(v)));
        } else if (g.epi == EPI_SWISH_BWD) {
          float4 yb = ldg4(g.E1 + row * g.Ns + col);
          float4 mean = ldg4(BNP_MEAN(g.ebnp, g.Ns) + col), rstd = ldg4(BNP_RSTD(g.ebnp, g.Ns) + col);
          float4 scale = ldg4(BNP_SCALE(g.ebnp, g.Ns) + col), beta = ldg4(BNP_BETA(g.ebnp, g.Ns) + col);
          float4 gt = make_float4(1.f, 1.f, 1.f, 1.f);
          if (g.egate) gt = ldg4(g.egate + (row / g.rows_per_sample) * g.Ns + col);
          float yv[4] = {yb.x, yb.y, yb.z, yb.w}, mv[4] = {mean.x, mean.y, mean.z, mean.w};
          float rv[4] = {rstd.x, rstd.y, rstd.z, rstd.w}, sv[4] = {scale.x, scale.y, scale.z, scale.w};
          float bv[4] = {beta.x, beta.y, beta.z, beta.w}, gv[4] = {gt.x, gt.y, gt.z, gt.w};
          float av4[4] = {v.x, v.y, v.z, v.w}, du[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float xc = yv[j] - mv[j];
            float zh = xc * rv[j];
            float u = fmaf(xc, sv[j], bv[j]) * gv[j];
            du[j] = av4[j] * swish_gradf_(u);
            if (!straddle) { s1[j] += du[j]; s2[j] += du[j] * zh; }
            else if (has_stats) {
              long long sp = row / g.rows_per_sample;
              atomicAdd(g.stats + (sp * 2 + 0) * g.Ns + col + j, (double)du[j]);
              atomicAdd(g.stats + (sp * 2 + 1) * g.Ns + col + j, (double)(du[j] * zh));
            }
          }
          st4(g.Y + row * g.Ns + col, make_float4(du[0], du[1], du[2], du[3]));
        } else if (g.epi == EPI_ADD2) {
          if (g.E1) v = f4add(v, ldg4(g.E1 + row * g.Ns + col));
          if (g.E2 && !(m.oh & 1) && !(m.ow & 1)) {
            long long hr = (long long)m.img * (a.OHW >> 2) + (long long)(m.oh >> 1) * (OW >> 1) + (m.ow >> 1);
            v = f4add(v, ldg4(g.E2 + hr * g.Ns + col));
          }
          st4(g.Y + row * g.Ns + col, v);
        } else if (g.epi == EPI_ABSDIFF_BWD) {
          // backward of |x0 - x1|: Y (grad of x0) += sign * acc, Y2 (grad of x1) -= sign * acc, in place
          long long pix = (long long)(m.oh * OW + m.ow) * g.Ns + col;
          long long addr = (long long)m.img * g.out_img_stride + pix;
          long long eaddr = (long long)m.img * g.e1_img_stride + pix;
          const float4 x0 = ldg4(g.E1 + eaddr), x1 = ldg4(g.E2 + eaddr);
          float4 t;
          t.x = x0.x > x1.x ? v.x : (x0.x < x1.x ? -v.x : 0.f); t.y = x0.y > x1.y ? v.y : (x0.y < x1.y ? -v.y : 0.f);
          t.z = x0.z > x1.z ? v.z : (x0.z < x1.z ? -v.z : 0.f); t.w = x0.w > x1.w ? v.w : (x0.w < x1.w ? -v.w : 0.f);
          float4 a0 = *reinterpret_cast<const float4*>(g.Y + addr), a1 = *reinterpret_cast<const float4*>(g.Y2 + addr);
          st4(g.Y + addr, f4add(a0, t));
          st4(g.Y2 + addr, make_float4(a1.x - t.x, a1.y - t.y, a1.z - t.z, a1.w - t.w));
        }
      }
      if (has_stats && (g.epi == EPI_STORE || (g.epi == EPI_SWISH_BWD && !straddle))) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 16);
          s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 16);
        }
        if ((tid & 31) < 16 && col_ok) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            atomicAdd(&s_stat[nc * 64 + 4 * tx + j], s1[j]);
            atomicAdd(&s_stat[NB + nc * 64 + 4 * tx + j], s2[j]);
          }
        }
      }
    }
  }
  __syncthreads();
  if (has_stats && cur_samp >= 0) {
    for (int c = tid; c < NB; c += 256) {
      if (n0 + c < g.Ns) {
        atomicAdd(g.stats + (cur_samp * 2 + 0) * g.Ns + n0 + c, (double)s_stat[c]);
        atomicAdd(g.stats + (cur_samp * 2 + 1) * g.Ns + n0 + c, (double)s_stat[NB + c]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// wgrad: dW[n][k] += sum_r P[r][n] * Q[r][k]
// ------------------------------------------------------------------------------------------
struct WgradArgs {
  TileSrc p, q;
  long long M;
  float* dW; long long dw_sn, dw_sk;
  int N, K;          // logical extents written
  int n_chunks;      // grid.y = n_chunks * k_chunks
  int n_per, k_per;  // columns of P / Q handled per CTA (n_per = 16*NI_T, k_per = 64*KC_T)
};

template <int NI_T, int KC_T>
__global__ void __launch_bounds__(256) pw_wgrad_kernel(const WgradArgs g) {
  constexpr int BR = 32, ldP = BR + 4;
  constexpr int NP = NI_T * 16, KQ = KC_T * 64, ldQ = KQ + 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* Pt = reinterpret_cast<float*>(smem_raw);        // [NP][ldP]
  float* Qs = Pt + NP * ldP;                             // [BR][ldQ]
  RowMeta* metaP = reinterpret_cast<RowMeta*>(Qs + BR * ldQ);
  RowMeta* metaQ = metaP + BR;

  const int tid = threadIdx.x, tk = tid & 15, tn = tid >> 4;
  const int nch = blockIdx.y % g.n_chunks, kch = blockIdx.y / g.n_chunks;
  const int nbase = nch * NP, kbase = kch * KQ;
  // staged widths for this CTA (multiples of 4, clipped to the operand widths)
  const int pw = max(0, min(NP, g.p.K - nbase));
  const int qw = max(0, min(KQ, g.q.K - kbase));

  float acc[NI_T][KC_T][4];
#pragma unroll
  for (int i = 0; i < NI_T; ++i)
#pragma unroll
    for (int c = 0; c < KC_T; ++c) { acc[i][c][0] = acc[i][c][1] = acc[i][c][2] = acc[i][c][3] = 0.f; }

  // zero the smem once so columns beyond pw/qw stay zero
  for (int i = tid; i < NP * ldP + BR * ldQ; i += 256) Pt[i] = 0.f;

  const long long ntiles = (g.M + BR - 1) / BR;
  const long long tpc = (ntiles + gridDim.x - 1) / gridDim.x;
  const long long t_begin = (long long)blockIdx.x * tpc;
  const long long t_end = t_begin + tpc < ntiles ? t_begin + tpc : ntiles;

  for (long long tile = t_begin; tile < t_end; ++tile) {
    const long long row0 = tile * BR;
    __syncthreads();
    tile_row_meta(g.p, row0, g.M, BR, metaP);
    tile_row_meta(g.q, row0, g.M, BR, metaQ);
    __syncthreads();
    {
      const int pq4 = pw >> 2;
      for (int idx = tid; idx < BR * pq4; idx += 256) {
        int r = idx / pq4, q = idx - r * pq4;
        float4 v = tile_fetch(g.p, metaP[r], nbase + 4 * q);
        float* d = Pt + (4 * q) * ldP + r;
        d[0] = v.x; d[ldP] = v.y; d[2 * ldP] = v.z; d[3 * ldP] = v.w;
      }
      const int qq4 = qw >> 2;
      for (int idx = tid; idx < BR * qq4; idx += 256) {
        int r = idx / qq4, q = idx - r * qq4;
        st4(Qs + r * ldQ + 4 * q, tile_fetch(g.q, metaQ[r], kbase + 4 * q));
      }
    }
    __syncthreads();
#pragma unroll 2
    for (int r4 = 0; r4 < BR; r4 += 4) {
      float4 av[NI_T];
#pragma unroll
      for (int i = 0; i < NI_T; ++i) av[i] = *reinterpret_cast<const float4*>(Pt + (tn + 16 * i) * ldP + r4);
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
#pragma unroll
        for (int c = 0; c < KC_T; ++c) {
          float4 b = *reinterpret_cast<const float4*>(Qs + (r4 + rr) * ldQ + c * 64 + 4 * tk);
#pragma unroll
          for (int i = 0; i < NI_T; ++i) {
            float s = rr == 0 ? av[i].x : rr == 1 ? av[i].y : rr == 2 ? av[i].z : av[i].w;
            acc[i][c][0] = fmaf(s, b.x, acc[i][c][0]); acc[i][c][1] = fmaf(s, b.y, acc[i][c][1]);
            acc[i][c][2] = fmaf(s, b.z, acc[i][c][2]); acc[i][c][3] = fmaf(s, b.w, acc[i][c][3]);
          }
        }
      }
    }
  }
  if (t_begin >= t_end) return;
#pragma unroll
  for (int i = 0; i < NI_T; ++i) {
    const int n = nbase + tn + 16 * i;
    if (n >= g.N) continue;
#pragma unroll
    for (int c = 0; c < KC_T; ++c)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = kbase + c * 64 + 4 * tk + j;
        if (k < g.K) atomicAdd(g.dW + (long long)n * g.dw_sn + (long long)k * g.dw_sk, acc[i][c][j]);
      }
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
int c3d_launch_pw_gemm_tc(const GemmArgs& g, int num_sms, cudaStream_t stream, int lbo_is_k, int compact, int use_tma);   // pw_gemm_tc.cu
int c3d_launch_pw_wgrad_tc(const TileSrc& p, const TileSrc& q, long long M, float* dW, long long dw_sn, long long dw_sk,
                           int N, int K, int num_sms, cudaStream_t stream, int desc_swap);          // pw_wgrad_tc.cu
int c3d_launch_pw_wgrad_mn(const TileSrc& p, const TileSrc& q, long long M, float* dW, long long dw_sn, long long dw_sk,
                           int N, int K, int num_sms, cudaStream_t stream, int swap_lbo);           // pw_wgrad_mn.cu

// C3D_TC=0 forces the FFMA inner product; C3D_TC_LBO=0 swaps the LBO/SBO descriptor convention (bring-up aid).
static int env_flag(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

static int g_num_sms = 0;
static int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

static void fill_src(TileSrc& t, const c3d_operand& o) {
  t.A = o.A; t.A2 = o.A2; t.bnp = o.bnp; t.coef = o.coef; t.gate = o.gate;
  t.mode = o.mode; t.map = o.map; t.ld = o.ld;
  t.K = o.ld;
  t.OHW = o.OH * o.OW; t.OW = o.OW; t.IH = o.IH; t.IW = o.IW;
  t.img_stride = o.img_stride; t.img_stride2 = o.img_stride2 ? o.img_stride2 : o.img_stride;
  t.frames_per_sample = o.frames_per_sample > 0 ? o.frames_per_sample : 1;
}

static int check_src(const c3d_operand& o) {
  if (!o.A || o.ld <= 0 || (o.ld & 3) || o.OH <= 0 || o.OW <= 0 || o.IH <= 0 || o.IW <= 0) return C3D_ERR_ARG;
  if ((o.mode == PRO_BN_RELU || o.mode == PRO_BN_GATE_SWISH || o.mode == PRO_BNBWD) && !o.bnp) return C3D_ERR_ARG;
  if (o.mode == PRO_BNBWD && (!o.A2 || !o.coef)) return C3D_ERR_ARG;
  if ((o.mode == PRO_ABSDIFF || o.mode == PRO_MASK_POS) && !o.A2) return C3D_ERR_ARG;
  if (o.mode < 0 || o.mode > PRO_MASK_POS || (o.map != MAP_DENSE && o.map != MAP_SUB2)) return C3D_ERR_ARG;
  return C3D_OK;
}

extern "C" int c3d_pw_gemm(const c3d_gemm_desc* d, void* stream_) {
  if (!d || !d->W || !d->Y || d->M <= 0 || d->N <= 0 || d->Ns < d->N || (d->Ns & 3)) return C3D_ERR_ARG;
  if (int e = check_src(d->a)) return e;
  if (d->epi < 0 || d->epi > EPI_ABSDIFF_BWD || d->epi == EPI_RESERVED4) return C3D_ERR_ARG;
  if (d->epi == EPI_ABSDIFF_BWD && (!d->E1 || !d->E2 || !d->Y2)) return C3D_ERR_ARG;
  if (d->epi == EPI_SWISH_BWD && (!d->E1 || !d->ebnp)) return C3D_ERR_ARG;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  GemmArgs g;
  fill_src(g.a, d->a);
  g.W = d->W; g.w_sr = d->w_sr; g.w_so = d->w_so;
  g.Kred = d->Kred > 0 ? d->Kred : g.a.K;
  g.N = d->N; g.Ns = d->Ns; g.M = d->M; g.Y = d->Y;
  g.out_img_stride = d->out_img_stride ? d->out_img_stride : (long long)g.a.OHW * d->Ns;
  g.epi = d->epi; g.stats = d->stats;
  g.E1 = d->E1; g.e1_img_stride = d->e1_img_stride; g.E2 = d->E2;
  g.ebnp = d->ebnp; g.egate = d->egate; g.Y2 = d->Y2;
  g.rows_per_sample = d->rows_per_sample > 0 ? d->rows_per_sample : (1LL << 62);
  g.w_const = (d->flags & C3D_GEMM_W_CONSTANT) ? 1 : 0;
  if (env_flag("C3D_TC", 1)) {
    g.NB = 0; g.nsplit = 1;
    const int r = c3d_launch_pw_gemm_tc(g, num_sms(), stream, env_flag("C3D_TC_LBO", 1) | (env_flag("C3D_TC_HINT", 0) << 1),
                                            env_flag("C3D_TC_COMPACT", 1), env_flag("C3D_TC_TMA", 1));
    if (r >= 0) return r;
  }

  const int K = g.a.K;
  const int NBp = (d->Ns + 63) / 64 * 64;
  const size_t budget = 216 * 1024;
  int BM = 128, nsplit = 1, NB = NBp;
  auto smem_need = [&](int bm, int nb) -> size_t {
    return (size_t)K * (nb + 4) * 4 + (size_t)bm * (K + 4) * 4 + (size_t)2 * nb * 4 + (size_t)bm * sizeof(RowMeta);
  };
  while (true) {
    NB = ((NBp / 64 + nsplit - 1) / nsplit) * 64;
    if (smem_need(BM, NB) <= budget) break;
    if (BM == 128) { BM = 64; continue; }
    if (NB > 64) { nsplit++; BM = 128; continue; }
    if (BM == 64) { BM = 32; continue; }
    return C3D_ERR_SMEM;
  }
  nsplit = (NBp + NB - 1) / NB;
  if (d->M < 64 * 148 && BM == 128) BM = 64;   // small problems: more tiles
  g.NB = NB; g.nsplit = nsplit;
  const size_t smem = smem_need(BM, NB);
  const long long ntiles = (d->M + BM - 1) / BM;
  // persistent CTAs: as many as fit per SM by shared memory (registers cap at 2-3)
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  per_sm = per_sm < 1 ? 1 : per_sm > 3 ? 3 : per_sm;
  long long gx = (long long)num_sms() * per_sm;
  if (nsplit > 1) gx = (gx + nsplit - 1) / nsplit;
  if (gx > ntiles) gx = ntiles;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)nsplit);
  cudaError_t e;
  if (BM == 128) {
    e = cudaFuncSetAttribute(pw_gemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return C3D_ERR_CUDA;
    pw_gemm_kernel<128><<<grid, 256, smem, stream>>>(g);
  } else if (BM == 32) {
    e = cudaFuncSetAttribute(pw_gemm_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return C3D_ERR_CUDA;
    pw_gemm_kernel<32><<<grid, 256, smem, stream>>>(g);
  } else {
    e = cudaFuncSetAttribute(pw_gemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return C3D_ERR_CUDA;
    pw_gemm_kernel<64><<<grid, 256, smem, stream>>>(g);
  }
  return c3d_check_last(cudaGetLastError());
}

template <int NI_T, int KC_T>
static int launch_wgrad(WgradArgs& g, cudaStream_t stream) {
  constexpr int BR = 32;
  const size_t smem = (size_t)(NI_T * 16) * (BR + 4) * 4 + (size_t)BR * (KC_T * 64 + 4) * 4 + 2 * BR * sizeof(RowMeta);
  g.n_per = NI_T * 16; g.k_per = KC_T * 64;
  g.n_chunks = (g.p.K + g.n_per - 1) / g.n_per;
  const int k_chunks = (g.q.K + g.k_per - 1) / g.k_per;
  const long long ntiles = (g.M + BR - 1) / BR;
  long long gx = (long long)num_sms() * 2 / (g.n_chunks * k_chunks);
  if (gx < 1) gx = 1;
  if (gx > ntiles) gx = ntiles;
  dim3 grid((unsigned)gx, (unsigned)(g.n_chunks * k_chunks));
  cudaError_t e = cudaFuncSetAttribute(pw_wgrad_kernel<NI_T, KC_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return C3D_ERR_CUDA;
  pw_wgrad_kernel<NI_T, KC_T><<<grid, 256, smem, stream>>>(g);
  return c3d_check_last(cudaGetLastError());
}

extern "C" int c3d_pw_wgrad(const c3d_wgrad_desc* d, void* stream_) {
  if (!d || !d->dW || d->M <= 0 || d->N <= 0 || d->K <= 0) return C3D_ERR_ARG;
  if (int e = check_src(d->p)) return e;
  if (int e = check_src(d->q)) return e;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  WgradArgs g;
  fill_src(g.p, d->p);
  fill_src(g.q, d->q);
  g.M = d->M; g.dW = d->dW; g.dw_sn = d->dw_sn; g.dw_sk = d->dw_sk; g.N = d->N; g.K = d->K;
  if (env_flag("C3D_TC", 1) && env_flag("C3D_TC_WGRAD", 1) && env_flag("C3D_TC_WMN", 1)) {
    // dense operands: TMA-fed MN-major kernel (no transposes); C3D_TC_WMN=2 swaps the descriptor strides (bring-up)
    const int r = c3d_launch_pw_wgrad_mn(g.p, g.q, g.M, g.dW, g.dw_sn, g.dw_sk, g.N, g.K, num_sms(), stream,
                                         env_flag("C3D_TC_WMN", 1) == 2);
    if (r >= 0) return r;
  }
  if (env_flag("C3D_TC", 1) && env_flag("C3D_TC_WGRAD", 1)) {
    const int r = c3d_launch_pw_wgrad_tc(g.p, g.q, g.M, g.dW, g.dw_sn, g.dw_sk, g.N, g.K, num_sms(), stream,
                                         env_flag("C3D_TC_WSWAP", 0));
    if (r >= 0) return r;
  }
  const int np = g.p.K, kq = g.q.K;
  const bool k1 = kq <= 64;
  if (np <= 32) return k1 ? launch_wgrad<2, 1>(g, stream) : launch_wgrad<2, 2>(g, stream);
  if (np <= 64) return k1 ? launch_wgrad<4, 1>(g, stream) : launch_wgrad<4, 2>(g, stream);
  if (np <= 112) return k1 ? launch_wgrad<7, 1>(g, stream) : launch_wgrad<7, 2>(g, stream);
  return k1 ? launch_wgrad<14, 1>(g, stream) : launch_wgrad<14, 2>(g, stream);
}
