// tcgen05 pointwise GEMM, transposed formulation with the weights resident in TENSOR MEMORY:
//
//     D^T[out channel (TMEM lane)][pixel (TMEM column)] = W[out][k] (A operand, TMEM) * X^T[k][pixel] (B operand, smem)
//
// Why this orientation (measured on B200, profiles/r01_summary.md): keeping both halves of the 3xTF32-split weight
// matrix in shared memory (pw_gemm_tc.cu) costs up to 166 KB at res4, forces an N-split that makes every producer
// warp redo the activation prologue twice, and leaves 3-6 pipeline stages.  Here the weights sit in TMEM
// (2*K columns per 128-channel block), shared memory holds only activation stages (up to 12 in flight), the
// activation tile is staged exactly as before (K-major, pixels as rows), and the epilogue owns one output channel per
// thread: global stores are coalesced across lanes without staging and the BatchNorm statistics are per-thread
// register sums (no shuffles, no shared memory).
//
// Warps: 0-3 epilogue (+ initial W -> TMEM load), 4 TMEM allocation + the MMA-issuing thread, 5-16 producers.
// A tile is P pixels (128/64/32, chosen so that weights + accumulators fit the 512 TMEM columns); a pipeline stage is
// P pixels x KC channels with P*KC = 2048 (16 KB for both split halves).
#include "tc_common.cuh"

namespace tw {

using namespace tc;

constexpr int NPROD = 12;
constexpr int NTHREADS = (5 + NPROD) * 32;
constexpr int STAGE_ELEMS = 2048;            // per split half

struct Params {
  GemmArgs g;
  int P;             // pixels per tile (UMMA N)
  int KC;            // channels per stage (STAGE_ELEMS / P)
  int nblk;          // 128-lane blocks of output channels
  int nsets;         // accumulator sets (1 or 2)
  int nstage;        // == producer warps in use
  int wcols;         // TMEM columns holding the weights: nblk * 2 * K
  int dense_contig;
  int dbg;           // bring-up switches: 1 skip epilogue body, 2 skip MMA issue, 4 skip producer loads
};

__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
  const uint32_t* u = reinterpret_cast<const uint32_t*>(v);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]) : "memory");
}

// One producer warp fills one stage: P pixels x KC channels.  lane = (pixel % 8, 16-byte chunk within a group of 4);
// 16 iterations of 8 pixels x 16 channels, 8 loads (16 with a second source) in flight per lane.
template <int MODE>
__device__ __forceinline__ void produce_stage(const Params& P, const TileSrc& s, long long row0, int k0, float* a_hi,
                                              float* a_lo, int lane) {
  constexpr bool HAS2 = (MODE == PRO_BNBWD || MODE == PRO_ABSDIFF || MODE == PRO_MASK_POS);
  constexpr int BATCH = 8;
  const int qq = lane >> 3, rl = lane & 7;
  const long long M = P.g.M;
  const int rgs = P.P >> 3;                    // row groups per stage
  const bool fast = P.dense_contig && (row0 + P.P <= M);
  // SE gate: rows of a tile belong to at most two consecutive samples when a sample has >= P rows
  const uint32_t rps = (uint32_t)s.OHW * (uint32_t)s.frames_per_sample;
  const bool gate_fast = (MODE == PRO_BN_GATE_SWISH) && s.gate && fast && rps >= (uint32_t)P.P;
  uint32_t samp0 = 0;
  int gsplit = P.P;
  if (gate_fast) {
    samp0 = (uint32_t)row0 / rps;
    gsplit = (int)((samp0 + 1) * rps - (uint32_t)row0);
  }
#pragma unroll 1
  for (int b0 = 0; b0 < 16; b0 += BATCH) {
    float4 v[BATCH], v2[BATCH];
    uint32_t imgs[BATCH];
#pragma unroll
    for (int i = 0; i < BATCH; ++i) {
      const int it = b0 + i;
      const int qg = it / rgs, rg = it - qg * rgs;
      const int k = k0 + (qg * 4 + qq) * 4;
      const int r = rg * 8 + rl;
      v[i] = f4zero(); v2[i] = f4zero(); imgs[i] = 0;
      if (k < s.K) {
        if (fast) {
          const long long off = (row0 + r) * (long long)s.ld + k;
          v[i] = ldg4(s.A + off);
          if (HAS2) v2[i] = ldg4(s.A2 + off);
        } else if (row0 + r < M) {
          long long off, off2;
          if (P.dense_contig) { off = (row0 + r) * (long long)s.ld; off2 = off; imgs[i] = (uint32_t)(row0 + r) / (uint32_t)s.OHW; }
          else row_offsets(s, (uint32_t)(row0 + r), off, off2, imgs[i]);
          v[i] = ldg4(s.A + off + k);
          if (HAS2) v2[i] = ldg4(s.A2 + off2 + k);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < BATCH; ++i) {
      const int it = b0 + i;
      const int qg = it / rgs, rg = it - qg * rgs;
      const int quad = qg * 4 + qq;
      const int k = k0 + quad * 4;
      const int r = rg * 8 + rl;
      if (k >= s.K) continue;                  // never read by the MMA (K-steps stop at K)
      float4 x = f4zero();
      if (fast || row0 + r < M) {
        float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (MODE == PRO_BN_GATE_SWISH && s.gate) {
          uint32_t sp;
          if (gate_fast) sp = samp0 + (r >= gsplit ? 1u : 0u);
          else if (fast) sp = ((uint32_t)(row0 + r) / (uint32_t)s.OHW) / (uint32_t)s.frames_per_sample;
          else sp = imgs[i] / (uint32_t)s.frames_per_sample;
          g4 = ldg4(s.gate + (long long)sp * s.ld + k);
        }
        const ChanParams cp = load_chan_params<MODE>(s, k);
        x = prologue<MODE>(cp, v[i], v2[i], g4);
      }
      float4 hi, lo;
      split4(x, hi, lo);
      const int o = ((quad * rgs + rg) * 8 + rl) * 4;
      *reinterpret_cast<float4*>(a_hi + o) = hi;
      *reinterpret_cast<float4*>(a_lo + o) = lo;
    }
  }
}

__global__ void __launch_bounds__(NTHREADS, 1) pw_gemm_tw_kernel(const Params P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const GemmArgs& g = P.g;
  const TileSrc a = g.a;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = a.K;
  const int nstg_per_tile = (K + P.KC - 1) / P.KC;

  float* stages = reinterpret_cast<float*>(smem_raw);                       // nstage x (hi, lo) x STAGE_ELEMS
  uint64_t* bars = reinterpret_cast<uint64_t*>(stages + (size_t)P.nstage * 2 * STAGE_ELEMS);
  uint64_t* full = bars;                  // [NPROD]
  uint64_t* empty = bars + NPROD;         // [NPROD]
  uint64_t* tfull = empty + NPROD;        // [2]
  uint64_t* tempty = tfull + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  if (threadIdx.x == 0) {
    for (int i = 0; i < P.nstage; ++i) { mbar_init(smem_u32(full + i), 1); mbar_init(smem_u32(empty + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(tfull + i), 1); mbar_init(smem_u32(tempty + i), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc(smem_u32(tmem_slot), 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- weights -> TMEM (warps 0-3: thread = output channel within a block; hi at column k, lo at K + k) ----
  if (warp < 4) {
    for (int blk = 0; blk < P.nblk; ++blk) {
      const int n = blk * 128 + warp * 32 + lane;
      const uint32_t t_w = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(blk * 2 * K);
      for (int k8 = 0; k8 < K; k8 += 8) {
        float hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int red = k8 + j;
          const float w = (red < g.Kred && n < g.N) ? __ldg(g.W + (long long)red * g.w_sr + (long long)n * g.w_so) : 0.f;
          split_tf32(w, hi[j], lo[j]);
        }
        tmem_st8(t_w + (uint32_t)k8, hi);
        tmem_st8(t_w + (uint32_t)(K + k8), lo);
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const long long ntiles = (g.M + P.P - 1) / P.P;
  const long long tpc = (ntiles + gridDim.x - 1) / gridDim.x;
  const long long t_begin = (long long)blockIdx.x * tpc;
  const long long t_end = t_begin + tpc < ntiles ? t_begin + tpc : ntiles;
  const long long my_tiles = t_end > t_begin ? t_end - t_begin : 0;
  const uint32_t d_base = tmem_base + (uint32_t)P.wcols;          // accumulators: [set][blk][main P | corr P]

  if (warp >= 5) {
    // ===================== producers (one warp per stage) =====================
    const int p = warp - 5;
    const long long total = (p < P.nstage) ? my_tiles * nstg_per_tile : 0;
    uint32_t use = 0;
    for (long long c = p; c < total; c += P.nstage, ++use) {
      const long long ti = c / nstg_per_tile;
      const int sidx = (int)(c - ti * nstg_per_tile);
      mbar_wait(smem_u32(empty + p), (use & 1) ^ 1);
      float* a_hi = stages + (size_t)p * 2 * STAGE_ELEMS;
      float* a_lo = a_hi + STAGE_ELEMS;
      const long long row0 = (t_begin + ti) * P.P;
      const int k0 = sidx * P.KC;
      if (!(P.dbg & 4))
      switch (a.mode) {
        case PRO_NONE: produce_stage<PRO_NONE>(P, a, row0, k0, a_hi, a_lo, lane); break;
        case PRO_BN_RELU: produce_stage<PRO_BN_RELU>(P, a, row0, k0, a_hi, a_lo, lane); break;
        case PRO_BN_GATE_SWISH: produce_stage<PRO_BN_GATE_SWISH>(P, a, row0, k0, a_hi, a_lo, lane); break;
        case PRO_BNBWD: produce_stage<PRO_BNBWD>(P, a, row0, k0, a_hi, a_lo, lane); break;
        case PRO_ABSDIFF: produce_stage<PRO_ABSDIFF>(P, a, row0, k0, a_hi, a_lo, lane); break;
        default: produce_stage<PRO_MASK_POS>(P, a, row0, k0, a_hi, a_lo, lane); break;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(full + p));
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(P.P >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t lbo = (uint32_t)(P.P >> 3) * 128, sbo = 128;   // stage layout [16-byte chunk][row group][8][16 B]
      long long c = 0;
      for (long long ti = 0; ti < my_tiles; ++ti) {
        const int set = (P.nsets == 2) ? (int)(ti & 1) : 0;
        const uint32_t use = (P.nsets == 2) ? (uint32_t)(ti >> 1) : (uint32_t)ti;
        mbar_wait(smem_u32(tempty + set), (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_set = d_base + (uint32_t)(set * P.nblk * 2 * P.P);
        for (int sidx = 0; sidx < nstg_per_tile; ++sidx, ++c) {
          const int stage = (int)(c % P.nstage);
          mbar_wait(smem_u32(full + stage), (uint32_t)(c / P.nstage) & 1);
          tc_fence_after();
          const uint32_t xhi = smem_u32(stages + (size_t)stage * 2 * STAGE_ELEMS), xlo = xhi + STAGE_ELEMS * 4;
          const int k0 = sidx * P.KC;
          const int ksteps = (K - k0 >= P.KC ? P.KC : K - k0) >> 3;
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t dxh = make_desc(xhi + (uint32_t)ks * 2 * lbo, lbo, sbo);
            const uint64_t dxl = make_desc(xlo + (uint32_t)ks * 2 * lbo, lbo, sbo);
            const uint32_t first = (sidx == 0 && ks == 0) ? 0u : 1u;
            for (int blk = 0; blk < P.nblk; ++blk) {
              const uint32_t w_hi = tmem_base + (uint32_t)(blk * 2 * K + k0 + ks * 8), w_lo = w_hi + (uint32_t)K;
              const uint32_t d_main = d_set + (uint32_t)(blk * 2 * P.P), d_corr = d_main + (uint32_t)P.P;
              if (!(P.dbg & 2)) {
                umma_tf32_ts(d_main, w_hi, dxh, idesc, first);
                umma_tf32_ts(d_corr, w_lo, dxh, idesc, first);
                umma_tf32_ts(d_corr, w_hi, dxl, idesc, 1u);
              }
            }
          }
          umma_commit(smem_u32(empty + stage));
        }
        umma_commit(smem_u32(tfull + set));
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue: thread = output channel (TMEM lane), columns = pixels =====================
    const bool has_stats = g.stats != nullptr;
    const bool track = has_stats || (g.epi == EPI_SWISH_BWD && g.egate != nullptr);   // needs the sample index
    const long long rps = g.rows_per_sample;
    float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};      // per-block register sums for the block's current sample
    long long cur_samp[2] = {-1, -1}, next_b[2] = {0, 0};
    for (long long ti = 0; ti < my_tiles; ++ti) {
      const long long row0 = (t_begin + ti) * P.P;
      const int set = (P.nsets == 2) ? (int)(ti & 1) : 0;
      const uint32_t use = (P.nsets == 2) ? (uint32_t)(ti >> 1) : (uint32_t)ti;
      mbar_wait(smem_u32(tfull + set), use & 1);
      tc_fence_after();
      const uint32_t t_set = d_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(set * P.nblk * 2 * P.P);
#pragma unroll
      for (int blk = 0; blk < 2; ++blk) {
        if (blk >= P.nblk) break;
        const int n = blk * 128 + warp * 32 + lane;
        const bool n_ok = n < g.Ns;
        float mean = 0.f, rstd = 0.f, scale = 0.f, beta = 0.f, gt = 1.f;
        if (g.epi == EPI_SWISH_BWD && n_ok) {
          mean = __ldg(BNP_MEAN(g.ebnp, g.Ns) + n); rstd = __ldg(BNP_RSTD(g.ebnp, g.Ns) + n);
          scale = __ldg(BNP_SCALE(g.ebnp, g.Ns) + n); beta = __ldg(BNP_BETA(g.ebnp, g.Ns) + n);
        }
        if (track) {
          if (cur_samp[blk] < 0) { cur_samp[blk] = row0 / rps; next_b[blk] = (cur_samp[blk] + 1) * rps; }
          if (g.egate && n_ok) gt = __ldg(g.egate + cur_samp[blk] * g.Ns + n);
        }
        const long long o0 = row0 * (long long)g.Ns + n;      // element offset of (row0, n)
        for (int c0 = 0; c0 < P.P; c0 += 32) {
          float r[32], r2[32];
          tmem_ld32(t_set + (uint32_t)(blk * 2 * P.P + c0), r);
          tmem_ld32(t_set + (uint32_t)(blk * 2 * P.P + P.P + c0), r2);
          if (blk == P.nblk - 1 && c0 + 32 >= P.P) {       // accumulator set fully read by this warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(tempty + set));
          }
          // EPI_ADD2 with the x2-upsampled shortcut gradient: lane j decodes pixel c0+j once, broadcast below
          long long hr_mine = -1;
          if (g.epi == EPI_ADD2 && g.E2) {
            const long long row = row0 + c0 + lane;
            if (row < g.M) {
              const uint32_t img = (uint32_t)row / (uint32_t)a.OHW;
              const uint32_t rem = (uint32_t)row - img * (uint32_t)a.OHW;
              const uint32_t oh = rem / (uint32_t)a.OW, ow = rem - oh * (uint32_t)a.OW;
              if (!(oh & 1) && !(ow & 1)) hr_mine = (long long)img * (a.OHW >> 2) + (long long)(oh >> 1) * (a.OW >> 1) + (ow >> 1);
            }
          }
          int jmax = (g.M - (row0 + c0)) < 32 ? (int)(g.M - (row0 + c0)) : 32;      // warp-uniform
          if (P.dbg & 1) jmax = 0;
          const float* e1p = g.E1 ? g.E1 + o0 + (long long)c0 * g.Ns : nullptr;
          float* yp = g.Y + o0 + (long long)c0 * g.Ns;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            long long hr = -1;
            if (g.epi == EPI_ADD2 && g.E2) hr = __shfl_sync(0xffffffffu, hr_mine, j);
            if (j >= jmax) continue;
            if (track && row0 + c0 + j >= next_b[blk]) {        // warp-uniform: the pixel starts the next sample
              if (has_stats && n_ok) {
                atomicAdd(g.stats + (cur_samp[blk] * 2 + 0) * g.Ns + n, (double)s1[blk]);
                atomicAdd(g.stats + (cur_samp[blk] * 2 + 1) * g.Ns + n, (double)s2[blk]);
              }
              s1[blk] = 0.f; s2[blk] = 0.f;
              cur_samp[blk] += 1; next_b[blk] += rps;
              if (g.egate && n_ok) gt = __ldg(g.egate + cur_samp[blk] * g.Ns + n);
            }
            if (!n_ok) continue;
            float v = r[j] + r2[j];
            const long long oj = (long long)j * g.Ns;
            if (g.epi == EPI_STORE) {
              yp[oj] = v;
              s1[blk] += v;
              s2[blk] = fmaf(v, v, s2[blk]);
            } else if (g.epi == EPI_SWISH_BWD) {
              const float xc = __ldg(e1p + oj) - mean;
              const float zh = xc * rstd;
              const float u = fmaf(xc, scale, beta) * gt;
              const float du = v * swish_gradf_(u);
              yp[oj] = du;
              s1[blk] += du;
              s2[blk] = fmaf(du, zh, s2[blk]);
            } else {   // EPI_ADD2
              if (e1p) v += __ldg(e1p + oj);
              if (hr >= 0) v += __ldg(g.E2 + hr * g.Ns + n);
              yp[oj] = v;
            }
          }
        }
      }
    }
    if (has_stats) {
#pragma unroll
      for (int blk = 0; blk < 2; ++blk) {
        const int n = blk * 128 + warp * 32 + lane;
        if (blk < P.nblk && cur_samp[blk] >= 0 && n < g.Ns) {
          atomicAdd(g.stats + (cur_samp[blk] * 2 + 0) * g.Ns + n, (double)s1[blk]);
          atomicAdd(g.stats + (cur_samp[blk] * 2 + 1) * g.Ns + n, (double)s2[blk]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

}  // namespace tw

// Host-side eligibility + launch.  Returns -1 when the shape / mode is not handled here.
int c3d_launch_pw_gemm_tw(const GemmArgs& g, int num_sms, cudaStream_t stream, int dbg) {
  if (g.a.map != MAP_DENSE && g.a.map != MAP_SUB2) return -1;
  if (g.epi != EPI_STORE && g.epi != EPI_SWISH_BWD && g.epi != EPI_ADD2) return -1;
  if (g.out_img_stride != (long long)g.a.OHW * g.Ns) return -1;
  if ((g.a.K & 7) || g.a.K < 8 || (g.Ns & 3)) return -1;
  if (g.M >= (1LL << 31) || g.M < 128) return -1;
  if (g.stats && g.epi == EPI_ADD2) return -1;
  tw::Params P;
  P.g = g;
  P.dbg = dbg;
  const int K = g.a.K;
  P.nblk = (g.Ns + 127) / 128;
  if (P.nblk > 2) return -1;
  P.wcols = P.nblk * 2 * K;
  const int left = 512 - P.wcols;
  // largest tile that fits, preferring two accumulator sets (epilogue of tile i overlaps the MMAs of tile i+1)
  int bestP = 0, bestSets = 0;
  for (int sets = 2; sets >= 1 && !bestP; --sets)
    for (int p = 128; p >= 32; p >>= 1)
      if (sets * P.nblk * 2 * p <= left) { bestP = p; bestSets = sets; break; }
  if (bestSets == 2 && bestP < 64) {           // a single set with a 2x larger tile beats two tiny sets
    for (int p = 128; p >= 32; p >>= 1)
      if (P.nblk * 2 * p <= left && p > bestP) { bestP = p; bestSets = 1; break; }
  }
  if (!bestP) return -1;
  P.P = bestP;
  P.nsets = bestSets;
  P.KC = tw::STAGE_ELEMS / P.P;
  P.nstage = tw::NPROD;
  P.dense_contig = (g.a.map == MAP_DENSE && g.a.img_stride == (long long)g.a.OHW * g.a.ld &&
                    (g.a.A2 == nullptr || g.a.img_stride2 == (long long)g.a.OHW * g.a.ld)) ? 1 : 0;
  const size_t smem = (size_t)P.nstage * 2 * tw::STAGE_ELEMS * 4 + 512;
  cudaError_t e = cudaFuncSetAttribute(tw::pw_gemm_tw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return C3D_ERR_SMEM;
  const long long ntiles = (g.M + P.P - 1) / P.P;
  long long gx = num_sms;
  if (gx > ntiles) gx = ntiles;
  tw::pw_gemm_tw_kernel<<<(unsigned)gx, tw::NTHREADS, smem, stream>>>(P);
  return c3d_check_last(cudaGetLastError());
}
