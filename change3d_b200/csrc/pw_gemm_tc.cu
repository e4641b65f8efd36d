// tcgen05 inner product for the pointwise-GEMM family (sm_100a).
//
//   D[128 pixels x N] (fp32, TMEM) = A[128 x K] * B[K x N]        with fp32-class accuracy from a
//   3-term TF32 split:  A = Ah + Al, B = Bh + Bl (h = top 19 bits, l = exact remainder),
//   D += Ah*Bh + Al*Bh + Ah*Bl  (tcgen05.mma.kind::tf32, fp32 accumulate).
//
// Warp roles (one CTA per 128-pixel tile stream, persistent over a contiguous tile range):
//   warps 0-7   epilogue, two warpgroups alternating tiles (even tiles -> warps 0-3 / accumulator set 0,
//               odd tiles -> warps 4-7 / set 1): tcgen05.ld -> registers -> shared (row per thread) ->
//               coalesced global stores, fused epilogue math (BN statistics / Swish'*SE*BN_b backward /
//               residual join)
//   warp  8     TMEM allocation + the single MMA-issuing thread
//   warps 9-16  producers: global float4 loads of one 16-channel K-chunk of the tile, fused prologue
//               (BN apply / ReLU / SE gate*Swish / BN-backward / |a-b| / mask), hi/lo split, stores in the
//               UMMA canonical K-major (no-swizzle) layout, mbarrier arrive
// The weight matrix (both split halves) stays resident in shared memory in the same canonical layout.
// Each tile owns two TMEM accumulators: the main term Ah*Bh and the correction terms Al*Bh + Ah*Bl are
// accumulated separately and added in the epilogue — the tensor core's fp32 accumulation rounds toward
// zero, and keeping the 2^-11-sized terms out of the main chain cuts that bias ~3x.
#include <stdio.h>
#include <stdlib.h>

#include "pw_gemm.cuh"

#include "tc_common.cuh"

namespace tc {

constexpr int BM = 128;        // pixels per tile (UMMA M)
constexpr int KC = 16;         // channels per pipeline stage (two K=8 tf32 MMA steps)
// Role counts are template parameters: <8,8> (17 warps, one CTA per SM) or the compact <4,4> (9 warps) that fits
// two CTAs per SM for the small-channel layers — per-tile latency (producer -> MMA -> commit -> epilogue hand-offs,
// ~2 us) is what bounds these short-K GEMMs, so tiles in flight per SM matter more than anything else.
constexpr int STAGE_FLOATS = BM * KC;          // per split half
constexpr int EPI_LD = 36;                     // staging pitch (floats) of the epilogue tile: 32 columns + 4

struct Params {
  GemmArgs g;
  int NpB;           // accumulator width of this CTA (multiple of 16, <= 128)
  int NpA;           // TMEM column stride per accumulator (multiple of 32, <= 128)
  int tmem_cols;     // allocation: 4 accumulators (2 tile sets x {main, correction}), power of two >= 32
  int nstage;        // pipeline stages; stage s belongs to producer warp s % nprod (exactly one producer per stage,
                     // so the parity-tracked empty/full mbarriers always see a single sequential producer)
  int nprod;         // producer warps in use (<= NPROD, <= nstage)
  int tma;           // 1: the raw operand rows arrive by TMA into the stage itself (64B-swizzled) and the
                     // producer warps transform them in place; 0: producers gather rows with global loads
  int dense_contig;  // rows are contiguous pixels: offset = row * ld
  int lbo_is_k;      // descriptor convention switch (1: LBO = stride between K chunks)
  int wait_hint;     // mbarrier try_wait suspend hint in ns (0 = none)
  int ncat;          // 1: W hi / W lo column groups are interleaved per K chunk so that Ah x [Wh | Wl] is ONE MMA of 2 * NpB
                     // columns writing [main | correction] (correction at column NpB): A hi is read twice, not three times
  int epi_bufs;      // staging tiles per epilogue warp: 2 when the Swish-backward statistics need the second one
  int early;         // TMA path: the first copies of every producer warp are issued BEFORE the weights are staged (the
                     // staging is ~5 K cycles per launch and used to sit in front of the first DRAM round trip)
  int epar;          // 1: the Swish-backward epilogue reads its per-column parameters from shared memory (staged at setup)
  int pf_dist;       // TMA path: L2 prefetch distance in tiles (0 = off): the boxes of tile t + pf_dist are requested
                     // when tile t's copies are issued, so HBM latency is paid ahead of the shared-memory pipeline
  uint32_t rps;      // rows per batch sample, clamped to 2^31 - 1 (M < 2^31: all row arithmetic fits 32 bits)
  unsigned long long* dbg;   // C3D_TC_DBG=1: per-warp wait/work cycle counters of one CTA ([32 warps][8]), else null
};

// One producer warp fills one stage: 128 rows x 16 channels, lane = (row % 8, 16-byte chunk), 16 row groups.
template <int MODE>
__device__ __forceinline__ void produce_chunk(const Params& P, const TileSrc& s, long long row0, int chunk, float* a_hi,
                                              float* a_lo, int lane, int bg0, int bg1) {
  constexpr bool HAS2 = (MODE == PRO_BNBWD || MODE == PRO_ABSDIFF || MODE == PRO_MASK_POS);
  constexpr int BATCH = 8;
  const int qq = lane >> 3, rl = lane & 7;
  const int k = chunk * KC + 4 * qq;
  const bool kvalid = k < s.K;
  const long long M = P.g.M;
  const bool fast = P.dense_contig && (row0 + BM <= M);
  float* dst_hi = a_hi + (qq * 16 * 8 + rl) * 4;
  float* dst_lo = a_lo + (qq * 16 * 8 + rl) * 4;
  if (!kvalid) {      // K % 16 == 8: the upper half of the last chunk is never read by the MMA
    return;
  }
  const ChanParams cp = load_chan_params<MODE>(s, k);
  // SE gate: rows of a tile belong to at most two consecutive samples when a sample has >= 128 rows
  float4 g0 = make_float4(1.f, 1.f, 1.f, 1.f), g1 = g0;
  int gsplit = BM;          // first tile row that belongs to the second sample
  bool gate_fast = false;
  if (MODE == PRO_BN_GATE_SWISH && s.gate && fast) {
    const uint32_t rps = (uint32_t)s.OHW * (uint32_t)s.frames_per_sample;
    if (rps >= (uint32_t)BM) {
      const uint32_t samp0 = (uint32_t)row0 / rps;
      gsplit = (int)((samp0 + 1) * rps - (uint32_t)row0);
      g0 = ldg4(s.gate + (long long)samp0 * s.ld + k);
      if (gsplit < BM) g1 = ldg4(s.gate + (long long)(samp0 + 1) * s.ld + k);
      gate_fast = true;
    }
  }
  const float* baseA = s.A + row0 * s.ld + k;
  const float* baseA2 = HAS2 ? s.A2 + row0 * s.ld + k : nullptr;
#pragma unroll 1
  for (int b0 = bg0; b0 < bg1; b0 += BATCH) {
    float4 v[BATCH], v2[BATCH];
    uint32_t imgs[BATCH];
#pragma unroll
    for (int i = 0; i < BATCH; ++i) {
      const int r = (b0 + i) * 8 + rl;
      v2[i] = f4zero();
      imgs[i] = 0;
      if (fast) {
        v[i] = ldg4(baseA + r * s.ld);
        if (HAS2) v2[i] = ldg4(baseA2 + r * s.ld);
      } else {
        v[i] = f4zero();
        const long long row = row0 + r;
        if (row < M) {
          long long off, off2;
          if (P.dense_contig) { off = row * s.ld; off2 = off; imgs[i] = (uint32_t)row / (uint32_t)s.OHW; }
          else row_offsets(s, (uint32_t)row, off, off2, imgs[i]);
          v[i] = ldg4(s.A + off + k);
          if (HAS2) v2[i] = ldg4(s.A2 + off2 + k);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < BATCH; ++i) {
      const int r = (b0 + i) * 8 + rl;
      float4 x;
      if (fast) {
        float4 g4 = g0;
        if (MODE == PRO_BN_GATE_SWISH && s.gate) {
          if (gate_fast) g4 = (r >= gsplit) ? g1 : g0;
          else g4 = ldg4(s.gate + (long long)(((uint32_t)(row0 + r) / (uint32_t)s.OHW) / (uint32_t)s.frames_per_sample) * s.ld + k);
        }
        x = prologue<MODE>(cp, v[i], v2[i], g4);
      } else {
        x = f4zero();
        if (row0 + r < M) {
          float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f);
          if (MODE == PRO_BN_GATE_SWISH && s.gate) g4 = ldg4(s.gate + (long long)(imgs[i] / (uint32_t)s.frames_per_sample) * s.ld + k);
          x = prologue<MODE>(cp, v[i], v2[i], g4);
        }
      }
      float4 hi, lo;
      split4(x, hi, lo);
      *reinterpret_cast<float4*>(dst_hi + (b0 + i) * 32) = hi;
      *reinterpret_cast<float4*>(dst_lo + (b0 + i) * 32) = lo;
    }
  }
}

// TMA path: the stage already holds the raw rows (A in the hi buffer, the second operand in the lo buffer), as
// 128 rows x 64 bytes with the 16-byte units of row r XOR-swizzled by (r >> 1) & 3.  Apply the prologue and the
// hi/lo split in place; a lane keeps its logical channel quad for all rows, so the physical offset inside an
// 8-row group is a per-lane constant and a warp touches each 512-byte group exactly once (conflict-free).
template <int MODE>
__device__ __forceinline__ void transform_chunk(const Params& P, const TileSrc& s, long long row0, int chunk, float* a_hi,
                                                float* a_lo, int lane, int bg0, int bg1) {
  constexpr bool HAS2 = (MODE == PRO_BNBWD || MODE == PRO_ABSDIFF || MODE == PRO_MASK_POS);
  const int qq = lane >> 3, rl = lane & 7;
  const int k = chunk * KC + 4 * qq;
  if (k >= s.K) return;      // K % 16 == 8: the upper half of the last chunk is never read by the MMA
  const ChanParams cp = load_chan_params<MODE>(s, k);
  float4 g0 = make_float4(1.f, 1.f, 1.f, 1.f), g1 = g0;
  int gsplit = BM;
  bool gate_fast = true;
  if (MODE == PRO_BN_GATE_SWISH && s.gate) {
    const uint32_t rps = (uint32_t)s.OHW * (uint32_t)s.frames_per_sample;
    if (rps >= (uint32_t)BM) {
      const uint32_t samp0 = (uint32_t)row0 / rps;
      gsplit = (int)((samp0 + 1) * rps - (uint32_t)row0);
      g0 = ldg4(s.gate + (long long)samp0 * s.ld + k);
      if (gsplit < BM && (long long)(samp0 + 1) * rps < P.g.M) g1 = ldg4(s.gate + (long long)(samp0 + 1) * s.ld + k);
    } else {
      gate_fast = false;
    }
  }
  const long long left = P.g.M - row0;
  const int nvalid = left < BM ? (int)left : BM;
  const int po = rl * 16 + ((qq ^ ((rl >> 1) & 3)) << 2);
  const bool slow_gate = (MODE == PRO_BN_GATE_SWISH && s.gate && !gate_fast);
  // batches of 4 row groups: all shared-memory loads of a batch are issued before its first store (the in-place
  // stores would otherwise order every load behind the previous iteration's stores)
  if (nvalid == BM && !slow_gate && gsplit >= BM) {
    // common case: a full tile inside one batch sample -- no per-row selects (gate choice, tail zeroing) in the loop
#pragma unroll 2
    for (int b0 = bg0; b0 < bg1; b0 += 4) {
      float4 v[4], v2[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[i] = *reinterpret_cast<const float4*>(a_hi + (b0 + i) * 128 + po);
        v2[i] = HAS2 ? *reinterpret_cast<const float4*>(a_lo + (b0 + i) * 128 + po) : f4zero();
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 hi, lo;
        split4(prologue<MODE>(cp, v[i], v2[i], g0), hi, lo);
        *reinterpret_cast<float4*>(a_hi + (b0 + i) * 128 + po) = hi;
        *reinterpret_cast<float4*>(a_lo + (b0 + i) * 128 + po) = lo;
      }
    }
    return;
  }
#pragma unroll 2
  for (int b0 = bg0; b0 < bg1; b0 += 4) {
    float4 v[4], v2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[i] = *reinterpret_cast<const float4*>(a_hi + (b0 + i) * 128 + po);
      v2[i] = HAS2 ? *reinterpret_cast<const float4*>(a_lo + (b0 + i) * 128 + po) : f4zero();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = (b0 + i) * 8 + rl;
      float4 g4 = (r >= gsplit) ? g1 : g0;
      if (slow_gate && r < nvalid)
        g4 = ldg4(s.gate + (long long)(((uint32_t)(row0 + r) / (uint32_t)s.OHW) / (uint32_t)s.frames_per_sample) * s.ld + k);
      float4 x = prologue<MODE>(cp, v[i], v2[i], g4);
      if (r >= nvalid) x = f4zero();     // TMA zero-fills rows past M; keep them zero through the prologue
      float4 hi, lo;
      split4(x, hi, lo);
      *reinterpret_cast<float4*>(a_hi + (b0 + i) * 128 + po) = hi;
      *reinterpret_cast<float4*>(a_lo + (b0 + i) * 128 + po) = lo;
    }
  }
}

// WPS = producer warps per pipeline stage (each transforms 128 / WPS rows of the stage's chunk); CPS = CTAs per SM
template <int NEPI, int NPROD, int WPS, int CPS>
__global__ void __launch_bounds__((NEPI + 1 + NPROD) * 32, CPS) pw_gemm_tc_kernel(const Params P, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2) {
  constexpr int MMA_WARP = NEPI;
  constexpr int PROD_WARP0 = NEPI + 1;
  constexpr int NTHREADS = (NEPI + 1 + NPROD) * 32;
  constexpr int NWG = NEPI / 4;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const GemmArgs& g = P.g;
  TileSrc a = g.a;
  // warp index through a shuffle: the compiler then knows it is warp-uniform and keeps everything derived from it
  // (stage cursors, descriptors, barrier addresses) in uniform registers — tcgen05.mma / TMA take uniform operands,
  // and values of unknown uniformity cost a R2UR waterfall loop around every single instruction
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const bool dbg_on = P.dbg != nullptr && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0;
  long long d_t0 = dbg_on ? clock64() : 0, d_a = 0, d_b = 0, d_c = 0, d_n = 0;
#define DBG_T(acc, stmt) do { if (dbg_on) { const long long t_ = clock64(); stmt; acc += clock64() - t_; } else { stmt; } } while (0)
  const int n0 = blockIdx.y * P.NpB;
  const int K = a.K;                      // multiple of 8
  const int NpB = P.NpB;
  const int nchunks = (K + KC - 1) / KC;

  // ---- shared memory carve-up ----
  float* B_hi = reinterpret_cast<float*>(smem_raw);                 // [K/4][NpB/8][8][4]
  float* B_lo = P.ncat ? B_hi + (size_t)NpB * 4 : B_hi + (size_t)NpB * K;   // ncat: [K/4][hi: NpB/8 | lo: NpB/8][8][4]
  float* stages = B_hi + (size_t)2 * NpB * K;                       // nstage x (A_hi, A_lo)
  float* epi = stages + (size_t)P.nstage * 2 * STAGE_FLOATS;        // NEPI warps x epi_bufs x [32][EPI_LD]
  float* s_stat = epi + NEPI * P.epi_bufs * 32 * EPI_LD;            // [NEPI warps][2][NpA]
  float* s_epar = s_stat + NEPI * 2 * P.NpA;                       // [6][NpA] Swish-backward epilogue: mean, rstd, scale, beta of
                                                                    // this CTA's columns + the SE gates of its first two samples
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_epar + 6 * P.NpA);
  uint64_t* full = bars;                  // [nstage]
  uint64_t* empty = bars + P.nstage;      // [nstage]
  uint64_t* rawfull = empty + P.nstage;   // [nstage] TMA bytes of the raw rows have landed
  uint64_t* tfull = rawfull + P.nstage;   // [2]  accumulator set ready for its epilogue warpgroup
  uint64_t* tempty = tfull + 2;           // [2]  accumulator set drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  // ---- one-time setup ----
  if (threadIdx.x == 0) {
    for (int i = 0; i < P.nstage; ++i) {
      mbar_init(smem_u32(full + i), WPS); mbar_init(smem_u32(empty + i), 1); mbar_init(smem_u32(rawfull + i), 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(tfull + i), 1); mbar_init(smem_u32(tempty + i), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) tmem_alloc(smem_u32(tmem_slot), (uint32_t)P.tmem_cols);

  const long long ntiles = (g.M + BM - 1) / BM;
  const long long tpc = (ntiles + gridDim.x - 1) / gridDim.x;
  const long long t_begin = (long long)blockIdx.x * tpc;
  const long long t_end = t_begin + tpc < ntiles ? t_begin + tpc : ntiles;
  const int my_tiles = t_end > t_begin ? (int)(t_end - t_begin) : 0;
  const uint32_t stages_u32 = smem_u32(stages);
  constexpr uint32_t STAGE_BYTES = 2 * STAGE_FLOATS * 4;     // hi + lo

  // ---- producer cursors and the TMA issue step (used by the producer warps only) ----
  // Chunk c of this CTA (tile-major, then K chunk) uses stage c % nstage.  nstage is a multiple of nprod, so warp p
  // handles chunks p, p + nprod, ... and cycles through its own stages p, p + nprod, ...; (tile, chunk, stage, phase)
  // advance incrementally — no divisions in the loop.
  // with WPS > 1 the warps of a group share the group's stages; warp `half` 0 issues the TMA copies
  const int p = warp >= PROD_WARP0 ? (warp - PROD_WARP0) / WPS : 0, half = warp >= PROD_WARP0 ? (warp - PROD_WARP0) % WPS : 0;
  const bool is_prod = warp >= PROD_WARP0 && p < P.nprod && my_tiles > 0;
  const int own = P.nstage / P.nprod;
  const bool has2 = (a.mode == PRO_BNBWD || a.mode == PRO_ABSDIFF || a.mode == PRO_MASK_POS);
  struct Cursor { int ti, chunk, stage; uint32_t phase; };
  auto advance = [&](Cursor& c) {
    c.chunk += P.nprod;
    while (c.chunk >= nchunks) { c.chunk -= nchunks; ++c.ti; }
    c.stage += P.nprod;
    if (c.stage >= P.nstage) { c.stage -= P.nstage; c.phase ^= 1u; }
  };
  auto issue = [&](const Cursor& c) {      // wait for the stage to drain, then start the TMA copy of the raw rows
    DBG_T(d_a, mbar_wait(smem_u32(empty + c.stage), c.phase ^ 1u, (uint32_t)P.wait_hint));
    if (lane == 0) {
      const int r0 = (int)((t_begin + c.ti) * BM);
      const uint32_t dst = stages_u32 + (uint32_t)c.stage * STAGE_BYTES;
      const uint32_t bar = smem_u32(rawfull + c.stage);
      mbar_expect_tx(bar, (uint32_t)(STAGE_FLOATS * 4 * (has2 ? 2 : 1)));
      tma_load_2d(dst, &tmA, c.chunk * KC, r0, bar);
      if (has2) tma_load_2d(dst + STAGE_FLOATS * 4, &tmA2, c.chunk * KC, r0, bar);
      if (P.pf_dist) {
        // the first tile also requests the tiles in between (nothing was requested for them yet)
        for (int pt = c.ti == 0 ? 1 : c.ti + P.pf_dist; pt <= c.ti + P.pf_dist && pt < my_tiles; ++pt) {
          const int pr = (int)((t_begin + pt) * BM);
          tma_prefetch_2d(&tmA, c.chunk * KC, pr);
          if (has2) tma_prefetch_2d(&tmA2, c.chunk * KC, pr);
        }
      }
    }
    __syncwarp();
  };
  Cursor cur, nxt;
  cur.ti = 0; cur.chunk = p; cur.stage = p; cur.phase = 0;
  while (cur.chunk >= nchunks) { cur.chunk -= nchunks; ++cur.ti; }
  nxt = cur;
  auto pre_issue = [&]() {                 // the first own - 1 stages of this producer warp
    for (int i = 0; i < own - 1; ++i) { if (nxt.ti < my_tiles && half == 0) issue(nxt); advance(nxt); }
  };
  const bool early = P.tma && P.early;
  if (early) {
    __syncthreads();                       // the mbarrier inits are visible to the lanes that arm them
    if (is_prod) {
      pdl_wait();                          // the operand rows are an earlier kernel's output
      pre_issue();                         // in flight while all warps stage the weights below
    }
  }

  // Programmatic dependent launch: everything up to here touches no global memory.  Model parameters are constant
  // within a step, so their staging below may also overlap the previous kernel's tail; other weights wait first.
  if (!g.w_const) pdl_wait();
  const long long d_tw0 = dbg_on ? clock64() : 0;
  // resident weights, split and laid out for UMMA (zero outside the logical matrix)
  {
    // Batches of WU quads per thread with every global load of a batch issued before the first split / store: the
    // staging is L2-latency bound (a handful of dependent round trips per thread used to cost ~9-12 K cycles per
    // launch, profiles/r02_summary.md), so the loads of a batch must be in flight together.
    const uint32_t kq4 = (uint32_t)K >> 2;
    const uint32_t total = (uint32_t)NpB * kq4;
    const float* Wg = g.W;
    const bool k_contig = g.w_sr == 1;
    const bool vec_ok = k_contig && (g.w_so & 3) == 0 && (g.Kred & 3) == 0 && ((reinterpret_cast<uintptr_t>(Wg) & 15) == 0);
    constexpr int WU = 8;
    for (uint32_t base = threadIdx.x; base < total; base += NTHREADS * WU) {
      float4 wv[WU];
      int oo[WU];
#pragma unroll
      for (int u = 0; u < WU; ++u) {
        const uint32_t idx = base + (uint32_t)u * NTHREADS;
        wv[u] = f4zero();
        oo[u] = -1;
        if (idx < total) {
          uint32_t n, q;
          // lanes = 8 consecutive columns x 4 consecutive K quads: the shared stores of a quarter warp then cover 128
          // contiguous bytes (with K fastest all 32 lanes of a store hit one bank group -- a 32-way conflict that cost
          // 8 K cycles per launch for the 112 x 96 block), and the global reads are still 64-byte runs per column
          if (k_contig) { const uint32_t t = idx >> 3; const uint32_t nh = t / kq4; q = t - nh * kq4; n = nh * 8 + (idx & 7); }
          else { q = idx / (uint32_t)NpB; n = idx - q * (uint32_t)NpB; }
          oo[u] = (int)(((q * (uint32_t)((P.ncat ? 2 * NpB : NpB) >> 3) + (n >> 3)) * 8 + (n & 7)) * 4);
          if ((int)(n0 + n) < g.N) {
            const float* src = Wg + (long long)(4 * q) * g.w_sr + (long long)(n0 + n) * g.w_so;
            if (vec_ok) {
              if ((int)(4 * q) < g.Kred) wv[u] = ldg4(src);
            } else {
              const int red = (int)(4 * q);
              if (red + 0 < g.Kred) wv[u].x = __ldg(src);
              if (red + 1 < g.Kred) wv[u].y = __ldg(src + (long long)g.w_sr);
              if (red + 2 < g.Kred) wv[u].z = __ldg(src + 2LL * g.w_sr);
              if (red + 3 < g.Kred) wv[u].w = __ldg(src + 3LL * g.w_sr);
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < WU; ++u) {
        if (oo[u] >= 0) {
          float4 hi, lo;
          split4(wv[u], hi, lo);
          *reinterpret_cast<float4*>(B_hi + oo[u]) = hi;
          *reinterpret_cast<float4*>(B_lo + oo[u]) = lo;
        }
      }
    }
    for (int i = threadIdx.x; i < NEPI * 2 * P.NpA; i += NTHREADS) s_stat[i] = 0.f;
  }
  // C3D_TC_EPAR=1: Swish-backward epilogue parameters of this CTA's columns staged once in shared memory instead of being
  // fetched (L1 / L2) in every (tile, column chunk) step; the SE gates of the CTA's first two batch samples ride along
  // (a CTA's tiles are consecutive rows), later samples fall back to global loads.  Measured neutral (21.62 / 21.65 vs
  // 21.62 / 21.59 ms over the family) although removing the loads altogether is worth 0.16 ms: kept as a switch, off.
  const uint32_t esamp0 = (uint32_t)(t_begin * BM) / P.rps;
  if (g.epi == EPI_SWISH_BWD && P.epar) {
    pdl_wait();                              // BN blocks / gates are an earlier kernel's output
    for (int i = threadIdx.x; i < 6 * P.NpA; i += NTHREADS) {
      const int r = i / P.NpA, cidx = i - r * P.NpA, col = n0 + cidx;
      float v = r == 1 || r == 2 || r >= 4 ? 1.f : 0.f;
      if (cidx < NpB && col < g.Ns) {
        if (r == 0) v = __ldg(BNP_MEAN(g.ebnp, g.Ns) + col);
        else if (r == 1) v = __ldg(BNP_RSTD(g.ebnp, g.Ns) + col);
        else if (r == 2) v = __ldg(BNP_SCALE(g.ebnp, g.Ns) + col);
        else if (r == 3) v = __ldg(BNP_BETA(g.ebnp, g.Ns) + col);
        else if (g.egate && (long long)(esamp0 + (uint32_t)(r - 4)) * P.rps < g.M)
          v = __ldg(g.egate + (long long)(esamp0 + (uint32_t)(r - 4)) * g.Ns + col);
      }
      s_epar[i] = v;
    }
  }
  const long long d_tw1 = dbg_on ? clock64() : 0;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const long long d_t1 = dbg_on ? clock64() : 0;
  pdl_trigger();     // tensor memory is allocated: the next kernel's CTAs may start their own setup
  pdl_wait();        // operands, BN blocks, statistics: produced by earlier kernels

  if (warp >= PROD_WARP0) {
    // ===================== producers =====================
    const int bg0 = half * (16 / WPS), bg1 = bg0 + 16 / WPS;          // this warp's 8-row groups of a chunk
    if (is_prod) {
      if (P.tma && !early) pre_issue();
      while (cur.ti < my_tiles) {
        float* a_hi = stages + (size_t)cur.stage * 2 * STAGE_FLOATS;
        float* a_lo = a_hi + STAGE_FLOATS;
        const long long row0 = (t_begin + cur.ti) * BM;
        if (P.tma) {
          if (nxt.ti < my_tiles && half == 0) issue(nxt);
          advance(nxt);
          DBG_T(d_b, mbar_wait(smem_u32(rawfull + cur.stage), cur.phase, (uint32_t)P.wait_hint));
          ++d_n;
          const long long t_x = dbg_on ? clock64() : 0;
          switch (a.mode) {
            case PRO_NONE: transform_chunk<PRO_NONE>(P, a, row0, cur.chunk, a_hi, a_lo, lane, bg0, bg1); break;
            case PRO_BN_RELU: transform_chunk<PRO_BN_RELU>(P, a, row0, cur.chunk, a_hi, a_lo, lane, bg0, bg1); break;
            case PRO_BN_GATE_SWISH: transform_chunk<PRO_BN_GATE_SWISH>(P, a, row0, cur.chunk, a_hi, a_lo, lane, bg0, bg1); break;
            case PRO_BNBWD: transform_chunk<PRO_BNBWD>(P, a, row0, cur.chunk, a_hi, a_lo, lane, bg0, bg1); break;
            case PRO_ABSDIFF: transform_chunk<PRO_ABSDIFF>(P, a, row0, cur.chunk, a_hi, a_lo, lane, bg0, bg1); break;
            default: transform_chunk<PRO_MASK_POS>(P, a, row0, cur.chunk, a_hi, a_lo, lane, bg0, bg1); break;
          }
          if (dbg_on) d_c += clock64() - t_x;
        } else {
          DBG_T(d_a, mbar_wait(smem_u32(empty + cur.stage), cur.phase ^ 1u, (uint32_t)P.wait_hint));
          ++d_n;
          const long long t_x = dbg_on ? clock64() : 0;
          switch (a.mode) {
            case PRO_NONE: produce_chunk<PRO_NONE>(P, a, row0, cur.chunk, a_hi, a_lo, lane, bg0, bg1); break;
            case PRO_BN_RELU: produce_chunk<PRO_BN_RELU>(P, a, row0, cur.chunk, a_hi, a_lo, lane, bg0, bg1); break;
            case PRO_BN_GATE_SWISH: produce_chunk<PRO_BN_GATE_SWISH>(P, a, row0, cur.chunk, a_hi, a_lo, lane, bg0, bg1); break;
            case PRO_BNBWD: produce_chunk<PRO_BNBWD>(P, a, row0, cur.chunk, a_hi, a_lo, lane, bg0, bg1); break;
            case PRO_ABSDIFF: produce_chunk<PRO_ABSDIFF>(P, a, row0, cur.chunk, a_hi, a_lo, lane, bg0, bg1); break;
            default: produce_chunk<PRO_MASK_POS>(P, a, row0, cur.chunk, a_hi, a_lo, lane, bg0, bg1); break;
          }
          if (dbg_on) d_c += clock64() - t_x;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(full + cur.stage));
        advance(cur);
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================== MMA issuer =====================
    // The whole warp runs the loop (uniform control flow, uniform operands); lane 0 alone issues the MMAs and
    // commits.  Everything that does not change per chunk is hoisted (descriptor templates, barrier addresses) and
    // the stage / phase counters advance incrementally: this instruction stream is the pipeline's pacemaker.
    {
      const uint32_t leader = lane == 0 ? 1u : 0u;
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NpB >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t lboA = 16 * 128, sboA = 128;                       // stage layout [k-chunk][row group][8][16 B]
      const uint32_t lboB = (uint32_t)((P.ncat ? 2 * NpB : NpB) >> 3) * 128, sboB = 128;     // weights     [k-chunk][col group][8][16 B]
      const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * NpB) >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

      // descriptor templates with a zero start address; the address field (bits 0-13, 16-byte units) is added per use
      uint64_t tA, tB;
      uint32_t a_kstep;                                                  // byte offset of the second K step inside a stage half
      if (P.tma) { tA = make_desc_sw64(0, 512); a_kstep = 32; }
      else if (P.lbo_is_k) { tA = make_desc(0, lboA, sboA); a_kstep = 2 * lboA; }
      else { tA = make_desc(0, sboA, lboA); a_kstep = 2 * lboA; }
      tB = (P.tma || P.lbo_is_k) ? make_desc(0, lboB, sboB) : make_desc(0, sboB, lboB);
      const uint64_t dB_hi0 = tB + (uint64_t)(smem_u32(B_hi) >> 4), dB_lo0 = tB + (uint64_t)(smem_u32(B_lo) >> 4);
      const uint32_t b_kstep16 = (2 * lboB) >> 4;                       // descriptor units per K step of 8
      const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);
      int stage = 0;
      uint32_t phase = 0;
      for (int ti = 0; ti < my_tiles; ++ti) {
        const int set = ti & 1;
        DBG_T(d_a, mbar_wait(smem_u32(tempty + set), ((uint32_t)(ti >> 1) & 1) ^ 1, (uint32_t)P.wait_hint));
        ++d_n;
        tc_fence_after();
        const uint32_t d_main = tmem_base + (uint32_t)(set * 2 * P.NpA);
        const uint32_t d_corr = d_main + (uint32_t)(P.ncat ? NpB : P.NpA);
        uint64_t dbh = dB_hi0, dbl = dB_lo0;
        int kleft = K;
        uint32_t accum = 0;
        for (int chunk = 0; chunk < nchunks; ++chunk, kleft -= KC) {
          DBG_T(d_b, mbar_wait(full0 + (uint32_t)stage * 8, phase, (uint32_t)P.wait_hint));
          tc_fence_after();
          const uint32_t ahi = stages_u32 + (uint32_t)stage * STAGE_BYTES;
          const uint64_t dah = tA + (uint64_t)(ahi >> 4), dal = dah + (uint64_t)((STAGE_FLOATS * 4) >> 4);
          const uint32_t two = (leader && kleft >= KC) ? 1u : 0u;
          const uint64_t dah2 = dah + (a_kstep >> 4), dal2 = dal + (a_kstep >> 4), dbh2 = dbh + b_kstep16, dbl2 = dbl + b_kstep16;
          // folded: [main | corr] (+)= Ah x [Wh | Wl], corr += Al x Wh;  round-1 scheme: three products per K step
          if (P.ncat) {
            umma_tf32_if(leader, d_main, dah, dbh, idesc2, accum);
            umma_tf32_if(leader, d_corr, dal, dbh, idesc, 1u);
            umma_tf32_if(two, d_main, dah2, dbh2, idesc2, 1u);
            umma_tf32_if(two, d_corr, dal2, dbh2, idesc, 1u);
          } else {
            umma_tf32_if(leader, d_main, dah, dbh, idesc, accum);
            umma_tf32_if(leader, d_corr, dal, dbh, idesc, accum);
            umma_tf32_if(leader, d_corr, dah, dbl, idesc, 1u);
            umma_tf32_if(two, d_main, dah2, dbh2, idesc, 1u);
            umma_tf32_if(two, d_corr, dal2, dbh2, idesc, 1u);
            umma_tf32_if(two, d_corr, dah2, dbl2, idesc, 1u);
          }
          umma_commit_if(leader, empty0 + (uint32_t)stage * 8);
          accum = 1u;
          dbh += 2 * b_kstep16; dbl += 2 * b_kstep16;
          if (++stage == P.nstage) { stage = 0; phase ^= 1u; }
        }
        umma_commit_if(leader, smem_u32(tfull + set));
      }
    }
  } else {
    // ===================== epilogue: warpgroup `set` (warps 4*set .. 4*set+3) handles tiles ti % 2 == set ===========
    const int wg = warp >> 2, wq = warp & 3;             // wq = TMEM lane quarter this warp may read
    float* S = epi + (size_t)warp * P.epi_bufs * 32 * EPI_LD;
    float* S2 = S + 32 * EPI_LD;      // only valid when P.epi_bufs == 2
    float* st = s_stat + (size_t)warp * 2 * P.NpA;
    const bool has_stats = g.stats != nullptr;
    const int ncc = (NpB + 31) / 32;
    const int bar_id = 1 + wg;
    long long cur_samp = -1;
    auto flush = [&](long long samp) {
      // the four warps of this warpgroup add their partial sums into the global statistics of `samp`
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      const float* wgs = s_stat + (size_t)(wg * 4) * 2 * P.NpA;
      for (int cidx = (threadIdx.x & 127); cidx < NpB; cidx += 128) {
        if (n0 + cidx < g.Ns) {
          float t1 = 0.f, t2 = 0.f;
#pragma unroll
          for (int w = 0; w < 4; ++w) { t1 += wgs[(w * 2 + 0) * P.NpA + cidx]; t2 += wgs[(w * 2 + 1) * P.NpA + cidx]; }
          atomicAdd(g.stats + (samp * 2 + 0) * g.Ns + n0 + cidx, (double)t1);
          atomicAdd(g.stats + (samp * 2 + 1) * g.Ns + n0 + cidx, (double)t2);
        }
      }
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      for (int i = lane; i < 2 * P.NpA; i += 32) st[i] = 0.f;
      __syncwarp();
    };
    // Fused epilogues read a second tensor (E1: y_b for the Swish backward, the residual gradient for the join).
    // Those reads do not depend on the accumulator, so they run one (tile, column chunk) step ahead: the loads of
    // the next step are in flight while this one waits for its MMAs and does its arithmetic.
    const bool pre_e1 = (g.epi == EPI_SWISH_BWD || g.epi == EPI_ADD2) && g.E1 != nullptr;
    float4 e1n[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) e1n[i] = f4zero();
    auto prefetch_e1 = [&](long long ti_, int cc_) {
      const long long wr0 = (t_begin + ti_) * BM + wq * 32;
      const int cl_ = cc_ * 32 + 4 * (lane & 7);
      const int col_ = n0 + cl_;
      const bool ok_ = col_ < g.Ns && cl_ < NpB;
      const float* src = g.E1 + (wr0 + (lane >> 3)) * (long long)g.Ns + col_;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = 4 * i + (lane >> 3);
        e1n[i] = (ok_ && wr0 + rr < g.M) ? ldg4(src + (long long)(4 * i) * g.Ns) : f4zero();
      }
    };
    if (pre_e1 && wg < my_tiles) prefetch_e1(wg, 0);
    for (long long ti = wg; ti < my_tiles; ti += NWG) {
      const int set = (int)(ti & 1);
      const long long row0 = (t_begin + ti) * BM;
      const long long last = row0 + BM - 1 < g.M - 1 ? row0 + BM - 1 : g.M - 1;
      // 32-bit divisions only: a 64-bit division compiles to a subroutine call, after which ptxas no longer keeps the
      // MMA warp's operands in uniform registers (every tcgen05.mma then sits in a R2UR waterfall loop)
      const long long samp0 = (long long)((uint32_t)row0 / P.rps);
      const bool straddle = (long long)((uint32_t)last / P.rps) != samp0;
      const bool full_tile = (row0 + BM <= g.M);
      // sample of this warp's first row and the first of its 32 rows that belongs to the next sample (>= 32: none)
      const long long wrow0 = row0 + wq * 32;
      const bool sp_fast = P.rps >= 32u;
      long long wsp0 = samp0;
      int wsplit = 32;
      if (g.epi == EPI_SWISH_BWD && sp_fast) {
        if (straddle) wsp0 = (long long)((uint32_t)wrow0 / P.rps);
        const long long nb = (wsp0 + 1) * (long long)P.rps - wrow0;
        wsplit = nb < 32 ? (int)nb : 32;
      }
      if (has_stats && samp0 != cur_samp) {
        if (cur_samp >= 0) flush(cur_samp);
        cur_samp = samp0;
      }
      DBG_T(d_a, mbar_wait(smem_u32(tfull + set), (uint32_t)(ti >> 1) & 1, (uint32_t)P.wait_hint));
      ++d_n;
      tc_fence_after();
      const uint32_t t_main = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(set * 2 * P.NpA);
      const uint32_t t_corr = t_main + (uint32_t)(P.ncat ? NpB : P.NpA);
      for (int cc = 0; cc < ncc; ++cc) {
        {
          float r[32], r2[32];
          tmem_ld32x2(t_main + (uint32_t)(cc * 32), t_corr + (uint32_t)(cc * 32), r, r2);
          if (cc == ncc - 1) {     // both accumulators fully read by this warp: hand the set back to the MMA issuer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(tempty + set));
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(S + lane * EPI_LD + 4 * j) =
                make_float4(r[4 * j] + r2[4 * j], r[4 * j + 1] + r2[4 * j + 1], r[4 * j + 2] + r2[4 * j + 2], r[4 * j + 3] + r2[4 * j + 3]);
        }
        __syncwarp();
        // coalesced pass: 8 lanes cover the 32 columns of a row, 4 rows per instruction
        const int q = lane & 7;
        const int cl = cc * 32 + 4 * q;          // column inside this CTA's accumulator
        const int col = n0 + cl;
        const bool col_ok = col < g.Ns && cl < NpB;
        const long long tile_o = (wrow0 + (lane >> 3)) * (long long)g.Ns + col;
        if (g.epi == EPI_STORE) {
          float4 sv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) sv[i] = *reinterpret_cast<const float4*>(S + (4 * i + (lane >> 3)) * EPI_LD + 4 * q);
          if (col_ok) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = 4 * i + (lane >> 3);
              if (full_tile || row0 + wq * 32 + rr < g.M) st4(g.Y + tile_o + (long long)(4 * i) * g.Ns, sv[i]);
            }
          }
          if (has_stats) {
            // column sums from the quads already in registers (rows past M and pad columns hold zeros): 8 rows per
            // lane, then the four lanes that share a column quad are folded -- no second pass over the staged tile
            // (shared-memory bandwidth, not issue slots, is what these kernels run out of: profiles/r02_summary.md)
            float4 t1 = f4zero(), t2 = f4zero();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              t1 = f4add(t1, sv[i]);
              t2.x = fmaf(sv[i].x, sv[i].x, t2.x); t2.y = fmaf(sv[i].y, sv[i].y, t2.y);
              t2.z = fmaf(sv[i].z, sv[i].z, t2.z); t2.w = fmaf(sv[i].w, sv[i].w, t2.w);
            }
#pragma unroll
            for (int o = 8; o <= 16; o <<= 1) {
              t1.x += __shfl_xor_sync(0xffffffffu, t1.x, o); t1.y += __shfl_xor_sync(0xffffffffu, t1.y, o);
              t1.z += __shfl_xor_sync(0xffffffffu, t1.z, o); t1.w += __shfl_xor_sync(0xffffffffu, t1.w, o);
              t2.x += __shfl_xor_sync(0xffffffffu, t2.x, o); t2.y += __shfl_xor_sync(0xffffffffu, t2.y, o);
              t2.z += __shfl_xor_sync(0xffffffffu, t2.z, o); t2.w += __shfl_xor_sync(0xffffffffu, t2.w, o);
            }
            if (lane < 8) {
              float4* s1 = reinterpret_cast<float4*>(st + cc * 32 + 4 * q);
              float4* s2 = reinterpret_cast<float4*>(st + P.NpA + cc * 32 + 4 * q);
              *s1 = f4add(*s1, t1);
              *s2 = f4add(*s2, t2);
            }
          }
        } else {
          float4 mean = f4zero(), rstd = f4zero(), scale = f4zero(), beta = f4zero();
          float4 gt0 = make_float4(1.f, 1.f, 1.f, 1.f), gt1 = gt0;
          if (g.epi == EPI_SWISH_BWD && col_ok && !P.epar) {
            mean = ldg4(BNP_MEAN(g.ebnp, g.Ns) + col); rstd = ldg4(BNP_RSTD(g.ebnp, g.Ns) + col);
            scale = ldg4(BNP_SCALE(g.ebnp, g.Ns) + col); beta = ldg4(BNP_BETA(g.ebnp, g.Ns) + col);
            if (g.egate && sp_fast) {
              gt0 = ldg4(g.egate + wsp0 * g.Ns + col);
              if (wsplit < 32 && wrow0 + wsplit < g.M) gt1 = ldg4(g.egate + (wsp0 + 1) * g.Ns + col);
            }
          } else if (g.epi == EPI_SWISH_BWD && col_ok) {
            const float* ep = s_epar + cl;                    // staged at setup: [mean | rstd | scale | beta | gate s0 | gate s0 + 1]
            mean = *reinterpret_cast<const float4*>(ep); rstd = *reinterpret_cast<const float4*>(ep + P.NpA);
            scale = *reinterpret_cast<const float4*>(ep + 2 * P.NpA); beta = *reinterpret_cast<const float4*>(ep + 3 * P.NpA);
            if (g.egate && sp_fast) {
              const long long rel = wsp0 - (long long)esamp0;
              if (rel >= 0 && rel <= 1) gt0 = *reinterpret_cast<const float4*>(ep + (4 + rel) * P.NpA);
              else gt0 = ldg4(g.egate + wsp0 * g.Ns + col);
              if (wsplit < 32 && wrow0 + wsplit < g.M) {
                if (rel == 0) gt1 = *reinterpret_cast<const float4*>(ep + 5 * P.NpA);
                else gt1 = ldg4(g.egate + (wsp0 + 1) * g.Ns + col);
              }
            }
          }
          // This step's E1 values were requested one step ago (e1n).  The next step's are requested quad by quad right
          // after the matching quad has been consumed, so no second register copy of the 8 quads is needed.
          long long nti = -1;
          int ncc_next = 0;
          if (pre_e1) {
            if (cc + 1 < ncc) { nti = ti; ncc_next = cc + 1; }
            else if (ti + NWG < my_tiles) { nti = ti + NWG; ncc_next = 0; }
          }
          const long long nwr0 = nti >= 0 ? (t_begin + nti) * BM + wq * 32 : 0;
          const int ncl = ncc_next * 32 + 4 * (lane & 7);
          const bool nok = nti >= 0 && n0 + ncl < g.Ns && ncl < NpB;
          const float* nsrc = pre_e1 ? g.E1 + (nwr0 + (lane >> 3)) * (long long)g.Ns + n0 + ncl : nullptr;
#define E1_TAKE(i, dst)                                                                                          \
          const float4 dst = e1n[i];                                                                            \
          e1n[i] = (nok && nwr0 + 4 * (i) + (lane >> 3) < g.M) ? ldg4(nsrc + (long long)(4 * (i)) * g.Ns) : f4zero();
          // the staged accumulator rows first: all shared loads issue back to back and the 8 row quads below are
          // independent instruction streams (no shared-memory stores between them)
          float4 vv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) vv[i] = *reinterpret_cast<const float4*>(S + (4 * i + (lane >> 3)) * EPI_LD + 4 * q);
          if (g.epi == EPI_SWISH_BWD && sp_fast) {
            // per-lane partial sums of (du, du * zhat) for the warp's first sample (a0, z0) and, when its 32 rows
            // straddle a sample boundary, the next one (a1, z1)
            float4 a0 = f4zero(), z0 = f4zero(), a1 = f4zero(), z1 = f4zero();
            if (wsplit >= 32 && wrow0 + 32 <= g.M) {
              // common case: the warp's 32 rows are all valid and belong to one sample -- no per-row predicates
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                E1_TAKE(i, yb)
                const float4 acc = vv[i];
                float4 du, dz;
                {
                  const float xc = yb.x - mean.x, u = fmaf(xc, scale.x, beta.x) * gt0.x;
                  du.x = acc.x * swish_gradf_(u); dz.x = du.x * (xc * rstd.x);
                }
                {
                  const float xc = yb.y - mean.y, u = fmaf(xc, scale.y, beta.y) * gt0.y;
                  du.y = acc.y * swish_gradf_(u); dz.y = du.y * (xc * rstd.y);
                }
                {
                  const float xc = yb.z - mean.z, u = fmaf(xc, scale.z, beta.z) * gt0.z;
                  du.z = acc.z * swish_gradf_(u); dz.z = du.z * (xc * rstd.z);
                }
                {
                  const float xc = yb.w - mean.w, u = fmaf(xc, scale.w, beta.w) * gt0.w;
                  du.w = acc.w * swish_gradf_(u); dz.w = du.w * (xc * rstd.w);
                }
                if (col_ok) {
                  st4(g.Y + tile_o + (long long)(4 * i) * g.Ns, du);
                  a0 = f4add(a0, du); z0 = f4add(z0, dz);
                }
              }
            } else
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = 4 * i + (lane >> 3);
              const bool valid = col_ok && wrow0 + rr < g.M;
              const bool second = rr >= wsplit;
              E1_TAKE(i, yb)
              const float4 gt = second ? gt1 : gt0, acc = vv[i];
              float4 du, dz;
              {
                const float xc = yb.x - mean.x, zh = xc * rstd.x, u = fmaf(xc, scale.x, beta.x) * gt.x;
                du.x = acc.x * swish_gradf_(u); dz.x = du.x * zh;
              }
              {
                const float xc = yb.y - mean.y, zh = xc * rstd.y, u = fmaf(xc, scale.y, beta.y) * gt.y;
                du.y = acc.y * swish_gradf_(u); dz.y = du.y * zh;
              }
              {
                const float xc = yb.z - mean.z, zh = xc * rstd.z, u = fmaf(xc, scale.z, beta.z) * gt.z;
                du.z = acc.z * swish_gradf_(u); dz.z = du.z * zh;
              }
              {
                const float xc = yb.w - mean.w, zh = xc * rstd.w, u = fmaf(xc, scale.w, beta.w) * gt.w;
                du.w = acc.w * swish_gradf_(u); dz.w = du.w * zh;
              }
              if (valid) {
                st4(g.Y + tile_o + (long long)(4 * i) * g.Ns, du);
                if (second) { a1 = f4add(a1, du); z1 = f4add(z1, dz); } else { a0 = f4add(a0, du); z0 = f4add(z0, dz); }
              }
            }
            if (has_stats) {
              // fold the four lanes that share a column quad (lanes q, q + 8, q + 16, q + 24)
              auto fold = [](float4& t) {
                t.x += __shfl_xor_sync(0xffffffffu, t.x, 8); t.y += __shfl_xor_sync(0xffffffffu, t.y, 8);
                t.z += __shfl_xor_sync(0xffffffffu, t.z, 8); t.w += __shfl_xor_sync(0xffffffffu, t.w, 8);
                t.x += __shfl_xor_sync(0xffffffffu, t.x, 16); t.y += __shfl_xor_sync(0xffffffffu, t.y, 16);
                t.z += __shfl_xor_sync(0xffffffffu, t.z, 16); t.w += __shfl_xor_sync(0xffffffffu, t.w, 16);
              };
              fold(a0); fold(z0);
              if (!straddle) {
                if (lane < 8) {
                  float* s1 = st + cc * 32 + 4 * q;
                  float* s2 = st + P.NpA + cc * 32 + 4 * q;
                  s1[0] += a0.x; s1[1] += a0.y; s1[2] += a0.z; s1[3] += a0.w;
                  s2[0] += z0.x; s2[1] += z0.y; s2[2] += z0.z; s2[3] += z0.w;
                }
              } else {       // tile straddles two samples: this warp's sums go straight to the global statistics
                fold(a1); fold(z1);
                if (lane < 8 && col_ok) {
                  double* d0 = g.stats + (wsp0 * 2) * g.Ns + col;
                  atomicAdd(d0 + 0, (double)a0.x); atomicAdd(d0 + 1, (double)a0.y); atomicAdd(d0 + 2, (double)a0.z); atomicAdd(d0 + 3, (double)a0.w);
                  atomicAdd(d0 + g.Ns + 0, (double)z0.x); atomicAdd(d0 + g.Ns + 1, (double)z0.y);
                  atomicAdd(d0 + g.Ns + 2, (double)z0.z); atomicAdd(d0 + g.Ns + 3, (double)z0.w);
                  if (wsplit < 32 && wrow0 + wsplit < g.M) {
                    double* d1 = d0 + 2 * g.Ns;
                    atomicAdd(d1 + 0, (double)a1.x); atomicAdd(d1 + 1, (double)a1.y); atomicAdd(d1 + 2, (double)a1.z); atomicAdd(d1 + 3, (double)a1.w);
                    atomicAdd(d1 + g.Ns + 0, (double)z1.x); atomicAdd(d1 + g.Ns + 1, (double)z1.y);
                    atomicAdd(d1 + g.Ns + 2, (double)z1.z); atomicAdd(d1 + g.Ns + 3, (double)z1.w);
                  }
                }
              }
            }
          } else if (g.epi == EPI_SWISH_BWD) {
            // samples shorter than a warp's 32 rows (tiny images): per-element statistics atomics
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = 4 * i + (lane >> 3);
              const long long row = wrow0 + rr;
              if (row < g.M && col_ok) {
                const long long sp = (long long)((uint32_t)row / P.rps);
                float4 gt = make_float4(1.f, 1.f, 1.f, 1.f);
                if (g.egate) gt = ldg4(g.egate + sp * g.Ns + col);
                const float4 yb = e1n[i];
                float yv[4] = {yb.x, yb.y, yb.z, yb.w}, mv[4] = {mean.x, mean.y, mean.z, mean.w};
                float rv[4] = {rstd.x, rstd.y, rstd.z, rstd.w}, sv[4] = {scale.x, scale.y, scale.z, scale.w};
                float bv[4] = {beta.x, beta.y, beta.z, beta.w}, gv[4] = {gt.x, gt.y, gt.z, gt.w};
                float av[4] = {vv[i].x, vv[i].y, vv[i].z, vv[i].w}, du[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float xc = yv[j] - mv[j];
                  const float zh = xc * rv[j];
                  const float u = fmaf(xc, sv[j], bv[j]) * gv[j];
                  du[j] = av[j] * swish_gradf_(u);
                  if (has_stats) {
                    atomicAdd(g.stats + (sp * 2 + 0) * g.Ns + col + j, (double)du[j]);
                    atomicAdd(g.stats + (sp * 2 + 1) * g.Ns + col + j, (double)(du[j] * zh));
                  }
                }
                st4(g.Y + tile_o + (long long)(4 * i) * g.Ns, make_float4(du[0], du[1], du[2], du[3]));
              }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) { E1_TAKE(i, unused_) (void)unused_; }
          } else if (g.epi == EPI_RELU_ADD || g.epi == EPI_ABSDIFF_BWD) {
            // Encoder.enhance forward / backward (model/trainer.py:88-108): the output is a frame slice of a larger
            // tensor (out_img_stride between images), updated in place
            long long addr[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const long long row = wrow0 + 4 * i + (lane >> 3);
              addr[i] = -1;
              if (col_ok && row < g.M) {
                const uint32_t img = (uint32_t)row / (uint32_t)a.OHW;
                const uint32_t pix = (uint32_t)row - img * (uint32_t)a.OHW;
                addr[i] = (long long)pix * g.Ns + col + (long long)img * g.out_img_stride;
              }
            }
            if (g.epi == EPI_RELU_ADD) {
              float4 old[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) old[i] = addr[i] >= 0 ? *reinterpret_cast<const float4*>(g.Y + addr[i]) : f4zero();
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                if (addr[i] >= 0) {
                  if (g.Y2) st4(g.Y2 + tile_o + (long long)(4 * i) * g.Ns, old[i]);
                  st4(g.Y + addr[i], f4add(old[i], f4relu(vv[i])));
                }
              }
            } else {
              // backward of |x0 - x1|: Y (grad of x0) += sign * acc, Y2 (grad of x1) -= sign * acc, in place
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                float4 x0[4], x1[4], g0v[4], g1v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int i = 4 * h + j;
                  x0[j] = x1[j] = g0v[j] = g1v[j] = f4zero();
                  if (addr[i] >= 0) {
                    const long long row = wrow0 + 4 * i + (lane >> 3);
                    const uint32_t img = (uint32_t)row / (uint32_t)a.OHW;
                    const long long eaddr = addr[i] - (long long)img * g.out_img_stride + (long long)img * g.e1_img_stride;
                    x0[j] = ldg4(g.E1 + eaddr); x1[j] = ldg4(g.E2 + eaddr);
                    g0v[j] = *reinterpret_cast<const float4*>(g.Y + addr[i]);
                    g1v[j] = *reinterpret_cast<const float4*>(g.Y2 + addr[i]);
                  }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int i = 4 * h + j;
                  if (addr[i] >= 0) {
                    const float4 v = vv[i];
                    float4 t;
                    t.x = x0[j].x > x1[j].x ? v.x : (x0[j].x < x1[j].x ? -v.x : 0.f);
                    t.y = x0[j].y > x1[j].y ? v.y : (x0[j].y < x1[j].y ? -v.y : 0.f);
                    t.z = x0[j].z > x1[j].z ? v.z : (x0[j].z < x1[j].z ? -v.z : 0.f);
                    t.w = x0[j].w > x1[j].w ? v.w : (x0[j].w < x1[j].w ? -v.w : 0.f);
                    st4(g.Y + addr[i], f4add(g0v[j], t));
                    st4(g.Y2 + addr[i], make_float4(g1v[j].x - t.x, g1v[j].y - t.y, g1v[j].z - t.z, g1v[j].w - t.w));
                  }
                }
              }
            }
          } else {   // EPI_ADD2: residual join (+ the stride-2 shortcut gradient at even pixels)
            float4 e2[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              e2[i] = f4zero();
              const long long row = wrow0 + 4 * i + (lane >> 3);
              if (g.E2 && col_ok && row < g.M) {
                const uint32_t img = (uint32_t)row / (uint32_t)a.OHW;
                const uint32_t rem = (uint32_t)row - img * (uint32_t)a.OHW;
                const uint32_t oh = rem / (uint32_t)a.OW, ow = rem - oh * (uint32_t)a.OW;
                if (!(oh & 1) && !(ow & 1)) {
                  const long long hr = (long long)img * (a.OHW >> 2) + (long long)(oh >> 1) * (a.OW >> 1) + (ow >> 1);
                  e2[i] = ldg4(g.E2 + hr * g.Ns + col);
                }
              }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              E1_TAKE(i, res)
              if (col_ok && wrow0 + 4 * i + (lane >> 3) < g.M)
                st4(g.Y + tile_o + (long long)(4 * i) * g.Ns, f4add(f4add(vv[i], res), e2[i]));
            }
          }
#undef E1_TAKE
        }
        __syncwarp();
      }
    }
    if (has_stats && cur_samp >= 0) flush(cur_samp);
  }
  if (dbg_on && lane == 0) {
    unsigned long long* o = P.dbg + warp * 8;
    const long long t_end = clock64();
    o[0] = (unsigned long long)(d_t1 - d_t0); o[1] = (unsigned long long)(t_end - d_t1);
    o[2] = (unsigned long long)d_a; o[3] = (unsigned long long)d_b; o[4] = (unsigned long long)d_c; o[5] = (unsigned long long)d_n;
    o[6] = (unsigned long long)my_tiles; o[7] = (unsigned long long)(d_tw1 - d_tw0) | ((unsigned long long)(d_tw0 - d_t0) << 32);
  }
#undef DBG_T
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
  }
}

}  // namespace tc

// Size one kernel variant for this GEMM: N split across grid.y so that the accumulators fit `max_npa` TMEM columns
// each and the resident weights (both halves) fit in `budget` next to >= `min_stage` pipeline stages.
template <int NEPI, int NPROD, int WPS>
static bool tc_plan(tc::Params& P, size_t budget, int max_npa, int min_stage, int& nsplit, size_t& smem) {
  const GemmArgs& g = P.g;
  const int K = g.a.K;
  const int Np = (g.Ns + 15) / 16 * 16;
  int NpB = Np;
  size_t fixed = 0;
  for (nsplit = 1;; ++nsplit) {
    NpB = ((Np / 16 + nsplit - 1) / nsplit) * 16;
    const int NpA = (NpB + 31) / 32 * 32;
    if (NpA > max_npa) continue;
    fixed = (size_t)2 * NpB * K * 4 + (size_t)NEPI * P.epi_bufs * 32 * tc::EPI_LD * 4 + (size_t)NEPI * 2 * NpA * 4 + (size_t)6 * NpA * 4 + 512;
    if (fixed + (size_t)min_stage * 2 * tc::STAGE_FLOATS * 4 <= budget) break;
    if (NpB <= 16) return false;
  }
  nsplit = (Np + NpB - 1) / NpB;
  P.NpB = NpB;
  P.NpA = (NpB + 31) / 32 * 32;
  int cols = 32;
  while (cols < 4 * P.NpA) cols <<= 1;
  P.tmem_cols = cols;
  int nstage = (int)((budget - fixed) / ((size_t)2 * tc::STAGE_FLOATS * 4));
  // gather path: one stage per producer warp; TMA path: spare stages hold copies in flight (up to 3 per warp)
  constexpr int NGRP = NPROD / WPS;                        // producer groups (one stage in work per group)
  const int cap = P.tma ? (3 * NGRP < 16 ? 3 * NGRP : 16) : NGRP;
  if (nstage > cap) nstage = cap;
  if (nstage > NGRP) nstage = nstage / NGRP * NGRP;        // producers cycle through whole multiples of their count
  P.nstage = nstage;
  P.nprod = nstage < NGRP ? nstage : NGRP;
  smem = fixed + (size_t)nstage * 2 * tc::STAGE_FLOATS * 4;
  return true;
}

template <int NEPI, int NPROD, int WPS, int CPS>
static int tc_launch(const tc::Params& P, const CUtensorMap& tmA, const CUtensorMap& tmA2, int ctas, int nsplit, size_t smem,
                     cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(tc::pw_gemm_tc_kernel<NEPI, NPROD, WPS, CPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return C3D_ERR_SMEM;
  const long long ntiles = (P.g.M + tc::BM - 1) / tc::BM;
  long long gx = ctas / nsplit;
  if (gx < 1) gx = 1;
  if (gx > ntiles) gx = ntiles;
  dim3 grid((unsigned)gx, (unsigned)nsplit);
  static const bool dbg = getenv("C3D_TC_DBG") && atoi(getenv("C3D_TC_DBG")) != 0;
  if (dbg) {      // bring-up aid: cycle counters of one CTA's warps, printed after a synchronising copy
    static unsigned long long* dbuf = nullptr;
    if (!dbuf) cudaMalloc(&dbuf, 32 * 8 * sizeof(unsigned long long));
    cudaMemsetAsync(dbuf, 0, 32 * 8 * sizeof(unsigned long long), stream);
    tc::Params Pd = P;
    Pd.dbg = dbuf;
    tc::pw_gemm_tc_kernel<NEPI, NPROD, WPS, CPS><<<grid, (NEPI + 1 + NPROD) * 32, smem, stream>>>(Pd, tmA, tmA2);
    unsigned long long h[32 * 8];
    cudaMemcpyAsync(h, dbuf, sizeof(h), cudaMemcpyDeviceToHost, stream);
    cudaStreamSynchronize(stream);
    fprintf(stderr, "[tcdbg] M=%lld K=%d N=%d mode=%d epi=%d grid=%ux%u nepi=%d nprod=%d/%d nstage=%d tma=%d NpB=%d smem=%zu\n",
            P.g.M, P.g.a.K, P.g.Ns, P.g.a.mode, P.g.epi, grid.x, grid.y, NEPI, P.nprod * WPS, NPROD, P.nstage, P.tma, P.NpB, smem);
    for (int w = 0; w < NEPI + 1 + NPROD; ++w) {
      const unsigned long long* o = h + w * 8;
      const char* role = w < NEPI ? "epi " : w == NEPI ? "mma " : "prod";
      fprintf(stderr, "[tcdbg]  w%02d %s setup=%llu (pre %llu, weights %llu) total=%llu waitA=%llu waitB=%llu work=%llu n=%llu tiles=%llu\n", w, role,
              o[0], o[7] >> 32, o[7] & 0xffffffffull, o[1], o[2], o[3], o[4], o[5], o[6]);
    }
    return c3d_check_last(cudaGetLastError());
  }
  return c3d_check_last(c3d_launch_pdl(tc::pw_gemm_tc_kernel<NEPI, NPROD, WPS, CPS>, grid, dim3((NEPI + 1 + NPROD) * 32), smem,
                                       stream, P, tmA, tmA2));
}

// Host-side eligibility + launch.  Returns -1 when the shape / mode is not handled here (caller falls back
// to the FFMA kernel), otherwise a C3D status.  `compact`: 0 = always the 17-warp kernel, 1 = the 9-warp kernel
// (two CTAs per SM) when it needs no finer N split than the big one, 2 = whenever it fits.  `use_tma`: feed dense
// operands by TMA (C3D_TC_TMA, default on).
int c3d_launch_pw_gemm_tc(const GemmArgs& g0, int num_sms, cudaStream_t stream, int lbo_is_k, int compact, int use_tma) {
  const GemmArgs& g = g0;
  if (g.a.map != MAP_DENSE && g.a.map != MAP_SUB2) return -1;
  const bool enhance_epi = g.epi == EPI_RELU_ADD || g.epi == EPI_ABSDIFF_BWD;      // strided, in-place output
  if (g.epi != EPI_STORE && g.epi != EPI_SWISH_BWD && g.epi != EPI_ADD2 && !enhance_epi) return -1;
  if (!enhance_epi && g.out_img_stride != (long long)g.a.OHW * g.Ns) return -1;
  if (enhance_epi && g.stats) return -1;
  if ((g.a.K & 7) || g.a.K < 8 || (g.Ns & 3)) return -1;
  if (g.M >= (1LL << 31) || g.M < 1) return -1;      // a single partial tile is fine: TMA zero-fills, the epilogue masks
  if (g.M * (long long)(g.a.ld > g.Ns ? g.a.ld : g.Ns) >= (1LL << 40)) return -1;
  tc::Params P;
  P.g = g;
  P.dbg = nullptr;
  P.rps = g.rows_per_sample >= 0x7fffffffLL ? 0x7fffffffu : (uint32_t)(g.rows_per_sample > 0 ? g.rows_per_sample : 1);
  P.lbo_is_k = lbo_is_k & 1;
  P.wait_hint = lbo_is_k >> 1;      // upper bits of the bring-up flag carry the wait hint (C3D_TC_HINT)
  P.epi_bufs = 1;
  {
    static const int cat_env = getenv("C3D_TC_CAT") ? atoi(getenv("C3D_TC_CAT")) : 1;
    P.ncat = cat_env ? 1 : 0;
  }
  P.dense_contig = (g.a.map == MAP_DENSE && g.a.img_stride == (long long)g.a.OHW * g.a.ld &&
                    (g.a.A2 == nullptr || g.a.img_stride2 == (long long)g.a.OHW * g.a.ld)) ? 1 : 0;
  // TMA feed: rows must be uniformly strided (dense map); the second operand shares the row geometry
  CUtensorMap tmA, tmA2;
  memset(&tmA, 0, sizeof(tmA));
  memset(&tmA2, 0, sizeof(tmA2));
  P.tma = 0;
  if (use_tma && P.dense_contig) {
    const bool has2 = (g.a.mode == PRO_BNBWD || g.a.mode == PRO_ABSDIFF || g.a.mode == PRO_MASK_POS);
    bool ok = tc::make_tmap_rows(&tmA, g.a.A, g.a.K, g.M, g.a.ld, tc::BM);
    if (ok && has2) ok = tc::make_tmap_rows(&tmA2, g.a.A2, g.a.K, g.M, g.a.ld, tc::BM);
    P.tma = ok ? 1 : 0;
  }
  {
    static const int early_env = getenv("C3D_TC_EARLY") ? atoi(getenv("C3D_TC_EARLY")) : 0;   // measured neutral: off
    P.early = early_env ? 1 : 0;
  }
  { static const int epar_env = getenv("C3D_TC_EPAR") ? atoi(getenv("C3D_TC_EPAR")) : 0;   /* measured neutral: off */ P.epar = epar_env ? 1 : 0; }
  P.pf_dist = 0;
  if (P.tma) {
    static const int pf_env = getenv("C3D_L2PF") ? atoi(getenv("C3D_L2PF")) : 0;       // 0 off (default: measured 1-4 % slower, profiles/r02_summary.md), -1 auto, n tiles
    const bool has2 = (g.a.mode == PRO_BNBWD || g.a.mode == PRO_ABSDIFF || g.a.mode == PRO_MASK_POS);
    const long long tile_bytes = (long long)tc::BM * g.a.K * 4 * (has2 ? 2 : 1);
    int d = (int)((128 * 1024 + tile_bytes - 1) / tile_bytes);
    if (d > 8) d = 8;
    P.pf_dist = pf_env < 0 ? d : (pf_env > 16 ? 16 : pf_env);
  }
  tc::Params Pb = P, Pc = P, Ph = P;
  int ns_b = 0, ns_c = 0, ns_h = 0;
  size_t smem_b = 0, smem_c = 0, smem_h = 0;
  const bool ok_b = tc_plan<8, 7, 1>(Pb, 224 * 1024, 128, 4, ns_b, smem_b);
  // two CTAs per SM: half the shared memory (minus the 1 KB the driver reserves per CTA) and half the TMEM columns
  const bool ok_c = compact > 0 && tc_plan<4, 3, 1>(Pc, 111 * 1024, 64, 3, ns_c, smem_c) && (long long)(g.M + tc::BM - 1) / tc::BM >= 2LL * num_sms;
  // producer-heavy shapes (a transform with transcendental / two-operand arithmetic in front of a narrow output):
  // one epilogue warpgroup, five producer groups of two warps each.  C3D_TC_HEAVY: 0 off, 1 where the compact kernel
  // does not apply, 2 also instead of the compact kernel, 3 for every shape (tuning aid)
  // default 0 since round 2: with the cheaper split / folded BN-backward prologues the 8 + 7 kernel is the faster one for
  // these shapes too (res4 conv_a dgrad 113 -> 90 us, step 54.78 -> 53.83 ms, same box; CC unchanged)
  static const int heavy_on = getenv("C3D_TC_HEAVY") ? atoi(getenv("C3D_TC_HEAVY")) : 0;
  const bool heavy_shape = (g.a.mode == PRO_BN_GATE_SWISH || g.a.mode == PRO_BNBWD) && g.epi != EPI_SWISH_BWD && g.a.K >= 2 * g.N;
  const bool ok_h = heavy_on && (heavy_shape || heavy_on >= 3) && tc_plan<4, 10, 2>(Ph, 224 * 1024, 128, 4, ns_h, smem_h) && (!ok_b || ns_h <= ns_b);
  if (ok_h && heavy_on >= 2) return tc_launch<4, 10, 2, 1>(Ph, tmA, tmA2, num_sms, ns_h, smem_h, stream);
  if (ok_c && (compact >= 2 || !ok_b || ns_c <= ns_b)) return tc_launch<4, 3, 1, 2>(Pc, tmA, tmA2, 2 * num_sms, ns_c, smem_c, stream);
  if (ok_h) return tc_launch<4, 10, 2, 1>(Ph, tmA, tmA2, num_sms, ns_h, smem_h, stream);
  if (!ok_b) return -1;
  return tc_launch<8, 7, 1, 1>(Pb, tmA, tmA2, num_sms, ns_b, smem_b, stream);
}
