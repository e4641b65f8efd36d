// Weight gradient of the pointwise convolutions on tcgen05, MN-major operands fed by TMA:
//     dW[n][k] += sum_pixels P[pix][n] * Q[pix][k]
//
// Pixels are the MMA K dimension.  Activations are channel-contiguous in HBM, i.e. an operand tile "32 channels x PT
// pixels" IS the MN-major canonical UMMA layout (128-byte rows along M/N, 4 rows = one 512-byte swizzle atom along K)
// when TMA writes it with the 128-byte swizzle of 32-byte atoms.  So the tiles go HBM -> shared by cp.async.bulk.tensor, the producer
// warps apply the fused prologue and the TF32 hi/lo split IN PLACE (no transposes, no register staging of global
// loads), and tcgen05.mma reads both operands with a_major = b_major = MN.
//   stage  = [P hi | P lo | Q hi | Q lo], each [channel group of 32][PT pixels][128 B]; the raw tile lands in the hi
//            buffer, the prologue's second tensor (y for the BN backward, the ReLU pre-activation) in the lo buffer
//   MMA    = M 128 (a block of 4 channel groups of the wider operand) x N (the narrower operand) x K 8 pixels, three
//            terms of the 3xTF32 split into two TMEM accumulators (main + correction) that live for the whole CTA
//   warps  = 0-3 final epilogue (TMEM -> fp32 atomics; until then warp 0 issues the MMAs and warp 1 the TMA copies),
//            4-15 producers (all of them work on every stage, one (operand, channel group, 16 rows) job at a time)
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

#include "tc_common.cuh"

namespace tcm {

using namespace tc;

constexpr int NPW = 12;                    // producer warps
constexpr int NTHREADS = (4 + NPW) * 32;   // 16 warps: 128 registers per thread
constexpr int MAX_STAGES = 6;

struct Params {
  TileSrc big, small;      // operands after role assignment (big: TMEM lanes, small: accumulator columns)
  long long M;
  float* dW;
  long long dw_sb, dw_ss;  // element strides of dW along the big / small channel index
  int Nb, Ns_;             // logical channel counts written (big, small)
  int Gb, Gs;              // 32-channel groups staged per operand
  int nblocks;             // 128-lane blocks of the big operand
  int NsP;                 // accumulator columns (small channels rounded to 32)
  int PT;                  // pixels per stage (multiple of 8)
  int nstage;
  int tmem_cols;
  int swap_lbo;            // bring-up switch: exchange the LBO / SBO fields of the MN-major descriptors
  int pf_dist;             // L2 prefetch distance in stages (0 = off), counted from the stage being loaded
  int ncat;                // 1: [S hi | S lo] is read as ONE operand of 2 * NsP columns (Bh x [Sh | Sl] in one MMA)
  int stack;               // 1 (needs ncat, 2 * Gb <= 4): [B hi ; B lo] fills the 128 lanes, one MMA per k-step
  int nacc;                // accumulator sets used round-robin over the k-steps (power of two; tuning option, default 1),
                           // summed by the final epilogue
  uint32_t stage_bytes, off_blo, off_shi, off_slo, grp_bytes;
  // raw landing ring (TMA destinations), decoupled from the operand ring the MMAs read: a landing slot is re-requested as
  // soon as the producers have READ it, so the loads run ahead of the transform instead of behind the MMAs
  // (profiles/r02_summary.md section 2.2: with in-place stages ~1 stage per SM was in flight).  nraw == 0: in place.
  int nraw;
  uint32_t raw_bytes, raw_b2, raw_s, raw_s2;      // slot size; offsets of B's second tensor, S, S's second tensor
  unsigned long long* dbg;
};

// MN-major TF32 operands have exactly one legal shared-memory layout (CUTLASS sm100_common.inl: "for mn-major tf32
// operands, SW128_32B is the only available smem layout"): 128-byte rows of 32 consecutive M/N elements whose four
// 32-byte chunks are XOR-swizzled with the row index mod 4 (layout type SWIZZLE_128B_BASE32B, what TMA writes with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B); 4 K-rows form a 512-byte atom.
// LBO = bytes between consecutive 32-element groups along M/N, SBO = bytes between 4-row groups along K.
__device__ __forceinline__ uint64_t make_desc_mn128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;       // layout_type SWIZZLE_128B_BASE32B
  return d;
}

// The rows of one job with everything per-channel / per-sample already in registers (the producer loop hoists them:
// a warp's first job is the same (operand, channel group, 16-row chunk) in every stage).  FULL: all PT rows valid and
// inside one batch sample -- no per-row selects.
template <int MODE, bool FULL>
__device__ __forceinline__ void transform_rows(const ChanParams& cp, float4 g0, float4 g1, int gsplit, int nvalid, bool kvalid,
                                               const float* src, const float* src2, float* hi, float* lo, int lane, int q0) {
  constexpr bool HAS2 = (MODE == PRO_BNBWD || MODE == PRO_ABSDIFF || MODE == PRO_MASK_POS);
  const int u = lane & 7, rsub = lane >> 3;
  int o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = 4 * (q0 + i) + rsub;
    o[i] = r * 32 + (((((u >> 1) ^ (r & 3)) << 1) | (u & 1)) << 2);
  }
  if (!kvalid) {      // channels past K: TMA zero-filled the raw tile, the split halves must both read as zero
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      *reinterpret_cast<float4*>(hi + o[i]) = f4zero();
      *reinterpret_cast<float4*>(lo + o[i]) = f4zero();
    }
    return;
  }
  float4 v[4], v2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i] = *reinterpret_cast<const float4*>(src + o[i]);
    v2[i] = HAS2 ? *reinterpret_cast<const float4*>(src2 + o[i]) : f4zero();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float4 x;
    if (FULL) {
      x = prologue<MODE>(cp, v[i], v2[i], g0);
    } else {
      const int r = 4 * (q0 + i) + rsub;
      x = prologue<MODE>(cp, v[i], v2[i], (r >= gsplit) ? g1 : g0);
      if (r >= nvalid) x = f4zero();
    }
    float4 h, l;
    split4(x, h, l);
    *reinterpret_cast<float4*>(hi + o[i]) = h;
    *reinterpret_cast<float4*>(lo + o[i]) = l;
  }
}

// In-place prologue + hi/lo split of one (operand, channel group) job: PT rows of 128 bytes, 16-byte unit u of row r
// at physical unit (((u >> 1) ^ (r & 3)) << 1) | (u & 1).  A lane keeps its logical unit (4 channels) for all rows; one instruction covers four
// consecutive rows = 512 contiguous bytes, each lane its own 16 bytes (conflict-free).
template <int MODE>
__device__ __forceinline__ void transform_group(const TileSrc& s, long long M, long long row0, int PT, int g, const float* src,
                                                const float* src2, float* hi, float* lo,
                                                int lane, int q0) {      // q0: first 4-row quad of this job's 16 rows
  constexpr bool HAS2 = (MODE == PRO_BNBWD || MODE == PRO_ABSDIFF || MODE == PRO_MASK_POS);
  const int u = lane & 7, rsub = lane >> 3;
  int k = g * 32 + 4 * u;
  const bool kvalid = k < s.K;
  if (!kvalid) k = 0;                       // parameters are read in bounds; the columns are ignored downstream
  const ChanParams cp = load_chan_params<MODE>(s, k);
  float4 g0 = make_float4(1.f, 1.f, 1.f, 1.f), g1 = g0;
  int gsplit = PT;
  bool gate_fast = true;
  if (MODE == PRO_BN_GATE_SWISH && s.gate) {
    const uint32_t rps = (uint32_t)s.OHW * (uint32_t)s.frames_per_sample;
    if (rps >= (uint32_t)PT) {
      const uint32_t samp0 = (uint32_t)row0 / rps;
      gsplit = (int)((samp0 + 1) * rps - (uint32_t)row0);
      g0 = ldg4(s.gate + (long long)samp0 * s.ld + k);
      if (gsplit < PT && (long long)(samp0 + 1) * rps < M) g1 = ldg4(s.gate + (long long)(samp0 + 1) * s.ld + k);
    } else {
      gate_fast = false;
    }
  }
  const long long left = M - row0;
  const int nvalid = left < PT ? (int)left : PT;
  {
    const int i0 = q0;
    float4 v[4], v2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = 4 * (i0 + i) + rsub;
      const int o = r * 32 + (((((u >> 1) ^ (r & 3)) << 1) | (u & 1)) << 2);
      v[i] = (i0 + i < PT / 4) ? *reinterpret_cast<const float4*>(src + o) : f4zero();
      v2[i] = (HAS2 && i0 + i < PT / 4) ? *reinterpret_cast<const float4*>(src2 + o) : f4zero();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i0 + i >= PT / 4) break;
      const int r = 4 * (i0 + i) + rsub;
      const int o = r * 32 + (((((u >> 1) ^ (r & 3)) << 1) | (u & 1)) << 2);
      float4 g4 = (r >= gsplit) ? g1 : g0;
      if (MODE == PRO_BN_GATE_SWISH && s.gate && !gate_fast && r < nvalid)
        g4 = ldg4(s.gate + (long long)(((uint32_t)(row0 + r) / (uint32_t)s.OHW) / (uint32_t)s.frames_per_sample) * s.ld + k);
      float4 x = prologue<MODE>(cp, v[i], v2[i], g4);
      if (r >= nvalid || !kvalid) x = f4zero();      // TMA zero-fills out-of-range rows / channels: keep them zero
      float4 h, l;
      split4(x, h, l);
      *reinterpret_cast<float4*>(hi + o) = h;
      *reinterpret_cast<float4*>(lo + o) = l;
    }
  }
}

__device__ __forceinline__ void transform_group_any(const TileSrc& s, long long M, long long row0, int PT, int g, const float* src,
                                                    const float* src2, float* hi, float* lo, int lane, int q0) {
  switch (s.mode) {
    case PRO_NONE: transform_group<PRO_NONE>(s, M, row0, PT, g, src, src2, hi, lo, lane, q0); break;
    case PRO_BN_RELU: transform_group<PRO_BN_RELU>(s, M, row0, PT, g, src, src2, hi, lo, lane, q0); break;
    case PRO_BN_GATE_SWISH: transform_group<PRO_BN_GATE_SWISH>(s, M, row0, PT, g, src, src2, hi, lo, lane, q0); break;
    case PRO_BNBWD: transform_group<PRO_BNBWD>(s, M, row0, PT, g, src, src2, hi, lo, lane, q0); break;
    case PRO_ABSDIFF: transform_group<PRO_ABSDIFF>(s, M, row0, PT, g, src, src2, hi, lo, lane, q0); break;
    default: transform_group<PRO_MASK_POS>(s, M, row0, PT, g, src, src2, hi, lo, lane, q0); break;
  }
}

__global__ void __launch_bounds__(NTHREADS, 1) pw_wgrad_mn_kernel(const Params P, const __grid_constant__ CUtensorMap tmB,
                                                                  const __grid_constant__ CUtensorMap tmB2,
                                                                  const __grid_constant__ CUtensorMap tmS,
                                                                  const __grid_constant__ CUtensorMap tmS2) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const bool dbg_on = P.dbg != nullptr && blockIdx.x == gridDim.x / 2;
  long long d_t0 = dbg_on ? clock64() : 0, d_a = 0, d_b = 0, d_c = 0, d_n = 0;
#define DBG_T(acc, stmt) do { if (dbg_on) { const long long t_ = clock64(); stmt; acc += clock64() - t_; } else { stmt; } } while (0)
  // 1024-byte alignment of the stages (swizzle atoms): the dynamic shared window may start at any 16-byte boundary
  unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // operand ring [nstage], then (decoupled mode) the raw landing ring [nraw]
  unsigned char* rawbase = base + (size_t)P.nstage * P.stage_bytes;
  const int nslots = P.nraw > 0 ? P.nraw : P.nstage;       // landing slots (in-place mode: the stages themselves)
  uint64_t* bars = reinterpret_cast<uint64_t*>(rawbase + (size_t)P.nraw * P.raw_bytes);
  uint64_t* rawfull = bars;                  // [nslots] TMA bytes have landed
  uint64_t* full = bars + MAX_STAGES;        // [nstage] NPW producer warps have transformed the stage
  uint64_t* empty = full + MAX_STAGES;       // [nstage] the MMAs reading the stage have retired
  uint64_t* rawempty = empty + MAX_STAGES;   // [nraw]   NPW producer warps have read the landing slot
  uint64_t* done = rawempty + MAX_STAGES;    // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  if (threadIdx.x == 0) {
    for (int i = 0; i < P.nstage; ++i) { mbar_init(smem_u32(full + i), NPW); mbar_init(smem_u32(empty + i), 1); }
    for (int i = 0; i < nslots; ++i) { mbar_init(smem_u32(rawfull + i), 1); mbar_init(smem_u32(rawempty + i), NPW); }
    mbar_init(smem_u32(done), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), (uint32_t)P.tmem_cols);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const long long d_t1 = dbg_on ? clock64() : 0;
  pdl_trigger();     // programmatic dependent launch: setup done (tensor memory allocated), operands need the
  pdl_wait();        // earlier kernels' results

  const long long ntiles = (P.M + P.PT - 1) / P.PT;
  const long long tpc = (ntiles + gridDim.x - 1) / gridDim.x;
  const long long t_begin = (long long)blockIdx.x * tpc;
  const long long t_end = t_begin + tpc < ntiles ? t_begin + tpc : ntiles;
  const int my_tiles = t_end > t_begin ? (int)(t_end - t_begin) : 0;
  const uint32_t base_u32 = smem_u32(base);
  const bool big2 = (P.big.mode == PRO_BNBWD || P.big.mode == PRO_ABSDIFF || P.big.mode == PRO_MASK_POS);
  const bool small2 = (P.small.mode == PRO_BNBWD || P.small.mode == PRO_ABSDIFF || P.small.mode == PRO_MASK_POS);

  if (warp >= 4) {
    // ===================== producers: in-place prologue + split of (operand, channel group) jobs ====================
    const int pw = warp - 4;
    const int nrc = P.PT / 16;                       // 16-row chunks per channel group
    const int njobs = (P.Gb + P.Gs) * nrc;
    // A warp's first job (job index pw) is the same (operand, channel group, row chunk) in every stage: its channel
    // parameters are loaded once, and the SE gate of the current batch sample is tracked incrementally (the tiles of
    // a CTA are consecutive rows), so the per-stage work is the four row quads themselves.  Further jobs of the warp
    // (more than 12 jobs per stage: 432-channel layers) take the generic path.
    const bool has0 = pw < njobs;
    const int grp0 = has0 ? pw / nrc : 0, q00 = has0 ? (pw - grp0 * nrc) * 4 : 0;
    const bool big0 = grp0 < P.Gb;
    const TileSrc& s0 = big0 ? P.big : P.small;
    const int g_0 = big0 ? grp0 : grp0 - P.Gb;
    const uint32_t hi_off0 = big0 ? (uint32_t)g_0 * (P.grp_bytes >> 2) : (P.off_shi >> 2) + (uint32_t)g_0 * (P.grp_bytes >> 2);
    const uint32_t lo_rel0 = big0 ? (P.off_blo >> 2) : ((P.off_slo - P.off_shi) >> 2);
    // landing-slot offsets (floats) of this job's tensors
    const uint32_t src_off0 = (big0 ? 0u : (P.raw_s >> 2)) + (uint32_t)g_0 * (P.grp_bytes >> 2);
    const uint32_t src2_off0 = (big0 ? (P.raw_b2 >> 2) : (P.raw_s2 >> 2)) + (uint32_t)g_0 * (P.grp_bytes >> 2);
    const bool decoupled = P.nraw > 0;
    int k0 = g_0 * 32 + 4 * (lane & 7);
    const bool kvalid0 = k0 < s0.K;
    if (!kvalid0) k0 = 0;
    const uint32_t rps0 = (uint32_t)s0.OHW * (uint32_t)s0.frames_per_sample;
    const bool gated0 = s0.mode == PRO_BN_GATE_SWISH && s0.gate != nullptr;
    const bool hoist0 = has0 && (!gated0 || rps0 >= (uint32_t)P.PT);
    auto run = [&](auto mode_tag) {
      constexpr int MODE = decltype(mode_tag)::value;
      const ChanParams cp = hoist0 ? load_chan_params<MODE>(s0, k0) : ChanParams();
      const float4 ones = make_float4(1.f, 1.f, 1.f, 1.f);
      float4 g0 = ones, g1 = ones;
      uint32_t samp = 0, samp_end = 0xffffffffu;       // first row of the next batch sample (no gate: never reached)
      if (hoist0 && gated0 && my_tiles > 0) {
        const uint32_t r0 = (uint32_t)(t_begin * P.PT);
        samp = r0 / rps0;
        samp_end = (samp + 1) * rps0;
        g0 = ldg4(s0.gate + (long long)samp * s0.ld + k0);
      }
      int stage = 0, rslot = 0;
      uint32_t phase = 0, rphase = 0;
      for (int ti = 0; ti < my_tiles; ++ti) {
        DBG_T(d_a, mbar_wait(smem_u32(rawfull + rslot), rphase));
        if (decoupled) DBG_T(d_b, mbar_wait(smem_u32(empty + stage), phase ^ 1u));      // operand slot drained by its MMAs
        ++d_n;
        const long long t_x = dbg_on ? clock64() : 0;
        const long long row0 = (t_begin + ti) * P.PT;
        float* st = reinterpret_cast<float*>(base + (size_t)stage * P.stage_bytes);
        const float* rw = decoupled ? reinterpret_cast<const float*>(rawbase + (size_t)rslot * P.raw_bytes) : nullptr;
        if (hoist0) {
          int gsplit = P.PT;
          if (gated0) {
            while ((uint32_t)row0 >= samp_end) { ++samp; samp_end += rps0; g0 = ldg4(s0.gate + (long long)samp * s0.ld + k0); }
            if (samp_end - (uint32_t)row0 < (uint32_t)P.PT) {
              gsplit = (int)(samp_end - (uint32_t)row0);
              g1 = (long long)samp_end < P.M ? ldg4(s0.gate + (long long)(samp + 1) * s0.ld + k0) : ones;
            }
          }
          const long long left = P.M - row0;
          const int nvalid = left < P.PT ? (int)left : P.PT;
          float* hi = st + hi_off0;
          const float* src = decoupled ? rw + src_off0 : hi;
          const float* src2 = decoupled ? rw + src2_off0 : hi + lo_rel0;
          if (nvalid == P.PT && gsplit >= P.PT) transform_rows<MODE, true>(cp, g0, g1, gsplit, nvalid, kvalid0, src, src2, hi, hi + lo_rel0, lane, q00);
          else transform_rows<MODE, false>(cp, g0, g1, gsplit, nvalid, kvalid0, src, src2, hi, hi + lo_rel0, lane, q00);
        }
        for (int job = hoist0 ? pw + NPW : pw; job < njobs; job += NPW) {
          const int grp = job / nrc, q0 = (job - grp * nrc) * 4;      // (operand, channel group), 16-row chunk
          if (grp < P.Gb) {
            float* hi = st + (size_t)grp * (P.grp_bytes >> 2);
            float* lo = hi + (P.off_blo >> 2);
            const float* src = decoupled ? rw + (size_t)grp * (P.grp_bytes >> 2) : hi;
            const float* src2 = decoupled ? rw + (P.raw_b2 >> 2) + (size_t)grp * (P.grp_bytes >> 2) : lo;
            transform_group_any(P.big, P.M, row0, P.PT, grp, src, src2, hi, lo, lane, q0);
          } else {
            const int g = grp - P.Gb;
            float* hi = st + (P.off_shi >> 2) + (size_t)g * (P.grp_bytes >> 2);
            float* lo = hi + ((P.off_slo - P.off_shi) >> 2);
            const float* src = decoupled ? rw + (P.raw_s >> 2) + (size_t)g * (P.grp_bytes >> 2) : hi;
            const float* src2 = decoupled ? rw + (P.raw_s2 >> 2) + (size_t)g * (P.grp_bytes >> 2) : lo;
            transform_group_any(P.small, P.M, row0, P.PT, g, src, src2, hi, lo, lane, q0);
          }
        }
        if (dbg_on) d_c += clock64() - t_x;
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(smem_u32(full + stage));
          if (decoupled) mbar_arrive(smem_u32(rawempty + rslot));       // this warp has read its part of the landing slot
        }
        if (++stage == P.nstage) { stage = 0; phase ^= 1u; }
        if (decoupled) { if (++rslot == P.nraw) { rslot = 0; rphase ^= 1u; } }
        else { rslot = stage; rphase = phase; }
      }
    };
    switch (s0.mode) {
      case PRO_NONE: run(std::integral_constant<int, PRO_NONE>()); break;
      case PRO_BN_RELU: run(std::integral_constant<int, PRO_BN_RELU>()); break;
      case PRO_BN_GATE_SWISH: run(std::integral_constant<int, PRO_BN_GATE_SWISH>()); break;
      case PRO_BNBWD: run(std::integral_constant<int, PRO_BNBWD>()); break;
      case PRO_ABSDIFF: run(std::integral_constant<int, PRO_ABSDIFF>()); break;
      default: run(std::integral_constant<int, PRO_MASK_POS>()); break;
    }
  } else {
    // ===================== TMA issuer: epilogue warp 1 (idle until the end) =====================
    if (warp == 1) {
    const uint32_t bytes = (uint32_t)P.PT * 128u * (uint32_t)(P.Gb * (big2 ? 2 : 1) + P.Gs * (small2 ? 2 : 1));
    int stage = 0;
    uint32_t phase = 0;
    const bool decoupled = P.nraw > 0;
    const int nslots_ = decoupled ? P.nraw : P.nstage;
    const uint32_t rawbase_u32 = base_u32 + (uint32_t)P.nstage * P.stage_bytes;
    for (int ti = 0; ti < my_tiles; ++ti) {
      // in place: the stage's MMAs have retired; decoupled: the producers have read the landing slot
      DBG_T(d_a, mbar_wait(smem_u32((decoupled ? rawempty : empty) + stage), phase ^ 1u));
      if (lane == 0) {
        const int r0 = (int)((t_begin + ti) * P.PT);
        const uint32_t st = decoupled ? rawbase_u32 + (uint32_t)stage * P.raw_bytes : base_u32 + (uint32_t)stage * P.stage_bytes;
        const uint32_t o_b2 = decoupled ? P.raw_b2 : P.off_blo, o_s = decoupled ? P.raw_s : P.off_shi, o_s2 = decoupled ? P.raw_s2 : P.off_slo;
        const uint32_t bar = smem_u32(rawfull + stage);
        mbar_expect_tx(bar, bytes);
        for (int g = 0; g < P.Gb; ++g) {
          tma_load_2d(st + (uint32_t)g * P.grp_bytes, &tmB, g * 32, r0, bar);
          if (big2) tma_load_2d(st + o_b2 + (uint32_t)g * P.grp_bytes, &tmB2, g * 32, r0, bar);
        }
        for (int g = 0; g < P.Gs; ++g) {
          tma_load_2d(st + o_s + (uint32_t)g * P.grp_bytes, &tmS, g * 32, r0, bar);
          if (small2) tma_load_2d(st + o_s2 + (uint32_t)g * P.grp_bytes, &tmS2, g * 32, r0, bar);
        }
        if (P.pf_dist) {      // L2 prefetch: stages beyond the ring (the first stage also requests the ones in between)
          for (int pt = ti == 0 ? nslots_ : ti + P.pf_dist; pt <= ti + P.pf_dist && pt < my_tiles; ++pt) {
            const int pr = (int)((t_begin + pt) * P.PT);
            for (int g = 0; g < P.Gb; ++g) {
              tma_prefetch_2d(&tmB, g * 32, pr);
              if (big2) tma_prefetch_2d(&tmB2, g * 32, pr);
            }
            for (int g = 0; g < P.Gs; ++g) {
              tma_prefetch_2d(&tmS, g * 32, pr);
              if (small2) tma_prefetch_2d(&tmS2, g * 32, pr);
            }
          }
        }
      }
      __syncwarp();
      if (++stage == nslots_) { stage = 0; phase ^= 1u; }
    }
    }
    // ===================== MMA issuer: lane 0 of warp 0 (the epilogue warps have nothing to do until the end) ==========
    if (warp == 0) {
      // the whole warp runs the loop (uniform operands stay in uniform registers); lane 0 alone issues
      const uint32_t leader = lane == 0 ? 1u : 0u;
      // M = 128 (big channels), N = NsP (small channels), K = 8 pixels; both operands MN-major
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(P.NsP >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
      // The issue rate of these MMAs is bound by their shared-memory operand reads (MN-major TF32: ~140-170 cycles per
      // M128 x N x K8 instruction measured, profiles/r02_summary.md), so the three split products are folded:
      //   ncat : S hi and S lo are adjacent column groups -> Bh x [Sh | Sl] is ONE instruction of 2 * NsP columns that
      //          writes [main | correction]; Bl x Sh then accumulates into the correction columns (2 reads of B, not 3)
      //   stack: with <= 2 channel groups the 128 lanes hold [B hi ; B lo] (the layout that used to be padding), so the
      //          same instruction also yields Bl x [Sh | Sl] in lanes 32 * Gb ..: one MMA per k-step, lanes folded by
      //          the final atomics
      const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)((2 * P.NsP) >> 3) << 17) |
                              ((uint32_t)(128 >> 4) << 24);
      const uint32_t ncat = P.ncat ? 1u : 0u, second = (P.ncat && !P.stack) ? 1u : 0u;
      const uint32_t lbo = P.swap_lbo ? 512u : P.grp_bytes, sbo = P.swap_lbo ? P.grp_bytes : 512u;
      const uint64_t tmpl = make_desc_mn128(0, lbo, sbo);      // start address (16-byte units, bits 0-13) is added per use
      const uint32_t blk16 = (4u * P.grp_bytes) >> 4;           // descriptor units between 128-channel blocks
      const int ksteps = P.PT / 8;
      const uint32_t set_cols = (uint32_t)(P.nblocks * 2 * P.NsP);
      uint32_t kcount = 0;                                       // k-steps issued so far: set = kcount % nacc
      int stage = 0;
      uint32_t phase = 0;
      for (int ti = 0; ti < my_tiles; ++ti) {
        DBG_T(d_b, mbar_wait(smem_u32(full + stage), phase));
        ++d_n;
        tc_fence_after();
        const uint32_t st = base_u32 + (uint32_t)stage * P.stage_bytes;
        uint64_t dsh = tmpl + (uint64_t)((st + P.off_shi) >> 4), dsl = tmpl + (uint64_t)((st + P.off_slo) >> 4);
        uint64_t dbh = tmpl + (uint64_t)(st >> 4), dbl = tmpl + (uint64_t)((st + P.off_blo) >> 4);
        for (int ks = 0; ks < ksteps; ++ks, ++kcount) {
          const uint32_t set = kcount & (uint32_t)(P.nacc - 1);
          const uint32_t first = kcount >= (uint32_t)P.nacc ? 1u : 0u;      // a set's first MMA overwrites
          for (int b = 0; b < P.nblocks; ++b) {
            const uint32_t db = tmem_base + set * set_cols + (uint32_t)(b * 2 * P.NsP);
            const uint64_t bo = (uint64_t)((uint32_t)b * blk16);
            if (ncat) {
              umma_tf32_if(leader, db, dbh + bo, dsh, idesc2, first);
              umma_tf32_if(leader & second, db + (uint32_t)P.NsP, dbl + bo, dsh, idesc, 1u);
            } else {
              umma_tf32_if(leader, db, dbh + bo, dsh, idesc, first);
              umma_tf32_if(leader, db + (uint32_t)P.NsP, dbl + bo, dsh, idesc, first);
              umma_tf32_if(leader, db + (uint32_t)P.NsP, dbh + bo, dsl, idesc, 1u);
            }
          }
          dsh += 64; dsl += 64; dbh += 64; dbl += 64;           // next 8 pixels: two 512-byte atoms
        }
        umma_commit_if(leader, smem_u32(empty + stage));
        if (++stage == P.nstage) { stage = 0; phase ^= 1u; }
      }
      umma_commit_if(leader, smem_u32(done));
    }
    __syncwarp();
    // ===================== final epilogue: lane = big channel, columns = small channels =====================
    // A warp's 32 x 32 block is added to dW with 32-lane contiguous atomics either way round: directly when the big
    // channel index is the contiguous one (dw_sb == 1), through a transposing pass over shared memory when the small
    // one is (dw_ss == 1: lane-strided atomics cost one L2 transaction per lane -- 63 K cycles of a 143 K-cycle launch
    // for the 216 x 96 conv_a gradient).  The pipeline stages are free by now (every MMA has retired).
    if (my_tiles > 0) {
      DBG_T(d_a, mbar_wait(smem_u32(done), 0));
      tc_fence_after();
      float* S = reinterpret_cast<float*>(base) + warp * (32 * 33);
      const bool transpose = P.dw_ss == 1 && P.dw_sb != 1;
      const long long total_ks = (long long)my_tiles * (P.PT / 8);
      const int sets_used = total_ks < P.nacc ? (int)total_ks : P.nacc;      // sets that received at least one MMA
      for (int b = 0; b < P.nblocks; ++b) {
        // stacked operand: lanes 32 * Gb .. 64 * Gb hold the B-lo rows of channels 0 .. 32 * Gb (their products with
        // [Sh | Sl] are added to the same dW entries by the atomics below); lanes past 64 * Gb hold nothing
        if (P.stack && warp >= 2 * P.Gb) break;
        const int bc0 = P.stack ? (warp >= P.Gb ? warp - P.Gb : warp) * 32 : b * 128 + warp * 32;   // first big channel of this warp's lanes
        const int bc = bc0 + lane;
        const uint32_t t_main = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(b * 2 * P.NsP);
        for (int c0 = 0; c0 < P.NsP; c0 += 32) {
          float r[32];
          {
            float r2[32];
            tmem_ld32(t_main + (uint32_t)c0, r);
            tmem_ld32(t_main + (uint32_t)(P.NsP + c0), r2);
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] += r2[j];
          }
          for (int sidx = 1; sidx < sets_used; ++sidx) {        // further round-robin accumulator sets
            const uint32_t so = (uint32_t)(sidx * P.nblocks * 2 * P.NsP);
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
              float q[32];
              tmem_ld32(t_main + so + (uint32_t)(half * P.NsP + c0), q);
#pragma unroll
              for (int j = 0; j < 32; ++j) r[j] += q[j];
            }
          }
          if (!transpose) {
            if (bc < P.Nb) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int sc = c0 + j;
                if (sc < P.Ns_) atomicAdd(P.dW + (long long)bc * P.dw_sb + (long long)sc * P.dw_ss, r[j]);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) S[lane * 33 + j] = r[j];
            __syncwarp();
            const int sc = c0 + lane;
            if (sc < P.Ns_) {
              const int nrow = P.Nb - bc0 < 32 ? P.Nb - bc0 : 32;
              for (int rr = 0; rr < nrow; ++rr) atomicAdd(P.dW + (long long)(bc0 + rr) * P.dw_sb + sc, S[rr * 33 + lane]);
            }
            __syncwarp();
          }
        }
      }
    }
  }
  if (dbg_on && lane == 0) {
    unsigned long long* o = P.dbg + warp * 8;
    const long long t_end_ = clock64();
    o[0] = (unsigned long long)(d_t1 - d_t0); o[1] = (unsigned long long)(t_end_ - d_t1);
    o[2] = (unsigned long long)d_a; o[3] = (unsigned long long)d_b; o[4] = (unsigned long long)d_c; o[5] = (unsigned long long)d_n;
    o[6] = (unsigned long long)my_tiles; o[7] = 0;
  }
#undef DBG_T
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
  }
}

// Tensor map over a row-major fp32 matrix [rows][ld] (first `inner` columns addressable): boxes of 32 columns x
// box_rows rows land as 128-byte rows with the 128-byte swizzle.
static bool make_tmap_mn(CUtensorMap* tm, const float* base, long long inner, long long rows, long long ld, int box_rows) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return false;
    fn = (EncodeFn)p;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld & 3) || inner < 4 || box_rows > 256) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tcm

// Returns -1 when the shape is not handled here (caller falls back to the transposing kernel), else a C3D status.
int c3d_launch_pw_wgrad_mn(const TileSrc& p, const TileSrc& q, long long M, float* dW, long long dw_sn, long long dw_sk,
                           int N, int K, int num_sms, cudaStream_t stream, int swap_lbo) {
  auto dense = [](const TileSrc& s) {
    return s.map == MAP_DENSE && s.img_stride == (long long)s.OHW * s.ld &&
           (s.A2 == nullptr || s.img_stride2 == (long long)s.OHW * s.ld);
  };
  if (!dense(p) || !dense(q)) return -1;
  if ((p.K & 3) || (q.K & 3) || M >= (1LL << 31) || M < 256) return -1;
  tcm::Params P;
  const bool p_big = p.K >= q.K;
  P.big = p_big ? p : q;
  P.small = p_big ? q : p;
  P.Nb = p_big ? N : K;
  P.Ns_ = p_big ? K : N;
  P.dw_sb = p_big ? dw_sn : dw_sk;
  P.dw_ss = p_big ? dw_sk : dw_sn;
  P.M = M;
  P.dW = dW;
  P.Gb = (P.big.K + 31) / 32;
  P.Gs = (P.small.K + 31) / 32;
  P.nblocks = (P.Gb + 3) / 4;
  P.NsP = P.Gs * 32;
  if (P.NsP > 256 || P.nblocks > 4) return -1;
  {      // C3D_WMN_CAT: 0 = three MMAs per k-step (round-1 scheme), 1 = folded split products without lane stacking, 2 (default) = both
    static const int cat_env = getenv("C3D_WMN_CAT") ? atoi(getenv("C3D_WMN_CAT")) : 2;
    P.ncat = (cat_env >= 1 && 2 * P.NsP <= 256) ? 1 : 0;
    P.stack = (cat_env >= 2 && P.ncat && 2 * P.Gb <= 4) ? 1 : 0;
  }
  if (P.nblocks * 2 * P.NsP > 512) return -1;
  {      // C3D_WMN_NACC: round-robin accumulator sets (default 1: 1 / 2 / 4 sets measured identical at every stage,
         // profiles/r02_summary.md -- the MMA chain is not what bounds this kernel; kept as a tuning option)
    static const int nacc_env = getenv("C3D_WMN_NACC") ? atoi(getenv("C3D_WMN_NACC")) : 1;
    int nacc = 1;
    while (2 * nacc <= nacc_env && 2 * nacc * P.nblocks * 2 * P.NsP <= 512) nacc *= 2;
    P.nacc = nacc;
  }
  int cols = 32;
  while (cols < P.nacc * P.nblocks * 2 * P.NsP) cols <<= 1;
  P.tmem_cols = cols;
  P.swap_lbo = swap_lbo;
  P.dbg = nullptr;
  // pixels per stage: larger for narrow operands (per-stage hand-off costs are fixed), bounded by shared memory;
  // the big operand's last block may read (never use) up to 3 groups past its own: keep the small operand behind it
  const int groups = P.Gb + P.Gs;
  int PT = groups <= 3 ? 64 : groups <= 6 ? 32 : 16;
  {      // tuning override: C3D_WMN_PT = pixels per stage (multiple of 16)
    static const int pt_env = getenv("C3D_WMN_PT") ? atoi(getenv("C3D_WMN_PT")) : 0;
    if (pt_env >= 16 && (pt_env & 15) == 0 && pt_env <= 256) PT = pt_env;
  }
  P.PT = PT;
  P.grp_bytes = (uint32_t)PT * 128u;
  P.off_blo = (uint32_t)P.Gb * P.grp_bytes;
  P.off_shi = 2u * P.off_blo;
  P.off_slo = P.off_shi + (uint32_t)P.Gs * P.grp_bytes;
  P.stage_bytes = P.off_slo + (uint32_t)P.Gs * P.grp_bytes;
  // the last big block reads 4 groups from its first one: stay inside the dynamic shared window
  const uint32_t overrun = (uint32_t)(P.nblocks * 4 - P.Gb) * P.grp_bytes;
  const size_t tail = 1024 /* alignment slack */ + 320 /* barriers */ + overrun;
  auto two_ = [](const TileSrc& s) { return s.mode == PRO_BNBWD || s.mode == PRO_ABSDIFF || s.mode == PRO_MASK_POS; };
  P.raw_b2 = (uint32_t)P.Gb * P.grp_bytes;
  P.raw_s = (uint32_t)(P.Gb * (two_(P.big) ? 2 : 1)) * P.grp_bytes;
  P.raw_s2 = P.raw_s + (uint32_t)P.Gs * P.grp_bytes;
  P.raw_bytes = P.raw_s + (uint32_t)(P.Gs * (two_(P.small) ? 2 : 1)) * P.grp_bytes;
  // C3D_WMN_RAW: 1 = landing ring decoupled from a 2-deep operand ring where at least 2 landing slots fit; 0 (default) in
  // place.  Measured neutral (277 vs 275 us at res2, 54.97 vs 55.10 ms per step on the same box): with the loads running
  // ahead the producers wait for operand slots instead of for data -- the per-stage hand-off chain, not the load depth,
  // sets the rate (profiles/r02_summary.md section 2.2).
  static const int raw_env = getenv("C3D_WMN_RAW") ? atoi(getenv("C3D_WMN_RAW")) : 0;
  const size_t budget = 226 * 1024 - tail;
  int nstage, nraw = 0;
  if (raw_env && 2 * (size_t)P.stage_bytes + 2 * (size_t)P.raw_bytes <= budget) {
    nstage = 2;
    nraw = (int)((budget - 2 * (size_t)P.stage_bytes) / P.raw_bytes);
    if (nraw > tcm::MAX_STAGES) nraw = tcm::MAX_STAGES;
  } else {
    nstage = (int)(budget / P.stage_bytes);
    if (nstage > tcm::MAX_STAGES) nstage = tcm::MAX_STAGES;
    {      // C3D_WMN_STAGES caps the ring (tuning: a shallower ring leaves shared memory to kernels of the other stream)
      static const int cap_env = getenv("C3D_WMN_STAGES") ? atoi(getenv("C3D_WMN_STAGES")) : tcm::MAX_STAGES;
      if (cap_env >= 2 && nstage > cap_env) nstage = cap_env;
    }
    if (nstage < 2) return -1;
  }
  P.nstage = nstage;
  P.nraw = nraw;
  const size_t smem = (size_t)nstage * P.stage_bytes + (size_t)nraw * P.raw_bytes + tail;
  {
    static const int pf_env = getenv("C3D_L2PF") ? atoi(getenv("C3D_L2PF")) : 0;       // 0 off (default), -1 auto, n stages past the ring
    auto two = [](const TileSrc& s) { return s.mode == PRO_BNBWD || s.mode == PRO_ABSDIFF || s.mode == PRO_MASK_POS; };
    const long long raw = (long long)PT * 128 * (P.Gb * (two(P.big) ? 2 : 1) + P.Gs * (two(P.small) ? 2 : 1));
    int d = (int)((128 * 1024 + raw - 1) / raw);
    if (d > 12) d = 12;
    P.pf_dist = pf_env == 0 ? 0 : (nraw > 0 ? nraw : nstage) + (pf_env < 0 ? d : pf_env);
  }

  CUtensorMap tmB, tmB2, tmS, tmS2;
  memset(&tmB, 0, sizeof(tmB)); memset(&tmB2, 0, sizeof(tmB2)); memset(&tmS, 0, sizeof(tmS)); memset(&tmS2, 0, sizeof(tmS2));
  auto has2 = [](const TileSrc& s) { return s.mode == PRO_BNBWD || s.mode == PRO_ABSDIFF || s.mode == PRO_MASK_POS; };
  bool ok = tcm::make_tmap_mn(&tmB, P.big.A, P.big.K, M, P.big.ld, PT) && tcm::make_tmap_mn(&tmS, P.small.A, P.small.K, M, P.small.ld, PT);
  if (ok && has2(P.big)) ok = tcm::make_tmap_mn(&tmB2, P.big.A2, P.big.K, M, P.big.ld, PT);
  if (ok && has2(P.small)) ok = tcm::make_tmap_mn(&tmS2, P.small.A2, P.small.K, M, P.small.ld, PT);
  if (!ok) return -1;

  cudaError_t e = cudaFuncSetAttribute(tcm::pw_wgrad_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return C3D_ERR_SMEM;
  const long long ntiles = (M + PT - 1) / PT;
  long long gx = num_sms;
  if (gx > ntiles) gx = ntiles;
  static const bool dbg = getenv("C3D_TC_DBG") && atoi(getenv("C3D_TC_DBG")) != 0;
  if (dbg) {
    static unsigned long long* dbuf = nullptr;
    if (!dbuf) cudaMalloc(&dbuf, 32 * 8 * sizeof(unsigned long long));
    cudaMemsetAsync(dbuf, 0, 32 * 8 * sizeof(unsigned long long), stream);
    P.dbg = dbuf;
    tcm::pw_wgrad_mn_kernel<<<(unsigned)gx, tcm::NTHREADS, smem, stream>>>(P, tmB, tmB2, tmS, tmS2);
    unsigned long long h[32 * 8];
    cudaMemcpyAsync(h, dbuf, sizeof(h), cudaMemcpyDeviceToHost, stream);
    cudaStreamSynchronize(stream);
    fprintf(stderr, "[wmdbg] M=%lld big=%d(mode %d) small=%d(mode %d) nblocks=%d NsP=%d PT=%d nstage=%d nraw=%d smem=%zu\n", M, P.big.K,
            P.big.mode, P.small.K, P.small.mode, P.nblocks, P.NsP, PT, P.nstage, P.nraw, smem);
    for (int w = 0; w < tcm::NTHREADS / 32; ++w) {
      const unsigned long long* o = h + w * 8;
      const char* role = w == 0 ? "mma+epi" : w == 1 ? "tma+epi" : w < 4 ? "epi " : "prod";
      fprintf(stderr, "[wmdbg]  w%02d %s setup=%llu total=%llu waitA=%llu waitB=%llu work=%llu n=%llu tiles=%llu\n", w, role, o[0], o[1],
              o[2], o[3], o[4], o[5], o[6]);
    }
    return c3d_check_last(cudaGetLastError());
  }
  return c3d_check_last(c3d_launch_pdl(tcm::pw_wgrad_mn_kernel, dim3((unsigned)gx), dim3(tcm::NTHREADS), smem, stream, P, tmB,
                                       tmB2, tmS, tmS2));
}
