// Weight gradient of the pointwise convolutions on tcgen05:  dW[n][k] += sum_pixels P[pix][n] * Q[pix][k].
//
// Pixels are the MMA K dimension.  Activations arrive channel-contiguous (a float4 = 4 channels of one pixel), so
// the producers transpose 4x4 blocks across lanes (4 shuffles) and store both operands K-major with K = pixels:
//   tile[pixel/4][channel/8][channel%8][pixel%4]   (the no-swizzle canonical layout the GEMM kernel also uses).
// (Reading the channel-contiguous tile through MN-major descriptors returned zeros on B200 bring-up.)
// The operand with more channels ("big") sits on the 128 TMEM lanes (1-2 blocks), the other ("small", <= 128
// channels) on the accumulator columns.  Accumulators (main + correction terms of the 3xTF32 split) stay in
// TMEM for the CTA's whole pixel range; they are flushed once with fp32 atomics.
//
// Warps: 0-3 final epilogue (TMEM -> global atomics; lane 0 of warp 0 is the MMA issuer until then), 4-15 producers
// (4 warps per pipeline stage, fused prologues identical to the GEMM kernel's).
#include <stdio.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace tcw {

using namespace tc;

constexpr int PT = 32;                 // pixels per stage (4 MMA K-steps of 8)
constexpr int WPS = 4;                 // producer warps per stage
constexpr int MAX_STAGES = 3;
constexpr int NTHREADS = (4 + WPS * MAX_STAGES) * 32;      // 16 warps: 128 registers per thread

struct Params {
  TileSrc big, small;      // operands after role assignment
  long long M;
  float* dW;
  long long dw_sb, dw_ss;  // element strides of dW along the big / small channel index
  int Nb, Ns_;             // logical channel counts written (big, small)
  int nblocks;             // 128-lane blocks of the big operand
  int NsP;                 // accumulator columns (small channels rounded to 16)
  int big_chunks, small_chunks;   // 16-byte channel chunks actually staged (operand K / 4)
  int nstage;
  int tmem_cols;
  int desc_swap;           // bring-up switch for the MN-major LBO/SBO convention
  int big_dense, small_dense;
  unsigned long long* dbg; // C3D_TC_DBG=1: per-warp wait/work cycle counters of one CTA, else null
};

// stage layout (floats): [big hi][big lo][small hi][small lo], each [PT/4][channels_alloc/8][8][4]
__device__ __forceinline__ size_t stage_floats(const Params& P) {
  return (size_t)2 * (P.nblocks * 32 + P.NsP / 4) * PT * 4;
}

// A producer batch = one quad of 16-byte channel chunks (16 channels; lane = (pixel % 8, chunk)) x the stage's 32
// pixels: four float4 per lane and operand.  Loading and transforming are separate steps so the loads of the NEXT
// batch (possibly of the next tile) are in flight while the current one is transformed: a warp always has one
// batch of global reads outstanding, whatever the pipeline stages are doing.
struct RawBatch { float4 v[4], v2[4]; uint32_t img[4]; };

__device__ __forceinline__ void load_batch(const TileSrc& s, bool dense, long long M, long long row0, int cq, int lane, RawBatch& rb) {
  const bool has2 = (s.mode == PRO_BNBWD || s.mode == PRO_ABSDIFF || s.mode == PRO_MASK_POS);
  const int sub = lane >> 3, pl = lane & 7;
  const int k = (cq * 4 + sub) * 4;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long long row = row0 + u * 8 + pl;
    rb.v[u] = f4zero(); rb.v2[u] = f4zero(); rb.img[u] = 0;
    if (k < s.K && row < M) {
      long long off, off2;
      if (dense) { off = row * s.ld; off2 = off; rb.img[u] = (uint32_t)row / (uint32_t)s.OHW; }
      else row_offsets(s, (uint32_t)row, off, off2, rb.img[u]);
      rb.v[u] = ldg4(s.A + off + k);
      if (has2) rb.v2[u] = ldg4(s.A2 + off2 + k);
    }
  }
}

template <int MODE>
__device__ __forceinline__ void store_batch(const TileSrc& s, long long M, long long row0, int cq, int cgroups, float* hi,
                                            float* lo, int lane, const RawBatch& rb) {
  const int sub = lane >> 3, pl = lane & 7;
  const int k = (cq * 4 + sub) * 4;
  const bool kvalid = k < s.K;
  ChanParams cp = load_chan_params<MODE>(s, kvalid ? k : 0);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int pix = u * 8 + pl;
    float4 x = f4zero();
    if (kvalid && row0 + pix < M) {
      float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f);
      if (MODE == PRO_BN_GATE_SWISH && s.gate) g4 = ldg4(s.gate + (long long)(rb.img[u] / (uint32_t)s.frames_per_sample) * s.ld + k);
      x = prologue<MODE>(cp, rb.v[u], rb.v2[u], g4);
    }
    // 4x4 transpose across the 4 lanes that hold 4 consecutive pixels of this channel chunk (all lanes shuffle):
    // afterwards lane j = pixel % 4 holds channel 4*chunk + j at pixels p0..p0+3
    const int j = pl & 3;
    {
      const bool up = (j & 2) != 0;
      const float s0 = up ? x.x : x.z, s1 = up ? x.y : x.w;
      const float r0 = __shfl_xor_sync(0xffffffffu, s0, 2), r1 = __shfl_xor_sync(0xffffffffu, s1, 2);
      if (up) { x.x = r0; x.y = r1; } else { x.z = r0; x.w = r1; }
    }
    {
      const bool odd = (j & 1) != 0;
      const float s0 = odd ? x.x : x.y, s1 = odd ? x.z : x.w;
      const float r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
      if (odd) { x.x = r0; x.z = r1; } else { x.y = r0; x.w = r1; }
    }
    if (!kvalid) continue;      // channel chunks past the operand's width stay at the zeros written at kernel start
    float4 h, l;
    split4(x, h, l);
    const int ch = k + j;                                  // channel this lane now owns
    const int o = (((pix >> 2) * cgroups + (ch >> 3)) * 8 + (ch & 7)) * 4;
    *reinterpret_cast<float4*>(hi + o) = h;
    *reinterpret_cast<float4*>(lo + o) = l;
  }
}

__device__ __forceinline__ void store_batch_any(const TileSrc& s, long long M, long long row0, int cq, int cgroups, float* hi,
                                                float* lo, int lane, const RawBatch& rb) {
  switch (s.mode) {
    case PRO_NONE: store_batch<PRO_NONE>(s, M, row0, cq, cgroups, hi, lo, lane, rb); break;
    case PRO_BN_RELU: store_batch<PRO_BN_RELU>(s, M, row0, cq, cgroups, hi, lo, lane, rb); break;
    case PRO_BN_GATE_SWISH: store_batch<PRO_BN_GATE_SWISH>(s, M, row0, cq, cgroups, hi, lo, lane, rb); break;
    case PRO_BNBWD: store_batch<PRO_BNBWD>(s, M, row0, cq, cgroups, hi, lo, lane, rb); break;
    case PRO_ABSDIFF: store_batch<PRO_ABSDIFF>(s, M, row0, cq, cgroups, hi, lo, lane, rb); break;
    default: store_batch<PRO_MASK_POS>(s, M, row0, cq, cgroups, hi, lo, lane, rb); break;
  }
}

__global__ void __launch_bounds__(NTHREADS, 1) pw_wgrad_tc_kernel(const Params P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool dbg_on = P.dbg != nullptr && blockIdx.x == gridDim.x / 2;
  long long d_t0 = dbg_on ? clock64() : 0, d_a = 0, d_b = 0, d_c = 0, d_n = 0;
#define DBG_T(acc, stmt) do { if (dbg_on) { const long long t_ = clock64(); stmt; acc += clock64() - t_; } else { stmt; } } while (0)
  float* stages = reinterpret_cast<float*>(smem_raw);
  const size_t sfl = stage_floats(P);
  uint64_t* bars = reinterpret_cast<uint64_t*>(stages + (size_t)P.nstage * sfl);
  uint64_t* full = bars;                   // [nstage], WPS arrivals
  uint64_t* empty = bars + MAX_STAGES;     // [nstage]
  uint64_t* done = empty + MAX_STAGES;     // [1] all MMAs of this CTA finished
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int big_alloc = P.nblocks * 32, small_alloc = P.NsP / 4;    // chunks allocated per operand half
  if (threadIdx.x == 0) {
    for (int i = 0; i < P.nstage; ++i) { mbar_init(smem_u32(full + i), WPS); mbar_init(smem_u32(empty + i), 1); }
    mbar_init(smem_u32(done), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), (uint32_t)P.tmem_cols);
  // zero every stage once: channel chunks beyond the operands' real width are never written by the producers
  for (size_t i = threadIdx.x; i < (size_t)P.nstage * sfl / 4; i += NTHREADS)
    reinterpret_cast<float4*>(stages)[i] = f4zero();
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long d_t1 = dbg_on ? clock64() : 0;
  pdl_trigger();     // programmatic dependent launch: setup done (tensor memory allocated), operands need the
  pdl_wait();        // earlier kernels' results

  const long long ntiles = (P.M + PT - 1) / PT;
  const long long tpc = (ntiles + gridDim.x - 1) / gridDim.x;
  const long long t_begin = (long long)blockIdx.x * tpc;
  const long long t_end = t_begin + tpc < ntiles ? t_begin + tpc : ntiles;
  const long long my_tiles = t_end > t_begin ? t_end - t_begin : 0;

  if (warp >= 4) {
    // ===================== producers: stage = (warp - 4) / WPS, part = (warp - 4) % WPS =====================
    const int pw = warp - 4;
    const int stage = pw / WPS, part = pw % WPS;
    if (stage < P.nstage) {
      float* base = stages + (size_t)stage * sfl;
      float* big_hi = base;
      float* big_lo = big_hi + (size_t)big_alloc * PT * 4;
      float* small_hi = big_lo + (size_t)big_alloc * PT * 4;
      float* small_lo = small_hi + (size_t)small_alloc * PT * 4;
      const int bq = (P.big_chunks + 3) / 4, sq = (P.small_chunks + 3) / 4;    // chunk quads per operand
      // flat cursor over this warp's batches: tiles stage, stage + nstage, ...; per tile the big operand's chunk
      // quads part, part + WPS, ... then the small operand's
      struct Cur { long long ti; int opnd, cq; };
      auto normalize = [&](Cur& c) {
        while (c.ti < my_tiles && c.cq >= (c.opnd ? sq : bq)) {
          if (c.opnd == 0) { c.opnd = 1; c.cq = part; } else { c.opnd = 0; c.cq = part; c.ti += P.nstage; }
        }
      };
      auto load = [&](const Cur& c, RawBatch& rb) {
        const long long row0 = (t_begin + c.ti) * PT;
        if (c.opnd == 0) load_batch(P.big, P.big_dense != 0, P.M, row0, c.cq, lane, rb);
        else load_batch(P.small, P.small_dense != 0, P.M, row0, c.cq, lane, rb);
      };
      Cur nxt;
      nxt.ti = stage; nxt.opnd = 0; nxt.cq = part;
      normalize(nxt);
      RawBatch cur_rb, nxt_rb;
      if (nxt.ti < my_tiles) load(nxt, cur_rb);
      uint32_t use = 0;
      for (long long ti = stage; ti < my_tiles; ti += P.nstage, ++use) {
        DBG_T(d_a, mbar_wait(smem_u32(empty + stage), (use & 1) ^ 1));
        ++d_n;
        const long long t_x = dbg_on ? clock64() : 0;
        const long long row0 = (t_begin + ti) * PT;
        while (nxt.ti == ti) {
          const Cur c = nxt;
          nxt.cq += WPS;
          normalize(nxt);
          if (nxt.ti < my_tiles) load(nxt, nxt_rb);
          if (c.opnd == 0) store_batch_any(P.big, P.M, row0, c.cq, big_alloc / 2, big_hi, big_lo, lane, cur_rb);
          else store_batch_any(P.small, P.M, row0, c.cq, small_alloc / 2, small_hi, small_lo, lane, cur_rb);
          cur_rb = nxt_rb;
        }
        if (dbg_on) d_c += clock64() - t_x;
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(full + stage));
      }
    }
  } else {
    // ===================== MMA issuer: lane 0 of warp 0 (the epilogue warps have nothing to do until the end) ==========
    if (warp == 0 && lane == 0) {
      // M = 128 (big channels), N = NsP (small channels), K = 8 pixels; both operands K-major (K = pixels)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(P.NsP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t lbo_big = (uint32_t)(big_alloc / 2) * 128, lbo_small = (uint32_t)(small_alloc / 2) * 128, sbo = 128;
      uint32_t first = 0;
      int stage = 0;
      uint32_t phase = 0;
      for (long long ti = 0; ti < my_tiles; ++ti) {
        DBG_T(d_b, mbar_wait(smem_u32(full + stage), phase));
        ++d_n;
        tc_fence_after();
        const uint32_t base = smem_u32(stages + (size_t)stage * sfl);
        const uint32_t bhi = base, blo = bhi + (uint32_t)big_alloc * PT * 16;
        const uint32_t shi = blo + (uint32_t)big_alloc * PT * 16, slo = shi + (uint32_t)small_alloc * PT * 16;
        for (int ks = 0; ks < PT / 8; ++ks) {
          const uint64_t dsh = make_desc(shi + (uint32_t)ks * 2 * lbo_small, lbo_small, sbo);
          const uint64_t dsl = make_desc(slo + (uint32_t)ks * 2 * lbo_small, lbo_small, sbo);
          for (int b = 0; b < P.nblocks; ++b) {
            const uint32_t boff = (uint32_t)ks * 2 * lbo_big + (uint32_t)b * 16 * 128;     // block b = channel groups 16b..
            const uint64_t dbh = make_desc(bhi + boff, lbo_big, sbo), dbl = make_desc(blo + boff, lbo_big, sbo);
            const uint32_t d_main = tmem_base + (uint32_t)(b * 2 * P.NsP);
            const uint32_t d_corr = d_main + (uint32_t)P.NsP;
            umma_tf32(d_main, dbh, dsh, idesc, first);
            umma_tf32(d_corr, dbl, dsh, idesc, first);
            umma_tf32(d_corr, dbh, dsl, idesc, 1u);
          }
          first = 1u;
        }
        umma_commit(smem_u32(empty + stage));
        if (++stage == P.nstage) { stage = 0; phase ^= 1u; }
      }
      umma_commit(smem_u32(done));
    }
    __syncwarp();
    // ===================== final epilogue: lane = big channel, columns = small channels =====================
    if (my_tiles > 0) {
      DBG_T(d_a, mbar_wait(smem_u32(done), 0));
      tc_fence_after();
      for (int b = 0; b < P.nblocks; ++b) {
        const int bc = b * 128 + warp * 32 + lane;          // big channel index of this thread
        const uint32_t t_main = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(b * 2 * P.NsP);
        for (int c0 = 0; c0 < P.NsP; c0 += 32) {
          float r[32], r2[32];
          tmem_ld32(t_main + (uint32_t)c0, r);
          tmem_ld32(t_main + (uint32_t)(P.NsP + c0), r2);
          if (bc < P.Nb) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int sc = c0 + j;
              if (sc < P.Ns_) atomicAdd(P.dW + (long long)bc * P.dw_sb + (long long)sc * P.dw_ss, r[j] + r2[j]);
            }
          }
        }
      }
    }
  }
  if (dbg_on && lane == 0) {
    unsigned long long* o = P.dbg + warp * 8;
    const long long t_end = clock64();
    o[0] = (unsigned long long)(d_t1 - d_t0); o[1] = (unsigned long long)(t_end - d_t1);
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

}  // namespace tcw

struct WgradArgsTC {
  TileSrc p, q;
  long long M;
  float* dW;
  long long dw_sn, dw_sk;
  int N, K;
};

// Returns -1 when the shape is not handled here (caller falls back to the FFMA kernel), else a C3D status.
int c3d_launch_pw_wgrad_tc(const TileSrc& p, const TileSrc& q, long long M, float* dW, long long dw_sn, long long dw_sk,
                           int N, int K, int num_sms, cudaStream_t stream, int desc_swap) {
  if ((p.map != MAP_DENSE && p.map != MAP_SUB2) || (q.map != MAP_DENSE && q.map != MAP_SUB2)) return -1;
  if ((p.K & 3) || (q.K & 3) || M >= (1LL << 31) || M < 4 * tcw::PT) return -1;
  tcw::Params P;
  const bool p_big = p.K >= q.K;
  P.big = p_big ? p : q;
  P.small = p_big ? q : p;
  P.Nb = p_big ? N : K;
  P.Ns_ = p_big ? K : N;
  P.dw_sb = p_big ? dw_sn : dw_sk;
  P.dw_ss = p_big ? dw_sk : dw_sn;
  P.M = M;
  P.dW = dW;
  P.big_chunks = P.big.K / 4;
  P.small_chunks = P.small.K / 4;
  P.nblocks = (P.big.K + 127) / 128;
  P.NsP = (P.small.K + 15) / 16 * 16;
  if (P.NsP < 16 || P.NsP > 128 || P.nblocks > 2) return -1;
  if (P.NsP & 31) P.NsP = (P.NsP + 31) / 32 * 32;      // epilogue reads 32-column groups
  int cols = 32;
  while (cols < P.nblocks * 2 * P.NsP) cols <<= 1;
  if (cols > 512) return -1;
  P.tmem_cols = cols;
  P.desc_swap = desc_swap;
  P.dbg = nullptr;
  auto dense = [](const TileSrc& s) {
    return (s.map == MAP_DENSE && s.img_stride == (long long)s.OHW * s.ld &&
            (s.A2 == nullptr || s.img_stride2 == (long long)s.OHW * s.ld)) ? 1 : 0;
  };
  P.big_dense = dense(P.big);
  P.small_dense = dense(P.small);
  const size_t stage_bytes = (size_t)2 * (P.nblocks * 32 + P.NsP / 4) * tcw::PT * 16;
  int nstage = (int)((220 * 1024 - 256) / stage_bytes);
  if (nstage > tcw::MAX_STAGES) nstage = tcw::MAX_STAGES;
  if (nstage < 2) return -1;
  P.nstage = nstage;
  const size_t smem = (size_t)nstage * stage_bytes + 256;
  cudaError_t e = cudaFuncSetAttribute(tcw::pw_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return C3D_ERR_SMEM;
  const long long ntiles = (M + tcw::PT - 1) / tcw::PT;
  long long gx = num_sms;
  if (gx > ntiles) gx = ntiles;
  static const bool dbg = getenv("C3D_TC_DBG") && atoi(getenv("C3D_TC_DBG")) != 0;
  if (dbg) {
    static unsigned long long* dbuf = nullptr;
    if (!dbuf) cudaMalloc(&dbuf, 32 * 8 * sizeof(unsigned long long));
    cudaMemsetAsync(dbuf, 0, 32 * 8 * sizeof(unsigned long long), stream);
    P.dbg = dbuf;
    tcw::pw_wgrad_tc_kernel<<<(unsigned)gx, tcw::NTHREADS, smem, stream>>>(P);
    unsigned long long h[32 * 8];
    cudaMemcpyAsync(h, dbuf, sizeof(h), cudaMemcpyDeviceToHost, stream);
    cudaStreamSynchronize(stream);
    fprintf(stderr, "[wgdbg] M=%lld big=%d(mode %d) small=%d(mode %d) nblocks=%d NsP=%d nstage=%d smem=%zu\n", M, P.big.K, P.big.mode,
            P.small.K, P.small.mode, P.nblocks, P.NsP, P.nstage, smem);
    for (int w = 0; w < tcw::NTHREADS / 32; ++w) {
      const unsigned long long* o = h + w * 8;
      const char* role = w == 0 ? "mma+epi" : w < 4 ? "epi " : "prod";
      fprintf(stderr, "[wgdbg]  w%02d %s setup=%llu total=%llu waitA=%llu waitB=%llu work=%llu n=%llu tiles=%llu\n", w, role, o[0], o[1],
              o[2], o[3], o[4], o[5], o[6]);
    }
    return c3d_check_last(cudaGetLastError());
  }
  return c3d_check_last(c3d_launch_pdl(tcw::pw_wgrad_tc_kernel, dim3((unsigned)gx), dim3(tcw::NTHREADS), smem, stream, P));
}
