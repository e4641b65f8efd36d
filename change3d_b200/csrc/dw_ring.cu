// Depthwise 3x3x3 Conv3d, stride 1 (X3D conv_b of every non-first block, model/x3d.py:184-193): row-streaming
// kernels that read every input element from HBM once (plus a one-column halo per 32-column strip).
//
// Work unit = (sample, block of 32 channels, strip of 32 image columns); a CTA marches down the rows of a unit.
// The T frames of strip row r (34 pixels x 32 channels each, halo columns included) land in a ring of shared
// memory row slots by cp.async (16 bytes = 4 channels per request), D rows ahead of the row being computed.
// Forward: the thread that requested a piece applies relu(bn_a(.)) to it in place — once per element instead of
// once per tap — then one __syncthreads per row and every thread computes its outputs from three ring rows.
//   thread  = (channel, group of 4 adjacent columns), all T frames and the channel's 27 taps in registers;
//             a warp reads 32 consecutive channels of one pixel: 128-byte rows, conflict-free, and because the
//             slot geometry is a compile-time constant every shared load is base + immediate
//   work    = (unit, row) steps, linearised and cut into equal contiguous spans, one span per CTA
// Zero padding lives in the ring: pieces outside the image are zero-filled by cp.async and skipped by the
// transform (padding is zero in the ACTIVATED domain).
#include <stdlib.h>

#include "c3d_common.cuh"

namespace dwr {

constexpr int NT = 256;
constexpr int CB = 32;                // channels per block
constexpr int SEGW = 32;              // image columns per strip
constexpr int PIX = SEGW + 2;         // ring pixels per row (with halo columns)
constexpr int TS = PIX * CB;          // floats per frame of a ring row

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// all but the `d` most recent groups have landed (d = rows requested ahead of the computed row, 1..4)
__device__ __forceinline__ void cp_async_wait_depth(int d) {
  if (d >= 4) cp_async_wait<4>(); else if (d == 3) cp_async_wait<3>(); else if (d == 2) cp_async_wait<2>(); else cp_async_wait<1>();
}

// Packed fp32 pairs (fma.rn.f32x2 -> FFMA2: two FMAs per issued instruction).  The depthwise kernels are bound by issue
// slots, not by the FMA pipe (profiles/r02_summary.md: 62 % issue-active, FFMA 49 % of the instructions), so a thread
// that owns two ADJACENT CHANNELS gets its inputs as natural 8-byte pairs from the channel-contiguous ring (LDS.64),
// its weights as pairs of two channels' taps, and halves FFMA, LDS and STG counts alike -- no packing moves.
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t f2_pack(float lo, float hi) {
  f2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(f2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2_t f2_fma(f2_t a, f2_t b, f2_t c) {
  f2_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

struct Geo {
  int N, IH, IW, C, Cs;
  int nCB, nSEG;       // channel blocks, column strips
  int R, D;            // ring slots, rows requested ahead of the computed row (R >= D + 4)
  int R2;              // slots of the second ring of the fused backward (D + 2)
  long long total_steps, steps_per_cta;
};

// Per-thread description of the 16-byte pieces it moves for every ring row (the same for all rows of a unit)
template <int T>
struct Pieces {
  static constexpr int E = T * PIX * (CB / 4);
  static constexpr int NE = (E + NT - 1) / NT;
  int s_off[NE];       // float offset inside a slot
  int g_off[NE];       // element offset from (sample, frame 0, row 0, strip column -1, block channel 0)
  uint32_t ok;         // bit i: piece i exists and lies inside the image / channel range of the current unit
  uint32_t exists;
  __device__ __forceinline__ void init(int tid, int IH, int IW, int Cs) {
    exists = 0;
#pragma unroll
    for (int i = 0; i < NE; ++i) {
      const int e = tid + i * NT;
      s_off[i] = 0; g_off[i] = 0;
      if (e < E) {
        const int c4 = e & 7, r = e >> 3;
        const int p = r % PIX, ti = r / PIX;
        s_off[i] = (ti * PIX + p) * CB + c4 * 4;
        g_off[i] = (ti * IH * IW + p) * Cs + c4 * 4;
        exists |= 1u << i;
      }
    }
    ok = 0;
  }
  __device__ __forceinline__ void set_unit(int tid, int seg_start, int c0, int IW, int Cs) {
    ok = 0;
#pragma unroll
    for (int i = 0; i < NE; ++i) {
      if (exists >> i & 1) {
        const int e = tid + i * NT;
        const int c4 = e & 7, p = (e >> 3) % PIX;
        const int iw = seg_start - 1 + p;
        if (iw >= 0 && iw < IW && c0 + c4 * 4 < Cs) ok |= 1u << i;
      }
    }
  }
};

struct Unit { int n, cb, sg, r0, r1; };     // rows [r0, r1) of (sample, channel block, strip)

__device__ __forceinline__ Unit decode(long long step, long long s1, const Geo& G) {
  Unit u;
  const int unit = (int)(step / G.IH);
  u.r0 = (int)(step - (long long)unit * G.IH);
  const long long left = s1 - step;
  u.r1 = (long long)(G.IH - u.r0) < left ? G.IH : u.r0 + (int)left;
  u.sg = unit % G.nSEG;
  const int t = unit / G.nSEG;
  u.cb = t % G.nCB;
  u.n = t / G.nCB;
  return u;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
// PAIR: thread = (pair of adjacent channels, group of 2 columns) with packed FFMA2 arithmetic; else (channel, 4 columns)
template <int T, bool PAIR, bool ONE>
__global__ void __launch_bounds__(NT, ONE ? 1 : 2) dw_fwd_ring_kernel(const float* __restrict__ X, const float* __restrict__ bnp,
                                                            const float* __restrict__ w, float* __restrict__ Y,
                                                            double* __restrict__ stats, const Geo G) {
  pdl_trigger();   // programmatic dependent launch: let the next kernel start its setup,
  pdl_wait();      // then wait for the earlier kernels whose results this one reads

  constexpr int SF = T * TS;                          // floats per ring slot
  extern __shared__ __align__(16) float sm[];
  float* ring = sm;                                   // [R][T][PIX][CB]
  float* s_bn = ring + (size_t)G.R * SF;              // [3][CB] mean, scale, beta of the current channel block
  double* s_st = reinterpret_cast<double*>(s_bn + 4 * CB);     // [2][CB]
  const int tid = threadIdx.x;
  const int cl = tid & 31, grp = tid >> 5;            // channel inside the block, column group (4 columns)
  const uint32_t ring_u32 = smem_u32(ring);
  const long long img = (long long)G.IH * G.IW * G.Cs;
  const long long rowstride = (long long)G.IW * G.Cs;

  Pieces<T> pc;
  pc.init(tid, G.IH, G.IW, G.Cs);

  const long long s0 = (long long)blockIdx.x * G.steps_per_cta;
  long long s1 = s0 + G.steps_per_cta;
  if (s1 > G.total_steps) s1 = G.total_steps;

  long long step = s0;
  while (step < s1) {
    const Unit u = decode(step, s1, G);
    const int c0 = u.cb * CB, seg_start = u.sg * SEGW;
    const int c = c0 + cl;
    const bool c_ok = c < G.Cs;

    __syncthreads();                                  // the previous unit's compute and flush are finished
    if (tid < 3 * CB) {
      const int k = tid >> 5, cc = tid & 31;
      const int src = k == 0 ? 0 : k == 1 ? 2 : 3;    // mean, scale, beta rows of the parameter block
      s_bn[tid] = (c0 + cc < G.Cs) ? __ldg(bnp + src * G.Cs + c0 + cc) : 0.f;
    }
    if (tid < 2 * CB) s_st[tid] = 0.0;
    pc.set_unit(tid, seg_start, c0, G.IW, G.Cs);
    // PAIR geometry: channel pair cp2 (channels c0 + 2 cp2, + 1), column group g2 (2 columns)
    const int cp2 = tid & 15, g2 = tid >> 4;
    const int cA = c0 + 2 * cp2;
    float wr[PAIR ? 1 : 27];
    f2_t wr2[PAIR ? 27 : 1];
    if (PAIR) {
#pragma unroll
      for (int t = 0; t < 27; ++t)
        wr2[t] = f2_pack((cA < G.C) ? __ldg(w + cA * 27 + t) : 0.f, (cA + 1 < G.C) ? __ldg(w + (cA + 1) * 27 + t) : 0.f);
    } else {
#pragma unroll
      for (int t = 0; t < 27; ++t) wr[t] = (c < G.C) ? __ldg(w + c * 27 + t) : 0.f;
    }
    // element (sample, frame 0, row 0, strip column -1, block channel 0): only dereferenced where pc.ok says so
    const float* Xu = X + (long long)u.n * T * img + (long long)(seg_start - 1) * G.Cs + c0;
    float* Yt = PAIR ? Y + (long long)u.n * T * img + ((long long)u.r0 * G.IW + seg_start + 2 * g2) * G.Cs + cA
                     : Y + (long long)u.n * T * img + ((long long)u.r0 * G.IW + seg_start + 4 * grp) * G.Cs + c;   // frame 0, row r0
    int nvalid = G.IW - (seg_start + (PAIR ? 2 * g2 : 4 * grp));        // valid columns of this thread's group
    nvalid = nvalid < 0 ? 0 : nvalid > (PAIR ? 2 : 4) ? (PAIR ? 2 : 4) : nvalid;
    if (PAIR ? (cA >= G.Cs) : !c_ok) nvalid = 0;
    double st_s = 0.0, st_q = 0.0, st_s1 = 0.0, st_q1 = 0.0;      // (PAIR: second channel of the pair)

    // ring bookkeeping: slot of row r is (r - r0 + 1) mod R, tracked incrementally
    int issue_row = u.r0 - 1, issue_slot = 0;
    auto issue = [&]() {
      const bool row_ok = issue_row >= 0 && issue_row < G.IH;
      const uint32_t dst0 = ring_u32 + (uint32_t)(issue_slot * SF) * 4;
      const float* src0 = Xu + (long long)issue_row * rowstride;
#pragma unroll
      for (int i = 0; i < Pieces<T>::NE; ++i) {
        if (pc.exists >> i & 1) {
          const bool ok = row_ok && (pc.ok >> i & 1);
          cp_async16(dst0 + (uint32_t)pc.s_off[i] * 4, ok ? (const void*)(src0 + pc.g_off[i]) : (const void*)X, ok ? 16u : 0u);
        }
      }
      cp_async_commit();
      ++issue_row;
      if (++issue_slot == G.R) issue_slot = 0;
    };
    auto transform = [&](int row, int slot) {         // relu(bn_a(.)) in place on this thread's own pieces
      if (row < 0 || row >= G.IH) return;
      float* base = ring + (size_t)slot * SF;
#pragma unroll
      for (int i = 0; i < Pieces<T>::NE; ++i) {
        if (pc.ok >> i & 1) {
          const int c4 = ((tid + i * NT) & 7) * 4;
          float4* p = reinterpret_cast<float4*>(base + pc.s_off[i]);
          const float4 mean = *reinterpret_cast<const float4*>(s_bn + c4);
          const float4 scale = *reinterpret_cast<const float4*>(s_bn + CB + c4);
          const float4 beta = *reinterpret_cast<const float4*>(s_bn + 2 * CB + c4);
          *p = f4relu(f4bn(*p, mean, scale, beta));
        }
      }
    };

    // prologue: rows r0-1 .. r0+D requested (slots 0 .. D+1), the first two transformed
    for (int i = 0; i < G.D + 2; ++i) issue();
    __syncthreads();                                  // s_bn is visible
    cp_async_wait_depth(G.D);
    transform(u.r0 - 1, 0);
    transform(u.r0, 1);
    int sl = 0;                                       // slot of row oh - 1
    for (int oh = u.r0; oh < u.r1; ++oh) {
      issue();
      cp_async_wait_depth(G.D);
      const int sl1 = sl + 1 >= G.R ? sl + 1 - G.R : sl + 1;
      const int sl2 = sl1 + 1 >= G.R ? sl1 + 1 - G.R : sl1 + 1;
      transform(oh + 1, sl2);
      __syncthreads();
      if (PAIR && nvalid > 0) {
        const int toff = (2 * g2) * CB + 2 * cp2;
        const float* b0 = ring + sl * SF + toff;
        const float* b1 = ring + sl1 * SF + toff;
        const float* b2 = ring + sl2 * SF + toff;
        f2_t acc[T][2];
#pragma unroll
        for (int t = 0; t < T; ++t) { acc[t][0] = 0ull; acc[t][1] = 0ull; }
#pragma unroll
        for (int ti = 0; ti < T; ++ti) {
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const float* rp = (kh == 0 ? b0 : kh == 1 ? b1 : b2) + ti * TS;
            f2_t in[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) in[j] = *reinterpret_cast<const f2_t*>(rp + j * CB);
#pragma unroll
            for (int kt = 0; kt < 3; ++kt) {
              const int to = ti - kt + 1;
              if (to < 0 || to >= T) continue;
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) {
                const f2_t wv = wr2[kt * 9 + kh * 3 + kw];
                acc[to][0] = f2_fma(wv, in[kw], acc[to][0]);
                acc[to][1] = f2_fma(wv, in[kw + 1], acc[to][1]);
              }
            }
          }
        }
        const f2_t ones = f2_pack(1.f, 1.f);
        f2_t sf2 = 0ull, qf2 = 0ull;
#pragma unroll
        for (int t = 0; t < T; ++t) {
          float* yp = Yt + (long long)t * img;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (q < nvalid) {
              *reinterpret_cast<f2_t*>(yp + q * G.Cs) = acc[t][q];
              sf2 = f2_fma(acc[t][q], ones, sf2);
              qf2 = f2_fma(acc[t][q], acc[t][q], qf2);
            }
          }
        }
        float s_lo, s_hi, q_lo, q_hi;
        f2_unpack(sf2, s_lo, s_hi);
        f2_unpack(qf2, q_lo, q_hi);
        st_s += (double)s_lo; st_s1 += (double)s_hi;
        st_q += (double)q_lo; st_q1 += (double)q_hi;
      }
      if (!PAIR && nvalid > 0) {
        const int toff = (4 * grp) * CB + cl;
        const float* b0 = ring + sl * SF + toff;
        const float* b1 = ring + sl1 * SF + toff;
        const float* b2 = ring + sl2 * SF + toff;
        float acc[T][4];
#pragma unroll
        for (int t = 0; t < T; ++t)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[t][q] = 0.f;
#pragma unroll
        for (int ti = 0; ti < T; ++ti) {
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const float* rp = (kh == 0 ? b0 : kh == 1 ? b1 : b2) + ti * TS;
            float in[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) in[j] = rp[j * CB];
#pragma unroll
            for (int kt = 0; kt < 3; ++kt) {
              const int to = ti - kt + 1;
              if (to < 0 || to >= T) continue;
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) {
                const float wv = wr[kt * 9 + kh * 3 + kw];
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[to][q] = fmaf(wv, in[q + kw], acc[to][q]);
              }
            }
          }
        }
        float sf = 0.f, qf = 0.f;
#pragma unroll
        for (int t = 0; t < T; ++t) {
          float* yp = Yt + (long long)t * img;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (q < nvalid) {
              yp[q * G.Cs] = acc[t][q];
              sf += acc[t][q];
              qf = fmaf(acc[t][q], acc[t][q], qf);
            }
          }
        }
        st_s += (double)sf;
        st_q += (double)qf;
      }
      Yt += rowstride;
      sl = sl1;
    }
    cp_async_wait<0>();
    if (stats) {      // per-sample statistics of this unit: fold the 8 column groups in shared memory, then one atomic each
      if (nvalid > 0) {
        if (PAIR) {
          atomicAdd(&s_st[2 * cp2], st_s); atomicAdd(&s_st[2 * cp2 + 1], st_s1);
          atomicAdd(&s_st[CB + 2 * cp2], st_q); atomicAdd(&s_st[CB + 2 * cp2 + 1], st_q1);
        } else {
          atomicAdd(&s_st[cl], st_s); atomicAdd(&s_st[CB + cl], st_q);
        }
      }
      __syncthreads();
      if (tid < 2 * CB) {
        const int k = tid >> 5, cc = tid & 31;
        if (c0 + cc < G.Cs) atomicAdd(stats + ((long long)u.n * 2 + k) * G.Cs + c0 + cc, s_st[tid]);
      }
    }
    step += u.r1 - u.r0;
  }
}

// ------------------------------------------------------------------------------------------------
// backward: dr = conv_transpose(dy, w) * (a > 0), dW[tap] += sum a * dy, stats_a += (sum dr, sum dr * yhat_a)
//   dy  (N,T,H,W,Cs)  gradient w.r.t. the raw conv_b output (BN_b / SE / Swish backward already applied)
//   y_a (N,T,H,W,Cs)  raw conv_a output; a = relu(bn_a(y_a)) is recomputed per thread for its own pixels
// The ring holds dy rows ih-1 .. ih+1 (input row ih receives taps from output rows ih+1-kh); y_a is only needed at
// the thread's own pixels and is read straight from global memory while the row's ring data is awaited.  The 27
// weight-gradient accumulators of the thread's channel stay in registers for a whole unit.
// ------------------------------------------------------------------------------------------------
// FUSED: DY is the raw gradient du of the Swish input and the BN_b / SE backward transform
//     dy = scale_b * (du * gate + dpool - c1 - zhat_b * c2),  zhat_b = (y_b - mean_b) * rstd_b
// is applied in place to each ring piece by the thread that requested it (y_b pieces land in a small second ring),
// which removes the elementwise pre-pass over du / y_b.
struct FuseArgs {
  const float* YB;       // raw conv_b output
  const float* bnp_b;    // [4][Cs]
  const float* gate;     // [N][Cs] or null
  const float* dpool;    // [N][Cs] or null
  const float* coef_b;   // [2][Cs]
};

// ONE: one CTA per SM with the whole register file (no 128-register cap).  T >= 4 always runs that way (its ring does not fit
// twice); measured on the T = 5 / T = 4 configurations: 11.4 -> 7.5 ms and 15.2 -> 11.6 ms per step against the capped build.
template <int T, bool FUSED, bool ONE>
__global__ void __launch_bounds__(NT, ONE ? 1 : 2) dw_bwd_ring_kernel(const float* __restrict__ DY, const float* __restrict__ YA,
                                                            const float* __restrict__ bnp_a, const float* __restrict__ w,
                                                            float* __restrict__ DR, float* __restrict__ dW,
                                                            double* __restrict__ stats_a, const Geo G, const FuseArgs F) {
  pdl_trigger();   // programmatic dependent launch: let the next kernel start its setup,
  pdl_wait();      // then wait for the earlier kernels whose results this one reads

  constexpr int SF = T * TS;
  extern __shared__ __align__(16) float sm[];
  float* ring = sm;                                                  // [R][T][PIX][CB]
  float* ring2 = ring + (size_t)G.R * SF;                            // [R2][T][PIX][CB] y_b pieces (FUSED)
  float* s_dw = ring2 + (size_t)(FUSED ? G.R2 : 0) * SF;             // [27][CB]
  double* s_st = reinterpret_cast<double*>(s_dw + 28 * CB);          // [2][CB]
  float* s_cb = reinterpret_cast<float*>(s_st + 2 * CB);             // [7][CB] scale, gate, dpool, c1, mean, rstd, c2 (FUSED)
  const int tid = threadIdx.x;
  const int cl = tid & 31, grp = tid >> 5;
  const uint32_t ring_u32 = smem_u32(ring);
  const long long img = (long long)G.IH * G.IW * G.Cs;
  const long long rowstride = (long long)G.IW * G.Cs;

  Pieces<T> pc;
  pc.init(tid, G.IH, G.IW, G.Cs);

  const long long s0 = (long long)blockIdx.x * G.steps_per_cta;
  long long s1 = s0 + G.steps_per_cta;
  if (s1 > G.total_steps) s1 = G.total_steps;

  long long step = s0;
  while (step < s1) {
    const Unit u = decode(step, s1, G);
    const int c0 = u.cb * CB, seg_start = u.sg * SEGW;
    const int c = c0 + cl;
    const bool c_ok = c < G.Cs;

    __syncthreads();
    for (int i = tid; i < 27 * CB; i += NT) s_dw[i] = 0.f;
    if (tid < 2 * CB) s_st[tid] = 0.0;
    pc.set_unit(tid, seg_start, c0, G.IW, G.Cs);
    float wr[27], dwacc[27];
#pragma unroll
    for (int t = 0; t < 27; ++t) { wr[t] = (c < G.C) ? __ldg(w + c * 27 + t) : 0.f; dwacc[t] = 0.f; }
    float mean_a = 0.f, rstd_a = 0.f, scale_a = 0.f, beta_a = 0.f;
    if (c_ok) {
      mean_a = __ldg(bnp_a + c); rstd_a = __ldg(bnp_a + G.Cs + c);
      scale_a = __ldg(bnp_a + 2 * G.Cs + c); beta_a = __ldg(bnp_a + 3 * G.Cs + c);
    }
    if (FUSED && tid < CB) {
      const int cc = c0 + tid;
      const bool ok = cc < G.Cs;
      s_cb[0 * CB + tid] = ok ? __ldg(F.bnp_b + 2 * G.Cs + cc) : 0.f;                                   // scale_b
      s_cb[1 * CB + tid] = (ok && F.gate) ? __ldg(F.gate + (long long)u.n * G.Cs + cc) : 1.f;
      s_cb[2 * CB + tid] = (ok && F.gate) ? __ldg(F.dpool + (long long)u.n * G.Cs + cc) : 0.f;
      s_cb[3 * CB + tid] = ok ? __ldg(F.coef_b + cc) : 0.f;                                             // c1
      s_cb[4 * CB + tid] = ok ? __ldg(F.bnp_b + cc) : 0.f;                                              // mean_b
      s_cb[5 * CB + tid] = ok ? __ldg(F.bnp_b + G.Cs + cc) : 0.f;                                       // rstd_b
      s_cb[6 * CB + tid] = ok ? __ldg(F.coef_b + G.Cs + cc) : 0.f;                                      // c2
    }
    const float* DYu = DY + (long long)u.n * T * img + (long long)(seg_start - 1) * G.Cs + c0;
    const float* YBu = FUSED ? F.YB + (long long)u.n * T * img + (long long)(seg_start - 1) * G.Cs + c0 : nullptr;
    const long long toff_g = (long long)u.n * T * img + ((long long)u.r0 * G.IW + seg_start + 4 * grp) * G.Cs + c;
    const float* YAt = YA + toff_g;                   // frame 0, row r0, this thread's first column
    float* DRt = DR + toff_g;
    int nvalid = G.IW - (seg_start + 4 * grp);
    nvalid = nvalid < 0 ? 0 : nvalid > 4 ? 4 : nvalid;
    if (!c_ok) nvalid = 0;
    double st_s = 0.0, st_t = 0.0;

    int issue_row = u.r0 - 1, issue_slot = 0, issue_slot2 = 0;
    auto issue = [&]() {
      const bool row_ok = issue_row >= 0 && issue_row < G.IH;
      const uint32_t dst0 = ring_u32 + (uint32_t)(issue_slot * SF) * 4;
      const uint32_t dst2 = ring_u32 + (uint32_t)((G.R + issue_slot2) * SF) * 4;
      const float* src0 = DYu + (long long)issue_row * rowstride;
      const float* src2 = FUSED ? YBu + (long long)issue_row * rowstride : nullptr;
#pragma unroll
      for (int i = 0; i < Pieces<T>::NE; ++i) {
        if (pc.exists >> i & 1) {
          const bool ok = row_ok && (pc.ok >> i & 1);
          cp_async16(dst0 + (uint32_t)pc.s_off[i] * 4, ok ? (const void*)(src0 + pc.g_off[i]) : (const void*)DY, ok ? 16u : 0u);
          if (FUSED && ok) cp_async16(dst2 + (uint32_t)pc.s_off[i] * 4, (const void*)(src2 + pc.g_off[i]), 16u);
        }
      }
      cp_async_commit();
      ++issue_row;
      if (++issue_slot == G.R) issue_slot = 0;
      if (FUSED && ++issue_slot2 == G.R2) issue_slot2 = 0;
    };
    // FUSED: du -> dy in place on this thread's own pieces (same formula and operation order as dw_dy_kernel)
    auto transform = [&](int row, int slot, int slot2) {
      if (!FUSED || row < 0 || row >= G.IH) return;
      float* base = ring + (size_t)slot * SF;
      const float* base2 = ring2 + (size_t)slot2 * SF;
      const int c4 = (tid & 7) * 4;                    // NT is a multiple of 8: every piece of a thread has this quad
      const float4 scale = *reinterpret_cast<const float4*>(s_cb + 0 * CB + c4), g = *reinterpret_cast<const float4*>(s_cb + 1 * CB + c4);
      const float4 dp = *reinterpret_cast<const float4*>(s_cb + 2 * CB + c4), c1 = *reinterpret_cast<const float4*>(s_cb + 3 * CB + c4);
      const float4 mean = *reinterpret_cast<const float4*>(s_cb + 4 * CB + c4), rstd = *reinterpret_cast<const float4*>(s_cb + 5 * CB + c4);
      const float4 c2 = *reinterpret_cast<const float4*>(s_cb + 6 * CB + c4);
#pragma unroll
      for (int i = 0; i < Pieces<T>::NE; ++i) {
        if (pc.ok >> i & 1) {
          float4* p = reinterpret_cast<float4*>(base + pc.s_off[i]);
          const float4 d = *p, y = *reinterpret_cast<const float4*>(base2 + pc.s_off[i]);
          float4 o;
          o.x = scale.x * (fmaf(d.x, g.x, dp.x) - c1.x - (y.x - mean.x) * rstd.x * c2.x);
          o.y = scale.y * (fmaf(d.y, g.y, dp.y) - c1.y - (y.y - mean.y) * rstd.y * c2.y);
          o.z = scale.z * (fmaf(d.z, g.z, dp.z) - c1.z - (y.z - mean.z) * rstd.z * c2.z);
          o.w = scale.w * (fmaf(d.w, g.w, dp.w) - c1.w - (y.w - mean.w) * rstd.w * c2.w);
          *p = o;
        }
      }
    };

    for (int i = 0; i < G.D + 2; ++i) issue();
    int sl = 0, sq = 0;                               // slot of row ih - 1; second-ring slot of row ih + 1
    if (FUSED) {
      __syncthreads();                                // s_cb is visible
      cp_async_wait_depth(G.D);
      transform(u.r0 - 1, 0, 0);
      transform(u.r0, 1, 1);
      sq = 2;
    }
    for (int ih = u.r0; ih < u.r1; ++ih) {
      issue();
      // this thread's own y_a values: in flight while the ring row is awaited
      float ya[T][4];
#pragma unroll
      for (int t = 0; t < T; ++t)
#pragma unroll
        for (int q = 0; q < 4; ++q) ya[t][q] = (q < nvalid) ? __ldg(YAt + (long long)t * img + q * G.Cs) : 0.f;
      cp_async_wait_depth(G.D);
      const int sl1 = sl + 1 >= G.R ? sl + 1 - G.R : sl + 1;
      const int sl2 = sl1 + 1 >= G.R ? sl1 + 1 - G.R : sl1 + 1;
      transform(ih + 1, sl2, sq);
      if (FUSED && ++sq == G.R2) sq = 0;
      __syncthreads();
      if (nvalid > 0) {
        const int toff = (4 * grp) * CB + cl;
        const float* b0 = ring + sl2 * SF + toff;      // kh = 0: output row ih + 1
        const float* b1 = ring + sl1 * SF + toff;      // kh = 1: output row ih
        const float* b2 = ring + sl * SF + toff;       // kh = 2: output row ih - 1
        float a[T][4], da[T][4];
#pragma unroll
        for (int t = 0; t < T; ++t)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            a[t][q] = (q < nvalid) ? fmaxf(fmaf(ya[t][q] - mean_a, scale_a, beta_a), 0.f) : 0.f;
            da[t][q] = 0.f;
          }
#pragma unroll
        for (int to = 0; to < T; ++to) {
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const float* rp = (kh == 0 ? b0 : kh == 1 ? b1 : b2) + to * TS;
            float dyw[6];                              // output columns (first column of the group) - 1 .. + 4
#pragma unroll
            for (int j = 0; j < 6; ++j) dyw[j] = rp[j * CB];
#pragma unroll
            for (int kt = 0; kt < 3; ++kt) {
              const int ti = to + kt - 1;
              if (ti < 0 || ti >= T) continue;
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) {
                const int tap = kt * 9 + kh * 3 + kw;
                const float wv = wr[tap];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float d = dyw[q + 2 - kw];     // ow = iw + 1 - kw
                  da[ti][q] = fmaf(wv, d, da[ti][q]);
                  dwacc[tap] = fmaf(a[ti][q], d, dwacc[tap]);
                }
              }
            }
          }
        }
        float sf = 0.f, tf = 0.f;
#pragma unroll
        for (int t = 0; t < T; ++t) {
          float* dp = DRt + (long long)t * img;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (q < nvalid) {
              const float v = a[t][q] > 0.f ? da[t][q] : 0.f;
              dp[q * G.Cs] = v;
              sf += v;
              tf = fmaf(v, (ya[t][q] - mean_a) * rstd_a, tf);
            }
          }
        }
        st_s += (double)sf;
        st_t += (double)tf;
      }
      YAt += rowstride;
      DRt += rowstride;
      sl = sl1;
    }
    cp_async_wait<0>();
    // ---- unit flush: weight gradient and BN_a backward statistics of this channel block ----
    if (nvalid > 0) {
#pragma unroll
      for (int t = 0; t < 27; ++t) atomicAdd(&s_dw[t * CB + cl], dwacc[t]);
      atomicAdd(&s_st[cl], st_s);
      atomicAdd(&s_st[CB + cl], st_t);
    }
    __syncthreads();
    for (int i = tid; i < 27 * CB; i += NT) {
      const int t = i >> 5, cc = i & 31;
      if (c0 + cc < G.C) atomicAdd(dW + (c0 + cc) * 27 + t, s_dw[i]);
    }
    if (tid < 2 * CB) {
      const int k = tid >> 5, cc = tid & 31;
      if (c0 + cc < G.Cs) atomicAdd(stats_a + (long long)k * G.Cs + c0 + cc, s_st[tid]);
    }
    step += u.r1 - u.r0;
  }
}

// ------------------------------------------------------------------------------------------------
// stride (1,2,2): first block of every stage.  Same ring machinery; the forward consumes two input rows per
// output row (strip = 16 output columns = 33 input columns), the backward produces two input rows per step from
// two dy rows (strip = 32 input columns = 17 dy columns) — an input pixel only receives the taps whose parity
// matches: even rows kh = 1, odd rows kh = 0 and 2, same for columns.
// ------------------------------------------------------------------------------------------------
struct Geo2 {
  int N, IH, IW, OH, OW, C, Cs;
  int nCB, nSEG;
  int R;
  long long total_steps, steps_per_cta;
};

struct Unit2 { int n, cb, sg, r0, r1; };

__device__ __forceinline__ Unit2 decode2(long long step, long long s1, int rows, int nSEG, int nCB) {
  Unit2 u;
  const int unit = (int)(step / rows);
  u.r0 = (int)(step - (long long)unit * rows);
  const long long left = s1 - step;
  u.r1 = (long long)(rows - u.r0) < left ? rows : u.r0 + (int)left;
  u.sg = unit % nSEG;
  const int t = unit / nSEG;
  u.cb = t % nCB;
  u.n = t / nCB;
  return u;
}

template <int T>
__global__ void __launch_bounds__(NT, 2) dw_fwd_ring2_kernel(const float* __restrict__ X, const float* __restrict__ bnp,
                                                             const float* __restrict__ w, float* __restrict__ Y,
                                                             double* __restrict__ stats, const Geo2 G) {
  pdl_trigger();   // programmatic dependent launch: let the next kernel start its setup,
  pdl_wait();      // then wait for the earlier kernels whose results this one reads

  constexpr int SF = T * TS;
  extern __shared__ __align__(16) float sm[];
  float* ring = sm;                                   // [R][T][PIX][CB] input rows
  float* s_bn = ring + (size_t)G.R * SF;              // [3][CB]
  double* s_st = reinterpret_cast<double*>(s_bn + 4 * CB);
  const int tid = threadIdx.x;
  const int cl = tid & 31, grp = tid >> 5;            // channel, pair of output columns
  const uint32_t ring_u32 = smem_u32(ring);
  const long long img_in = (long long)G.IH * G.IW * G.Cs, img_out = (long long)G.OH * G.OW * G.Cs;
  const long long rowstride = (long long)G.IW * G.Cs;

  Pieces<T> pc;
  pc.init(tid, G.IH, G.IW, G.Cs);

  const long long s0 = (long long)blockIdx.x * G.steps_per_cta;
  long long s1 = s0 + G.steps_per_cta;
  if (s1 > G.total_steps) s1 = G.total_steps;

  long long step = s0;
  while (step < s1) {
    const Unit2 u = decode2(step, s1, G.OH, G.nSEG, G.nCB);
    const int c0 = u.cb * CB, ow0 = u.sg * 16;
    const int c = c0 + cl;
    const bool c_ok = c < G.Cs;

    __syncthreads();
    if (tid < 3 * CB) {
      const int k = tid >> 5, cc = tid & 31;
      const int src = k == 0 ? 0 : k == 1 ? 2 : 3;
      s_bn[tid] = (c0 + cc < G.Cs) ? __ldg(bnp + src * G.Cs + c0 + cc) : 0.f;
    }
    if (tid < 2 * CB) s_st[tid] = 0.0;
    pc.set_unit(tid, 2 * ow0, c0, G.IW, G.Cs);        // ring pixel p <-> input column 2*ow0 - 1 + p
    float wr[27];
#pragma unroll
    for (int t = 0; t < 27; ++t) wr[t] = (c < G.C) ? __ldg(w + c * 27 + t) : 0.f;
    const float* Xu = X + (long long)u.n * T * img_in + (long long)(2 * ow0 - 1) * G.Cs + c0;
    float* Yt = Y + (long long)u.n * T * img_out + ((long long)u.r0 * G.OW + ow0 + 2 * grp) * G.Cs + c;
    int nvalid = G.OW - (ow0 + 2 * grp);
    nvalid = nvalid < 0 ? 0 : nvalid > 2 ? 2 : nvalid;
    if (!c_ok) nvalid = 0;
    double st_s = 0.0, st_q = 0.0;

    // input row r lives in slot (r - (2 r0 - 1)) mod R
    int issue_row = 2 * u.r0 - 1, issue_slot = 0;
    auto issue_one = [&]() {
      const bool row_ok = issue_row >= 0 && issue_row < G.IH;
      const uint32_t dst0 = ring_u32 + (uint32_t)(issue_slot * SF) * 4;
      const float* src0 = Xu + (long long)issue_row * rowstride;
#pragma unroll
      for (int i = 0; i < Pieces<T>::NE; ++i) {
        if (pc.exists >> i & 1) {
          const bool ok = row_ok && (pc.ok >> i & 1);
          cp_async16(dst0 + (uint32_t)pc.s_off[i] * 4, ok ? (const void*)(src0 + pc.g_off[i]) : (const void*)X, ok ? 16u : 0u);
        }
      }
      ++issue_row;
      if (++issue_slot == G.R) issue_slot = 0;
    };
    auto transform = [&](int row, int slot) {
      if (row < 0 || row >= G.IH) return;
      float* base = ring + (size_t)slot * SF;
      const int c4 = (tid & 7) * 4;
      const float4 mean = *reinterpret_cast<const float4*>(s_bn + c4);
      const float4 scale = *reinterpret_cast<const float4*>(s_bn + CB + c4);
      const float4 beta = *reinterpret_cast<const float4*>(s_bn + 2 * CB + c4);
#pragma unroll
      for (int i = 0; i < Pieces<T>::NE; ++i) {
        if (pc.ok >> i & 1) {
          float4* p = reinterpret_cast<float4*>(base + pc.s_off[i]);
          *p = f4relu(f4bn(*p, mean, scale, beta));
        }
      }
    };
    auto wrap = [&](int x) { return x >= G.R ? x - G.R : x; };

    // group 0 = rows 2r0-1, 2r0, 2r0+1; group k = rows 2(r0+k), 2(r0+k)+1
    issue_one(); issue_one(); issue_one();
    cp_async_commit();
    __syncthreads();                                  // s_bn visible
    int sl = 0;                                       // slot of row 2*oh - 1
    for (int oh = u.r0; oh < u.r1; ++oh) {
      issue_one(); issue_one();
      cp_async_commit();
      cp_async_wait<1>();
      const int sl1 = wrap(sl + 1), sl2 = wrap(sl + 2);
      if (oh == u.r0) transform(2 * oh - 1, sl);
      transform(2 * oh, sl1);
      transform(2 * oh + 1, sl2);
      __syncthreads();
      if (nvalid > 0) {
        const int toff = (4 * grp) * CB + cl;         // output pair (2g, 2g+1) reads ring pixels 4g .. 4g+4
        const float* b0 = ring + sl * SF + toff;
        const float* b1 = ring + sl1 * SF + toff;
        const float* b2 = ring + sl2 * SF + toff;
        float acc[T][2];
#pragma unroll
        for (int t = 0; t < T; ++t) { acc[t][0] = 0.f; acc[t][1] = 0.f; }
#pragma unroll
        for (int ti = 0; ti < T; ++ti) {
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const float* rp = (kh == 0 ? b0 : kh == 1 ? b1 : b2) + ti * TS;
            float in[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) in[j] = rp[j * CB];
#pragma unroll
            for (int kt = 0; kt < 3; ++kt) {
              const int to = ti - kt + 1;
              if (to < 0 || to >= T) continue;
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) {
                const float wv = wr[kt * 9 + kh * 3 + kw];
                acc[to][0] = fmaf(wv, in[kw], acc[to][0]);
                acc[to][1] = fmaf(wv, in[2 + kw], acc[to][1]);
              }
            }
          }
        }
        float sf = 0.f, qf = 0.f;
#pragma unroll
        for (int t = 0; t < T; ++t) {
          float* yp = Yt + (long long)t * img_out;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (q < nvalid) {
              yp[q * G.Cs] = acc[t][q];
              sf += acc[t][q];
              qf = fmaf(acc[t][q], acc[t][q], qf);
            }
          }
        }
        st_s += (double)sf;
        st_q += (double)qf;
      }
      Yt += (long long)G.OW * G.Cs;
      sl = sl2;
    }
    cp_async_wait<0>();
    if (stats) {
      if (nvalid > 0) { atomicAdd(&s_st[cl], st_s); atomicAdd(&s_st[CB + cl], st_q); }
      __syncthreads();
      if (tid < 2 * CB) {
        const int k = tid >> 5, cc = tid & 31;
        if (c0 + cc < G.Cs) atomicAdd(stats + ((long long)u.n * 2 + k) * G.Cs + c0 + cc, s_st[tid]);
      }
    }
    step += u.r1 - u.r0;
  }
}

// backward, stride 2: step m produces input rows 2m, 2m+1 of a 32-column strip from dy rows m, m+1
template <int T>
__global__ void __launch_bounds__(NT, 1) dw_bwd_ring2_kernel(const float* __restrict__ DY, const float* __restrict__ YA,
                                                             const float* __restrict__ bnp_a, const float* __restrict__ w,
                                                             float* __restrict__ DR, float* __restrict__ dW,
                                                             double* __restrict__ stats_a, const Geo2 G) {
  pdl_trigger();   // programmatic dependent launch: let the next kernel start its setup,
  pdl_wait();      // then wait for the earlier kernels whose results this one reads

  constexpr int SF = T * TS;
  extern __shared__ __align__(16) float sm[];
  float* ring = sm;                                                  // [R][T][PIX][CB] dy rows
  float* s_dw = ring + (size_t)G.R * SF;                             // [27][CB]
  double* s_st = reinterpret_cast<double*>(s_dw + 28 * CB);          // [2][CB]
  const int tid = threadIdx.x;
  const int cl = tid & 31, grp = tid >> 5;                           // channel, group of 4 input columns
  const uint32_t ring_u32 = smem_u32(ring);
  const long long img_in = (long long)G.IH * G.IW * G.Cs, img_out = (long long)G.OH * G.OW * G.Cs;
  const long long rowstride_o = (long long)G.OW * G.Cs, rowstride_i = (long long)G.IW * G.Cs;
  const int nsteps = (G.IH + 1) >> 1;                                // input row pairs

  Pieces<T> pc;
  pc.init(tid, G.OH, G.OW, G.Cs);

  const long long s0 = (long long)blockIdx.x * G.steps_per_cta;
  long long s1 = s0 + G.steps_per_cta;
  if (s1 > G.total_steps) s1 = G.total_steps;

  long long step = s0;
  while (step < s1) {
    const Unit2 u = decode2(step, s1, nsteps, G.nSEG, G.nCB);
    const int c0 = u.cb * CB, iw0 = u.sg * SEGW, oc0 = iw0 >> 1;      // first dy column of the strip
    const int c = c0 + cl;
    const bool c_ok = c < G.Cs;

    __syncthreads();
    for (int i = tid; i < 27 * CB; i += NT) s_dw[i] = 0.f;
    if (tid < 2 * CB) s_st[tid] = 0.0;
    pc.set_unit(tid, oc0 + 1, c0, G.OW, G.Cs);                       // ring pixel p <-> dy column oc0 + p
    float wr[27], dwacc[27];
#pragma unroll
    for (int t = 0; t < 27; ++t) { wr[t] = (c < G.C) ? __ldg(w + c * 27 + t) : 0.f; dwacc[t] = 0.f; }
    float mean_a = 0.f, rstd_a = 0.f, scale_a = 0.f, beta_a = 0.f;
    if (c_ok) {
      mean_a = __ldg(bnp_a + c); rstd_a = __ldg(bnp_a + G.Cs + c);
      scale_a = __ldg(bnp_a + 2 * G.Cs + c); beta_a = __ldg(bnp_a + 3 * G.Cs + c);
    }
    const float* DYu = DY + (long long)u.n * T * img_out + (long long)oc0 * G.Cs + c0;
    const long long toff_g = (long long)u.n * T * img_in + ((long long)(2 * u.r0) * G.IW + iw0 + 4 * grp) * G.Cs + c;
    const float* YAt = YA + toff_g;                                  // frame 0, input row 2*r0
    float* DRt = DR + toff_g;
    int nvalid = G.IW - (iw0 + 4 * grp);
    nvalid = nvalid < 0 ? 0 : nvalid > 4 ? 4 : nvalid;
    if (!c_ok) nvalid = 0;
    double st_s = 0.0, st_t = 0.0;

    // dy row r lives in slot (r - r0) mod R
    int issue_row = u.r0, issue_slot = 0;
    auto issue = [&]() {
      const bool row_ok = issue_row >= 0 && issue_row < G.OH;
      const uint32_t dst0 = ring_u32 + (uint32_t)(issue_slot * SF) * 4;
      const float* src0 = DYu + (long long)issue_row * rowstride_o;
#pragma unroll
      for (int i = 0; i < Pieces<T>::NE; ++i) {
        if (pc.exists >> i & 1) {
          const bool ok = row_ok && (pc.ok >> i & 1);
          cp_async16(dst0 + (uint32_t)pc.s_off[i] * 4, ok ? (const void*)(src0 + pc.g_off[i]) : (const void*)DY, ok ? 16u : 0u);
        }
      }
      cp_async_commit();
      ++issue_row;
      if (++issue_slot == G.R) issue_slot = 0;
    };

    issue(); issue(); issue();                        // rows r0, r0+1, r0+2
    int sl = 0;                                       // slot of dy row m
    for (int m = u.r0; m < u.r1; ++m) {
      issue();                                        // row m + 3
      const int nrows = (2 * m + 1 < G.IH) ? 2 : 1;   // odd image height: the last pair has one row
      float ya[T][2][4];
#pragma unroll
      for (int t = 0; t < T; ++t)
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
          for (int x = 0; x < 4; ++x)
            ya[t][e][x] = (x < nvalid && e < nrows) ? __ldg(YAt + (long long)t * img_in + (long long)e * rowstride_i + x * G.Cs) : 0.f;
      cp_async_wait<2>();                             // rows m, m + 1 have landed
      const int sl1 = sl + 1 >= G.R ? sl + 1 - G.R : sl + 1;
      __syncthreads();
      if (nvalid > 0) {
        const int toff = (2 * grp) * CB + cl;          // dy columns oc0 + 2g .. oc0 + 2g + 2
        const float* b0 = ring + sl * SF + toff;       // dy row m
        const float* b1 = ring + sl1 * SF + toff;      // dy row m + 1
        float a[T][2][4], da[T][2][4];
#pragma unroll
        for (int t = 0; t < T; ++t)
#pragma unroll
          for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int x = 0; x < 4; ++x) {
              a[t][e][x] = (x < nvalid && e < nrows) ? fmaxf(fmaf(ya[t][e][x] - mean_a, scale_a, beta_a), 0.f) : 0.f;
              da[t][e][x] = 0.f;
            }
#pragma unroll
        for (int to = 0; to < T; ++to) {
          float d0[3], d1[3];                          // dy rows m / m + 1 at columns h, h + 1, h + 2
#pragma unroll
          for (int j = 0; j < 3; ++j) { d0[j] = b0[to * TS + j * CB]; d1[j] = b1[to * TS + j * CB]; }
#pragma unroll
          for (int kt = 0; kt < 3; ++kt) {
            const int ti = to + kt - 1;
            if (ti < 0 || ti >= T) continue;
            // (input row e, kh, dy row): (0, 1, m), (1, 0, m + 1), (1, 2, m)
            // (input col x, kw, dy col j): (0,1,0) (1,0,1) (1,2,0) (2,1,1) (3,0,2) (3,2,1)
#define DW2_TAP(e, kh, dv, x, kw, j)                                             \
            {                                                                    \
              const int tap = kt * 9 + (kh) * 3 + (kw);                          \
              da[ti][e][x] = fmaf(wr[tap], dv[j], da[ti][e][x]);                 \
              dwacc[tap] = fmaf(a[ti][e][x], dv[j], dwacc[tap]);                 \
            }
#define DW2_ROW(e, kh, dv)                                                       \
            DW2_TAP(e, kh, dv, 0, 1, 0) DW2_TAP(e, kh, dv, 1, 0, 1) DW2_TAP(e, kh, dv, 1, 2, 0) \
            DW2_TAP(e, kh, dv, 2, 1, 1) DW2_TAP(e, kh, dv, 3, 0, 2) DW2_TAP(e, kh, dv, 3, 2, 1)
            DW2_ROW(0, 1, d0)
            DW2_ROW(1, 0, d1)
            DW2_ROW(1, 2, d0)
#undef DW2_ROW
#undef DW2_TAP
          }
        }
        float sf = 0.f, tf = 0.f;
#pragma unroll
        for (int t = 0; t < T; ++t)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            if (e >= nrows) continue;
            float* dp = DRt + (long long)t * img_in + (long long)e * rowstride_i;
#pragma unroll
            for (int x = 0; x < 4; ++x) {
              if (x < nvalid) {
                const float v = a[t][e][x] > 0.f ? da[t][e][x] : 0.f;
                dp[x * G.Cs] = v;
                sf += v;
                tf = fmaf(v, (ya[t][e][x] - mean_a) * rstd_a, tf);
              }
            }
          }
        st_s += (double)sf;
        st_t += (double)tf;
      }
      YAt += 2 * rowstride_i;
      DRt += 2 * rowstride_i;
      sl = sl1;
    }
    cp_async_wait<0>();
    if (nvalid > 0) {
#pragma unroll
      for (int t = 0; t < 27; ++t) atomicAdd(&s_dw[t * CB + cl], dwacc[t]);
      atomicAdd(&s_st[cl], st_s);
      atomicAdd(&s_st[CB + cl], st_t);
    }
    __syncthreads();
    for (int i = tid; i < 27 * CB; i += NT) {
      const int t = i >> 5, cc = i & 31;
      if (c0 + cc < G.C) atomicAdd(dW + (c0 + cc) * 27 + t, s_dw[i]);
    }
    if (tid < 2 * CB) {
      const int k = tid >> 5, cc = tid & 31;
      if (c0 + cc < G.Cs) atomicAdd(stats_a + (long long)k * G.Cs + c0 + cc, s_st[tid]);
    }
    step += u.r1 - u.r0;
  }
}

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

static bool plan(Geo& G, int T, int N, int IH, int IW, int C, int Cs, size_t extra_smem, size_t& smem, int& grid, int ctas_per_sm,
                 bool second_ring = false) {
  if (IW < 1 || (Cs & 3) || T < 3 || T > 5) return false;
  if ((long long)T * IH * IW * Cs >= (1LL << 31)) return false;        // per-sample offsets are 32-bit
  G.N = N; G.IH = IH; G.IW = IW; G.C = C; G.Cs = Cs;
  G.nCB = (Cs + CB - 1) / CB;
  G.nSEG = (IW + SEGW - 1) / SEGW;
  const size_t slot = (size_t)T * TS * 4;
  int dev = 0, sms = 148, max_smem = 227 * 1024;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const size_t budget = (size_t)(max_smem + 1024) / ctas_per_sm - 1024;
  // rows requested ahead of the computed row (C3D_DW_DEPTH, 1..4; default 2: depths 2 / 3 / 4 measured identical at every
  // stage, profiles/r02_summary.md section 2.4, and the shallower ring leaves shared memory to co-resident kernels)
  static const int dmax = getenv("C3D_DW_DEPTH") ? atoi(getenv("C3D_DW_DEPTH")) : 2;
  G.D = dmax < 1 ? 1 : dmax > 4 ? 4 : dmax;
  for (;; --G.D) {
    G.R = G.D + 4; G.R2 = second_ring ? G.D + 2 : 0;
    if ((G.R + G.R2) * slot + extra_smem <= budget) break;
    if (G.D == 1) return false;
  }
  smem = (G.R + G.R2) * slot + extra_smem;
  G.total_steps = (long long)N * G.nCB * G.nSEG * IH;
  long long g = (long long)sms * ctas_per_sm;
  if (g > G.total_steps) g = G.total_steps;
  G.steps_per_cta = (G.total_steps + g - 1) / g;
  grid = (int)((G.total_steps + G.steps_per_cta - 1) / G.steps_per_cta);
  return true;
}

}  // namespace dwr

// Returns -1 when the shape is not handled here (caller uses the generic kernel), else a C3D status.
int c3d_launch_dw_fwd_ring(const float* X, const float* bnp, const float* w, float* Y, double* stats, int N, int T, int IH,
                           int IW, int C, int Cs, cudaStream_t st) {
  if (!dwr::env_int("C3D_DW_RING", 1)) return -1;
  dwr::Geo G;
  size_t smem = 0;
  int grid = 0;
  const size_t extra = (size_t)(4 * dwr::CB * 4 + 2 * dwr::CB * 8);        // s_bn [4][CB] floats + s_st [2][CB] doubles
  // two CTAs per SM when the ring fits twice (T = 3 always; T = 4 with depth 2, T = 5 with depth 1), else one.
  // C3D_DW_OCC=1 restores the round-1 choice (one CTA per SM for T >= 4).
  // C3D_DW_FWD_OCC=1: one CTA per SM with the whole register file for every T (timing experiment switch)
  const bool occ1 = (dwr::env_int("C3D_DW_OCC", 2) < 2 && T != 3) || dwr::env_int("C3D_DW_FWD_OCC", 2) < 2;
  bool one = occ1;
  if (occ1 || !dwr::plan(G, T, N, IH, IW, C, Cs, extra, smem, grid, 2)) {
    if (!dwr::plan(G, T, N, IH, IW, C, Cs, extra, smem, grid, 1)) return -1;
    one = true;
  }
  // channel-pair FFMA2 arithmetic (C3D_DW_PAIR=0: one channel x four columns per thread, scalar FFMA); pairs need an even
  // channel stride and 8-byte aligned tensors
  const bool pair = dwr::env_int("C3D_DW_PAIR", 1) && (Cs & 1) == 0 && ((reinterpret_cast<uintptr_t>(Y) & 7) == 0);
  cudaError_t e = cudaSuccess;
#define FWD_GO(T_, P_, O_) do {                                                                                          \
    e = cudaFuncSetAttribute(dwr::dw_fwd_ring_kernel<T_, P_, O_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return C3D_ERR_SMEM;                                                                          \
    c3d_launch_pdl(dwr::dw_fwd_ring_kernel<T_, P_, O_>, dim3(grid), dim3(dwr::NT), smem, st, X, bnp, w, Y, stats, G); } while (0)
#define FWD_T(T_) do { if (pair) { if (one) FWD_GO(T_, true, true); else FWD_GO(T_, true, false); }                      \
                       else { if (one) FWD_GO(T_, false, true); else FWD_GO(T_, false, false); } } while (0)
  switch (T) {
    case 3: FWD_T(3); break;
    case 4: FWD_T(4); break;
    default: FWD_T(5); break;
  }
#undef FWD_T
#undef FWD_GO
  return c3d_check_last(cudaGetLastError());
}

// fuse == nullptr: dy = du already transformed by the elementwise pre-pass (dw_dy_kernel); else the raw du plus the
// BN_b / SE backward operands (transform fused into the ring fill).  Returns -1 when not handled here.
template <int T, bool ONE>
static int launch_bwd(const float* dy, const float* ya, const float* bnp_a, const float* w, float* dr, float* dW, double* stats_a,
                      const dwr::Geo& G, const dwr::FuseArgs* fuse, int grid, size_t smem, cudaStream_t st) {
  cudaError_t e;
  if (fuse) {
    e = cudaFuncSetAttribute(dwr::dw_bwd_ring_kernel<T, true, ONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return C3D_ERR_SMEM;
    c3d_launch_pdl(dwr::dw_bwd_ring_kernel<T, true, ONE>, dim3(grid), dim3(dwr::NT), smem, st, dy, ya, bnp_a, w, dr, dW, stats_a, G, *fuse);
  } else {
    dwr::FuseArgs none = {nullptr, nullptr, nullptr, nullptr, nullptr};
    e = cudaFuncSetAttribute(dwr::dw_bwd_ring_kernel<T, false, ONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return C3D_ERR_SMEM;
    c3d_launch_pdl(dwr::dw_bwd_ring_kernel<T, false, ONE>, dim3(grid), dim3(dwr::NT), smem, st, dy, ya, bnp_a, w, dr, dW, stats_a, G, none);
  }
  return c3d_check_last(cudaGetLastError());
}

int c3d_launch_dw_bwd_ring(const float* dy, const float* ya, const float* bnp_a, const float* w, float* dr, float* dW,
                           double* stats_a, int N, int T, int IH, int IW, int C, int Cs, cudaStream_t st, const float* yb,
                           const float* bnp_b, const float* gate, const float* dpool, const float* coef_b) {
  if (!dwr::env_int("C3D_DW_RING", 1)) return -1;
  dwr::Geo G;
  size_t smem = 0;
  int grid = 0;
  const bool fused = yb != nullptr;
  // s_dw [27][CB] floats (+ pad) + s_st [2][CB] doubles + s_cb [7][CB] floats
  const size_t extra = (size_t)(28 * dwr::CB * 4 + 2 * dwr::CB * 8 + 8 * dwr::CB * 4);
  // C3D_DW_BWD_OCC: CTAs per SM of the T = 3 kernel (2: 128-register build, 1: whole register file, one CTA per SM)
  const bool one3 = dwr::env_int("C3D_DW_BWD_OCC", 1) < 2;      // measured: 9.67 -> 9.34 ms per step (profiles/r02_summary.md section 2.7)
  if (!dwr::plan(G, T, N, IH, IW, C, Cs, extra, smem, grid, (T == 3 && !one3) ? 2 : 1, fused)) return -1;
  dwr::FuseArgs F = {yb, bnp_b, gate, dpool, coef_b};
  const dwr::FuseArgs* fp = fused ? &F : nullptr;
  switch (T) {
    case 3:
      if (one3) return launch_bwd<3, true>(dy, ya, bnp_a, w, dr, dW, stats_a, G, fp, grid, smem, st);
      return launch_bwd<3, false>(dy, ya, bnp_a, w, dr, dW, stats_a, G, fp, grid, smem, st);
    case 4: return launch_bwd<4, true>(dy, ya, bnp_a, w, dr, dW, stats_a, G, fp, grid, smem, st);
    default: return launch_bwd<5, true>(dy, ya, bnp_a, w, dr, dW, stats_a, G, fp, grid, smem, st);
  }
}

// ---- stride-2 launchers (first block of each stage).  Return -1 when the shape is not handled here. ----
static bool plan2(dwr::Geo2& G, int T, int N, int IH, int IW, int C, int Cs, int R, int rows, int nseg, size_t extra, size_t& smem,
                  int& grid, int ctas_per_sm) {
  if (IW < 1 || IH < 1 || (Cs & 3) || T < 3 || T > 5) return false;
  if ((long long)T * IH * IW * Cs >= (1LL << 31)) return false;
  G.N = N; G.IH = IH; G.IW = IW; G.OH = (IH - 1) / 2 + 1; G.OW = (IW - 1) / 2 + 1; G.C = C; G.Cs = Cs;
  G.nCB = (Cs + dwr::CB - 1) / dwr::CB;
  G.nSEG = nseg;
  G.R = R;
  int dev = 0, sms = 148, max_smem = 227 * 1024;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  smem = (size_t)R * T * dwr::TS * 4 + extra;
  if (smem > (size_t)max_smem) return false;
  if (ctas_per_sm == 2 && 2 * (smem + 1024) > (size_t)max_smem + 1024) ctas_per_sm = 1;
  G.total_steps = (long long)N * G.nCB * nseg * rows;
  long long g = (long long)sms * ctas_per_sm;
  if (g > G.total_steps) g = G.total_steps;
  G.steps_per_cta = (G.total_steps + g - 1) / g;
  grid = (int)((G.total_steps + G.steps_per_cta - 1) / G.steps_per_cta);
  return true;
}

int c3d_launch_dw_fwd_ring2(const float* X, const float* bnp, const float* w, float* Y, double* stats, int N, int T, int IH,
                            int IW, int C, int Cs, cudaStream_t st) {
  if (!dwr::env_int("C3D_DW_RING", 1) || !dwr::env_int("C3D_DW_RING2", 1)) return -1;
  dwr::Geo2 G;
  size_t smem = 0;
  int grid = 0;
  const int OH = (IH - 1) / 2 + 1, OW = (IW - 1) / 2 + 1;
  const size_t extra = (size_t)(4 * dwr::CB * 4 + 2 * dwr::CB * 8);
  if (!plan2(G, T, N, IH, IW, C, Cs, 7, OH, (OW + 15) / 16, extra, smem, grid, 2)) return -1;
  cudaError_t e;
#define C3D_LAUNCH_FWD2(TT)                                                                                                 \
  e = cudaFuncSetAttribute(dwr::dw_fwd_ring2_kernel<TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
  if (e != cudaSuccess) return C3D_ERR_SMEM;                                                                                 \
  c3d_launch_pdl(dwr::dw_fwd_ring2_kernel<TT>, dim3(grid), dim3(dwr::NT), smem, st, X, bnp, w, Y, stats, G);
  if (T == 3) { C3D_LAUNCH_FWD2(3) } else if (T == 4) { C3D_LAUNCH_FWD2(4) } else { C3D_LAUNCH_FWD2(5) }
#undef C3D_LAUNCH_FWD2
  return c3d_check_last(cudaGetLastError());
}

int c3d_launch_dw_bwd_ring2(const float* dy, const float* ya, const float* bnp_a, const float* w, float* dr, float* dW,
                            double* stats_a, int N, int T, int IH, int IW, int C, int Cs, cudaStream_t st) {
  if (!dwr::env_int("C3D_DW_RING", 1) || !dwr::env_int("C3D_DW_RING2", 1)) return -1;
  dwr::Geo2 G;
  size_t smem = 0;
  int grid = 0;
  const size_t extra = (size_t)(28 * dwr::CB * 4 + 2 * dwr::CB * 8);
  if (!plan2(G, T, N, IH, IW, C, Cs, 5, (IH + 1) / 2, (IW + dwr::SEGW - 1) / dwr::SEGW, extra, smem, grid, 1)) return -1;
  cudaError_t e;
#define C3D_LAUNCH_BWD2(TT)                                                                                                 \
  e = cudaFuncSetAttribute(dwr::dw_bwd_ring2_kernel<TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
  if (e != cudaSuccess) return C3D_ERR_SMEM;                                                                                 \
  c3d_launch_pdl(dwr::dw_bwd_ring2_kernel<TT>, dim3(grid), dim3(dwr::NT), smem, st, dy, ya, bnp_a, w, dr, dW, stats_a, G);
  if (T == 3) { C3D_LAUNCH_BWD2(3) } else if (T == 4) { C3D_LAUNCH_BWD2(4) } else { C3D_LAUNCH_BWD2(5) }
#undef C3D_LAUNCH_BWD2
  return c3d_check_last(cudaGetLastError());
}
