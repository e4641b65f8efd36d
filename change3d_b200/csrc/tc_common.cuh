// Shared pieces of the tcgen05 kernels (pointwise GEMM and weight gradient): PTX wrappers for mbarrier / TMEM /
// tcgen05.mma, the round-to-nearest TF32 hi/lo split, and the fused operand prologues.
#pragma once
#include <cuda.h>   // CUtensorMap (type only: the encoder is fetched with cudaGetDriverEntryPoint, no libcuda link)

#include "pw_gemm.cuh"

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Wait for the phase with the given parity.  `hint_ns` > 0 lets the hardware suspend the thread for up to that
// long per probe (fewer issue slots burnt by spinning, but the wake-up can lag by as much); 0 = plain try_wait.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t hint_ns = 0) {
  uint32_t done;
  if (hint_ns) {
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(bar), "r"(parity), "r"(hint_ns) : "memory");
    } while (!done);
  } else {
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
  }
}
// TMA: one elected thread arms the barrier with the byte count, then issues the tile copy (global -> shared)
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
// L2 prefetch of one box of a tensor map (no shared memory, no barrier): issued a few tiles ahead of the loads so that
// the bytes in flight from HBM are not bounded by the pipeline stages that fit in shared memory
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* tm, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tm), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// Predicated forms for warp-uniform issue loops: every lane executes the statement (uniform control flow, so the
// descriptors can live in uniform registers), the instruction itself only runs where `issue` is non-zero.
__device__ __forceinline__ void umma_tf32_if(uint32_t issue, uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
               "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(issue) : "memory");
}
__device__ __forceinline__ void umma_commit_if(uint32_t issue, uint32_t bar) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
               "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar), "r"(issue) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// TMEM -> registers, 32 lanes x 32 columns per warp.  The load and its tcgen05.wait::ld sit in ONE asm statement:
// the destination registers are only valid after the wait, and a separate wait statement would let the compiler
// hoist arithmetic on them above it (nothing ties the registers to the wait).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* r) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr) : "memory");
}
// two accumulators (main + correction) in flight behind a single wait
__device__ __forceinline__ void tmem_ld32x2(uint32_t taddr_a, uint32_t taddr_b, float* ra, float* rb) {
  uint32_t* u = reinterpret_cast<uint32_t*>(ra);
  uint32_t* v = reinterpret_cast<uint32_t*>(rb);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%64];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%65];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31]),
        "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr_a), "r"(taddr_b) : "memory");
}

// UMMA shared-memory descriptor, K-major, no swizzle: element (row, k) of a [rows x K] tf32 tile lives at
//   (k/4) * LBO + (row/8) * SBO + (row%8) * 16 + (k%4) * 4   bytes.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;      // descriptor version (sm_100)
  return d;                     // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

// K-major operand with 64-byte rows in the 64B-swizzle layout TMA writes (CU_TENSOR_MAP_SWIZZLE_64B): the 16-byte
// unit u of row r lives at unit u ^ ((r >> 1) & 3); 8-row groups are 512 bytes apart (SBO); LBO is unused.
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;       // layout_type SWIZZLE_64B
  return d;
}

// hi = x rounded to nearest TF32 (ties away from zero, what cvt.rna.tf32.f32 computes — done with two integer
// instructions: the cvt compiles to four with its NaN / overflow guards, and the split runs once per staged
// element, the producers' hottest arithmetic), lo = x - hi exactly (fp32, <= 13 significant bits).  The tensor core
// reads the top 19 bits of an operand, so hi passes through unchanged and lo is truncated to its top 11 bits:
// |error| <= 2^-22 |x|, of either sign (lo is symmetric around zero because hi is rounded to nearest).
__device__ __forceinline__ float rna_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = rna_tf32(x);
  lo = x - hi;
}
__device__ __forceinline__ void split4(float4 v, float4& hi, float4& lo) {
  split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y); split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
}

// prologue applied to one float4 of 4 consecutive channels; per-channel parameters are passed in registers.
// PRO_BNBWD keeps folded constants: scale * (v - c1 - (y - mean) * rstd * c2) = fma(y - mean, k2, fma(v, scale, k0)) with
// k0 = -c1 * scale (stored in c1) and k2 = -rstd * c2 * scale (stored in c2): 3 instructions per element instead of 6.
struct ChanParams { float4 mean, rstd, scale, beta, c1, c2; };

template <int MODE>
__device__ __forceinline__ ChanParams load_chan_params(const TileSrc& s, int c) {
  ChanParams p;
  p.mean = p.rstd = p.scale = p.beta = p.c1 = p.c2 = f4zero();
  if (MODE == PRO_BN_RELU || MODE == PRO_BN_GATE_SWISH || MODE == PRO_BNBWD) {
    p.mean = ldg4(BNP_MEAN(s.bnp, s.ld) + c);
    p.scale = ldg4(BNP_SCALE(s.bnp, s.ld) + c);
  }
  if (MODE == PRO_BN_RELU || MODE == PRO_BN_GATE_SWISH) p.beta = ldg4(BNP_BETA(s.bnp, s.ld) + c);
  if (MODE == PRO_BNBWD) {
    p.rstd = ldg4(BNP_RSTD(s.bnp, s.ld) + c);
    const float4 c1 = ldg4(s.coef + c), c2 = ldg4(s.coef + s.ld + c);
    p.c1 = make_float4(-c1.x * p.scale.x, -c1.y * p.scale.y, -c1.z * p.scale.z, -c1.w * p.scale.w);
    p.c2 = make_float4(-p.rstd.x * c2.x * p.scale.x, -p.rstd.y * c2.y * p.scale.y, -p.rstd.z * c2.z * p.scale.z,
                       -p.rstd.w * c2.w * p.scale.w);
  }
  return p;
}

template <int MODE>
__device__ __forceinline__ float4 prologue(const ChanParams& p, float4 v, float4 v2, float4 gate4) {
  if (MODE == PRO_NONE) return v;
  if (MODE == PRO_BN_RELU) return f4relu(f4bn(v, p.mean, p.scale, p.beta));
  if (MODE == PRO_BN_GATE_SWISH) {
    v = f4mul(f4bn(v, p.mean, p.scale, p.beta), gate4);
    return make_float4(swishf_(v.x), swishf_(v.y), swishf_(v.z), swishf_(v.w));
  }
  if (MODE == PRO_BNBWD) {
    v.x = fmaf(v2.x - p.mean.x, p.c2.x, fmaf(v.x, p.scale.x, p.c1.x));
    v.y = fmaf(v2.y - p.mean.y, p.c2.y, fmaf(v.y, p.scale.y, p.c1.y));
    v.z = fmaf(v2.z - p.mean.z, p.c2.z, fmaf(v.z, p.scale.z, p.c1.z));
    v.w = fmaf(v2.w - p.mean.w, p.c2.w, fmaf(v.w, p.scale.w, p.c1.w));
    return v;
  }
  if (MODE == PRO_ABSDIFF) return make_float4(fabsf(v.x - v2.x), fabsf(v.y - v2.y), fabsf(v.z - v2.z), fabsf(v.w - v2.w));
  // PRO_MASK_POS
  v.x = v2.x > 0.f ? v.x : 0.f; v.y = v2.y > 0.f ? v.y : 0.f; v.z = v2.z > 0.f ? v.z : 0.f; v.w = v2.w > 0.f ? v.w : 0.f;
  return v;
}

// general row addressing (frame slices, stride-2 subsample): row -> element offsets of A / A2, image index
__device__ __forceinline__ void row_offsets(const TileSrc& s, uint32_t row, long long& off, long long& off2, uint32_t& img) {
  img = row / (uint32_t)s.OHW;
  const uint32_t rem = row - img * (uint32_t)s.OHW;
  const uint32_t oh = rem / (uint32_t)s.OW, ow = rem - oh * (uint32_t)s.OW;
  const int mul = (s.map == MAP_SUB2) ? 2 : 1;
  const long long pix = (long long)(oh * mul) * s.IW + ow * mul;
  off = (long long)img * s.img_stride + pix * s.ld;
  off2 = (long long)img * s.img_stride2 + pix * s.ld;
}


// Host: tensor map over a row-major fp32 matrix [rows][ld] of which the first `inner` columns are addressable;
// boxes of box_rows x 16 columns land in shared memory as 64-byte rows, 64B-swizzled.
static inline bool make_tmap_rows(CUtensorMap* tm, const float* base, long long inner, long long rows, long long ld, int box_rows) {
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
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld & 3) || inner < 16) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {16, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tc
