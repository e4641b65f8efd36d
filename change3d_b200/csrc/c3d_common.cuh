// change3d_b200 — shared device helpers (sm_100a).
// Activations are fp32 NDHWC: [N, T, H, W, Cs] with channel stride Cs (multiple of 4; the
// bottleneck's inner width is padded 54->56, 108->112 and the pad lanes are kept at zero).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define C3D_OK 0
#define C3D_ERR_ARG 1
#define C3D_ERR_CUDA 2
#define C3D_ERR_SMEM 3

// BN parameter block written by c3d_bn_finalize: float[4][Cs] = mean, rstd, scale(=gamma*rstd), beta
#define BNP_MEAN(p, Cs) (p)
#define BNP_RSTD(p, Cs) ((p) + (Cs))
#define BNP_SCALE(p, Cs) ((p) + 2 * (Cs))
#define BNP_BETA(p, Cs) ((p) + 3 * (Cs))

// prologue modes for an operand tile (TileSrc.mode)
enum { PRO_NONE = 0, PRO_BN_RELU = 1, PRO_BN_GATE_SWISH = 2, PRO_BNBWD = 3, PRO_ABSDIFF = 4, PRO_MASK_POS = 5 };
// row mappings (TileSrc.map)
enum { MAP_DENSE = 0, MAP_SUB2 = 1 };
// GEMM epilogues
enum { EPI_STORE = 0, EPI_RELU_ADD = 1, EPI_SWISH_BWD = 2, EPI_ADD2 = 3, EPI_RESERVED4 = 4, EPI_ABSDIFF_BWD = 5 };

// 1 / (1 + e^-v) with the SFU exponential and reciprocal (<= 2 ulp each; inf-safe: e^-v = inf -> 0)
__device__ __forceinline__ float sigmoidf_(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4fma(float4 a, float4 b, float4 c) {
  return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 f4relu(float4 a) {
  return make_float4(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f));
}
// (x - mean) * scale + beta
__device__ __forceinline__ float4 f4bn(float4 x, float4 mean, float4 scale, float4 beta) {
  return make_float4(fmaf(x.x - mean.x, scale.x, beta.x), fmaf(x.y - mean.y, scale.y, beta.y),
                     fmaf(x.z - mean.z, scale.z, beta.z), fmaf(x.w - mean.w, scale.w, beta.w));
}
__device__ __forceinline__ float swishf_(float u) { return u * sigmoidf_(u); }
__device__ __forceinline__ float swish_gradf_(float u) {
  float s = sigmoidf_(u);
  return s * (1.0f + u * (1.0f - s));
}

static inline int c3d_check_last(cudaError_t e) { return e == cudaSuccess ? C3D_OK : C3D_ERR_CUDA; }

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------
// Kernels launched through c3d_launch_pdl may become resident while the previous kernel of the stream is still
// running; everything they do before pdl_wait() must neither read data written earlier in the step nor write
// global memory.  pdl_wait() returns once every earlier kernel has completed and its writes are visible.
// pdl_trigger() lets the NEXT kernel's CTAs start being scheduled; it must come after this CTA has acquired
// everything a co-resident early CTA could take away from it (tensor memory columns).
// Both are no-ops for a kernel launched without the attribute (the default, see c3d_pdl_enabled).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

#include <stdlib.h>
// Off by default: measured on B200 at batch 32 the step is 2.3 % SLOWER with it (64.19 vs 62.72 ms, same box,
// profiles/r01_summary.md section 5) -- the early-resident CTAs of the next kernel compete with the weight-gradient
// GEMMs that already fill the dependency bubbles from the side stream.  C3D_PDL=1 turns it on (read per launch).
// C3D_PDL=2: only the single-CTA finalizer kernels (BN / SE statistics -> parameters) are launched that way: they become
// resident while their producer is still running (the producers trigger early) and only pay their own latency after it.
static inline int c3d_pdl_mode() {
  const char* v = getenv("C3D_PDL");
  return v ? atoi(v) : 0;
}
static inline bool c3d_pdl_enabled() { return c3d_pdl_mode() == 1; }

template <typename... KArgs, typename... Args>
static inline cudaError_t c3d_launch_pdl_if(bool on, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = on ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
static inline cudaError_t c3d_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                         Args&&... args) {
  return c3d_launch_pdl_if(c3d_pdl_mode() == 1, kernel, grid, block, smem, st, static_cast<Args&&>(args)...);
}
// finalizer kernels (one or a few tiny CTAs between two grid-filling kernels)
template <typename... KArgs, typename... Args>
static inline cudaError_t c3d_launch_pdl_small(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                               Args&&... args) {
  return c3d_launch_pdl_if(c3d_pdl_mode() >= 1, kernel, grid, block, smem, st, static_cast<Args&&>(args)...);
}

// Operand descriptor for the pointwise-GEMM family (see pw_gemm.cu).
struct TileSrc {
  const float* A;      // primary tensor
  const float* A2;     // secondary (y for PRO_BNBWD, second frame for PRO_ABSDIFF)
  const float* bnp;    // [4][ld]
  const float* coef;   // [2][ld]: c1 = mean(d), c2 = mean(d * yhat)      (PRO_BNBWD)
  const float* gate;   // [samples][ld]                                   (PRO_BN_GATE_SWISH, may be null)
  int mode, map;
  int ld;              // row stride (channels incl. pad) of A / A2
  int K;               // staged width (= ld)
  int OHW, OW;         // GEMM row -> (img, oh, ow)
  int IH, IW;          // source spatial dims
  long long img_stride, img_stride2;  // elements between consecutive images of A / A2
  int frames_per_sample;
};
