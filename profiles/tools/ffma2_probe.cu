// Microbenchmark for the next depthwise-kernel step (profiles/r01_summary.md section 6): issue rate of scalar FFMA
// versus packed FFMA2 (fma.rn.f32x2) on sm_100a, with the register pressure of a 27-tap stencil (27 weights live).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe profiles/tools/ffma2_probe.cu && ./ffma2_probe
// Prints FMA/s per variant; on a B200 the scalar peak is 148 SMs x 128 lanes x SM clock.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pack(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

constexpr int TAPS = 27, ITERS = 4096, ACC = 8;     // 8 independent accumulator chains hide the FMA latency

__global__ void __launch_bounds__(256) scalar_kernel(const float* __restrict__ w, float* __restrict__ out, float x0) {
  float wt[TAPS];
#pragma unroll
  for (int t = 0; t < TAPS; ++t) wt[t] = w[t];
  float acc[ACC];
#pragma unroll
  for (int j = 0; j < ACC; ++j) acc[j] = threadIdx.x * 1e-3f + j;
  float x = x0;
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int t = 0; t < TAPS; ++t) {
#pragma unroll
      for (int j = 0; j < ACC; ++j) acc[j] = fmaf(wt[t], x, acc[j]);
    }
    x += 1e-7f;
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < ACC; ++j) s += acc[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) packed_kernel(const float* __restrict__ w, float* __restrict__ out, float x0) {
  unsigned long long wt[TAPS];
#pragma unroll
  for (int t = 0; t < TAPS; ++t) wt[t] = pack(w[t], w[t]);
  unsigned long long acc[ACC / 2];
#pragma unroll
  for (int j = 0; j < ACC / 2; ++j) acc[j] = pack(threadIdx.x * 1e-3f + 2 * j, threadIdx.x * 1e-3f + 2 * j + 1);
  float x = x0;
  for (int i = 0; i < ITERS; ++i) {
    const unsigned long long xx = pack(x, x);
#pragma unroll
    for (int t = 0; t < TAPS; ++t) {
#pragma unroll
      for (int j = 0; j < ACC / 2; ++j) acc[j] = ffma2(wt[t], xx, acc[j]);
    }
    x += 1e-7f;
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < ACC / 2; ++j) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[j]));
    s += lo + hi;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float *w, *out;
  cudaMalloc(&w, TAPS * sizeof(float));
  cudaMalloc(&out, (size_t)sms * 8 * 256 * sizeof(float));
  float hw[TAPS];
  for (int t = 0; t < TAPS; ++t) hw[t] = 1e-3f * (t + 1);
  cudaMemcpy(w, hw, sizeof(hw), cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const double fmas = (double)sms * 8 * 256 * ITERS * TAPS * ACC;
  for (int variant = 0; variant < 2; ++variant) {
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      if (variant == 0) scalar_kernel<<<sms * 8, 256>>>(w, out, 0.5f);
      else packed_kernel<<<sms * 8, 256>>>(w, out, 0.5f);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      best = ms < best ? ms : best;
    }
    printf("%s: %.3f ms, %.2f T FMA/s (%d SMs)\n", variant == 0 ? "FFMA  (scalar)" : "FFMA2 (f32x2) ", best, fmas / best / 1e9, sms);
  }
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
