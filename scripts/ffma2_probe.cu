// FP32 FMA issue rate on sm_100a: scalar FFMA (three distinct register operands) against the packed fma.rn.f32x2
// (FFMA2: two FMAs per instruction on 64-bit register pairs).  Each thread runs NCH independent chains for ITERS
// iterations; prints GFMA/s per form.   nvcc -gencode arch=compute_100a,code=sm_100a -o scripts/ffma2_probe.bin scripts/ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int NCH = 16, ITERS = 4096;

__global__ void __launch_bounds__(256) k_scalar(float* out, float a0, float b0) {
  float acc[NCH], a[4], b[4];
#pragma unroll
  for (int i = 0; i < NCH; ++i) acc[i] = threadIdx.x * 1e-6f + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) { a[i] = a0 + i * 1e-3f; b[i] = b0 + i * 1e-3f + threadIdx.x * 1e-7f; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) acc[i] = fmaf(a[i & 3], b[(i >> 2) & 3], acc[i]);     // outer-product pattern: 4 x 4 block
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ unsigned long long pk(float x, float y) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

__global__ void __launch_bounds__(256) k_packed(float* out, float a0, float b0) {
  unsigned long long acc[NCH / 2], a[4], b[2];
#pragma unroll
  for (int i = 0; i < NCH / 2; ++i) acc[i] = pk(threadIdx.x * 1e-6f + 2 * i, threadIdx.x * 1e-6f + 2 * i + 1);
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = pk(a0 + i * 1e-3f, a0 + i * 1e-3f);                 // broadcast operand: (a_i, a_i)
#pragma unroll
  for (int i = 0; i < 2; ++i) b[i] = pk(b0 + 2 * i * 1e-3f + threadIdx.x * 1e-7f, b0 + (2 * i + 1) * 1e-3f + threadIdx.x * 1e-7f);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NCH / 2; ++i) acc[i] = fma2(a[i & 3], b[(i >> 2) & 1], acc[i]);  // the same 4 x 4 block, two columns per instruction
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NCH / 2; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(acc[i])); s += x + y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float* out;
  const int blocks = 148 * 8, threads = 256;
  cudaMalloc(&out, sizeof(float) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int form = 0; form < 2; ++form) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (form == 0) k_scalar<<<blocks, threads>>>(out, 1.0001f, 0.9999f);
      else k_packed<<<blocks, threads>>>(out, 1.0001f, 0.9999f);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double fma = (double)blocks * threads * NCH * ITERS;
      if (rep == 2) printf("%s: %.3f ms, %.1f GFMA/s (%.1f FMA/clk/SM at 1.92 GHz)\n", form == 0 ? "scalar FFMA      " : "packed fma.f32x2 ", ms,
                           fma / ms / 1e6, fma / ms / 1e6 / 148 / 1.92);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
