// Shared device helpers for libcvae_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/cvae_b200.h"

#define CVAE_LAUNCH_CHECK()                                   \
  do {                                                        \
    if (cudaPeekAtLastError() != cudaSuccess) {               \
      cudaGetLastError();                                     \
      return CVAE_ERR_LAUNCH;                                 \
    }                                                         \
  } while (0)

static inline cudaStream_t as_stream(cvae_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

namespace cvae {

constexpr int kNumSMs = 148;  // B200

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

// a[0..3] += x * w: as two packed fma.rn.f32x2 (FFMA2: two FMAs per instruction on 64-bit register pairs, the same
// round-to-nearest results).  nvcc never emits FFMA2 by itself; measured on this B200 (scripts/ffma2_probe.cu) scalar FFMA
// reaches 112 FMA/clk/SM, the packed form 127 with half the issue slots -- and these kernels spend ~40 % of their issue
// slots on the loads, address arithmetic and moves between the FMAs.  -DCVAE_FFMA2=0 restores the scalar form.
#ifndef CVAE_FFMA2
#define CVAE_FFMA2 1
#endif
__device__ __forceinline__ void fma4(float (&a)[4], const float x, const float4& w) {
#if CVAE_FFMA2
  unsigned long long xx, w01, w23, a01, a23;
  asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(x));
  asm("mov.b64 %0, {%1, %2};" : "=l"(w01) : "f"(w.x), "f"(w.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(w23) : "f"(w.z), "f"(w.w));
  asm("mov.b64 %0, {%1, %2};" : "=l"(a01) : "f"(a[0]), "f"(a[1]));
  asm("mov.b64 %0, {%1, %2};" : "=l"(a23) : "f"(a[2]), "f"(a[3]));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a01) : "l"(xx), "l"(w01));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a23) : "l"(xx), "l"(w23));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a[0]), "=f"(a[1]) : "l"(a01));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a[2]), "=f"(a[3]) : "l"(a23));
#else
  a[0] = fmaf(x, w.x, a[0]); a[1] = fmaf(x, w.y, a[1]); a[2] = fmaf(x, w.z, a[2]); a[3] = fmaf(x, w.w, a[3]);
#endif
}


struct XformDev {
  const float* scale;
  const float* shift;
  const float* center;
  float slope;
  bool affine;
  bool act;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum of a double; result valid in thread 0. `red` needs >= 32 doubles of shared memory.
__device__ __forceinline__ double block_sum_d(double v, double* red) {
  v = warp_sum_d(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) red[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  double r = 0.0;
  if (w == 0) {
    r = lane < nw ? red[lane] : 0.0;
    r = warp_sum_d(r);
  }
  __syncthreads();
  return r;
}

// ---- counter-based RNG: Philox4x32-10 ------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi,
           c3 = (uint32_t)(ctr_hi >> 32);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// keep-mask for element `idx` of dropout site (seed, offset): one Philox call serves 4 elements.
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t offset, uint64_t idx, float p) {
  const uint4 r = philox4x32(seed, idx >> 2, offset);
  const uint32_t w = (idx & 3) == 0 ? r.x : (idx & 3) == 1 ? r.y : (idx & 3) == 2 ? r.z : r.w;
  return (float)(w >> 8) * (1.0f / 16777216.0f) >= p;
}

}  // namespace cvae
