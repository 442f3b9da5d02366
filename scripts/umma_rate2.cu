// tcgen05.mma kind::tf32 rate probe, second series (sm_100a): why does conv_halo_tc see 120-200 clk per MMA when
// scripts/umma_rate.cu measures 96.6?  One thread issues MMAs (M = 128, K = 8, A and B in shared memory, SWIZZLE_128B
// K-major) in the patterns the kernel uses, optionally while other warps hammer shared memory with 128-bit stores
// (the halo producers) or read tensor memory (the epilogue):
//   pat 0: one instruction descriptor, N = 128, same D                         (umma_rate.cu's case)
//   pat 1: alternating N = 128 / N = 64 descriptors, D / D + 64                (gather plans with cross-term columns)
//   pat 2: three MMAs N = 128 into D, D + 128, D                               (scatter plans with cross-term columns)
//   pat 3: N = 64 only
//   pat 4: N = 128 then N = 128 with different D (no descriptor change);  pat 5: N = 256;  pat 6: N = 32
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/umma_rate2.bin scripts/umma_rate2.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../causal_vae_b200/csrc/tc_common.cuh"
using namespace cvae::tc;

__global__ void rate(int pat, int iters, int writers, int readers, int uniform, long long* out) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  __shared__ volatile int s_stop;
  uint8_t* buf = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(buf);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (96 * 1024) / 4; i += blockDim.x) reinterpret_cast<float*>(buf)[i] = 1.0f;
  if (tid == 0) s_stop = 0;
  if (warp == 0) {
    if (lane == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
    __syncwarp();
    tmem_alloc(smem_u32(&s_tmem), 512);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (warp == 0 && (uniform || lane == 0)) {
    // uniform = 1: the whole warp runs the loop and one elected lane issues (descriptors in uniform registers, MMAs back
    // to back in SASS); uniform = 0: `if (lane == 0)`, where the compiler wraps every MMA in an ELECT / BRA.U.ANY loop
    const bool leader = uniform ? elect_one() : true;
    const uint32_t id128 = make_idesc_tf32(128, 128, 0, 0), id64 = make_idesc_tf32(128, 64, 0, 0);
    const uint32_t id256 = make_idesc_tf32(128, 256, 0, 0), id32 = make_idesc_tf32(128, 32, 0, 0);
    const uint64_t da0 = make_smem_desc(sbase, 16, 1280, kLayoutSw128);
    const uint64_t db0 = make_smem_desc(sbase + 48 * 1024, 16, 1024, kLayoutSw128);
    long long t0 = clock64();
    int n = 0;
    for (int i = 0; i < iters; ++i) {
      const uint64_t da = da0 + (uint64_t)((i & 7) * 8);     // shifted 128-byte windows
      if (pat == 0) { if (leader) { mma_tf32(tmem, da, db0, id128, 1u); mma_tf32(tmem, da + 2, db0 + 2, id128, 1u); } n += 2; }
      else if (pat == 1) { if (leader) { mma_tf32(tmem, da, db0, id128, 1u); mma_tf32(tmem + 64, da + 2, db0, id64, 1u); } n += 2; }
      else if (pat == 2) { if (leader) { mma_tf32(tmem + 128, da + 2, db0, id128, 1u); mma_tf32(tmem + 128, da, db0 + 2, id128, 1u);
                           mma_tf32(tmem, da, db0, id128, 1u); } n += 3; }
      else if (pat == 3) { if (leader) { mma_tf32(tmem, da, db0, id64, 1u); mma_tf32(tmem, da + 2, db0 + 2, id64, 1u); } n += 2; }
      else if (pat == 4) { if (leader) { mma_tf32(tmem, da, db0, id128, 1u); mma_tf32(tmem + 128, da + 2, db0 + 2, id128, 1u); } n += 2; }
      else if (pat == 5) { if (leader) { mma_tf32(tmem, da, db0, id256, 1u); mma_tf32(tmem, da + 2, db0 + 2, id256, 1u); } n += 2; }
      else { if (leader) { mma_tf32(tmem, da, db0, id32, 1u); mma_tf32(tmem, da + 2, db0 + 2, id32, 1u); } n += 2; }
    }
    long long t1 = clock64();
    if (leader) mma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    if (lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; out[2] = n; s_stop = 1; }
  } else if (warp >= 1 && warp <= writers) {
    // shared-memory writers: two 128-bit stores per iteration into the upper part of the buffer (not read by the MMAs)
    float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    uint8_t* dst = buf + 64 * 1024 + (size_t)(warp - 1) * 4096;
    int k = 0;
    while (!s_stop) {
      *reinterpret_cast<float4*>(dst + ((lane * 16 + k * 512) & 4095)) = v;
      *reinterpret_cast<float4*>(dst + ((lane * 16 + k * 512 + 2048) & 4095)) = v;
      ++k;
    }
  } else if (warp > 8 && warp <= 8 + readers) {
    // tensor-memory readers (the epilogue's 16x256b loads) from columns the MMAs do not write
    uint32_t r[8];
    float s = 0.f;
    while (!s_stop) {
      tmem_ld_16x256b_x2(tmem + 256 + ((uint32_t)((warp & 3) * 32) << 16), r);
      tmem_ld_wait();
      s += __uint_as_float(r[0]);
    }
    if (s == 123.456f) out[3] = 1;
  }
  __syncthreads();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 32);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 2000;
  for (int uniform : {1, 0})
  for (int writers : {0, 8})
    for (int readers : {0, 4})
      for (int pat = 0; pat < 7; ++pat) {
        if (!uniform && (writers || readers)) continue;
        rate<<<1, 13 * 32, 98 * 1024>>>(pat, iters, writers, readers, uniform, d);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        long long h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
        printf("%s writers %d tmem-readers %d pat %d: issue %.1f clk/MMA, complete %.1f clk/MMA\n", uniform ? "converged-warp issue" : "lane-0 issue        ", writers, readers, pat,
               (double)h[0] / h[2], (double)h[1] / h[2]);
      }
  return 0;
}
