// tcgen05.mma kind::tf32 issue / execution rate probe (sm_100a): one thread issues `iters` x 3 MMAs
// (M = 128, N, K = 8; A and B from shared memory, SWIZZLE_128B K-major) and the elapsed clocks until the
// final commit lands are reported per MMA.  Variants: same operands every time / A window shifted per MMA.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/umma_rate.bin scripts/umma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../causal_vae_b200/csrc/tc_common.cuh"
using namespace cvae::tc;

__global__ void rate(int N, int iters, int a_tmem_mode, long long* out) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  uint8_t* buf = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(buf);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (96 * 1024) / 4; i += blockDim.x) reinterpret_cast<float*>(buf)[i] = 1.0f;
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
  if (tid == 0) {
    const uint32_t idesc = make_idesc_tf32(128, N, 0, 0);
    const uint64_t da0 = make_smem_desc(sbase, 16, 1280, kLayoutSw128);
    const uint64_t db0 = make_smem_desc(sbase + 48 * 1024, 16, 1024, kLayoutSw128);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint64_t da = da0 + (uint64_t)((i & 7) * 8);     // shifted 128-byte windows
      if (a_tmem_mode) {
        mma_tf32_ts(tmem, tmem + 256, db0, idesc, 1u);
        mma_tf32_ts(tmem, tmem + 264, db0 + 2, idesc, 1u);
        mma_tf32_ts(tmem, tmem + 272, db0 + 4, idesc, 1u);
      } else {
        mma_tf32(tmem, da, db0, idesc, 1u);
        mma_tf32(tmem, da + 2, db0 + 2, idesc, 1u);
        mma_tf32(tmem, da + 4, db0 + 4, idesc, 1u);
      }
    }
    long long t1 = clock64();
    mma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  __syncthreads();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 2000;
  for (int mode = 0; mode < 2; ++mode)
    for (int N : {16, 32, 64, 128, 256}) {
      rate<<<1, 128, 98 * 1024>>>(N, iters, mode, d);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("%s N=%3d: issue %.1f clk/MMA, complete %.1f clk/MMA (floor 128*N/256 = %d)\n", mode ? "A=TMEM" : "A=smem", N,
             (double)h[0] / (3.0 * iters), (double)h[1] / (3.0 * iters), N / 2);
    }
  return 0;
}
