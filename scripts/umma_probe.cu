// Probe of tcgen05.mma shared-memory descriptor semantics on sm_100a (no public docs in this image):
// does a K-major SWIZZLE_128B operand tolerate (a) a start address that is a multiple of 128 B but
// not of 1024 B (a window shifted by whole pixel rows), (b) a stride-byte-offset (8-row group pitch)
// that is not a multiple of 1024 B, when the data was written with the ADDRESS-based swizzle
// (16-byte chunk index ^= address bits [7,10))?  And what does the descriptor's base_offset field do?
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/umma_probe.bin scripts/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../causal_vae_b200/csrc/tc_common.cuh"

using namespace cvae::tc;

__device__ __forceinline__ uint64_t desc_bo(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint64_t layout, uint32_t base_off) {
  return make_smem_desc(saddr, lbo, sbo, layout) | ((uint64_t)(base_off & 7u) << 49);
}

// A: window of 128 rows = 16 groups of 8 consecutive 128-byte slots; group g starts at slot
// base_slot + g * pitch.  slot s holds 32 floats, value(s, ch) = (s % 61) * 16 + (ch % 16) (tf32-exact).
// B: 16 rows x 32 k, B[n][k] = (k == n) -> D[r][n] = A[r][n], n < 16.
__global__ void probe(int base_slot, int pitch, int bo_mode, int nslots, float* out) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  uint8_t* buf = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(buf);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* bB = buf + (size_t)nslots * 128;      // nslots is a multiple of 8 -> 1024-aligned
  for (int i = tid; i < nslots * 32; i += blockDim.x) {
    const int s = i >> 5, ch = i & 31;
    const uint32_t addr = (uint32_t)s * 128u + (uint32_t)(((ch >> 2) ^ (s & 7)) << 4) + (ch & 3) * 4;
    *reinterpret_cast<float*>(buf + addr) = (float)((s % 61) * 16 + (ch % 16));
  }
  for (int i = tid; i < 16 * 32; i += blockDim.x) {
    const int n = i >> 5, k = i & 31;
    const uint32_t addr = (uint32_t)n * 128u + (uint32_t)(((k >> 2) ^ (n & 7)) << 4) + (k & 3) * 4;
    *reinterpret_cast<float*>(bB + addr) = (k == n) ? 1.0f : 0.0f;
  }
  if (warp == 0) {
    if (lane == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
    __syncwarp();
    tmem_alloc(smem_u32(&s_tmem), 32);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_tf32(128, 16, 0, 0);
    const uint32_t a0 = sbase + (uint32_t)base_slot * 128u;
    const uint32_t b0 = sbase + (uint32_t)nslots * 128u;
    const uint32_t bo = bo_mode ? ((a0 >> 7) & 7u) : 0u;
    for (int k = 0; k < 4; ++k) {
      const uint64_t da = desc_bo(a0 + k * 32, 16, (uint32_t)pitch * 128u, kLayoutSw128, bo);
      const uint64_t db = make_smem_desc(b0 + k * 32, 16, 1024, kLayoutSw128);
      mma_tf32(tmem, da, db, idesc, k > 0 ? 1u : 0u);
    }
    mma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  if (warp < 4) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
    for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 16 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

int main() {
  const int nslots = 400;   // 50 KB of A slots
  float* d_out;
  cudaMalloc(&d_out, 128 * 16 * 4);
  float h[128 * 16];
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int bases[] = {0, 1, 3, 7, 8, 9, 17, 26};
  const int pitches[] = {8, 10, 16, 18};
  for (int bo = 0; bo < 2; ++bo)
    for (int pitch : pitches)
      for (int base : bases) {
        cudaMemset(d_out, 0xff, sizeof(h));
        probe<<<1, 128, nslots * 128 + 16 * 128 + 2048>>>(base, pitch, bo, nslots, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("base %d pitch %d bo %d: CUDA error %s\n", base, pitch, bo, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
        int bad = 0, first = -1;
        for (int r = 0; r < 128; ++r) {
          const int s = base + (r / 8) * pitch + (r % 8);
          for (int n = 0; n < 16; ++n) {
            const float want = (float)((s % 61) * 16 + n);
            if (h[r * 16 + n] != want) { if (first < 0) first = r * 16 + n; ++bad; }
          }
        }
        printf("bo_mode %d pitch %2d base %2d : %s (%d bad", bo, pitch, base, bad ? "MISMATCH" : "ok", bad);
        if (bad) {
          const int r = first / 16, n = first % 16;
          const float got = h[first];
          printf("; first r=%d n=%d got %.0f = slot%%61 %d ch %d, want slot %d", r, n, got, (int)got / 16, (int)got % 16,
                 base + (r / 8) * pitch + (r % 8));
        }
        printf(")\n");
      }
  return 0;
}
