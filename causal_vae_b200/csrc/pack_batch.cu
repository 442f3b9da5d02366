// All weight packings of a training step in ONE launch.
//
// Every conv / linear layer needs its weights re-laid-out twice per step (forward operand and
// input-gradient operand; fp32 [tap][Cin][Cout] for the SIMT kernels, the tf32-split swizzled tile image
// for the tcgen05 kernels).  As per-use launches that was 90 kernels and 0.41 ms of a 10 ms step, almost
// all of it launch latency on tensors of a few KB.  The weights only change in the optimizer, so a trainer
// records the (source, destination, layout) jobs of its first step and replays them as one grid before
// each later forward (causal_vae_b200/ops.py::PackPlan).
#include "common.cuh"
#include "tc_common.cuh"

namespace cvae {

using namespace tc;

constexpr int kPbPerBlock = 2048;      // elements of the destination per block

__global__ void __launch_bounds__(256) pack_batch_kernel(const cvae_pack_job_t* __restrict__ jobs, const int njobs) {
  __shared__ float s_tile[64][33];
  // binary search: last job whose first block is <= blockIdx.x
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].block0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const cvae_pack_job_t j = jobs[lo];
  const int A = j.A, A_pad = j.A_pad, B = j.B, taps = j.taps, src_ld = j.src_ld;
  const bool bat = j.src_bat != 0;
  const float* __restrict__ src = j.src;
  float* __restrict__ dst = j.dst;
  const int blk = (int)blockIdx.x - j.block0;
  if (j.tc) {      // [tap][k-block of 32][hi | lo][B rows][32 floats, 128B-swizzled by (row & 7)]  (conv_tc.cu)
    // One thread = one 16-byte chunk of a row: four source values, split once, stored as one 128-bit vector into the
    // hi plane and one into the lo plane (the scalar form read every value twice and spent three runtime divisions per
    // float: 157 us per step for 14 M weights).  Threads run along the source's contiguous dimension.
    const int KB = (A_pad + 31) >> 5;
    const long long units = (long long)taps * KB * B * 8;
    const long long u = (long long)blk * (kPbPerBlock / 8) + threadIdx.x;
    if (u >= units) return;
    const int per = B * 8;
    const int g = (int)(u / per), idx = (int)(u - (long long)g * per);
    const int kb = g % KB, tap = g / KB;
    int n, c;
    if (bat) { c = idx & 7; n = idx >> 3; } else { n = idx % B; c = idx / B; }
    const int k0 = kb * 32 + ((c ^ (n & 7)) << 2);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bat || n < src_ld) {
      if (bat) {
        const float* p = src + ((size_t)n * src_ld + k0) * taps + tap;
        if (taps == 1 && k0 + 3 < A && (((size_t)p & 15) == 0)) {
          v = __ldg(reinterpret_cast<const float4*>(p));
        } else {
          if (k0 < A) v.x = __ldg(p);
          if (k0 + 1 < A) v.y = __ldg(p + taps);
          if (k0 + 2 < A) v.z = __ldg(p + 2 * taps);
          if (k0 + 3 < A) v.w = __ldg(p + 3 * taps);
        }
      } else {
        const size_t st = (size_t)src_ld * taps;
        const float* p = src + ((size_t)k0 * src_ld + n) * taps + tap;
        if (k0 < A) v.x = __ldg(p);
        if (k0 + 1 < A) v.y = __ldg(p + st);
        if (k0 + 2 < A) v.z = __ldg(p + 2 * st);
        if (k0 + 3 < A) v.w = __ldg(p + 3 * st);
      }
    }
    float4 vh, vl;
    split4(v, vh, vl);
    float* o = dst + ((size_t)(tap * KB + kb) * 2 * B + n) * 32 + (c << 2);
    *reinterpret_cast<float4*>(o) = vh;
    *reinterpret_cast<float4*>(o + (size_t)B * 32) = vl;
  } else if (taps == 1 && bat && (A_pad & 31) == 0 && (B & 63) == 0) {
    // fp32 [A_pad][B] from a Linear weight [B][src_ld]: a plain transposition (decoder_input: 16384 x 512 = 8.4 M
    // values).  A block owns a 32 (a) x 64 (b) tile: 128-byte row reads, shared-memory transposition, 256-byte row writes
    // (the per-element form issued one 32-byte sector read per value).
    const int tiles_b = B >> 6;
    const int a0 = (blk / tiles_b) << 5, b0 = (blk % tiles_b) << 6;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int bl = ty + 8 * i, a_ = a0 + tx;
      s_tile[bl][tx] = a_ < A ? __ldg(src + (size_t)(b0 + bl) * src_ld + a_) : 0.f;
    }
    __syncthreads();
    const int bx = threadIdx.x & 63, ay = threadIdx.x >> 6;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int al = ay + 4 * i;
      dst[(size_t)(a0 + al) * B + b0 + bx] = s_tile[bx][al];
    }
  } else {         // [tap][A_pad][B] fp32  (conv.cu::pack_weight_kernel)
    const size_t base = (size_t)blk * kPbPerBlock;
    const size_t total = (size_t)taps * A_pad * B;
    for (int e = threadIdx.x; e < kPbPerBlock; e += 256) {
      const size_t i = base + e;
      if (i >= total) break;
      const int b = (int)(i % B);
      const size_t r = i / B;
      const int a_ = (int)(r % A_pad), t = (int)(r / A_pad);
      float v = 0.f;
      if (a_ < A && (bat || b < src_ld))
        v = bat ? src[((size_t)b * src_ld + a_) * taps + t] : src[((size_t)a_ * src_ld + b) * taps + t];
      dst[i] = v;
    }
  }
}

}  // namespace cvae
using namespace cvae;

extern "C" int cvae_pack_batch_blocks(int A_pad, int B, int taps, int tc) {
  const size_t total = tc ? (size_t)taps * ((A_pad + 31) / 32) * 2 * B * 32 : (size_t)taps * A_pad * B;
  return (int)((total + kPbPerBlock - 1) / kPbPerBlock);
}

extern "C" int cvae_pack_batch(const cvae_pack_job_t* jobs_dev, int njobs, int nblocks, cvae_stream_t s) {
  if (!jobs_dev || njobs < 1 || nblocks < 1) return CVAE_ERR_BAD_ARG;
  pack_batch_kernel<<<nblocks, 256, 0, as_stream(s)>>>(jobs_dev, njobs);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}
