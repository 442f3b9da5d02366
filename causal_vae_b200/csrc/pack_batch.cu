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
  const size_t base = (size_t)((int)blockIdx.x - j.block0) * kPbPerBlock;
  if (j.tc) {      // [tap][k-block of 32][hi | lo][B rows][32 floats, 128B-swizzled by (row & 7)]  (conv_tc.cu)
    const int KB = (A_pad + 31) >> 5;
    const size_t total = (size_t)taps * KB * 2 * B * 32;
    for (int e = threadIdx.x; e < kPbPerBlock; e += 256) {
      const size_t i = base + e;
      if (i >= total) break;
      const int kk = (int)(i & 31);
      size_t r = i >> 5;
      const int n = (int)(r % B); r /= B;
      const int h = (int)(r & 1); r >>= 1;
      const int kb = (int)(r % KB);
      const int tap = (int)(r / KB);
      const int lc = (kk >> 2) ^ (n & 7);
      const int k = kb * 32 + lc * 4 + (kk & 3);
      float v = 0.f;
      if (k < A && (bat || n < src_ld))
        v = bat ? src[((size_t)n * src_ld + k) * taps + tap] : src[((size_t)k * src_ld + n) * taps + tap];
      float vh, vl;
      split_tf32(v, vh, vl);
      dst[i] = h ? vl : vh;
    }
  } else {         // [tap][A_pad][B] fp32  (conv.cu::pack_weight_kernel)
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
