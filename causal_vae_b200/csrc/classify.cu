// Treatment-label kernels of the MNIST adversarial step and the cascade model: argmax over one-hot
// rows, one-hot expansion of class indices (integer work: bit-exact), softmax cross-entropy against
// class indices and KL(Uniform || softmax) with their logit gradients.  One warp per row; the row
// width T is the number of treatments (10 / 19), so everything lives in registers.
#include "common.cuh"

namespace cvae {

// torch.argmax(t, dim=1): index of the FIRST maximum (ties resolve to the lowest index).
__global__ void argmax_rows_kernel(const float* __restrict__ t, int64_t rows, int T, int64_t* __restrict__ out) {
  const int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int c = lane; c < T; c += 32) {
    const float v = t[r * T + c];
    if (v > best || (v == best && c < bi)) { best = v; bi = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > best || (ov == best && oi < bi))) { best = ov; bi = oi; }
  }
  if (lane == 0) out[r] = bi == 0x7fffffff ? 0 : bi;
}

// F.one_hot(idx, T).float()
__global__ void one_hot_kernel(const int64_t* __restrict__ idx, int64_t rows, int T, float* __restrict__ out) {
  const int64_t n = rows * T;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (idx[i / T] == (int64_t)(i % T)) ? 1.0f : 0.0f;
}

// log-softmax pieces of one row held by a warp: returns (max, log-sum-exp)
__device__ __forceinline__ void row_lse(const float* row, int T, int lane, float& mx, float& lse) {
  mx = -INFINITY;
  for (int c = lane; c < T; c += 32) mx = fmaxf(mx, row[c]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int c = lane; c < T; c += 32) s += expf(row[c] - mx);
  s = warp_sum(s);
  lse = logf(s);
}

// mode 0: sum_r -log_softmax(l_r)[target_r]        (F.cross_entropy, caller scales by 1/rows)
// mode 1: sum_r sum_c u (log u - log_softmax(l_r)_c), u = 1/T   (F.kl_div(logp, U, 'batchmean') * rows)
__global__ void softmax_loss_fwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target,
                                        int64_t rows, int T, int mode, double* sum) {
  __shared__ double red[32];
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  double acc = 0.0;
  for (int64_t r = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * wpb) {
    const float* row = logits + r * T;
    float mx, lse;
    row_lse(row, T, lane, mx, lse);
    if (mode == 0) {
      if (lane == 0) acc += (double)(-(row[target[r]] - mx - lse));
    } else {
      const float u = 1.0f / (float)T, lu = logf(u);
      float s = 0.f;
      for (int c = lane; c < T; c += 32) s += u * (lu - (row[c] - mx - lse));
      acc += (double)s;
    }
  }
  const double t = block_sum_d(acc, red);
  if (threadIdx.x == 0) atomicAdd(sum, t);
}

// dlogits = (softmax - onehot(target)) * g*gmul  (mode 0)  |  (softmax - 1/T) * g*gmul  (mode 1)
__global__ void softmax_loss_bwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target,
                                        int64_t rows, int T, int mode, const float* g, float gmul,
                                        float* __restrict__ dlogits) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const float gs = (g ? *g : 1.f) * gmul;
  for (int64_t r = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * wpb) {
    const float* row = logits + r * T;
    float mx, lse;
    row_lse(row, T, lane, mx, lse);
    const int64_t tg = mode == 0 ? target[r] : -1;
    const float u = 1.0f / (float)T;
    for (int c = lane; c < T; c += 32) {
      const float p = expf(row[c] - mx - lse);
      dlogits[r * T + c] = (p - (mode == 0 ? (c == tg ? 1.f : 0.f) : u)) * gs;
    }
  }
}

}  // namespace cvae

using namespace cvae;

static inline int row_blocks(int64_t rows) {
  int64_t b = (rows + 7) / 8;
  if (b > kNumSMs * 8) b = kNumSMs * 8;
  return (int)(b < 1 ? 1 : b);
}

extern "C" int cvae_argmax_rows(const float* t, int64_t rows, int T, int64_t* out, cvae_stream_t s) {
  if (!t || !out || rows < 1 || T < 1) return CVAE_ERR_BAD_ARG;
  argmax_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, as_stream(s)>>>(t, rows, T, out);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_one_hot(const int64_t* idx, int64_t rows, int T, float* out, cvae_stream_t s) {
  if (!idx || !out || rows < 1 || T < 1) return CVAE_ERR_BAD_ARG;
  const int64_t n = rows * T;
  int64_t b = (n + 255) / 256;
  if (b > kNumSMs * 8) b = kNumSMs * 8;
  one_hot_kernel<<<(unsigned)b, 256, 0, as_stream(s)>>>(idx, rows, T, out);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_softmax_ce_fwd(const float* logits, const int64_t* target, int64_t rows, int T, double* sum,
                                   cvae_stream_t s) {
  if (!logits || !target || !sum || rows < 1 || T < 1) return CVAE_ERR_BAD_ARG;
  softmax_loss_fwd_kernel<<<row_blocks(rows), 256, 0, as_stream(s)>>>(logits, target, rows, T, 0, sum);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_softmax_ce_bwd(const float* logits, const int64_t* target, int64_t rows, int T, const float* g,
                                   float gmul, float* dlogits, cvae_stream_t s) {
  if (!logits || !target || !dlogits || rows < 1 || T < 1) return CVAE_ERR_BAD_ARG;
  softmax_loss_bwd_kernel<<<row_blocks(rows), 256, 0, as_stream(s)>>>(logits, target, rows, T, 0, g, gmul, dlogits);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_uniform_kl_fwd(const float* logits, int64_t rows, int T, double* sum, cvae_stream_t s) {
  if (!logits || !sum || rows < 1 || T < 1) return CVAE_ERR_BAD_ARG;
  softmax_loss_fwd_kernel<<<row_blocks(rows), 256, 0, as_stream(s)>>>(logits, nullptr, rows, T, 1, sum);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_uniform_kl_bwd(const float* logits, int64_t rows, int T, const float* g, float gmul,
                                   float* dlogits, cvae_stream_t s) {
  if (!logits || !dlogits || rows < 1 || T < 1) return CVAE_ERR_BAD_ARG;
  softmax_loss_bwd_kernel<<<row_blocks(rows), 256, 0, as_stream(s)>>>(logits, nullptr, rows, T, 1, g, gmul, dlogits);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}
