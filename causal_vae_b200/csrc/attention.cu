// Multi-head self-attention core for short sequences (S <= 128 tokens: 65 at 256x256, 17 at 128x128).
// One CTA per (batch, head): Q, K, V and the SxS probability tile live in shared memory, so the
// scores / softmax / dropout / PV chain never touches HBM except for the saved probabilities.
#include "common.cuh"

namespace cvae {

__global__ void __launch_bounds__(128) attention_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                            float* __restrict__ probs, int S, int H, int d,
                                                            float p_drop, uint64_t seed, uint64_t offset,
                                                            const int64_t* __restrict__ counter) {
  extern __shared__ float sm[];
  if (counter) seed += (uint64_t)(*counter) * 0x9E3779B97F4A7C15ull;
  const int ld = d + 1, lp = S + 1;
  float* Q = sm; float* K = Q + S * ld; float* V = K + S * ld; float* P = V + S * ld;
  const int b = blockIdx.x / H, h = blockIdx.x % H, D = H * d, tid = threadIdx.x;
  const float* base = qkv + (size_t)b * S * 3 * D + h * d;
  for (int i = tid; i < S * d; i += blockDim.x) {
    const int s = i / d, c = i % d;
    const float* r = base + (size_t)s * 3 * D + c;
    Q[s * ld + c] = r[0]; K[s * ld + c] = r[D]; V[s * ld + c] = r[2 * D];
  }
  __syncthreads();
  const float scale = rsqrtf((float)d);
  for (int i = tid; i < S * S; i += blockDim.x) {
    const int qi = i / S, kj = i % S;
    float acc = 0.f;
    for (int c = 0; c < d; ++c) acc = fmaf(Q[qi * ld + c], K[kj * ld + c], acc);
    P[qi * lp + kj] = acc * scale;
  }
  __syncthreads();
  const int lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  float* pg = probs + (size_t)blockIdx.x * S * S;
  for (int r = w; r < S; r += nw) {
    float mx = -INFINITY;
    for (int j = lane; j < S; j += 32) mx = fmaxf(mx, P[r * lp + j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < S; j += 32) { const float e = expf(P[r * lp + j] - mx); P[r * lp + j] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int j = lane; j < S; j += 32) {
      float pv = P[r * lp + j] * inv;
      pg[r * S + j] = pv;
      if (p_drop > 0.f) {
        const uint64_t idx = ((uint64_t)blockIdx.x * S + r) * S + j;
        pv = dropout_keep(seed, offset, idx, p_drop) ? pv * keep_scale : 0.f;
      }
      P[r * lp + j] = pv;
    }
  }
  __syncthreads();
  float* ob = out + (size_t)b * S * D + h * d;
  for (int i = tid; i < S * d; i += blockDim.x) {
    const int s = i / d, c = i % d;
    float acc = 0.f;
    for (int j = 0; j < S; ++j) acc = fmaf(P[s * lp + j], V[j * ld + c], acc);
    ob[(size_t)s * D + c] = acc;
  }
}

__global__ void __launch_bounds__(256) attention_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ probs,
                                                            const float* __restrict__ dout, float* __restrict__ dqkv,
                                                            int S, int H, int d, float p_drop, uint64_t seed,
                                                            uint64_t offset, const int64_t* __restrict__ counter) {
  extern __shared__ float sm[];
  if (counter) seed += (uint64_t)(*counter) * 0x9E3779B97F4A7C15ull;
  const int ld = d + 1, lp = S + 1;
  float* Q = sm; float* K = Q + S * ld; float* V = K + S * ld; float* dO = V + S * ld;
  float* P = dO + S * ld; float* dS = P + S * lp; float* PD = dS + S * lp;
  const int b = blockIdx.x / H, h = blockIdx.x % H, D = H * d, tid = threadIdx.x;
  const float* base = qkv + (size_t)b * S * 3 * D + h * d;
  const float* dob = dout + (size_t)b * S * D + h * d;
  for (int i = tid; i < S * d; i += blockDim.x) {
    const int s = i / d, c = i % d;
    const float* r = base + (size_t)s * 3 * D + c;
    Q[s * ld + c] = r[0]; K[s * ld + c] = r[D]; V[s * ld + c] = r[2 * D];
    dO[s * ld + c] = dob[(size_t)s * D + c];
  }
  const float* pg = probs + (size_t)blockIdx.x * S * S;
  for (int i = tid; i < S * S; i += blockDim.x) P[(i / S) * lp + (i % S)] = pg[i];
  __syncthreads();
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  // dP_raw = dO V^T, then (one Philox draw per probability) dP = mask*dP_raw and PD = mask*P
  for (int i = tid; i < S * S; i += blockDim.x) {
    const int qi = i / S, kj = i % S;
    float acc = 0.f;
    for (int c = 0; c < d; ++c) acc = fmaf(dO[qi * ld + c], V[kj * ld + c], acc);
    float mk = 1.f;
    if (p_drop > 0.f) {
      const uint64_t idx = ((uint64_t)blockIdx.x * S + qi) * S + kj;
      mk = dropout_keep(seed, offset, idx, p_drop) ? keep_scale : 0.f;
    }
    dS[qi * lp + kj] = acc * mk;
    PD[qi * lp + kj] = P[qi * lp + kj] * mk;
  }
  __syncthreads();
  float* gb = dqkv + (size_t)b * S * 3 * D + h * d;
  // dV[j][c] = sum_i PD[i][j] * dO[i][c]
  for (int i = tid; i < S * d; i += blockDim.x) {
    const int j = i / d, c = i % d;
    float acc = 0.f;
    for (int q = 0; q < S; ++q) acc = fmaf(PD[q * lp + j], dO[q * ld + c], acc);
    gb[(size_t)j * 3 * D + 2 * D + c] = acc;
  }
  __syncthreads();
  // dS = P * (dP - rowsum(dP * P)) * scale
  const int lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
  const float scale = rsqrtf((float)d);
  for (int r = w; r < S; r += nw) {
    float s = 0.f;
    for (int j = lane; j < S; j += 32) s = fmaf(dS[r * lp + j], P[r * lp + j], s);
    s = warp_sum(s);
    for (int j = lane; j < S; j += 32) dS[r * lp + j] = P[r * lp + j] * (dS[r * lp + j] - s) * scale;
  }
  __syncthreads();
  for (int i = tid; i < S * d; i += blockDim.x) {
    const int s = i / d, c = i % d;
    float aq = 0.f, ak = 0.f;
    for (int j = 0; j < S; ++j) {
      aq = fmaf(dS[s * lp + j], K[j * ld + c], aq);
      ak = fmaf(dS[j * lp + s], Q[j * ld + c], ak);
    }
    gb[(size_t)s * 3 * D + c] = aq;
    gb[(size_t)s * 3 * D + D + c] = ak;
  }
}

}  // namespace cvae
using namespace cvae;

extern "C" int cvae_attention_fwd(const float* qkv, float* out, float* probs, int B, int S, int H, int d,
                                  float dropout_p, uint64_t seed, uint64_t offset, const int64_t* counter,
                                  cvae_stream_t s) {
  if (!qkv || !out || !probs || B <= 0 || S <= 0 || H <= 0 || d <= 0) return CVAE_ERR_BAD_ARG;
  if (S > 128 || d > 64) return CVAE_ERR_UNSUPPORTED_SHAPE;
  const size_t smem = (size_t)(3 * S * (d + 1) + S * (S + 1)) * sizeof(float);
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return CVAE_ERR_LAUNCH;
  }
  attention_fwd_kernel<<<B * H, 128, smem, as_stream(s)>>>(qkv, out, probs, S, H, d, dropout_p, seed, offset, counter);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_attention_bwd(const float* qkv, const float* probs, const float* dout, float* dqkv, int B,
                                  int S, int H, int d, float dropout_p, uint64_t seed, uint64_t offset,
                                  const int64_t* counter, cvae_stream_t s) {
  if (!qkv || !probs || !dout || !dqkv || B <= 0 || S <= 0) return CVAE_ERR_BAD_ARG;
  if (S > 128 || d > 64) return CVAE_ERR_UNSUPPORTED_SHAPE;
  const size_t smem = (size_t)(4 * S * (d + 1) + 3 * S * (S + 1)) * sizeof(float);
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return CVAE_ERR_LAUNCH;
  }
  attention_bwd_kernel<<<B * H, 256, smem, as_stream(s)>>>(qkv, probs, dout, dqkv, S, H, d, dropout_p, seed, offset, counter);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}
