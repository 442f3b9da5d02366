// Multi-head self-attention core for short sequences (S <= 128 tokens: 65 at 256x256, 17 at 128x128).
// One CTA per (batch, head): Q, K, V and the SxS probability tile live in shared memory, so the
// scores / softmax / dropout / PV chain never touches HBM except for the saved probabilities.
//
// Every contraction is register-blocked (5x5 score tiles, 5x2 output tiles): a scalar inner product
// per thread needs two shared-memory loads per FMA and made the first version of this kernel
// shared-memory-issue bound (81 us forward / 126 us backward per ViT block at B = 64); the tiles
// bring it to 0.4 / 0.7 loads per FMA.
//
// The dropout mask is drawn once, in forward, and travels to backward in the SIGN BIT of the saved
// probability (softmax outputs are >= 0): -p means "p was dropped".  Backward therefore needs no
// generator calls and cannot disagree with forward about a mask.
#include "common.cuh"

namespace cvae {

constexpr int kAT = 5;   // tile edge (65 = 13 * 5 tokens at 256x256)

__device__ __forceinline__ int att_lp(int Sp) { return (Sp & 1) ? Sp : Sp + 1; }

// rows [0,S) of NT heads' [S, d] slices (row stride `stride[t]` floats) -> smem [Sp][ld] each, rows >= S zeroed.
// All loads of a batch (2 vectors per thread and tensor) are issued before the first shared-memory store: the
// one-tensor-at-a-time load -> store loop paid 3 dependent L2 round trips per tensor, 9 per forward launch.
template <int NT>
__device__ __forceinline__ void att_load_multi(float* const (&dst)[NT], const float* const (&src)[NT], const size_t (&stride)[NT],
                                               int S, int Sp, int d, int ld) {
  const int d4 = d >> 2, n = Sp * d4;
  for (int i0 = threadIdx.x; i0 < n; i0 += 2 * blockDim.x) {
    float4 v[2][NT];
    int so[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = i0 + u * blockDim.x;
      so[u] = -1;
      const int s = i / d4, c = (i - s * d4) << 2;
      if (i < n) so[u] = s * ld + c;
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        v[u][t] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n && s < S) v[u][t] = __ldg(reinterpret_cast<const float4*>(src[t] + (size_t)s * stride[t] + c));
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (so[u] < 0) continue;
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        float* o = dst[t] + so[u];
        o[0] = v[u][t].x; o[1] = v[u][t].y; o[2] = v[u][t].z; o[3] = v[u][t].w;
      }
    }
  }
}

// acc[i][j] = sum_c A[(5 bi + i)][c] * B[(5 bj + j)][c]
__device__ __forceinline__ void att_tile_nt(const float* __restrict__ A, const float* __restrict__ B, int ld, int d,
                                            float (&acc)[kAT][kAT]) {
#pragma unroll
  for (int i = 0; i < kAT; ++i)
#pragma unroll
    for (int j = 0; j < kAT; ++j) acc[i][j] = 0.f;
#pragma unroll 4
  for (int c = 0; c < d; ++c) {
    float a[kAT], b[kAT];
#pragma unroll
    for (int i = 0; i < kAT; ++i) { a[i] = A[i * ld + c]; b[i] = B[i * ld + c]; }
#pragma unroll
    for (int i = 0; i < kAT; ++i)
#pragma unroll
      for (int j = 0; j < kAT; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

__global__ void __launch_bounds__(256, 4) attention_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                            float* __restrict__ probs, int S, int H, int d,
                                                            float p_drop, uint64_t seed, uint64_t offset,
                                                            const int64_t* __restrict__ counter) {
  extern __shared__ float sm[];
  if (counter) seed += (uint64_t)(*counter) * 0x9E3779B97F4A7C15ull;
  const int nb = (S + kAT - 1) / kAT, Sp = nb * kAT, ld = d + 1, lp = att_lp(Sp);
  float* Q = sm; float* K = Q + Sp * ld; float* V = K + Sp * ld; float* P = V + Sp * ld;
  const int b = blockIdx.x / H, h = blockIdx.x % H, D = H * d, tid = threadIdx.x;
  const float* base = qkv + (size_t)b * S * 3 * D + h * d;
  {
    float* const dsts[3] = {Q, K, V};
    const float* const srcs[3] = {base, base + D, base + 2 * D};
    const size_t strides[3] = {(size_t)3 * D, (size_t)3 * D, (size_t)3 * D};
    att_load_multi<3>(dsts, srcs, strides, S, Sp, d, ld);
  }
  __syncthreads();
  const float scale = rsqrtf((float)d);
  for (int t = tid; t < nb * nb; t += blockDim.x) {
    const int bi = t / nb, bj = t - bi * nb;
    float acc[kAT][kAT];
    att_tile_nt(Q + bi * kAT * ld, K + bj * kAT * ld, ld, d, acc);
#pragma unroll
    for (int i = 0; i < kAT; ++i)
#pragma unroll
      for (int j = 0; j < kAT; ++j) P[(bi * kAT + i) * lp + bj * kAT + j] = acc[i][j] * scale;
  }
  __syncthreads();
  const int lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  float* pg = probs + (size_t)blockIdx.x * S * S;
  // A lane owns keys 4*lane .. 4*lane + 3 of the row (S <= 128): one Philox4x32 call yields the four
  // keep decisions (the per-element form spent ~100 instructions of generator per probability: 25 % of
  // the kernel's instructions in profiles/).
  for (int r = w; r < S; r += nw) {
    float e[4];
    float mx = -INFINITY;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = 4 * lane + u;
      e[u] = j < S ? P[r * lp + j] : -INFINITY;
      mx = fmaxf(mx, e[u]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) { e[u] = (4 * lane + u < S) ? expf(e[u] - mx) : 0.f; sum += e[u]; }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    uint4 rnd = make_uint4(~0u, ~0u, ~0u, ~0u);
    if (p_drop > 0.f) rnd = philox4x32(seed, ((uint64_t)blockIdx.x * S + r) * 32 + lane, offset);
    const uint32_t rw[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = 4 * lane + u;
      if (j < S) {
        const float pv = e[u] * inv;
        const bool keep = p_drop > 0.f ? (float)(rw[u] >> 8) * (1.0f / 16777216.0f) >= p_drop : true;
        pg[r * S + j] = keep ? pv : -pv;          // sign bit = dropped
        P[r * lp + j] = keep ? pv * keep_scale : 0.f;
      }
    }
  }
  __syncthreads();
  // out[s][c] = sum_j P[s][j] V[j][c]: 5 rows x 2 columns per thread
  float* ob = out + (size_t)b * S * D + h * d;
  const int d2 = d >> 1;
  for (int t = tid; t < nb * d2; t += blockDim.x) {
    const int bs = t / d2, c = (t - bs * d2) << 1;
    float acc[kAT][2];
#pragma unroll
    for (int i = 0; i < kAT; ++i) acc[i][0] = acc[i][1] = 0.f;
    const float* pr = P + bs * kAT * lp;
#pragma unroll 4
    for (int j = 0; j < S; ++j) {
      const float v0 = V[j * ld + c], v1 = V[j * ld + c + 1];
#pragma unroll
      for (int i = 0; i < kAT; ++i) {
        const float pv = pr[i * lp + j];
        acc[i][0] = fmaf(pv, v0, acc[i][0]);
        acc[i][1] = fmaf(pv, v1, acc[i][1]);
      }
    }
#pragma unroll
    for (int i = 0; i < kAT; ++i) {
      const int s = bs * kAT + i;
      if (s < S) *reinterpret_cast<float2*>(ob + (size_t)s * D + c) = make_float2(acc[i][0], acc[i][1]);
    }
  }
}

__global__ void __launch_bounds__(256, 4) attention_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ probs,
                                                            const float* __restrict__ dout, float* __restrict__ dqkv,
                                                            int S, int H, int d, float p_drop) {
  extern __shared__ float sm[];
  const int nb = (S + kAT - 1) / kAT, Sp = nb * kAT, ld = d + 1, lp = att_lp(Sp);
  float* A = sm;                 // dO, later Q
  float* Bm = A + Sp * ld;       // V, later K
  float* Ps = Bm + Sp * ld;      // signed saved probabilities (sign bit = dropped)
  float* dS = Ps + Sp * lp;
  const int b = blockIdx.x / H, h = blockIdx.x % H, D = H * d, tid = threadIdx.x;
  const float* base = qkv + (size_t)b * S * 3 * D + h * d;
  {
    float* const dsts[2] = {A, Bm};
    const float* const srcs[2] = {dout + (size_t)b * S * D + h * d, base + 2 * D};
    const size_t strides[2] = {(size_t)D, (size_t)3 * D};
    att_load_multi<2>(dsts, srcs, strides, S, Sp, d, ld);
  }
  const float* pg = probs + (size_t)blockIdx.x * S * S;
  for (int i = tid; i < Sp * lp; i += blockDim.x) Ps[i] = 0.f;
  __syncthreads();
  for (int i0 = tid; i0 < S * S; i0 += 6 * blockDim.x) {       // batches of 6 loads per thread before the first store
    float pv[6];
#pragma unroll
    for (int u = 0; u < 6; ++u) { const int i = i0 + u * blockDim.x; pv[u] = i < S * S ? __ldg(pg + i) : 0.f; }
#pragma unroll
    for (int u = 0; u < 6; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < S * S) { const int r = i / S; Ps[r * lp + (i - r * S)] = pv[u]; }
    }
  }
  __syncthreads();
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  // dP = mask * (dO V^T)
  for (int t = tid; t < nb * nb; t += blockDim.x) {
    const int bi = t / nb, bj = t - bi * nb;
    float acc[kAT][kAT];
    att_tile_nt(A + bi * kAT * ld, Bm + bj * kAT * ld, ld, d, acc);
#pragma unroll
    for (int i = 0; i < kAT; ++i)
#pragma unroll
      for (int j = 0; j < kAT; ++j) {
        const int o = (bi * kAT + i) * lp + bj * kAT + j;
        dS[o] = Ps[o] < 0.f ? 0.f : acc[i][j] * keep_scale;
      }
  }
  __syncthreads();
  float* gb = dqkv + (size_t)b * S * 3 * D + h * d;
  const int d2 = d >> 1;
  // dV[j][c] = sum_q PD[q][j] dO[q][c], PD = mask * P: 5 keys x 2 columns per thread
  for (int t = tid; t < nb * d2; t += blockDim.x) {
    const int bj = t / d2, c = (t - bj * d2) << 1;
    float acc[kAT][2];
#pragma unroll
    for (int i = 0; i < kAT; ++i) acc[i][0] = acc[i][1] = 0.f;
#pragma unroll 4
    for (int q = 0; q < S; ++q) {
      const float g0 = A[q * ld + c], g1 = A[q * ld + c + 1];
#pragma unroll
      for (int i = 0; i < kAT; ++i) {
        const float pd = fmaxf(Ps[q * lp + bj * kAT + i], 0.f);
        acc[i][0] = fmaf(pd, g0, acc[i][0]);
        acc[i][1] = fmaf(pd, g1, acc[i][1]);
      }
    }
#pragma unroll
    for (int i = 0; i < kAT; ++i) {
      const int j = bj * kAT + i;
      if (j < S)
        *reinterpret_cast<float2*>(gb + (size_t)j * 3 * D + 2 * D + c) = make_float2(acc[i][0] * keep_scale, acc[i][1] * keep_scale);
    }
  }
  // Q and K replace dO and V in shared memory after the next barrier: their loads are issued now (three vectors per
  // thread and tensor cover S <= 96 at 256 threads) so that the round trip overlaps the softmax-gradient pass
  const int d4 = d >> 2, nvec = Sp * d4;
  float4 pq[3], pk[3];
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    const int i = tid + u * blockDim.x;
    pq[u] = make_float4(0.f, 0.f, 0.f, 0.f); pk[u] = pq[u];
    if (i < nvec) {
      const int s_ = i / d4, c_ = (i - s_ * d4) << 2;
      if (s_ < S) {
        pq[u] = __ldg(reinterpret_cast<const float4*>(base + (size_t)s_ * 3 * D + c_));
        pk[u] = __ldg(reinterpret_cast<const float4*>(base + D + (size_t)s_ * 3 * D + c_));
      }
    }
  }
  // dS = P * (dP - rowsum(dP * P)) * scale   (rows only touch Ps / dS, which the dV pass does not write)
  const int lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
  const float scale = rsqrtf((float)d);
  for (int r = w; r < S; r += nw) {
    float s = 0.f;
    for (int j = lane; j < S; j += 32) s = fmaf(dS[r * lp + j], fabsf(Ps[r * lp + j]), s);
    s = warp_sum(s);
    for (int j = lane; j < S; j += 32) dS[r * lp + j] = fabsf(Ps[r * lp + j]) * (dS[r * lp + j] - s) * scale;
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    const int i = tid + u * blockDim.x;
    if (i < nvec) {
      const int s_ = i / d4, c_ = (i - s_ * d4) << 2;
      float* oq = A + s_ * ld + c_;
      float* ok = Bm + s_ * ld + c_;
      oq[0] = pq[u].x; oq[1] = pq[u].y; oq[2] = pq[u].z; oq[3] = pq[u].w;
      ok[0] = pk[u].x; ok[1] = pk[u].y; ok[2] = pk[u].z; ok[3] = pk[u].w;
    }
  }
  for (int i = tid + 3 * blockDim.x; i < nvec; i += blockDim.x) {      // S > 96 at 256 threads
    const int s_ = i / d4, c_ = (i - s_ * d4) << 2;
    float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f), k4 = q4;
    if (s_ < S) {
      q4 = __ldg(reinterpret_cast<const float4*>(base + (size_t)s_ * 3 * D + c_));
      k4 = __ldg(reinterpret_cast<const float4*>(base + D + (size_t)s_ * 3 * D + c_));
    }
    float* oq = A + s_ * ld + c_;
    float* ok = Bm + s_ * ld + c_;
    oq[0] = q4.x; oq[1] = q4.y; oq[2] = q4.z; oq[3] = q4.w;
    ok[0] = k4.x; ok[1] = k4.y; ok[2] = k4.z; ok[3] = k4.w;
  }
  __syncthreads();
  // dQ[s][c] = sum_j dS[s][j] K[j][c];  dK[s][c] = sum_j dS[j][s] Q[j][c]
  for (int t = tid; t < nb * d2; t += blockDim.x) {
    const int bs = t / d2, c = (t - bs * d2) << 1;
    float aq[kAT][2], ak[kAT][2];
#pragma unroll
    for (int i = 0; i < kAT; ++i) aq[i][0] = aq[i][1] = ak[i][0] = ak[i][1] = 0.f;
#pragma unroll 2
    for (int j = 0; j < S; ++j) {
      const float k0 = Bm[j * ld + c], k1 = Bm[j * ld + c + 1];
      const float q0 = A[j * ld + c], q1 = A[j * ld + c + 1];
#pragma unroll
      for (int i = 0; i < kAT; ++i) {
        const float a = dS[(bs * kAT + i) * lp + j], bt = dS[j * lp + bs * kAT + i];
        aq[i][0] = fmaf(a, k0, aq[i][0]); aq[i][1] = fmaf(a, k1, aq[i][1]);
        ak[i][0] = fmaf(bt, q0, ak[i][0]); ak[i][1] = fmaf(bt, q1, ak[i][1]);
      }
    }
#pragma unroll
    for (int i = 0; i < kAT; ++i) {
      const int s = bs * kAT + i;
      if (s < S) {
        *reinterpret_cast<float2*>(gb + (size_t)s * 3 * D + c) = make_float2(aq[i][0], aq[i][1]);
        *reinterpret_cast<float2*>(gb + (size_t)s * 3 * D + D + c) = make_float2(ak[i][0], ak[i][1]);
      }
    }
  }
}

static inline int att_threads(int S) { return S <= 20 ? 64 : S <= 40 ? 128 : 256; }

// attention_long.cu: strips of 32 rows with the other operand streamed through shared memory (S > 128)
int attention_long_fwd(const float* qkv, float* out, float* probs, int B, int S, int H, int d, float dropout_p, uint64_t seed,
                       uint64_t offset, const int64_t* counter, cudaStream_t st);
int attention_long_bwd(const float* qkv, const float* probs, const float* dout, float* dqkv, float* ws, int B, int S, int H, int d,
                       float dropout_p, cudaStream_t st);

}  // namespace cvae
using namespace cvae;

extern "C" int cvae_attention_fwd(const float* qkv, float* out, float* probs, int B, int S, int H, int d,
                                  float dropout_p, uint64_t seed, uint64_t offset, const int64_t* counter,
                                  cvae_stream_t s) {
  if (!qkv || !out || !probs || B <= 0 || S <= 0 || H <= 0 || d <= 0) return CVAE_ERR_BAD_ARG;
  if (S > 128) return attention_long_fwd(qkv, out, probs, B, S, H, d, dropout_p, seed, offset, counter, as_stream(s));
  if (d > 64 || (d & 3)) return CVAE_ERR_UNSUPPORTED_SHAPE;
  const int Sp = (S + kAT - 1) / kAT * kAT, lp = (Sp & 1) ? Sp : Sp + 1;
  const size_t smem = (size_t)(3 * Sp * (d + 1) + Sp * lp) * sizeof(float);
  static size_t attr = 48 * 1024;
  if (smem > attr) {
    if (cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return CVAE_ERR_LAUNCH;
    attr = smem;
  }
  attention_fwd_kernel<<<B * H, att_threads(S), smem, as_stream(s)>>>(qkv, out, probs, S, H, d, dropout_p, seed, offset, counter);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

// Workspace of cvae_attention_bwd_ws: 0 for S <= 128 (everything lives in shared memory), B*H*S floats above
// (delta_i = sum_j dP_ij P_ij per query row, written by the dQ kernel and read by the dK / dV kernel).
extern "C" int64_t cvae_attention_ws_bytes(int B, int S, int H, int d) {
  (void)d;
  return S > 128 ? (int64_t)B * H * S * (int64_t)sizeof(float) : 0;
}

extern "C" int cvae_attention_bwd_ws(const float* qkv, const float* probs, const float* dout, float* dqkv, float* ws,
                                     int64_t ws_bytes, int B, int S, int H, int d, float dropout_p, cvae_stream_t s) {
  if (!qkv || !probs || !dout || !dqkv || B <= 0 || S <= 0 || H <= 0 || d <= 0) return CVAE_ERR_BAD_ARG;
  if (S <= 128) return cvae_attention_bwd(qkv, probs, dout, dqkv, B, S, H, d, dropout_p, 0, 0, nullptr, s);
  if (!ws || ws_bytes < cvae_attention_ws_bytes(B, S, H, d)) return CVAE_ERR_BAD_ARG;
  return attention_long_bwd(qkv, probs, dout, dqkv, ws, B, S, H, d, dropout_p, as_stream(s));
}

// `seed`, `offset`, `counter` are accepted for ABI stability and ignored: the mask is read from the
// sign bits of `probs` (written by cvae_attention_fwd).  S <= 128 only: longer sequences need the workspace
// of cvae_attention_bwd_ws.
extern "C" int cvae_attention_bwd(const float* qkv, const float* probs, const float* dout, float* dqkv, int B,
                                  int S, int H, int d, float dropout_p, uint64_t seed, uint64_t offset,
                                  const int64_t* counter, cvae_stream_t s) {
  (void)seed; (void)offset; (void)counter;
  if (!qkv || !probs || !dout || !dqkv || B <= 0 || S <= 0) return CVAE_ERR_BAD_ARG;
  if (S > 128 || d > 64 || (d & 3)) return CVAE_ERR_UNSUPPORTED_SHAPE;
  const int Sp = (S + kAT - 1) / kAT * kAT, lp = (Sp & 1) ? Sp : Sp + 1;
  const size_t smem = (size_t)(2 * Sp * (d + 1) + 2 * Sp * lp) * sizeof(float);
  static size_t attr = 48 * 1024;
  if (smem > attr) {
    if (cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return CVAE_ERR_LAUNCH;
    attr = smem;
  }
  attention_bwd_kernel<<<B * H, att_threads(S), smem, as_stream(s)>>>(qkv, probs, dout, dqkv, S, H, d, dropout_p);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}
