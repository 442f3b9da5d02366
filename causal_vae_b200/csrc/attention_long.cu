// Multi-head self-attention core for LONG sequences (S > 128 tokens: 961 at the reference's default
// 768x1280 vessel image, vessel_analysis/00_core/config.py:10-11 with vit_backbone.py:62-66).
//
// attention.cu keeps a whole (batch, head) problem in shared memory, which stops at S = 128.  Here a CTA owns a
// strip of 32 query rows (forward, dQ) or 32 key rows (dK / dV) and streams the other operand through shared
// memory in tiles of 64 rows; the 32 x S strip of scores lives in shared memory (123 KB at S = 961), so the
// softmax needs one pass and the probabilities are written to HBM exactly once -- the reference materialises them
// too (nn.MultiheadAttention with need_weights=True, vit_backbone.py:43).  Same conventions as attention.cu: the
// dropout mask travels to backward in the SIGN BIT of the saved probability; backward needs no generator calls.
//
// Backward is two kernels: (A) per query strip: dP = dO V^T, delta_i = sum_j dP_ij P_ij, dS, dQ = dS K, and
// delta written to a B*H*S workspace; (B) per key strip: dV = PD^T dO and dK = dS^T Q with dS rebuilt from the
// saved probabilities, dO V^T and delta.
#include "common.cuh"

namespace cvae {

constexpr int kLQ = 32;        // strip height (query rows in fwd / bwd-A, key rows in bwd-B)
constexpr int kLK = 64;        // rows of the streamed tile
constexpr int kLThreads = 256;

// rows [r0, r0 + R) of one head's [S, d] slice (row stride `stride`) -> smem [R][ld]; rows >= S are zero
__device__ __forceinline__ void al_load(float* dst, const float* __restrict__ src, int r0, int R, int S, int d, int ld,
                                        size_t stride) {
  const int d4 = d >> 2;
  for (int i = threadIdx.x; i < R * d4; i += blockDim.x) {
    const int r = i / d4, c = (i - r * d4) << 2;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < S) v = __ldg(reinterpret_cast<const float4*>(src + (size_t)(r0 + r) * stride + c));
    float* o = dst + r * ld + c;
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
}

// strip[q][k0 + j] = A[q] . Bt[j] for the 32 x 64 tile: thread = 2 strip rows x 4 tile rows (lane-strided, so the
// tile reads are conflict-free with ld = d + 1 and the strip-row reads are half-warp broadcasts)
__device__ __forceinline__ void al_scores(const float* __restrict__ A, const float* __restrict__ Bt, float* __restrict__ strip,
                                          int lp, int k0, int d, int ld, float scale) {
  const int tq = threadIdx.x >> 4, tk = threadIdx.x & 15;
  float acc[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int u = 0; u < 4; ++u) acc[i][u] = 0.f;
  const float* a0 = A + (2 * tq) * ld;
  const float* b0 = Bt + tk * ld;
#pragma unroll 4
  for (int c = 0; c < d; ++c) {
    const float x0 = a0[c], x1 = a0[ld + c];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float y = b0[16 * u * ld + c];
      acc[0][u] = fmaf(x0, y, acc[0][u]);
      acc[1][u] = fmaf(x1, y, acc[1][u]);
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int u = 0; u < 4; ++u) strip[(2 * tq + i) * lp + k0 + tk + 16 * u] = acc[i][u] * scale;
}

// acc[r][c..] += sum_j strip[r][k0 + j] * T[j][c..]: thread = strip row tid / 8, 4-column groups tid % 8 (+ 8)
__device__ __forceinline__ void al_accum(const float* __restrict__ strip, const float* __restrict__ T, int lp, int k0, int nk,
                                         int d, int ld4, float (&acc)[2][4]) {
  const int r = threadIdx.x >> 3, c = (threadIdx.x & 7) << 2;
  const float* pr = strip + r * lp + k0;
#pragma unroll 4
  for (int j = 0; j < nk; ++j) {
    const float pv = pr[j];
    const float4 v0 = *reinterpret_cast<const float4*>(T + j * ld4 + c);
    acc[0][0] = fmaf(pv, v0.x, acc[0][0]); acc[0][1] = fmaf(pv, v0.y, acc[0][1]);
    acc[0][2] = fmaf(pv, v0.z, acc[0][2]); acc[0][3] = fmaf(pv, v0.w, acc[0][3]);
    if (d > 32) {
      const float4 v1 = *reinterpret_cast<const float4*>(T + j * ld4 + 32 + c);
      acc[1][0] = fmaf(pv, v1.x, acc[1][0]); acc[1][1] = fmaf(pv, v1.y, acc[1][1]);
      acc[1][2] = fmaf(pv, v1.z, acc[1][2]); acc[1][3] = fmaf(pv, v1.w, acc[1][3]);
    }
  }
}

__device__ __forceinline__ void al_store(float* __restrict__ dst, int q0, int S, int d, size_t stride, const float (&acc)[2][4],
                                         float mul) {
  const int r = threadIdx.x >> 3, c = (threadIdx.x & 7) << 2;
  if (q0 + r >= S) return;
  float* o = dst + (size_t)(q0 + r) * stride;
  if (c < d) *reinterpret_cast<float4*>(o + c) = make_float4(acc[0][0] * mul, acc[0][1] * mul, acc[0][2] * mul, acc[0][3] * mul);
  if (32 + c < d) *reinterpret_cast<float4*>(o + 32 + c) = make_float4(acc[1][0] * mul, acc[1][1] * mul, acc[1][2] * mul, acc[1][3] * mul);
}

__global__ void __launch_bounds__(kLThreads) attention_long_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                                       float* __restrict__ probs, int S, int H, int d, int lp,
                                                                       float p_drop, uint64_t seed, uint64_t offset,
                                                                       const int64_t* __restrict__ counter) {
  extern __shared__ __align__(16) float sm_l[];
  if (counter) seed += (uint64_t)(*counter) * 0x9E3779B97F4A7C15ull;
  const int ld = d + 1, ld4 = d + 4;
  float* T = sm_l;                       // streamed K / V tile: [kLK][ld] or [kLK][ld4]
  float* Qs = T + kLK * ld4;             // [kLQ][ld]
  float* P = Qs + kLQ * ld4;             // [kLQ][lp]
  const int bh = blockIdx.y, b = bh / H, h = bh % H, D = H * d, q0 = blockIdx.x * kLQ, tid = threadIdx.x;
  const float* base = qkv + (size_t)b * S * 3 * D + h * d;
  al_load(Qs, base, q0, kLQ, S, d, ld, 3 * D);
  const float scale = rsqrtf((float)d);
  const int ntile = (S + kLK - 1) / kLK;
  for (int kt = 0; kt < ntile; ++kt) {
    __syncthreads();
    al_load(T, base + D, kt * kLK, kLK, S, d, ld, 3 * D);
    __syncthreads();
    al_scores(Qs, T, P, lp, kt * kLK, d, ld, scale);
  }
  __syncthreads();
  const int lane = tid & 31, w = tid >> 5;
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  const int G = (S + 3) >> 2;
  for (int r = w; r < kLQ; r += kLThreads / 32) {
    const int q = q0 + r;
    if (q >= S) break;
    float* pr = P + r * lp;
    float mx = -INFINITY;
    for (int j = lane; j < S; j += 32) mx = fmaxf(mx, pr[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < S; j += 32) { const float e = expf(pr[j] - mx); pr[j] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    float* pg = probs + ((size_t)bh * S + q) * S;
    for (int g = lane; g < G; g += 32) {       // one Philox4x32 call per four keys
      uint4 rnd = make_uint4(~0u, ~0u, ~0u, ~0u);
      if (p_drop > 0.f) rnd = philox4x32(seed, ((uint64_t)bh * S + q) * G + g, offset);
      const uint32_t rw[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = 4 * g + u;
        if (j < S) {
          const float pv = pr[j] * inv;
          const bool keep = p_drop > 0.f ? (float)(rw[u] >> 8) * (1.0f / 16777216.0f) >= p_drop : true;
          pg[j] = keep ? pv : -pv;             // sign bit = dropped
          pr[j] = keep ? pv * keep_scale : 0.f;
        }
      }
    }
  }
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  for (int kt = 0; kt < ntile; ++kt) {
    __syncthreads();
    al_load(T, base + 2 * D, kt * kLK, kLK, S, d, ld4, 3 * D);
    __syncthreads();
    al_accum(P, T, lp, kt * kLK, min(kLK, S - kt * kLK), d, ld4, acc);
  }
  al_store(out + (size_t)b * S * D + h * d, q0, S, d, D, acc, 1.f);
}

// (A) per query strip: dQ and delta
__global__ void __launch_bounds__(kLThreads) attention_long_bwd_q_kernel(const float* __restrict__ qkv, const float* __restrict__ probs,
                                                                         const float* __restrict__ dout, float* __restrict__ dqkv,
                                                                         float* __restrict__ delta, int S, int H, int d, int lp,
                                                                         float p_drop) {
  extern __shared__ __align__(16) float sm_l[];
  const int ld = d + 1, ld4 = d + 4;
  float* T = sm_l;
  float* Gs = T + kLK * ld4;             // dO strip [kLQ][ld]
  float* P = Gs + kLQ * ld4;             // [kLQ][lp]: dO V^T, then dS
  const int bh = blockIdx.y, b = bh / H, h = bh % H, D = H * d, q0 = blockIdx.x * kLQ, tid = threadIdx.x;
  const float* base = qkv + (size_t)b * S * 3 * D + h * d;
  al_load(Gs, dout + (size_t)b * S * D + h * d, q0, kLQ, S, d, ld, D);
  const int ntile = (S + kLK - 1) / kLK;
  for (int kt = 0; kt < ntile; ++kt) {
    __syncthreads();
    al_load(T, base + 2 * D, kt * kLK, kLK, S, d, ld, 3 * D);
    __syncthreads();
    al_scores(Gs, T, P, lp, kt * kLK, d, ld, 1.f);
  }
  __syncthreads();
  const int lane = tid & 31, w = tid >> 5;
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  const float scale = rsqrtf((float)d);
  for (int r = w; r < kLQ; r += kLThreads / 32) {
    const int q = q0 + r;
    if (q >= S) break;
    float* pr = P + r * lp;
    const float* pg = probs + ((size_t)bh * S + q) * S;
    float s = 0.f;
    for (int j = lane; j < S; j += 32) {
      const float pv = __ldg(pg + j);
      const float dp = pv < 0.f ? 0.f : pr[j] * keep_scale;
      pr[j] = dp;
      s = fmaf(dp, fabsf(pv), s);
    }
    s = warp_sum(s);
    if (lane == 0) delta[(size_t)bh * S + q] = s;
    for (int j = lane; j < S; j += 32) pr[j] = fabsf(__ldg(pg + j)) * (pr[j] - s) * scale;
  }
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  for (int kt = 0; kt < ntile; ++kt) {
    __syncthreads();
    al_load(T, base + D, kt * kLK, kLK, S, d, ld4, 3 * D);
    __syncthreads();
    al_accum(P, T, lp, kt * kLK, min(kLK, S - kt * kLK), d, ld4, acc);
  }
  al_store(dqkv + (size_t)b * S * 3 * D + h * d, q0, S, d, 3 * D, acc, 1.f);
}

// (B) per key strip: dK and dV.  Streams tiles of 64 query rows; dS is rebuilt per tile.
__global__ void __launch_bounds__(kLThreads) attention_long_bwd_kv_kernel(const float* __restrict__ qkv, const float* __restrict__ probs,
                                                                          const float* __restrict__ dout, float* __restrict__ dqkv,
                                                                          const float* __restrict__ delta, int S, int H, int d,
                                                                          float p_drop) {
  extern __shared__ __align__(16) float sm_l[];
  const int ld = d + 1, ld4 = d + 4, lt = kLQ + 1;
  float* Vs = sm_l;                      // [kLQ][ld]   this strip's V rows
  float* Gs = Vs + kLQ * ld4;            // [kLK][ld4]  dO tile
  float* Qs = Gs + kLK * ld4;            // [kLK][ld4]  Q tile
  float* PD = Qs + kLK * ld4;            // [kLK][lt]   mask * P * keep_scale
  float* DS = PD + kLK * lt;             // [kLK][lt]   dS
  const int bh = blockIdx.y, b = bh / H, h = bh % H, D = H * d, j0 = blockIdx.x * kLQ, tid = threadIdx.x;
  const int lane = tid & 31, w = tid >> 5;
  const float* base = qkv + (size_t)b * S * 3 * D + h * d;
  const float* gbase = dout + (size_t)b * S * D + h * d;
  al_load(Vs, base + 2 * D, j0, kLQ, S, d, ld, 3 * D);
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  const float scale = rsqrtf((float)d);
  float av[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, ak[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  const int ntile = (S + kLK - 1) / kLK;
  const int jr = tid >> 3, c = (tid & 7) << 2;       // output mapping: key row jr, column group c
  for (int it = 0; it < ntile; ++it) {
    const int i0 = it * kLK;
    __syncthreads();
    al_load(Gs, gbase, i0, kLK, S, d, ld4, D);
    al_load(Qs, base, i0, kLK, S, d, ld4, 3 * D);
    __syncthreads();
    // dP[i][j] = dO_i . V_j: lane = key j, warp = 8 query rows (dO rows read as 128-bit broadcasts)
    {
      float acc[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] = 0.f;
      const float* vj = Vs + lane * ld;
      for (int cc = 0; cc < d; cc += 4) {
        const float y0 = vj[cc], y1 = vj[cc + 1], y2 = vj[cc + 2], y3 = vj[cc + 3];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float4 g = *reinterpret_cast<const float4*>(Gs + (w * 8 + u) * ld4 + cc);
          acc[u] = fmaf(g.x, y0, fmaf(g.y, y1, fmaf(g.z, y2, fmaf(g.w, y3, acc[u]))));
        }
      }
      const int j = j0 + lane;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int il = w * 8 + u, i = i0 + il;
        float pd = 0.f, ds = 0.f;
        if (i < S && j < S) {
          const float pv = __ldg(probs + ((size_t)bh * S + i) * S + j);
          const float a = fabsf(pv);
          const float dp = pv < 0.f ? 0.f : acc[u] * keep_scale;
          pd = pv < 0.f ? 0.f : a * keep_scale;
          ds = a * (dp - __ldg(delta + (size_t)bh * S + i)) * scale;
        }
        PD[il * lt + lane] = pd;
        DS[il * lt + lane] = ds;
      }
    }
    __syncthreads();
    const int ni = min(kLK, S - i0);
#pragma unroll 2
    for (int i = 0; i < ni; ++i) {
      const float pd = PD[i * lt + jr], ds = DS[i * lt + jr];
      const float4 g0 = *reinterpret_cast<const float4*>(Gs + i * ld4 + c);
      const float4 x0 = *reinterpret_cast<const float4*>(Qs + i * ld4 + c);
      av[0][0] = fmaf(pd, g0.x, av[0][0]); av[0][1] = fmaf(pd, g0.y, av[0][1]);
      av[0][2] = fmaf(pd, g0.z, av[0][2]); av[0][3] = fmaf(pd, g0.w, av[0][3]);
      ak[0][0] = fmaf(ds, x0.x, ak[0][0]); ak[0][1] = fmaf(ds, x0.y, ak[0][1]);
      ak[0][2] = fmaf(ds, x0.z, ak[0][2]); ak[0][3] = fmaf(ds, x0.w, ak[0][3]);
      if (d > 32) {
        const float4 g1 = *reinterpret_cast<const float4*>(Gs + i * ld4 + 32 + c);
        const float4 x1 = *reinterpret_cast<const float4*>(Qs + i * ld4 + 32 + c);
        av[1][0] = fmaf(pd, g1.x, av[1][0]); av[1][1] = fmaf(pd, g1.y, av[1][1]);
        av[1][2] = fmaf(pd, g1.z, av[1][2]); av[1][3] = fmaf(pd, g1.w, av[1][3]);
        ak[1][0] = fmaf(ds, x1.x, ak[1][0]); ak[1][1] = fmaf(ds, x1.y, ak[1][1]);
        ak[1][2] = fmaf(ds, x1.z, ak[1][2]); ak[1][3] = fmaf(ds, x1.w, ak[1][3]);
      }
    }
  }
  float* gb = dqkv + (size_t)b * S * 3 * D + h * d;
  al_store(gb + D, j0, S, d, 3 * D, ak, 1.f);
  al_store(gb + 2 * D, j0, S, d, 3 * D, av, 1.f);
}

static inline int al_lp(int S) { return ((S + kLK - 1) / kLK * kLK) + 4; }   // strip pitch: whole tiles (+4: rows land on different banks)
static inline size_t al_smem_strip(int S, int d) { return (size_t)((kLK + kLQ) * (d + 4) + kLQ * al_lp(S)) * sizeof(float); }
static inline size_t al_smem_kv(int d) { return (size_t)((kLQ + 2 * kLK) * (d + 4) + 2 * kLK * (kLQ + 1)) * sizeof(float); }

template <typename K>
static int al_attr(K kernel, size_t smem, size_t& cur) {
  if (smem <= cur) return CVAE_OK;
  if (smem > 227 * 1024) return CVAE_ERR_UNSUPPORTED_SHAPE;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return CVAE_ERR_LAUNCH;
  cur = smem;
  return CVAE_OK;
}

int attention_long_fwd(const float* qkv, float* out, float* probs, int B, int S, int H, int d, float dropout_p, uint64_t seed,
                       uint64_t offset, const int64_t* counter, cudaStream_t st) {
  if (d > 64 || (d & 3)) return CVAE_ERR_UNSUPPORTED_SHAPE;
  static size_t cur = 48 * 1024;
  const size_t smem = al_smem_strip(S, d);
  if (int rc = al_attr(attention_long_fwd_kernel, smem, cur)) return rc;
  dim3 grid((S + kLQ - 1) / kLQ, B * H);
  attention_long_fwd_kernel<<<grid, kLThreads, smem, st>>>(qkv, out, probs, S, H, d, al_lp(S), dropout_p, seed, offset, counter);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

int attention_long_bwd(const float* qkv, const float* probs, const float* dout, float* dqkv, float* ws, int B, int S, int H, int d,
                       float dropout_p, cudaStream_t st) {
  if (d > 64 || (d & 3)) return CVAE_ERR_UNSUPPORTED_SHAPE;
  static size_t cur_q = 48 * 1024, cur_kv = 48 * 1024;
  const size_t smem_q = al_smem_strip(S, d), smem_kv = al_smem_kv(d);
  if (int rc = al_attr(attention_long_bwd_q_kernel, smem_q, cur_q)) return rc;
  if (int rc = al_attr(attention_long_bwd_kv_kernel, smem_kv, cur_kv)) return rc;
  dim3 grid((S + kLQ - 1) / kLQ, B * H);
  attention_long_bwd_q_kernel<<<grid, kLThreads, smem_q, st>>>(qkv, probs, dout, dqkv, ws, S, H, d, al_lp(S), dropout_p);
  CVAE_LAUNCH_CHECK();
  attention_long_bwd_kv_kernel<<<grid, kLThreads, smem_kv, st>>>(qkv, probs, dout, dqkv, ws, S, H, d, dropout_p);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

}  // namespace cvae
