// BatchNorm bookkeeping (finalize / coefficients / apply passes) and LayerNorm for sm_100a.
// The heavy BN work (statistics, normalise + activation) is fused into the conv kernels; what is
// left here are O(C) finalisation kernels and 128-bit vectorised streaming passes.
#include "common.cuh"

namespace cvae {

__global__ void bn_finalize_kernel(const double* __restrict__ stats, int C, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* running_mean, float* running_var,
                                   int64_t* nbt, float* scale, float* shift, float* mean_out, float* rstd_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && nbt != nullptr) *nbt += 1;
  if (c >= C) return;
  const double mean = stats[c] / count;
  double var = stats[C + c] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const double rstd = 1.0 / sqrt(var + (double)eps);
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale[c] = (float)(g * rstd);   // consumer applies (v - mean) * scale + shift
  shift[c] = b;
  if (mean_out) mean_out[c] = (float)mean;
  if (rstd_out) rstd_out[c] = (float)rstd;
  if (running_mean) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

__global__ void bn_eval_coeffs_kernel(const float* rm, const float* rv, const float* gamma, const float* beta,
                                      float eps, int C, float* scale, float* shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float rstd = 1.0f / sqrtf(rv[c] + eps);
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale[c] = g * rstd;            // used with center = running_mean
  shift[c] = b;
}

// stats = (sum dz, sum dz*(y-mean)) ->  dy = ca*dz + cb*(y-mean) + cc
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ stats, int C, double count,
                                       const float* gamma, const float* mean, const float* rstd, float* ca,
                                       float* cb, float* cc, float* dgamma, float* dbeta, float* dbias_pre) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double sdz = stats[c], sdzc = stats[C + c];
  const double rs = rstd[c], g = gamma ? (double)gamma[c] : 1.0;
  const double sdzxh = rs * sdzc;  // sum dz * xhat
  if (dgamma) dgamma[c] = (float)sdzxh;
  if (dbeta) dbeta[c] = (float)sdz;
  const double k1 = sdz / count, k2 = sdzxh / count;
  ca[c] = (float)(g * rs); cb[c] = (float)(-g * rs * rs * k2); cc[c] = (float)(-g * rs * k1);
  // sum over the batch of dy = ca*sum(dz) + cb*sum(y-mean) + count*cc = 0: a bias feeding a
  // training-mode BatchNorm has an exactly zero gradient (the reference produces rounding noise)
  if (dbias_pre) dbias_pre[c] = 0.f;
}

// generic [rows, C] streaming kernels; C % 4 == 0 -> float4 path, else scalar
template <bool HAS_B>
__global__ void affine_act_kernel(const float* __restrict__ a, XformDev xa, const float* __restrict__ b,
                                  XformDev xb, float* __restrict__ out, int64_t total, int C) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if ((C & 3) == 0) {
    const int64_t n4 = total >> 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
      const int c = (int)((i << 2) % C);
      float4 v = reinterpret_cast<const float4*>(a)[i];
      float r[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (xa.affine) r[u] = fmaf(r[u] - (xa.center ? __ldg(xa.center + c + u) : 0.f), __ldg(xa.scale + c + u), __ldg(xa.shift + c + u));
        if (xa.act) r[u] = lrelu(r[u], xa.slope);
      }
      if (HAS_B) {
        const float4 w = reinterpret_cast<const float4*>(b)[i];
        float q[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (xb.affine) q[u] = fmaf(q[u] - (xb.center ? __ldg(xb.center + c + u) : 0.f), __ldg(xb.scale + c + u), __ldg(xb.shift + c + u));
          if (xb.act) q[u] = lrelu(q[u], xb.slope);
          r[u] += q[u];
        }
      }
      reinterpret_cast<float4*>(out)[i] = make_float4(r[0], r[1], r[2], r[3]);
    }
  } else {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
      const int c = (int)(i % C);
      float r = a[i];
      if (xa.affine) r = fmaf(r - (xa.center ? xa.center[c] : 0.f), xa.scale[c], xa.shift[c]);
      if (xa.act) r = lrelu(r, xa.slope);
      if (HAS_B) {
        float q = b[i];
        if (xb.affine) q = fmaf(q - (xb.center ? xb.center[c] : 0.f), xb.scale[c], xb.shift[c]);
        if (xb.act) q = lrelu(q, xb.slope);
        r += q;
      }
      out[i] = r;
    }
  }
}

__global__ void bn_bwd_apply_kernel(const float* __restrict__ dz, const float* __restrict__ y,
                                    const float* __restrict__ ca, const float* __restrict__ cb,
                                    const float* __restrict__ cc, const float* __restrict__ mean,
                                    float* __restrict__ out, int64_t total, int C) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if ((C & 3) == 0) {
    const int64_t n4 = total >> 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
      const int c = (int)((i << 2) % C);
      const float4 d = reinterpret_cast<const float4*>(dz)[i];
      const float4 v = reinterpret_cast<const float4*>(y)[i];
      const float4 A = __ldg(reinterpret_cast<const float4*>(ca + c));
      const float4 B = __ldg(reinterpret_cast<const float4*>(cb + c));
      const float4 Cc = __ldg(reinterpret_cast<const float4*>(cc + c));
      const float4 Mu = __ldg(reinterpret_cast<const float4*>(mean + c));
      float4 o;
      o.x = fmaf(A.x, d.x, fmaf(B.x, v.x - Mu.x, Cc.x)); o.y = fmaf(A.y, d.y, fmaf(B.y, v.y - Mu.y, Cc.y));
      o.z = fmaf(A.z, d.z, fmaf(B.z, v.z - Mu.z, Cc.z)); o.w = fmaf(A.w, d.w, fmaf(B.w, v.w - Mu.w, Cc.w));
      reinterpret_cast<float4*>(out)[i] = o;
    }
  } else {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
      const int c = (int)(i % C);
      out[i] = fmaf(ca[c], dz[i], fmaf(cb[c], y[i] - mean[c], cc[c]));
    }
  }
}

// bn_bwd_finalize + bn_bwd_apply in one launch (C <= 512, C % 4 == 0): every block derives the per-channel
// coefficients from the statistics into shared memory (2 C doubles from L2), block 0 also writes
// dgamma / dbeta / dbias; then the stream pass.  One launch less per BatchNorm on the backward critical path.
__global__ void __launch_bounds__(256) bn_bwd_fused_kernel(const float* __restrict__ dz, const float* __restrict__ y,
                                                           const double* __restrict__ stats, double count,
                                                           const float* __restrict__ gamma, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, float* __restrict__ out,
                                                           float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                           float* __restrict__ dbias_pre, int64_t total, int C) {
  __shared__ __align__(16) float sA[512], sB[512], sC[512], sM[512];
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double sdz = stats[c], sdzc = stats[C + c];
    const double rs = rstd[c], g = gamma ? (double)gamma[c] : 1.0;
    const double sdzxh = rs * sdzc;
    const double k1 = sdz / count, k2 = sdzxh / count;
    sA[c] = (float)(g * rs); sB[c] = (float)(-g * rs * rs * k2); sC[c] = (float)(-g * rs * k1); sM[c] = mean[c];
    if (blockIdx.x == 0) {
      if (dgamma) dgamma[c] = (float)sdzxh;
      if (dbeta) dbeta[c] = (float)sdz;
      if (dbias_pre) dbias_pre[c] = 0.f;     // a bias feeding a training-mode BatchNorm has an exactly zero gradient
    }
  }
  __syncthreads();
  // Four vector pairs per thread are requested before the first is used: inside the training step this pass runs beside
  // the weight-gradient kernels of the side stream, which leave room for ONE 256-thread block per SM - with two loads in
  // flight per thread that is 8 KB per SM and the pass ran at ~1.7 TB/s (2.5x its stand-alone time).
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, n4 = total >> 2;
  for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < n4; i0 += 4 * stride) {
    float4 d[4], v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n4) {
        d[u] = __ldg(reinterpret_cast<const float4*>(dz) + i);
        v[u] = __ldg(reinterpret_cast<const float4*>(y) + i);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n4) {
        const int c = (int)((i << 2) % C);
        const float4 A = *reinterpret_cast<const float4*>(sA + c), B = *reinterpret_cast<const float4*>(sB + c);
        const float4 Cc = *reinterpret_cast<const float4*>(sC + c), Mu = *reinterpret_cast<const float4*>(sM + c);
        float4 o;
        o.x = fmaf(A.x, d[u].x, fmaf(B.x, v[u].x - Mu.x, Cc.x)); o.y = fmaf(A.y, d[u].y, fmaf(B.y, v[u].y - Mu.y, Cc.y));
        o.z = fmaf(A.z, d[u].z, fmaf(B.z, v[u].z - Mu.z, Cc.z)); o.w = fmaf(A.w, d[u].w, fmaf(B.w, v[u].w - Mu.w, Cc.w));
        reinterpret_cast<float4*>(out)[i] = o;
      }
    }
  }
}

// column statistics of a [rows, C] matrix.  Block (32 x 8): threadIdx.x -> column, threadIdx.y
// strides rows; MODE 0: (sum y, sum y^2); MODE 1: dz = g*act'(xform(ref)) written out, (sum dz, sum dz*ref);
// MODE 2: plain column sum into a float output.
template <int MODE>
__global__ void col_reduce_kernel(const float* __restrict__ p0, const float* __restrict__ p1, XformDev x,
                                  float* __restrict__ out, double* __restrict__ stats, float* __restrict__ fsum,
                                  int64_t rows, int C, int accumulate) {
  __shared__ double r1[8][33], r2[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  double s1 = 0.0, s2 = 0.0;
  if (c < C) {
    const float sc = (MODE == 1 && x.affine) ? x.scale[c] : 1.f, sh = (MODE == 1 && x.affine) ? x.shift[c] : 0.f;
    const float ce = (MODE == 1 && x.affine && x.center) ? x.center[c] : 0.f;
    for (int64_t r = blockIdx.y * 8 + threadIdx.y; r < rows; r += (int64_t)gridDim.y * 8) {
      const int64_t i = r * C + c;
      if (MODE == 0) {
        const double v = (double)p0[i];
        s1 += v; s2 += v * v;
      } else if (MODE == 1) {
        const float ref = p1[i] - ce;
        const float z = fmaf(ref, sc, sh);
        float g = p0[i];
        g = z > 0.f ? g : g * x.slope;
        out[i] = g;
        s1 += (double)g; s2 += (double)g * (double)ref;
      } else {
        s1 += (double)p0[i];
      }
    }
  }
  r1[threadIdx.y][threadIdx.x] = s1; r2[threadIdx.y][threadIdx.x] = s2;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    double t1 = 0.0, t2 = 0.0;
#pragma unroll
    for (int y = 0; y < 8; ++y) { t1 += r1[y][threadIdx.x]; t2 += r2[y][threadIdx.x]; }
    if (MODE == 2) {
      atomicAdd(fsum + c, (float)t1);
    } else {
      atomicAdd(stats + c, t1);
      atomicAdd(stats + C + c, t2);
    }
  }
}

// ---- LayerNorm: one warp per row --------------------------------------------------------------
// s = a + b (written out) and y = LayerNorm(s) in one pass: the residual add in front of norm2 of a ViT block
// (vit_backbone.py:44-46).  D <= 1024, contiguous rows.
__global__ void add_layernorm_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                         float* __restrict__ sum, float* __restrict__ y, float* __restrict__ mean,
                                         float* __restrict__ rstd, int64_t rows, int D, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* ar = a + row * D;
  const float* br = b + row * D;
  float* sr = sum + row * D;
  float v[32];                                     // D <= 1024
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const int i = lane + 32 * k;
    v[k] = i < D ? ar[i] + br[i] : 0.f;
    s += v[k];
  }
  const float mu = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k) { const float d = (lane + 32 * k < D) ? v[k] - mu : 0.f; q = fmaf(d, d, q); }
  const float rs = rsqrtf(warp_sum(q) / D + eps);
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const int i = lane + 32 * k;
    if (i < D) { sr[i] = v[k]; y[row * D + i] = (v[k] - mu) * rs * gamma[i] + beta[i]; }
  }
  if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
}

__global__ void layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float* __restrict__ y, float* __restrict__ mean,
                                     float* __restrict__ rstd, int64_t rows, int D, int64_t xs, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * xs;
  float s = 0.f;
  for (int i = lane; i < D; i += 32) s += xr[i];
  const float mu = warp_sum(s) / D;
  float q = 0.f;
  for (int i = lane; i < D; i += 32) { const float d = xr[i] - mu; q = fmaf(d, d, q); }
  const float rs = rsqrtf(warp_sum(q) / D + eps);
  for (int i = lane; i < D; i += 32) y[row * D + i] = (xr[i] - mu) * rs * gamma[i] + beta[i];
  if (lane == 0) { if (mean) mean[row] = mu; if (rstd) rstd[row] = rs; }
}

// block = 256 threads = 8 warps; each block walks rows with stride; dgamma/dbeta partials in
// shared memory, flushed with one atomic per column per block.  D <= 1024.
__global__ void layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                     const float* __restrict__ gamma, const float* __restrict__ mean,
                                     const float* __restrict__ rstd, float* __restrict__ dx,
                                     float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t rows, int D,
                                     int64_t xs, int64_t dxs, int accumulate_dx) {
  extern __shared__ float sm[];  // [2][D]
  float* sg = sm; float* sb = sm + D;
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int64_t row = (int64_t)blockIdx.x * nw + w; row < rows; row += (int64_t)gridDim.x * nw) {
    const float* xr = x + row * xs;
    const float* dr = dy + row * D;
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
    for (int i = lane; i < D; i += 32) {
      const float g = dr[i] * gamma[i], xh = (xr[i] - mu) * rs;
      s1 += g; s2 = fmaf(g, xh, s2);
    }
    s1 = warp_sum(s1) / D; s2 = warp_sum(s2) / D;
    for (int i = lane; i < D; i += 32) {
      const float xh = (xr[i] - mu) * rs, d = dr[i];
      const float v = rs * (d * gamma[i] - s1 - xh * s2);
      float* o = dx + row * dxs + i;
      *o = accumulate_dx ? *o + v : v;
      atomicAdd(sg + i, d * xh);
      atomicAdd(sb + i, d);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    atomicAdd(dgamma + i, sg[i]);
    atomicAdd(dbeta + i, sb[i]);
  }
}

// D <= 256, D % 32 == 0: a lane owns columns lane + 32 k and keeps their dgamma / dbeta sums in registers
// across the rows of its warp (the generic kernel spends two shared-memory atomics per element, with
// all 8 warps of a block contending for the same 2 D words); x, dy and gamma are read once per row.
template <int NPL>
__global__ void __launch_bounds__(256) layernorm_bwd_reg_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ mean,
                                                                const float* __restrict__ rstd, float* __restrict__ dx,
                                                                float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                int64_t rows, int D, int64_t xs, int64_t dxs,
                                                                int accumulate_dx, const float* __restrict__ dadd = nullptr) {
  extern __shared__ float sm[];  // [2][D]
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int n = D >> 5;
  float ga[NPL], pg[NPL], pb[NPL];
#pragma unroll
  for (int k = 0; k < NPL; ++k) { ga[k] = k < n ? gamma[lane + 32 * k] : 0.f; pg[k] = 0.f; pb[k] = 0.f; }
  for (int64_t row = (int64_t)blockIdx.x * nw + w; row < rows; row += (int64_t)gridDim.x * nw) {
    const float* xr = x + row * xs;
    const float* dr = dy + row * D;
    const float mu = mean[row], rs = rstd[row];
    float d[NPL], xh[NPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NPL; ++k) {
      d[k] = k < n ? dr[lane + 32 * k] : 0.f;
      xh[k] = k < n ? (xr[lane + 32 * k] - mu) * rs : 0.f;
      const float g = d[k] * ga[k];
      s1 += g; s2 = fmaf(g, xh[k], s2);
    }
    s1 = warp_sum(s1) / D; s2 = warp_sum(s2) / D;
#pragma unroll
    for (int k = 0; k < NPL; ++k) {
      if (k < n) {
        float v = rs * (d[k] * ga[k] - s1 - xh[k] * s2);
        if (dadd != nullptr) v += dadd[row * D + lane + 32 * k];     // gradient arriving at the same tensor by the residual path
        float* o = dx + row * dxs + lane + 32 * k;
        *o = accumulate_dx ? *o + v : v;
        pg[k] = fmaf(d[k], xh[k], pg[k]); pb[k] += d[k];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NPL; ++k)
    if (k < n) { atomicAdd(sm + lane + 32 * k, pg[k]); atomicAdd(sm + D + lane + 32 * k, pb[k]); }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    atomicAdd(dgamma + i, sm[i]);
    atomicAdd(dbeta + i, sm[D + i]);
  }
}

// sum of a flat vector (column sum with C == 1, e.g. the bias gradient of the 1-channel image head):
// 128-bit loads, double accumulation per thread, one float atomic per block.
__global__ void __launch_bounds__(256) flat_sum_kernel(const float* __restrict__ p, int64_t n, float* __restrict__ out) {
  __shared__ double red[8];
  double s = 0.0;
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p) + i);
    s += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) s += (double)p[(n4 << 2) + threadIdx.x];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    atomicAdd(out, (float)t);
  }
}

static inline XformDev make_x(cvae_xform_t x) {
  XformDev d;
  d.scale = x.scale; d.shift = x.shift; d.center = x.center; d.slope = x.slope;
  d.affine = x.scale != nullptr; d.act = x.slope != 1.0f;
  return d;
}
static inline int stream_blocks(int64_t work_items) {
  int64_t b = (work_items + 255) / 256;
  if (b > kNumSMs * 16) b = kNumSMs * 16;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace cvae
using namespace cvae;

extern "C" int cvae_bn_finalize(const double* stats, int C, double count, const float* gamma, const float* beta,
                                float eps, float momentum, float* running_mean, float* running_var,
                                int64_t* nbt, float* scale, float* shift, float* mean, float* rstd,
                                cvae_stream_t s) {
  if (!stats || !scale || !shift || C <= 0 || count <= 0) return CVAE_ERR_BAD_ARG;
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, as_stream(s)>>>(stats, C, count, gamma, beta, eps, momentum,
                                                                running_mean, running_var, nbt, scale, shift,
                                                                mean, rstd);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_bn_eval_coeffs(const float* rm, const float* rv, const float* gamma, const float* beta,
                                   float eps, int C, float* scale, float* shift, cvae_stream_t s) {
  if (!rm || !rv || !scale || !shift || C <= 0) return CVAE_ERR_BAD_ARG;
  bn_eval_coeffs_kernel<<<(C + 127) / 128, 128, 0, as_stream(s)>>>(rm, rv, gamma, beta, eps, C, scale, shift);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_bn_bwd_finalize(const double* stats, int C, double count, const float* gamma,
                                    const float* mean, const float* rstd, float* ca, float* cb, float* cc,
                                    float* dgamma, float* dbeta, float* dbias_pre, cvae_stream_t s) {
  if (!stats || !mean || !rstd || !ca || !cb || !cc || C <= 0) return CVAE_ERR_BAD_ARG;
  bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, as_stream(s)>>>(stats, C, count, gamma, mean, rstd, ca, cb,
                                                                    cc, dgamma, dbeta, dbias_pre);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_affine_act(const float* a, cvae_xform_t xa, const float* b, cvae_xform_t xb, float* out,
                               int64_t rows, int C, cvae_stream_t s) {
  if (!a || !out || rows <= 0 || C <= 0) return CVAE_ERR_BAD_ARG;
  const int64_t total = rows * C;
  const int blocks = stream_blocks((C & 3) == 0 ? total / 4 : total);
  if (b) affine_act_kernel<true><<<blocks, 256, 0, as_stream(s)>>>(a, make_x(xa), b, make_x(xb), out, total, C);
  else affine_act_kernel<false><<<blocks, 256, 0, as_stream(s)>>>(a, make_x(xa), nullptr, make_x(xb), out, total, C);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_bn_bwd_apply(const float* dz, const float* y, const float* ca, const float* cb,
                                 const float* cc, const float* mean, float* out, int64_t rows, int C,
                                 cvae_stream_t s) {
  if (!dz || !y || !out || !mean || rows <= 0 || C <= 0) return CVAE_ERR_BAD_ARG;
  const int64_t total = rows * C;
  bn_bwd_apply_kernel<<<stream_blocks((C & 3) == 0 ? total / 4 : total), 256, 0, as_stream(s)>>>(dz, y, ca, cb, cc,
                                                                                                 mean, out, total, C);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

// 1 when cvae_bn_bwd covers the shape (otherwise use cvae_bn_bwd_finalize + cvae_bn_bwd_apply)
extern "C" int cvae_bn_bwd_fused_ok(int C) { return (C <= 512 && (C & 3) == 0) ? 1 : 0; }

extern "C" int cvae_bn_bwd(const float* dz, const float* y, const double* stats, double count, const float* gamma,
                           const float* mean, const float* rstd, float* out, float* dgamma, float* dbeta,
                           float* dbias_pre, int64_t rows, int C, cvae_stream_t s) {
  if (!dz || !y || !stats || !mean || !rstd || !out || rows <= 0 || count <= 0) return CVAE_ERR_BAD_ARG;
  if (!cvae_bn_bwd_fused_ok(C)) return CVAE_ERR_UNSUPPORTED_SHAPE;
  const int64_t total = rows * C;
  bn_bwd_fused_kernel<<<stream_blocks(total / 4), 256, 0, as_stream(s)>>>(dz, y, stats, count, gamma, mean, rstd, out,
                                                                           dgamma, dbeta, dbias_pre, total, C);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

static dim3 col_grid(int64_t rows, int C) {
  const int gx = (C + 31) / 32;
  int64_t gy = (rows + 63) / 64;
  const int64_t cap = (kNumSMs * 8 + gx - 1) / gx;
  if (gy > cap) gy = cap;
  if (gy < 1) gy = 1;
  return dim3(gx, (unsigned)gy, 1);
}

extern "C" int cvae_col_stats(const float* y, int64_t rows, int C, double* stats, cvae_stream_t s) {
  if (!y || !stats || rows <= 0 || C <= 0) return CVAE_ERR_BAD_ARG;
  XformDev x{};
  col_reduce_kernel<0><<<col_grid(rows, C), dim3(32, 8), 0, as_stream(s)>>>(y, nullptr, x, nullptr, stats, nullptr,
                                                                           rows, C, 0);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_dact_stats(const float* g, const float* ref, cvae_xform_t x, float* dz, double* stats,
                               int64_t rows, int C, cvae_stream_t s) {
  if (!g || !ref || !dz || !stats || rows <= 0 || C <= 0) return CVAE_ERR_BAD_ARG;
  col_reduce_kernel<1><<<col_grid(rows, C), dim3(32, 8), 0, as_stream(s)>>>(g, ref, make_x(x), dz, stats, nullptr,
                                                                           rows, C, 0);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_col_sum(const float* x, int64_t rows, int C, float* out, int accumulate, cvae_stream_t s) {
  if (!x || !out || rows <= 0 || C <= 0) return CVAE_ERR_BAD_ARG;
  if (!accumulate) {
    if (cudaMemsetAsync(out, 0, sizeof(float) * C, as_stream(s)) != cudaSuccess) return CVAE_ERR_LAUNCH;
  }
  if (C == 1 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    int64_t blocks = (rows / 4 + 255) / 256;
    if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
    if (blocks < 1) blocks = 1;
    flat_sum_kernel<<<(unsigned)blocks, 256, 0, as_stream(s)>>>(x, rows, out);
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
  }
  XformDev xf{};
  col_reduce_kernel<2><<<col_grid(rows, C), dim3(32, 8), 0, as_stream(s)>>>(x, nullptr, xf, nullptr, nullptr, out,
                                                                           rows, C, accumulate);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean,
                                  float* rstd, int64_t rows, int D, int64_t xs, float eps, cvae_stream_t s) {
  if (!x || !gamma || !beta || !y || rows <= 0 || D <= 0) return CVAE_ERR_BAD_ARG;
  const int64_t blocks = (rows + 7) / 8;
  layernorm_fwd_kernel<<<(unsigned)blocks, 256, 0, as_stream(s)>>>(x, gamma, beta, y, mean, rstd, rows, D, xs, eps);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_add_layernorm_fwd(const float* a, const float* b, const float* gamma, const float* beta, float* sum,
                                      float* y, float* mean, float* rstd, int64_t rows, int D, float eps, cvae_stream_t s) {
  if (!a || !b || !gamma || !beta || !sum || !y || !mean || !rstd || rows <= 0 || D <= 0) return CVAE_ERR_BAD_ARG;
  if (D > 1024) return CVAE_ERR_UNSUPPORTED_SHAPE;
  add_layernorm_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, as_stream(s)>>>(a, b, gamma, beta, sum, y, mean, rstd, rows, D, eps);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

// dx = LayerNorm-backward(dy) + dadd  (dadd: the gradient reaching the same tensor through the residual connection)
extern "C" int cvae_layernorm_bwd_add(const float* dy, const float* x, const float* gamma, const float* mean,
                                      const float* rstd, const float* dadd, float* dx, float* dgamma, float* dbeta,
                                      int64_t rows, int D, cvae_stream_t s) {
  if (!dy || !x || !gamma || !mean || !rstd || !dadd || !dx || !dgamma || !dbeta || rows <= 0 || D <= 0) return CVAE_ERR_BAD_ARG;
  if (D % 32 != 0 || D > 256) return CVAE_ERR_UNSUPPORTED_SHAPE;
  int64_t blocks = (rows + 15) / 16;
  if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
  layernorm_bwd_reg_kernel<8><<<(unsigned)blocks, 256, 2 * D * sizeof(float), as_stream(s)>>>(
      dy, x, gamma, mean, rstd, dx, dgamma, dbeta, rows, D, D, D, 0, dadd);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean,
                                  const float* rstd, float* dx, float* dgamma, float* dbeta, int64_t rows, int D,
                                  int64_t xs, int64_t dxs, int accumulate_dx, cvae_stream_t s) {
  if (!dy || !x || !gamma || !mean || !rstd || !dx || !dgamma || !dbeta || rows <= 0 || D <= 0) return CVAE_ERR_BAD_ARG;
  if (D > 1024) return CVAE_ERR_UNSUPPORTED_SHAPE;
  if (D % 32 == 0 && D <= 256) {
    int64_t blocks = (rows + 15) / 16;             // two rows per warp: the register sums amortise over them
    if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
    layernorm_bwd_reg_kernel<8><<<(unsigned)blocks, 256, 2 * D * sizeof(float), as_stream(s)>>>(
        dy, x, gamma, mean, rstd, dx, dgamma, dbeta, rows, D, xs, dxs, accumulate_dx);
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
  }
  int64_t blocks = (rows + 31) / 32;
  if (blocks > kNumSMs * 2) blocks = kNumSMs * 2;
  layernorm_bwd_kernel<<<(unsigned)blocks, 256, 2 * D * sizeof(float), as_stream(s)>>>(
      dy, x, gamma, mean, rstd, dx, dgamma, dbeta, rows, D, xs, dxs, accumulate_dx);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}
