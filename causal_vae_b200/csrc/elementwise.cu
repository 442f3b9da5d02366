// Streaming elementwise / layout kernels (HBM-bound: 128-bit accesses, grid-stride, grid sized in
// multiples of the SM count) and the small latent-space kernels.
#include "common.cuh"

namespace cvae {

static inline int ew_blocks(int64_t items) {
  int64_t b = (items + 255) / 256;
  if (b > kNumSMs * 16) b = kNumSMs * 16;
  if (b < 1) b = 1;
  return (int)b;
}

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}

template <int ACT>
__global__ void act_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, float slope) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, n4 = n >> 2;
  auto f = [&](float v) {
    if (ACT == CVAE_ACT_LRELU) return lrelu(v, slope);
    if (ACT == CVAE_ACT_GELU) return gelu_f(v);
    return 1.0f / (1.0f + expf(-v));
  };
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = reinterpret_cast<const float4*>(x)[i];
    v.x = f(v.x); v.y = f(v.y); v.z = f(v.z); v.w = f(v.w);
    reinterpret_cast<float4*>(y)[i] = v;
  }
  for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) y[i] = f(x[i]);
}

template <int ACT>
__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dx,
                               int64_t n, float slope) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, n4 = n >> 2;
  auto f = [&](float g, float v) {
    if (ACT == CVAE_ACT_LRELU) return v > 0.f ? g : g * slope;
    if (ACT == CVAE_ACT_GELU) return g * gelu_grad(v);
    return g * v * (1.0f - v);  // v = sigmoid output
  };
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 g = reinterpret_cast<const float4*>(dy)[i];
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    reinterpret_cast<float4*>(dx)[i] = make_float4(f(g.x, v.x), f(g.y, v.y), f(g.z, v.z), f(g.w, v.w));
  }
  for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) dx[i] = f(dy[i], x[i]);
}

__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, n4 = n >> 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 u = reinterpret_cast<const float4*>(a)[i], v = reinterpret_cast<const float4*>(b)[i];
    reinterpret_cast<float4*>(o)[i] = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
  }
  for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) o[i] = a[i] + b[i];
}

__global__ void fill_kernel(float* __restrict__ o, int64_t n, float v) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) o[i] = v;
}

__global__ void dropout_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, float p,
                               uint64_t seed, uint64_t offset, const int64_t* __restrict__ counter) {
  if (counter) seed += (uint64_t)(*counter) * 0x9E3779B97F4A7C15ull;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, n4 = (n + 3) >> 2;
  const float ks = 1.f / (1.f - p);
  const float thr = p;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const uint4 r = philox4x32(seed, (uint64_t)i, offset);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t e = (i << 2) + u;
      if (e < n) y[e] = ((float)(w[u] >> 8) * (1.0f / 16777216.0f) >= thr) ? x[e] * ks : 0.f;
    }
  }
}

// Fused activation + dropout (ViT MLP: GELU -> Dropout, vit_backbone.py:33-35) and dropout + residual add.
// The keep mask is the one dropout_kernel draws for the same (seed, offset): element e uses word e & 3 of
// Philox counter e >> 2, so fusing changes launches, not random streams.
// MODE 0: y = act(x) * m;   MODE 1 (backward): y = aux * act'(x) * m  (aux = dL/dy);   MODE 2: y = aux + x * m
template <int MODE, int ACT>
__global__ void dropout_fused_kernel(const float* __restrict__ x, const float* __restrict__ aux, float* __restrict__ y,
                                     int64_t n, float slope, float p, uint64_t seed, uint64_t offset,
                                     const int64_t* __restrict__ counter) {
  if (counter) seed += (uint64_t)(*counter) * 0x9E3779B97F4A7C15ull;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, n4 = (n + 3) >> 2;
  const float ks = 1.f / (1.f - p);
  auto act = [&](float v) {
    if (ACT == CVAE_ACT_LRELU) return lrelu(v, slope);
    if (ACT == CVAE_ACT_GELU) return gelu_f(v);
    return 1.0f / (1.0f + expf(-v));
  };
  auto dact = [&](float v) {
    if (ACT == CVAE_ACT_LRELU) return v > 0.f ? 1.f : slope;
    if (ACT == CVAE_ACT_GELU) return gelu_grad(v);
    const float sg = 1.0f / (1.0f + expf(-v));
    return sg * (1.0f - sg);
  };
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const uint4 r = philox4x32(seed, (uint64_t)i, offset);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    float xv[4] = {0.f, 0.f, 0.f, 0.f}, av[4] = {0.f, 0.f, 0.f, 0.f};
    const bool full = (i << 2) + 3 < n;
    if (full) {
      const float4 t = reinterpret_cast<const float4*>(x)[i];
      xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w;
      if (MODE != 0) { const float4 u = reinterpret_cast<const float4*>(aux)[i]; av[0] = u.x; av[1] = u.y; av[2] = u.z; av[3] = u.w; }
    } else {
      for (int u = 0; u < 4; ++u) {
        const int64_t e = (i << 2) + u;
        if (e < n) { xv[u] = x[e]; if (MODE != 0) av[u] = aux[e]; }
      }
    }
    float o[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float m = ((float)(w[u] >> 8) * (1.0f / 16777216.0f) >= p) ? ks : 0.f;
      if (MODE == 0) o[u] = act(xv[u]) * m;
      else if (MODE == 1) o[u] = av[u] * dact(xv[u]) * m;
      else o[u] = av[u] + xv[u] * m;
    }
    if (full) reinterpret_cast<float4*>(y)[i] = make_float4(o[0], o[1], o[2], o[3]);
    else for (int u = 0; u < 4; ++u) { const int64_t e = (i << 2) + u; if (e < n) y[e] = o[u]; }
  }
}

__global__ void tokens_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ cls,
                                  const float* __restrict__ pos, float* __restrict__ tok, int B, int n, int D) {
  const int64_t total = (int64_t)B * (n + 1) * D, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % D);
    const int64_t r = i / D;
    const int s = (int)(r % (n + 1));
    const int64_t b = r / (n + 1);
    const float v = s == 0 ? cls[c] : feat[(b * n + (s - 1)) * D + c];
    tok[i] = v + pos[(int64_t)s * D + c];
  }
}

// dfeat = dtok[:, 1:], dpos[s] = sum_b dtok[b, s], dcls = sum_b dtok[b, 0]; one thread per (s, c)
__global__ void tokens_bwd_kernel(const float* __restrict__ dtok, float* __restrict__ dfeat, float* __restrict__ dcls,
                                  float* __restrict__ dpos, int B, int n, int D) {
  const int total = (n + 1) * D;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % D, s = i / D;
    float acc = 0.f;
    for (int b = 0; b < B; ++b) {
      const float g = dtok[((int64_t)b * (n + 1) + s) * D + c];
      acc += g;
      if (s > 0 && dfeat) dfeat[((int64_t)b * n + (s - 1)) * D + c] = g;
    }
    if (dpos) dpos[i] = acc;
    if (s == 0 && dcls) dcls[c] = acc;
  }
}

// per-batch transpose [rows, cols] -> [cols, rows] through a padded 32x32 shared tile
__global__ void transpose_bc_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
  __shared__ float tile[32][33];
  const float* s = src + (size_t)blockIdx.z * rows * cols;
  float* d = dst + (size_t)blockIdx.z * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = s[(size_t)r * cols + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) d[(size_t)c * rows + r] = tile[threadIdx.x][j];
  }
}

__global__ void copy_cols_kernel(const float* __restrict__ src, int64_t sld, int sc0, float* __restrict__ dst,
                                 int64_t dld, int dc0, int64_t rows, int w, int accumulate) {
  const int64_t total = rows * w, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / w;
    const int c = (int)(i % w);
    const float v = src[r * sld + sc0 + c];
    float* o = dst + r * dld + dc0 + c;
    *o = accumulate ? *o + v : v;
  }
}

// ---- latent ------------------------------------------------------------------------------------
__global__ void latent_fwd_kernel(const float* __restrict__ h, const float* __restrict__ eps, float* __restrict__ mu,
                                  float* __restrict__ logvar, float* __restrict__ z, double* kld, int B, int Z,
                                  float mu_clamp, float lv_clamp) {
  __shared__ double red[32];
  const int total = B * Z;
  double acc = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / Z, j = i % Z;
    float m = h[(size_t)b * 2 * Z + j], lv = h[(size_t)b * 2 * Z + Z + j];
    if (mu_clamp > 0.f) m = fminf(fmaxf(m, -mu_clamp), mu_clamp);
    if (lv_clamp > 0.f) lv = fminf(fmaxf(lv, -lv_clamp), lv_clamp);
    mu[i] = m; logvar[i] = lv;
    if (z) z[i] = fmaf(eps[i], expf(0.5f * lv), m);
    acc += (double)(1.0f + lv - m * m - expf(lv));
  }
  const double t = block_sum_d(acc, red);
  if (threadIdx.x == 0 && kld) atomicAdd(kld, -0.5 * t);
}

__global__ void latent_bwd_kernel(const float* __restrict__ h, const float* __restrict__ eps,
                                  const float* __restrict__ dz, const float* __restrict__ dmu,
                                  const float* __restrict__ dlv, float* __restrict__ dh, int B, int Z, float mu_clamp,
                                  float lv_clamp) {
  const int total = B * Z;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / Z, j = i % Z;
    const float mr = h[(size_t)b * 2 * Z + j], lr = h[(size_t)b * 2 * Z + Z + j];
    // torch.clamp passes the gradient where min <= x <= max
    const bool mu_pass = !(mu_clamp > 0.f) || (mr >= -mu_clamp && mr <= mu_clamp);
    const bool lv_pass = !(lv_clamp > 0.f) || (lr >= -lv_clamp && lr <= lv_clamp);
    const float lv = lv_clamp > 0.f ? fminf(fmaxf(lr, -lv_clamp), lv_clamp) : lr;
    float gm = dmu ? dmu[i] : 0.f, gl = dlv ? dlv[i] : 0.f;
    if (dz) {
      const float g = dz[i];
      gm += g;
      gl += g * eps[i] * 0.5f * expf(0.5f * lv);
    }
    dh[(size_t)b * 2 * Z + j] = mu_pass ? gm : 0.f;
    dh[(size_t)b * 2 * Z + Z + j] = lv_pass ? gl : 0.f;
  }
}

__global__ void gauss_nll_fwd_kernel(const float* __restrict__ m, const float* __restrict__ mmu,
                                     const float* __restrict__ raw, float* __restrict__ lv_out, double* nll,
                                     int64_t n, float lv_clamp) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float lv = raw[i];
    if (lv_clamp > 0.f) lv = fminf(fmaxf(lv, -lv_clamp), lv_clamp);
    if (lv_out) lv_out[i] = lv;
    if (m) { const float e = m[i] - mmu[i]; acc += (double)(lv + e * e / expf(lv)); }
  }
  const double t = block_sum_d(acc, red);
  if (threadIdx.x == 0 && nll) atomicAdd(nll, 0.5 * t);
}

__global__ void gauss_nll_bwd_kernel(const float* __restrict__ m, const float* __restrict__ mmu,
                                     const float* __restrict__ raw, const float* __restrict__ gscale, float gmul,
                                     float* __restrict__ dmu, float* __restrict__ draw, int64_t n, float lv_clamp) {
  const float g = (gscale ? *gscale : 1.f) * gmul;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float lr = raw[i];
    const bool pass = !(lv_clamp > 0.f) || (lr >= -lv_clamp && lr <= lv_clamp);
    const float lv = lv_clamp > 0.f ? fminf(fmaxf(lr, -lv_clamp), lv_clamp) : lr;
    const float e = m[i] - mmu[i], iv = expf(-lv);
    dmu[i] = -g * e * iv;
    draw[i] = pass ? g * 0.5f * (1.f - e * e * iv) : 0.f;
  }
}

__global__ void kld_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv,
                               const float* __restrict__ gscale, float gmul, float* __restrict__ dmu,
                               float* __restrict__ dlv, int64_t n, int accumulate) {
  const float g = (gscale ? *gscale : 1.f) * gmul;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float a = g * mu[i], b = g * 0.5f * (expf(lv[i]) - 1.f);
    dmu[i] = accumulate ? dmu[i] + a : a;
    dlv[i] = accumulate ? dlv[i] + b : b;
  }
}

__global__ void clamp_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, float lo, float hi) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = fminf(fmaxf(x[i], lo), hi);
}
__global__ void clamp_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dx,
                                 int64_t n, float lo, float hi) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dx[i] = (x[i] >= lo && x[i] <= hi) ? dy[i] : 0.f;
}
__global__ void kld_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv, int64_t n, double* sum) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    acc += (double)(1.0f + lv[i] - mu[i] * mu[i] - expf(lv[i]));
  const double t = block_sum_d(acc, red);
  if (threadIdx.x == 0) atomicAdd(sum, -0.5 * t);
}

// ---- counterfactual helpers --------------------------------------------------------------------
__global__ void do_expand_kernel(const float* __restrict__ m, const float* __restrict__ z, float* __restrict__ out,
                                 int S, int K, int Z, int set_value, float v) {
  const int W = K + Z;
  const int64_t total = (int64_t)S * K * W, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % W);
    const int64_t row = i / W;
    const int k = (int)(row % K);
    const int64_t s = row / K;
    float o;
    if (c < K) {
      o = m[s * K + c];
      if (c == k) o = set_value ? v : o + v;
    } else {
      o = z[s * Z + (c - K)];
    }
    out[i] = o;
  }
}

__global__ void rowdiff_l2_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                                  int64_t rowlen, int group) {
  __shared__ double red[32];
  const int64_t row = blockIdx.x;
  const float* ar = a + row * rowlen;
  const float* br = b + (row / group) * rowlen;
  double acc = 0.0;
  if ((rowlen & 3) == 0) {
    for (int64_t i = threadIdx.x; i < (rowlen >> 2); i += blockDim.x) {
      const float4 u = reinterpret_cast<const float4*>(ar)[i], v = reinterpret_cast<const float4*>(br)[i];
      const float d0 = u.x - v.x, d1 = u.y - v.y, d2 = u.z - v.z, d3 = u.w - v.w;
      acc += (double)(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3);
    }
  } else {
    for (int64_t i = threadIdx.x; i < rowlen; i += blockDim.x) { const float d = ar[i] - br[i]; acc += (double)(d * d); }
  }
  const double t = block_sum_d(acc, red);
  if (threadIdx.x == 0) out[row] = (float)sqrt(t);
}

// torch.stack(preds).mean(0) and .std(0) (unbiased) over the reconstructions of up to 8 fold models
// (vessel_analysis/04_generate_counterfactual/ensemble_reconstruction.py:80-86; check_mechanism_z_perm.py:129): one pass,
// two-pass-in-registers variance (mean first, then squared deviations) so that nearly equal folds do not cancel.
struct EnsemblePtrs { const float* p[8]; };
__global__ void ensemble_mean_std_kernel(const __grid_constant__ EnsemblePtrs e, int n_models, float* __restrict__ mean,
                                         float* __restrict__ stdv, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v[8];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { v[k] = k < n_models ? __ldg(e.p[k] + i) : 0.f; s += v[k]; }
    const float mu = s / (float)n_models;
    mean[i] = mu;
    if (stdv != nullptr) {
      float q = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) { const float d = v[k] - mu; q = k < n_models ? fmaf(d, d, q) : q; }
      stdv[i] = n_models > 1 ? sqrtf(q / (float)(n_models - 1)) : nanf("");
    }
  }
}

// rows (i*N + j) = cat(m[i], scale * z[j]): every M source against every Z source (check_mechanism_z_perm.py:100-118)
__global__ void pair_expand_kernel(const float* __restrict__ m, const float* __restrict__ z, float* __restrict__ out, int N,
                                   int K, int Z, float scale) {
  const int64_t total = (int64_t)N * N * (K + Z);
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % (K + Z));
    const int64_t r = idx / (K + Z);
    const int i = (int)(r / N), j = (int)(r % N);
    out[idx] = c < K ? m[(int64_t)i * K + c] : z[(int64_t)j * Z + (c - K)] * scale;
  }
}

// nn.Upsample(scale_factor=2, mode='nearest') on NHWC (vessel_analysis/00_core/models.py:123-145, the CNN decoder):
// y[n, 2h+a, 2w+b, :] = x[n, h, w, :]; backward sums each 2x2 block.  V = 4 (C % 4 == 0) or 1.
template <int V>
__global__ void upsample2x_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n_in, int H, int W,
                                      int Cv) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_in; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cv);
    const int64_t p = i / Cv;
    const int w = (int)(p % W);
    const int64_t q = p / W;
    const int h = (int)(q % H);
    const int64_t n = q / H;
    const int64_t o = (((n * 2 * H + 2 * h) * 2 * W) + 2 * w) * Cv + c;
    if (V == 4) {
      const float4 v = reinterpret_cast<const float4*>(x)[i];
      float4* yo = reinterpret_cast<float4*>(y);
      yo[o] = v; yo[o + Cv] = v; yo[o + (int64_t)2 * W * Cv] = v; yo[o + (int64_t)2 * W * Cv + Cv] = v;
    } else {
      const float v = x[i];
      y[o] = v; y[o + Cv] = v; y[o + (int64_t)2 * W * Cv] = v; y[o + (int64_t)2 * W * Cv + Cv] = v;
    }
  }
}
template <int V>
__global__ void upsample2x_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int64_t n_in, int H, int W,
                                      int Cv) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_in; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cv);
    const int64_t p = i / Cv;
    const int w = (int)(p % W);
    const int64_t q = p / W;
    const int h = (int)(q % H);
    const int64_t n = q / H;
    const int64_t o = (((n * 2 * H + 2 * h) * 2 * W) + 2 * w) * Cv + c;
    const int64_t r = (int64_t)2 * W * Cv;
    if (V == 4) {
      const float4* g = reinterpret_cast<const float4*>(dy);
      const float4 a = g[o], b = g[o + Cv], cc = g[o + r], d = g[o + r + Cv];
      reinterpret_cast<float4*>(dx)[i] = make_float4((a.x + b.x) + (cc.x + d.x), (a.y + b.y) + (cc.y + d.y),
                                                     (a.z + b.z) + (cc.z + d.z), (a.w + b.w) + (cc.w + d.w));
    } else {
      dx[i] = (dy[o] + dy[o + Cv]) + (dy[o + r] + dy[o + r + Cv]);
    }
  }
}

}  // namespace cvae
using namespace cvae;

#define ST as_stream(s)

extern "C" int cvae_act_fwd(const float* x, float* y, int64_t n, int act, float slope, cvae_stream_t s) {
  if (!x || !y || n <= 0) return CVAE_ERR_BAD_ARG;
  const int b = ew_blocks(n / 4 + 1);
  if (act == CVAE_ACT_LRELU) act_fwd_kernel<CVAE_ACT_LRELU><<<b, 256, 0, ST>>>(x, y, n, slope);
  else if (act == CVAE_ACT_GELU) act_fwd_kernel<CVAE_ACT_GELU><<<b, 256, 0, ST>>>(x, y, n, slope);
  else if (act == CVAE_ACT_SIGMOID) act_fwd_kernel<CVAE_ACT_SIGMOID><<<b, 256, 0, ST>>>(x, y, n, slope);
  else return CVAE_ERR_BAD_ARG;
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_act_bwd(const float* dy, const float* x, float* dx, int64_t n, int act, float slope,
                            cvae_stream_t s) {
  if (!dy || !x || !dx || n <= 0) return CVAE_ERR_BAD_ARG;
  const int b = ew_blocks(n / 4 + 1);
  if (act == CVAE_ACT_LRELU) act_bwd_kernel<CVAE_ACT_LRELU><<<b, 256, 0, ST>>>(dy, x, dx, n, slope);
  else if (act == CVAE_ACT_GELU) act_bwd_kernel<CVAE_ACT_GELU><<<b, 256, 0, ST>>>(dy, x, dx, n, slope);
  else if (act == CVAE_ACT_SIGMOID) act_bwd_kernel<CVAE_ACT_SIGMOID><<<b, 256, 0, ST>>>(dy, x, dx, n, slope);
  else return CVAE_ERR_BAD_ARG;
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_add(const float* a, const float* b, float* out, int64_t n, cvae_stream_t s) {
  if (!a || !b || !out || n <= 0) return CVAE_ERR_BAD_ARG;
  add_kernel<<<ew_blocks(n / 4 + 1), 256, 0, ST>>>(a, b, out, n);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_fill(float* dst, int64_t n, float v, cvae_stream_t s) {
  if (!dst || n <= 0) return CVAE_ERR_BAD_ARG;
  fill_kernel<<<ew_blocks(n), 256, 0, ST>>>(dst, n, v);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_dropout(const float* x, float* y, int64_t n, float p, uint64_t seed, uint64_t offset,
                            const int64_t* counter, cvae_stream_t s) {
  if (!x || !y || n <= 0 || p < 0.f || p >= 1.f) return CVAE_ERR_BAD_ARG;
  if (p == 0.f) {
    if (x != y && cudaMemcpyAsync(y, x, n * sizeof(float), cudaMemcpyDeviceToDevice, ST) != cudaSuccess) return CVAE_ERR_LAUNCH;
    return CVAE_OK;
  }
  dropout_kernel<<<ew_blocks(n / 4 + 1), 256, 0, ST>>>(x, y, n, p, seed, offset, counter);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

// mode 0: y = act(x) * mask / (1-p);  mode 1: y = aux * act'(x) * mask / (1-p);  mode 2: y = aux + x * mask / (1-p)
extern "C" int cvae_dropout_fused(const float* x, const float* aux, float* y, int64_t n, int mode, int act, float slope,
                                  float p, uint64_t seed, uint64_t offset, const int64_t* counter, cvae_stream_t s) {
  if (!x || !y || n <= 0 || p <= 0.f || p >= 1.f || mode < 0 || mode > 2 || (mode != 0 && !aux)) return CVAE_ERR_BAD_ARG;
  if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(aux)) & 15) != 0)
    return CVAE_ERR_ALIGNMENT;
  const int b = ew_blocks(n / 4 + 1);
#define CVAE_DF(M, A) dropout_fused_kernel<M, A><<<b, 256, 0, ST>>>(x, aux, y, n, slope, p, seed, offset, counter)
  if (mode == 2) CVAE_DF(2, CVAE_ACT_LRELU);
  else if (act == CVAE_ACT_GELU) { if (mode == 0) CVAE_DF(0, CVAE_ACT_GELU); else CVAE_DF(1, CVAE_ACT_GELU); }
  else if (act == CVAE_ACT_LRELU) { if (mode == 0) CVAE_DF(0, CVAE_ACT_LRELU); else CVAE_DF(1, CVAE_ACT_LRELU); }
  else if (act == CVAE_ACT_SIGMOID) { if (mode == 0) CVAE_DF(0, CVAE_ACT_SIGMOID); else CVAE_DF(1, CVAE_ACT_SIGMOID); }
  else return CVAE_ERR_BAD_ARG;
#undef CVAE_DF
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_tokens_fwd(const float* feat, const float* cls, const float* pos, float* tok, int B, int n,
                               int D, cvae_stream_t s) {
  if (!feat || !cls || !pos || !tok || B <= 0) return CVAE_ERR_BAD_ARG;
  tokens_fwd_kernel<<<ew_blocks((int64_t)B * (n + 1) * D), 256, 0, ST>>>(feat, cls, pos, tok, B, n, D);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_tokens_bwd(const float* dtok, float* dfeat, float* dcls, float* dpos, int B, int n, int D,
                               cvae_stream_t s) {
  if (!dtok || B <= 0) return CVAE_ERR_BAD_ARG;
  tokens_bwd_kernel<<<ew_blocks((int64_t)(n + 1) * D), 256, 0, ST>>>(dtok, dfeat, dcls, dpos, B, n, D);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_transpose_bc(const float* src, float* dst, int B, int rows, int cols, cvae_stream_t s) {
  if (!src || !dst || B <= 0 || rows <= 0 || cols <= 0 || B > 65535) return CVAE_ERR_BAD_ARG;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, B);
  transpose_bc_kernel<<<grid, dim3(32, 8), 0, ST>>>(src, dst, rows, cols);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_copy_cols(const float* src, int64_t src_ld, int src_col0, float* dst, int64_t dst_ld,
                              int dst_col0, int64_t rows, int w, int accumulate, cvae_stream_t s) {
  if (!src || !dst || rows <= 0 || w <= 0) return CVAE_ERR_BAD_ARG;
  copy_cols_kernel<<<ew_blocks(rows * w), 256, 0, ST>>>(src, src_ld, src_col0, dst, dst_ld, dst_col0, rows, w, accumulate);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_latent_fwd(const float* h, const float* eps, float* mu, float* logvar, float* z,
                               double* kld_sum, int B, int Z, float mu_clamp, float lv_clamp, cvae_stream_t s) {
  if (!h || !mu || !logvar || (z && !eps) || B <= 0 || Z <= 0) return CVAE_ERR_BAD_ARG;
  latent_fwd_kernel<<<ew_blocks((int64_t)B * Z), 256, 0, ST>>>(h, eps, mu, logvar, z, kld_sum, B, Z, mu_clamp, lv_clamp);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_latent_bwd(const float* h, const float* eps, const float* dz, const float* dmu,
                               const float* dlogvar, float* dh, int B, int Z, float mu_clamp, float lv_clamp,
                               cvae_stream_t s) {
  if (!h || !dh || (dz && !eps) || B <= 0 || Z <= 0) return CVAE_ERR_BAD_ARG;
  latent_bwd_kernel<<<ew_blocks((int64_t)B * Z), 256, 0, ST>>>(h, eps, dz, dmu, dlogvar, dh, B, Z, mu_clamp, lv_clamp);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_gauss_nll_fwd(const float* m, const float* m_mu, const float* raw_lv, float* lv_out,
                                  double* nll_sum, int64_t n, float lv_clamp, cvae_stream_t s) {
  if (!raw_lv || n <= 0 || (m && !m_mu)) return CVAE_ERR_BAD_ARG;
  gauss_nll_fwd_kernel<<<ew_blocks(n), 256, 0, ST>>>(m, m_mu, raw_lv, lv_out, nll_sum, n, lv_clamp);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_gauss_nll_bwd(const float* m, const float* m_mu, const float* raw_lv, const float* gscale,
                                  float gmul, float* d_mu, float* d_raw_lv, int64_t n, float lv_clamp,
                                  cvae_stream_t s) {
  if (!m || !m_mu || !raw_lv || !d_mu || !d_raw_lv || n <= 0) return CVAE_ERR_BAD_ARG;
  gauss_nll_bwd_kernel<<<ew_blocks(n), 256, 0, ST>>>(m, m_mu, raw_lv, gscale, gmul, d_mu, d_raw_lv, n, lv_clamp);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_kld_bwd(const float* mu, const float* logvar, const float* gscale, float gmul, float* dmu,
                            float* dlogvar, int64_t n, int accumulate, cvae_stream_t s) {
  if (!mu || !logvar || !dmu || !dlogvar || n <= 0) return CVAE_ERR_BAD_ARG;
  kld_bwd_kernel<<<ew_blocks(n), 256, 0, ST>>>(mu, logvar, gscale, gmul, dmu, dlogvar, n, accumulate);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_upsample2x_fwd(const float* x, float* y, int N, int H, int W, int C, cvae_stream_t s) {
  if (!x || !y || N <= 0 || H <= 0 || W <= 0 || C <= 0) return CVAE_ERR_BAD_ARG;
  const bool v4 = (C & 3) == 0 && (((uintptr_t)x | (uintptr_t)y) & 15) == 0;
  const int Cv = v4 ? C / 4 : C;
  const int64_t n = (int64_t)N * H * W * Cv;
  if (v4) upsample2x_fwd_kernel<4><<<ew_blocks(n), 256, 0, ST>>>(x, y, n, H, W, Cv);
  else upsample2x_fwd_kernel<1><<<ew_blocks(n), 256, 0, ST>>>(x, y, n, H, W, Cv);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}
extern "C" int cvae_upsample2x_bwd(const float* dy, float* dx, int N, int H, int W, int C, cvae_stream_t s) {
  if (!dy || !dx || N <= 0 || H <= 0 || W <= 0 || C <= 0) return CVAE_ERR_BAD_ARG;
  const bool v4 = (C & 3) == 0 && (((uintptr_t)dy | (uintptr_t)dx) & 15) == 0;
  const int Cv = v4 ? C / 4 : C;
  const int64_t n = (int64_t)N * H * W * Cv;
  if (v4) upsample2x_bwd_kernel<4><<<ew_blocks(n), 256, 0, ST>>>(dy, dx, n, H, W, Cv);
  else upsample2x_bwd_kernel<1><<<ew_blocks(n), 256, 0, ST>>>(dy, dx, n, H, W, Cv);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}
extern "C" int cvae_clamp_fwd(const float* x, float* y, int64_t n, float lo, float hi, cvae_stream_t s) {
  if (!x || !y || n <= 0) return CVAE_ERR_BAD_ARG;
  clamp_fwd_kernel<<<ew_blocks(n), 256, 0, ST>>>(x, y, n, lo, hi);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}
extern "C" int cvae_clamp_bwd(const float* dy, const float* x, float* dx, int64_t n, float lo, float hi,
                              cvae_stream_t s) {
  if (!dy || !x || !dx || n <= 0) return CVAE_ERR_BAD_ARG;
  clamp_bwd_kernel<<<ew_blocks(n), 256, 0, ST>>>(dy, x, dx, n, lo, hi);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}
extern "C" int cvae_kld_fwd(const float* mu, const float* logvar, int64_t n, double* sum, cvae_stream_t s) {
  if (!mu || !logvar || !sum || n <= 0) return CVAE_ERR_BAD_ARG;
  kld_fwd_kernel<<<ew_blocks(n), 256, 0, ST>>>(mu, logvar, n, sum);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_do_expand(const float* m, const float* z, float* out, int S, int K, int Z, int set_value,
                              float v, cvae_stream_t s) {
  if (!m || !z || !out || S <= 0 || K <= 0 || Z <= 0) return CVAE_ERR_BAD_ARG;
  do_expand_kernel<<<ew_blocks((int64_t)S * K * (K + Z)), 256, 0, ST>>>(m, z, out, S, K, Z, set_value, v);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_ensemble_mean_std(const float* const* preds, int n_models, float* mean, float* stdv, int64_t n,
                                      cvae_stream_t s) {
  if (!preds || !mean || n_models < 1 || n_models > 8 || n <= 0) return CVAE_ERR_BAD_ARG;
  EnsemblePtrs e;
  for (int k = 0; k < 8; ++k) e.p[k] = k < n_models ? preds[k] : nullptr;
  for (int k = 0; k < n_models; ++k) if (!e.p[k]) return CVAE_ERR_BAD_ARG;
  const int blocks = (int)min((n + 255) / 256, (int64_t)kNumSMs * 16);
  ensemble_mean_std_kernel<<<blocks, 256, 0, ST>>>(e, n_models, mean, stdv, n);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_pair_expand(const float* m, const float* z, float* out, int N, int K, int Z, float scale, cvae_stream_t s) {
  if (!m || !z || !out || N <= 0 || K <= 0 || Z <= 0) return CVAE_ERR_BAD_ARG;
  const int64_t total = (int64_t)N * N * (K + Z);
  const int blocks = (int)min((total + 255) / 256, (int64_t)kNumSMs * 16);
  pair_expand_kernel<<<blocks, 256, 0, ST>>>(m, z, out, N, K, Z, scale);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_rowdiff_l2(const float* a, const float* b, float* out, int64_t rows, int64_t rowlen, int group,
                               cvae_stream_t s) {
  if (!a || !b || !out || rows <= 0 || rowlen <= 0 || group <= 0) return CVAE_ERR_BAD_ARG;
  rowdiff_l2_kernel<<<(unsigned)rows, 256, 0, ST>>>(a, b, out, rowlen, group);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}
