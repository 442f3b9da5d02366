// ELBO reconstruction terms and the fused clip + Adam optimizer step (HBM-bound streams; double
// accumulators so the 4M-pixel sums hold the 1e-5 loss tolerance).
#include "common.cuh"

namespace cvae {

static inline int red_blocks(int64_t items) {
  int64_t b = (items + 1023) / 1024;
  if (b > kNumSMs * 8) b = kNumSMs * 8;
  if (b < 1) b = 1;
  return (int)b;
}

__global__ void xsum_kernel(const float* __restrict__ x, int64_t n, double* sums) {
  __shared__ double red[32];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, n4 = n >> 2;
  float acc = 0.f;
  double dacc = 0.0;
  int cnt = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    acc += (v.x + v.y) + (v.z + v.w);
    if (++cnt == 64) { dacc += (double)acc; acc = 0.f; cnt = 0; }
  }
  for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) acc += x[i];
  dacc += (double)acc;
  const double t = block_sum_d(dacc, red);
  if (threadIdx.x == 0) atomicAdd(sums, t);
}

__device__ __forceinline__ float vessel_pos_weight(const double* sums, int64_t n) {
  // train.py:30-36 evaluated in fp32 like the reference: x.sum() (exact integer for binary x),
  // n_total + 1e-6 (== n_total in fp32), clamp to [1, 50].
  const float n_pos = (float)sums[0];
  const float n_total = (float)n + 1e-6f;
  const float pf = n_pos / n_total;
  const float w = (1.0f - pf) / (pf + 1e-6f);
  return fminf(fmaxf(w, 1.0f), 50.0f);
}

__global__ void vessel_recon_fwd_kernel(const float* __restrict__ r, const float* __restrict__ x, int64_t n,
                                        double* sums) {
  __shared__ double red[32];
  const float pw = vessel_pos_weight(sums, n) - 1.0f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, n4 = n >> 2;
  double a_rec = 0.0, a_sp = 0.0;
  auto term = [&](float rv, float xv, float& rec, float& sp) {
    const float d = rv - xv;
    rec = fmaf(d * d, fmaf(pw, xv, 1.0f), rec);
    sp += xv < 0.1f ? fabsf(rv) : 0.f;
  };
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 rv = reinterpret_cast<const float4*>(r)[i], xv = reinterpret_cast<const float4*>(x)[i];
    float rec = 0.f, sp = 0.f;
    term(rv.x, xv.x, rec, sp); term(rv.y, xv.y, rec, sp); term(rv.z, xv.z, rec, sp); term(rv.w, xv.w, rec, sp);
    a_rec += (double)rec; a_sp += (double)sp;
  }
  for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    float rec = 0.f, sp = 0.f;
    term(r[i], x[i], rec, sp);
    a_rec += (double)rec; a_sp += (double)sp;
  }
  const double t1 = block_sum_d(a_rec, red);
  const double t2 = block_sum_d(a_sp, red);
  if (threadIdx.x == 0) { atomicAdd(sums + 1, t1); atomicAdd(sums + 2, t2); }
}

__global__ void vessel_recon_bwd_kernel(const float* __restrict__ r, const float* __restrict__ x, int64_t n,
                                        const double* sums, const float* g_recon, const float* g_sp,
                                        float* __restrict__ dr) {
  const float pw = vessel_pos_weight(sums, n) - 1.0f;
  const float gr = 2.0f * (g_recon ? *g_recon : 1.f), gs = g_sp ? *g_sp : 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, n4 = n >> 2;
  auto f = [&](float rv, float xv) {
    const float sgn = rv > 0.f ? 1.f : (rv < 0.f ? -1.f : 0.f);
    return fmaf(gr * (rv - xv), fmaf(pw, xv, 1.0f), xv < 0.1f ? gs * sgn : 0.f);
  };
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 rv = reinterpret_cast<const float4*>(r)[i], xv = reinterpret_cast<const float4*>(x)[i];
    reinterpret_cast<float4*>(dr)[i] = make_float4(f(rv.x, xv.x), f(rv.y, xv.y), f(rv.z, xv.z), f(rv.w, xv.w));
  }
  for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) dr[i] = f(r[i], x[i]);
}

template <int KIND>  // 0 = squared error, 1 = BCE with log clamp -100
__global__ void pair_loss_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, double* sum) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (KIND == 0) {
      const float d = a[i] - b[i];
      acc += (double)(d * d);
    } else {
      const float p = a[i], y = b[i];
      const float l1 = fmaxf(logf(p), -100.f), l0 = fmaxf(log1pf(-p), -100.f);
      acc -= (double)(y * l1 + (1.f - y) * l0);
    }
  }
  const double t = block_sum_d(acc, red);
  if (threadIdx.x == 0) atomicAdd(sum, t);
}

template <int KIND>
__global__ void pair_loss_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                     const float* g, float gmul, float* __restrict__ da) {
  const float gg = (g ? *g : 1.f) * gmul;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (KIND == 0) {
      da[i] = 2.f * gg * (a[i] - b[i]);
    } else {
      // d/dp of -(y*max(log p,-100) + (1-y)*max(log(1-p),-100)); the clamped branch has zero slope
      const float p = a[i], y = b[i];
      const float t1 = logf(p) > -100.f ? y / p : 0.f;
      const float t0 = log1pf(-p) > -100.f ? (1.f - y) / (1.f - p) : 0.f;
      da[i] = gg * (t0 - t1);
    }
  }
}

__global__ void finish_scalar_kernel(const double* acc, float mul, float* out) { *out = (float)(*acc * (double)mul); }

// out = sum_i w[i] * x_i for up to 4 device scalars (total loss = recon + beta*kld + lambda*morph + 0.3*sparsity,
// vessel_analysis/01_train/train.py:82) and its backward g -> (w[i] * g): two launches instead of ~12 ATen ones
__global__ void scalar_combine_kernel(const float* a, const float* b, const float* c, const float* d, float wa, float wb,
                                      float wc, float wd, float* out) {
  float v = wa * *a;
  if (b) v += wb * *b;
  if (c) v += wc * *c;
  if (d) v += wd * *d;
  *out = v;
}
__global__ void scalar_scale4_kernel(const float* g, float wa, float wb, float wc, float wd, float* out4) {
  const float gv = *g;
  out4[0] = wa * gv; out4[1] = wb * gv; out4[2] = wc * gv; out4[3] = wd * gv;
}

// ---- optimizer ---------------------------------------------------------------------------------
__global__ void sumsq_kernel(const float* __restrict__ g, int64_t n, double* acc) {
  __shared__ double red[32];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, n4 = n >> 2;
  double d = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    d += (double)(v.x * v.x + v.y * v.y) + (double)(v.z * v.z + v.w * v.w);
  }
  for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) d += (double)(g[i] * g[i]);
  const double t = block_sum_d(d, red);
  if (threadIdx.x == 0) atomicAdd(acc, t);
}

__global__ void step_inc_kernel(int64_t* step) { *step += 1; }

__global__ void clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, int64_t n, const double* sumsq, float max_norm, float lr,
                                 float b1, float b2, float eps, float gscale, const int64_t* step) {
  // clip_grad_norm_: coef = min(1, max_norm / (norm + 1e-6)); the norm is of the (scaled) gradient
  float coef = gscale;
  if (max_norm > 0.f) {
    const float norm = (float)sqrt(*sumsq) * gscale;
    coef *= fminf(max_norm / (norm + 1e-6f), 1.0f);
  }
  const double t = (double)(*step);
  const float bc1 = (float)(1.0 - pow((double)b1, t));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, t));
  const float step_size = lr / bc1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, n4 = n >> 2;
  auto upd = [&](float& pw, float gw, float& mw, float& vw) {
    const float gc = gw * coef;
    mw = mw + (1.f - b1) * (gc - mw);            // torch: exp_avg.lerp_(grad, 1 - beta1)
    vw = vw * b2 + (1.f - b2) * gc * gc;         // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
    const float denom = sqrtf(vw) / bc2_sqrt + eps;
    pw = pw - step_size * (mw / denom);
  };
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pw = reinterpret_cast<float4*>(p)[i];
    const float4 gw = reinterpret_cast<const float4*>(g)[i];
    float4 mw = reinterpret_cast<float4*>(m)[i], vw = reinterpret_cast<float4*>(v)[i];
    upd(pw.x, gw.x, mw.x, vw.x); upd(pw.y, gw.y, mw.y, vw.y); upd(pw.z, gw.z, mw.z, vw.z); upd(pw.w, gw.w, mw.w, vw.w);
    reinterpret_cast<float4*>(p)[i] = pw; reinterpret_cast<float4*>(m)[i] = mw; reinterpret_cast<float4*>(v)[i] = vw;
  }
  for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride)
    upd(p[i], g[i], m[i], v[i]);
}

}  // namespace cvae
using namespace cvae;
#define ST as_stream(s)

extern "C" int cvae_vessel_xsum(const float* x, int64_t n, double* sums, cvae_stream_t s) {
  if (!x || !sums || n <= 0) return CVAE_ERR_BAD_ARG;
  xsum_kernel<<<red_blocks(n / 4 + 1), 256, 0, ST>>>(x, n, sums);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}
extern "C" int cvae_vessel_recon_fwd(const float* recon, const float* x, int64_t n, double* sums, cvae_stream_t s) {
  if (!recon || !x || !sums || n <= 0) return CVAE_ERR_BAD_ARG;
  vessel_recon_fwd_kernel<<<red_blocks(n / 4 + 1), 256, 0, ST>>>(recon, x, n, sums);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}
extern "C" int cvae_vessel_recon_bwd(const float* recon, const float* x, int64_t n, const double* sums,
                                     const float* g_recon, const float* g_sparsity, float* d_recon, cvae_stream_t s) {
  if (!recon || !x || !sums || !d_recon || n <= 0) return CVAE_ERR_BAD_ARG;
  vessel_recon_bwd_kernel<<<red_blocks(n / 4 + 1), 256, 0, ST>>>(recon, x, n, sums, g_recon, g_sparsity, d_recon);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}
extern "C" int cvae_mse_fwd(const float* a, const float* b, int64_t n, double* sum, cvae_stream_t s) {
  if (!a || !b || !sum || n <= 0) return CVAE_ERR_BAD_ARG;
  pair_loss_fwd_kernel<0><<<red_blocks(n), 256, 0, ST>>>(a, b, n, sum);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}
extern "C" int cvae_mse_bwd(const float* a, const float* b, int64_t n, const float* g, float gmul, float* da,
                            cvae_stream_t s) {
  if (!a || !b || !da || n <= 0) return CVAE_ERR_BAD_ARG;
  pair_loss_bwd_kernel<0><<<red_blocks(n), 256, 0, ST>>>(a, b, n, g, gmul, da);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}
extern "C" int cvae_bce_fwd(const float* p, const float* y, int64_t n, double* sum, cvae_stream_t s) {
  if (!p || !y || !sum || n <= 0) return CVAE_ERR_BAD_ARG;
  pair_loss_fwd_kernel<1><<<red_blocks(n), 256, 0, ST>>>(p, y, n, sum);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}
extern "C" int cvae_bce_bwd(const float* p, const float* y, int64_t n, const float* g, float gmul, float* dp,
                            cvae_stream_t s) {
  if (!p || !y || !dp || n <= 0) return CVAE_ERR_BAD_ARG;
  pair_loss_bwd_kernel<1><<<red_blocks(n), 256, 0, ST>>>(p, y, n, g, gmul, dp);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}
extern "C" int cvae_finish_scalar(const double* acc, float mul, float* out, cvae_stream_t s) {
  if (!acc || !out) return CVAE_ERR_BAD_ARG;
  finish_scalar_kernel<<<1, 1, 0, ST>>>(acc, mul, out);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}
extern "C" int cvae_scalar_combine(const float* a, const float* b, const float* c, const float* d, float wa, float wb,
                                   float wc, float wd, float* out, cvae_stream_t s) {
  if (!a || !out) return CVAE_ERR_BAD_ARG;
  scalar_combine_kernel<<<1, 1, 0, ST>>>(a, b, c, d, wa, wb, wc, wd, out);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}
extern "C" int cvae_scalar_scale4(const float* g, float wa, float wb, float wc, float wd, float* out4, cvae_stream_t s) {
  if (!g || !out4) return CVAE_ERR_BAD_ARG;
  scalar_scale4_kernel<<<1, 1, 0, ST>>>(g, wa, wb, wc, wd, out4);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}
extern "C" int cvae_sumsq(const float* g, int64_t n, double* acc, cvae_stream_t s) {
  if (!g || !acc || n <= 0) return CVAE_ERR_BAD_ARG;
  sumsq_kernel<<<red_blocks(n / 4 + 1), 256, 0, ST>>>(g, n, acc);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}
extern "C" int cvae_clip_adam(float* p, const float* g, float* m, float* v, int64_t n, const double* sumsq,
                              float max_norm, float lr, float beta1, float beta2, float eps, float grad_scale,
                              int64_t* step_count, cvae_stream_t s) {
  if (!p || !g || !m || !v || !step_count || n <= 0 || (max_norm > 0.f && !sumsq)) return CVAE_ERR_BAD_ARG;
  step_inc_kernel<<<1, 1, 0, ST>>>(step_count);
  clip_adam_kernel<<<red_blocks(n / 4 + 1) * 2, 256, 0, ST>>>(p, g, m, v, n, sumsq, max_norm, lr, beta1, beta2, eps,
                                                             grad_scale, step_count);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_version(void) { return 100; }
extern "C" int cvae_built_arch(void) { return 100; }
