// Backward of the image head (Conv2d C -> 1, 3x3, pad 1; vit_backbone.py:152-156) in ONE pass over its
// input: weight gradient and input gradient together.
//
//   forward:   out[p]   = bias + sum_{t,c} a[p + d_t][c] * w[c][t],      a = lrelu(BN(y))  (y = raw producer output)
//   backward:  G_t(q)   = g[q - d_t]                                      (g = dL/dout, one channel)
//              da[q][c] = sum_t G_t(q) * w[c][t];   dz[q][c] = da[q][c] * lrelu'(z[q][c])   (+ BN-backward sums)
//              dW[c][t] = sum_q a[q][c] * G_t(q)
//
// Both gradients need the same two things per pixel: the 16-channel vector y[q] (64 bytes) and the 3x3
// neighbourhood of g.  As separate kernels (wgrad_tile<16,1,1> + conv_cs1_tile<16,1>) the 268 MB tensor y was
// streamed twice (227 + 257 us at B = 64); here it is read once, dz is written once, and g lives in a
// shared-memory tile.  The loads of a patch's four y vectors are issued before the tile barrier.
#include "common.cuh"

namespace cvae {

constexpr int kHbThreads = 256;

template <int C>
__global__ void __launch_bounds__(kHbThreads, 2)
head_bwd_kernel(const float* __restrict__ g, const float* __restrict__ y, const float* __restrict__ x_scale,
                const float* __restrict__ x_shift, const float* __restrict__ x_center, const float slope,
                const float* __restrict__ w, float* __restrict__ dz, double* __restrict__ stats, float* __restrict__ dw,
                const int N, const int H, const int W, const int patches, const int tiles_h, const int tiles_w) {
  constexpr int TH = 8, TW = 32, NB = C / 4, PPP = kHbThreads / NB, NIT = TH * TW / PPP;
  constexpr int GR = TH + 2, GC = TW + 2;
  __shared__ float sG[GR * GC];
  __shared__ double s_red[kHbThreads][8];
  const int tid = threadIdx.x, c4 = tid % NB, c0 = c4 * 4, pp = tid / NB;
  float4 wv[9];
  int toff[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    wv[t] = make_float4(__ldg(w + (c0 + 0) * 9 + t), __ldg(w + (c0 + 1) * 9 + t), __ldg(w + (c0 + 2) * 9 + t), __ldg(w + (c0 + 3) * 9 + t));
    toff[t] = (2 - t / 3) * GC + (2 - t % 3);          // G_t(q) = g[q - (kh - 1, kw - 1)] in halo-tile coordinates
  }
  float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f), ce = sh;
  if (x_scale != nullptr) {
    sc = __ldg(reinterpret_cast<const float4*>(x_scale + c0));
    sh = __ldg(reinterpret_cast<const float4*>(x_shift + c0));
    if (x_center != nullptr) ce = __ldg(reinterpret_cast<const float4*>(x_center + c0));
  }
  float acc[9][4];
#pragma unroll
  for (int t = 0; t < 9; ++t) { acc[t][0] = 0.f; acc[t][1] = 0.f; acc[t][2] = 0.f; acc[t][3] = 0.f; }
  float f1[4] = {0.f, 0.f, 0.f, 0.f}, f2[4] = {0.f, 0.f, 0.f, 0.f};
  // fp64 running sums live in this thread's shared-memory slot (folded once per patch), not in registers:
  // the FMA loop already holds 36 weights + 36 weight-gradient accumulators per thread
#pragma unroll
  for (int j = 0; j < 8; ++j) s_red[tid][j] = 0.0;

  for (int patch = blockIdx.x; patch < patches; patch += gridDim.x) {
    const int tw = patch % tiles_w, tt = patch / tiles_w, th = tt % tiles_h, n = tt / tiles_h;
    const int h0 = th * TH, w0 = tw * TW;
    for (int idx = tid; idx < GR * GC; idx += kHbThreads) {
      const int gi = idx / GC, gj = idx - gi * GC, ih = h0 - 1 + gi, iw = w0 - 1 + gj;
      float v = 0.f;
      if ((unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W) v = __ldg(g + ((size_t)n * H + ih) * W + iw);
      sG[idx] = v;
    }
    float4 yv[NIT];
    int off[NIT];                                      // N*H*W*C < 2^31 (checked by the launcher)
    bool ok[NIT];
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      const int p = pp + i * PPP, r = p / TW, c = p - r * TW, qh = h0 + r, qw = w0 + c;
      ok[i] = qh < H && qw < W;
      off[i] = ((n * H + qh) * W + qw) * C + c0;
      yv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok[i]) yv[i] = __ldg(reinterpret_cast<const float4*>(y + off[i]));
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      if (!ok[i]) continue;
      const int p = pp + i * PPP, r = p / TW, c = p - r * TW;
      const float* gp = sG + r * GC + c;
      const float rf[4] = {yv[i].x - ce.x, yv[i].y - ce.y, yv[i].z - ce.z, yv[i].w - ce.w};
      const float z[4] = {fmaf(rf[0], sc.x, sh.x), fmaf(rf[1], sc.y, sh.y), fmaf(rf[2], sc.z, sh.z), fmaf(rf[3], sc.w, sh.w)};
      float av[4], o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 4; ++j) av[j] = z[j] > 0.f ? z[j] : z[j] * slope;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float gt = gp[toff[t]];
        o[0] = fmaf(gt, wv[t].x, o[0]); o[1] = fmaf(gt, wv[t].y, o[1]); o[2] = fmaf(gt, wv[t].z, o[2]); o[3] = fmaf(gt, wv[t].w, o[3]);
        acc[t][0] = fmaf(av[0], gt, acc[t][0]); acc[t][1] = fmaf(av[1], gt, acc[t][1]);
        acc[t][2] = fmaf(av[2], gt, acc[t][2]); acc[t][3] = fmaf(av[3], gt, acc[t][3]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        o[j] = z[j] > 0.f ? o[j] : o[j] * slope;
        f1[j] += o[j]; f2[j] = fmaf(o[j], rf[j], f2[j]);
      }
      *reinterpret_cast<float4*>(dz + off[i]) = make_float4(o[0], o[1], o[2], o[3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { s_red[tid][j] += (double)f1[j]; s_red[tid][4 + j] += (double)f2[j]; f1[j] = 0.f; f2[j] = 0.f; }
    __syncthreads();
  }
  // ---- BN-backward sums: threads with the same channel group -> one atomic per sum per block ----
  if (stats != nullptr) {
    __syncthreads();
    if (tid < NB * 8) {
      const int grp = tid / 8, k = tid % 8;
      double t = 0.0;
      for (int r = grp; r < kHbThreads; r += NB) t += s_red[r][k];
      atomicAdd(stats + (k < 4 ? 0 : C) + grp * 4 + (k & 3), t);
    }
    __syncthreads();
  }
  // ---- weight gradient: block sum per (tap, channel), one fp32 atomic each into dw[c][t] ----
  float* fr = reinterpret_cast<float*>(&s_red[0][0]);       // [256][4] floats fit in the double buffer
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    fr[tid * 4 + 0] = acc[t][0]; fr[tid * 4 + 1] = acc[t][1]; fr[tid * 4 + 2] = acc[t][2]; fr[tid * 4 + 3] = acc[t][3];
    __syncthreads();
    if (tid < C) {
      const int grp = tid >> 2, u = tid & 3;
      float s = 0.f;
      for (int r = grp; r < kHbThreads; r += NB) s += fr[r * 4 + u];
      atomicAdd(dw + (size_t)tid * 9 + t, s);
    }
    __syncthreads();
  }
}

}  // namespace cvae
using namespace cvae;

extern "C" int cvae_head_bwd_eligible(int N, int H, int W, int C) {
  return (C == 16 && (long long)N * H * W >= 65536 && (long long)N * H * W * C < (1ll << 31)) ? 1 : 0;
}

extern "C" int cvae_head_bwd(const float* g, const float* y, cvae_xform_t x, const float* w, float* dz, double* stats,
                             float* dw, int N, int H, int W, int C, cvae_stream_t s) {
  if (!g || !y || !w || !dz || !dw || N <= 0 || H <= 0 || W <= 0) return CVAE_ERR_BAD_ARG;
  if (!cvae_head_bwd_eligible(N, H, W, C)) return CVAE_ERR_UNSUPPORTED_SHAPE;
  if (x.scale != nullptr && x.shift == nullptr) return CVAE_ERR_BAD_ARG;
  cudaStream_t st = as_stream(s);
  if (cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)C * 9, st) != cudaSuccess) return CVAE_ERR_LAUNCH;
  const int tiles_h = (H + 7) / 8, tiles_w = (W + 31) / 32;
  const int patches = N * tiles_h * tiles_w;
  const int grid = patches < kNumSMs * 4 ? patches : kNumSMs * 4;
  head_bwd_kernel<16><<<grid, kHbThreads, 0, st>>>(g, y, x.scale, x.shift, x.center, x.slope, w, dz, stats, dw, N, H, W,
                                                   patches, tiles_h, tiles_w);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}
