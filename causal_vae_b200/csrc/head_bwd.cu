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
// shared-memory tile.  The global loads of patch i + 1 (four y vectors and the g halo values per thread) are
// issued before patch i is computed, so a block always has one patch of loads in flight.
#include "common.cuh"

namespace cvae {

constexpr int kHbThreads = 256;

template <int C>
__global__ void __launch_bounds__(kHbThreads, 2)
head_bwd_kernel(const float* __restrict__ g, const float* __restrict__ y, const float* __restrict__ x_scale,
                const float* __restrict__ x_shift, const float* __restrict__ x_center, const float slope,
                const float* __restrict__ w, float* __restrict__ dz, double* __restrict__ stats, float* __restrict__ dw,
                const int N, const int H, const int W, const int patches, const int tiles_h, const int tiles_w,
                const int tw_shift, const int th_shift) {
  constexpr int TH = 8, TW = 32, NB = C / 4, PPP = kHbThreads / NB, NIT = TH * TW / PPP;
  constexpr int GR = TH + 2, GC = TW + 2, GN = GR * GC, GIT = (GN + kHbThreads - 1) / kHbThreads;
  __shared__ float sG[GN];
  __shared__ float4 sW[9 * NB];                        // [tap][channel group]: broadcast reads (the weights used to
  __shared__ double s_red[kHbThreads][8];              // occupy 36 registers; they now pay for the patch prefetch)
  const int tid = threadIdx.x, c4 = tid % NB, c0 = c4 * 4, pp = tid / NB;
  if (tid < 9 * NB) {
    const int t = tid / NB, cc = (tid % NB) * 4;
    sW[tid] = make_float4(__ldg(w + (cc + 0) * 9 + t), __ldg(w + (cc + 1) * 9 + t), __ldg(w + (cc + 2) * 9 + t), __ldg(w + (cc + 3) * 9 + t));
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
  // fp64 running sums live in this thread's shared-memory slot, folded every fourth patch (16 fp32 terms per fold)
#pragma unroll
  for (int j = 0; j < 8; ++j) s_red[tid][j] = 0.0;

  // patch -> (image, tile origin); power-of-two tile grids (256 x 256: 32 x 8 tiles) avoid the runtime divisions
  auto decode = [&](int patch, int& n, int& h0, int& w0) {
    int tw, tt, th;
    if (tw_shift >= 0 && th_shift >= 0) { tw = patch & (tiles_w - 1); tt = patch >> tw_shift; th = tt & (tiles_h - 1); n = tt >> th_shift; }
    else { tw = patch % tiles_w; tt = patch / tiles_w; th = tt % tiles_h; n = tt / tiles_h; }
    h0 = th * TH; w0 = tw * TW;
  };
  // every global load of a patch (its four y vectors and its share of the g halo tile) in one batch
  auto fetch = [&](int patch, float4 (&yv)[NIT], float (&gv)[GIT]) {
    int n, h0, w0;
    decode(patch, n, h0, w0);
#pragma unroll
    for (int u = 0; u < GIT; ++u) {
      const int idx = tid + u * kHbThreads;
      const int gi = idx / GC, gj = idx - gi * GC, ih = h0 - 1 + gi, iw = w0 - 1 + gj;
      gv[u] = 0.f;
      if (idx < GN && (unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W) gv[u] = __ldg(g + ((size_t)n * H + ih) * W + iw);
    }
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      const int p = pp + i * PPP, r = p / TW, c = p - r * TW, qh = h0 + r, qw = w0 + c;
      yv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (qh < H && qw < W) yv[i] = __ldg(reinterpret_cast<const float4*>(y + ((n * H + qh) * W + qw) * C + c0));   // N*H*W*C < 2^31
    }
  };

  float4 yv[NIT];
  float gv[GIT];
  int patch = blockIdx.x, since = 0;
  if (patch < patches) fetch(patch, yv, gv);
  for (; patch < patches; patch += gridDim.x) {
    int n, h0, w0;
    decode(patch, n, h0, w0);
#pragma unroll
    for (int u = 0; u < GIT; ++u) {
      const int idx = tid + u * kHbThreads;
      if (idx < GN) sG[idx] = gv[u];
    }
    __syncthreads();
    // the next patch's loads are in flight while this one is computed
    float4 yn[NIT];
    float gn[GIT];
    const int next = patch + gridDim.x;
    if (next < patches) fetch(next, yn, gn);
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      const int p = pp + i * PPP, r = p / TW, c = p - r * TW, qh = h0 + r, qw = w0 + c;
      if (qh >= H || qw >= W) continue;
      const float* gp = sG + r * GC + c;
      const float rf[4] = {yv[i].x - ce.x, yv[i].y - ce.y, yv[i].z - ce.z, yv[i].w - ce.w};
      const float z[4] = {fmaf(rf[0], sc.x, sh.x), fmaf(rf[1], sc.y, sh.y), fmaf(rf[2], sc.z, sh.z), fmaf(rf[3], sc.w, sh.w)};
      float av[4], o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 4; ++j) av[j] = z[j] > 0.f ? z[j] : z[j] * slope;
      const float4 av4 = make_float4(av[0], av[1], av[2], av[3]);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float gt = gp[(2 - t / 3) * GC + (2 - t % 3)];      // G_t(q) = g[q - (kh - 1, kw - 1)] in halo-tile coordinates
        const float4 wt = sW[t * NB + c4];
        fma4(o, gt, wt);                                          // o[0..3] += gt * w_t[0..3]    (packed FMAs: common.cuh)
        fma4(acc[t], gt, av4);                                    // acc[t][0..3] += a[0..3] * gt
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        o[j] = z[j] > 0.f ? o[j] : o[j] * slope;
        f1[j] += o[j]; f2[j] = fmaf(o[j], rf[j], f2[j]);
      }
      *reinterpret_cast<float4*>(dz + ((n * H + qh) * W + qw) * C + c0) = make_float4(o[0], o[1], o[2], o[3]);
    }
    if (++since == 4) {
      since = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) { s_red[tid][j] += (double)f1[j]; s_red[tid][4 + j] += (double)f2[j]; f1[j] = 0.f; f2[j] = 0.f; }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NIT; ++i) yv[i] = yn[i];
#pragma unroll
    for (int u = 0; u < GIT; ++u) gv[u] = gn[u];
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) { s_red[tid][j] += (double)f1[j]; s_red[tid][4 + j] += (double)f2[j]; }
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
  // two resident blocks per SM (128 registers): one wave, every block walks its share of the patches
  const int grid = patches < kNumSMs * 2 ? patches : kNumSMs * 2;
  auto log2_or_neg = [](int v) { int s = 0; while ((1 << s) < v) ++s; return (1 << s) == v ? s : -1; };
  head_bwd_kernel<16><<<grid, kHbThreads, 0, st>>>(g, y, x.scale, x.shift, x.center, x.slope, w, dz, stats, dw, N, H, W,
                                                   patches, tiles_h, tiles_w, log2_or_neg(tiles_w), log2_or_neg(tiles_h));
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}
