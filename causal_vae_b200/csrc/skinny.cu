// One-channel layers of the conv stacks as coalesced HBM streams (sm_100a).
//
// The image head (16 -> 1 at 256x256), the stem head (1 -> 32), their input/weight gradients and the
// MNIST / cascade 1-channel layers have GEMM K or N of 1: a tensor-core (or even a tiled SIMT) GEMM
// wastes > 90 % of its lanes on them, while their traffic is large (4 MB/sample for the 16-channel
// 256x256 activation).  Here a group of TP = C/4 threads owns one pixel, every thread moves one
// 128-bit vector per tap, weights live in shared memory as float4 rows, and all per-channel sums stay
// in registers across the grid-stride loop.
#include <cstdlib>
#include "conv_args.cuh"

namespace cvae {

constexpr int kThreads = 256;

__device__ __forceinline__ float4 xform4(float4 v, const float4& sc, const float4& sh, const float4& ce, bool affine,
                                         bool act, float slope) {
  if (affine) {
    v.x = fmaf(v.x - ce.x, sc.x, sh.x); v.y = fmaf(v.y - ce.y, sc.y, sh.y);
    v.z = fmaf(v.z - ce.z, sc.z, sh.z); v.w = fmaf(v.w - ce.w, sc.w, sh.w);
  }
  if (act) { v.x = lrelu(v.x, slope); v.y = lrelu(v.y, slope); v.z = lrelu(v.z, slope); v.w = lrelu(v.w, slope); }
  return v;
}

// ---- Cs == 1 -> Cd (multiple of 4, <= 64): stem head forward, image-head input gradient ---------------
__global__ void __launch_bounds__(kThreads) conv_cs1_kernel(const __grid_constant__ GatherArgs a) {
  __shared__ __align__(16) float s_w[16 * 64];         // [widx][Cd]
  __shared__ double s_red[kThreads][8];
  const int tid = threadIdx.x, Cd = a.Cd, TP = Cd >> 2;
  for (int i = tid; i < a.wtaps * Cd; i += kThreads) s_w[i] = __ldg(a.wt + i);
  __syncthreads();
  const PhaseGeom& P = a.phase[blockIdx.z];
  const int M = a.N * P.Hq * P.Wq;
  const int cg = tid % TP, c0 = cg * 4;
  const float in_sc = a.in_affine ? __ldg(a.in_scale) : 1.f, in_sh = a.in_affine ? __ldg(a.in_shift) : 0.f;
  const float in_ce = (a.in_affine && a.in_center) ? __ldg(a.in_center) : 0.f;
  float4 bias = make_float4(0.f, 0.f, 0.f, 0.f), esc = make_float4(1.f, 1.f, 1.f, 1.f), esh = bias, ece = bias;
  if (a.bias) bias = __ldg(reinterpret_cast<const float4*>(a.bias + c0));
  if (a.e_affine) {
    esc = __ldg(reinterpret_cast<const float4*>(a.e_scale + c0));
    esh = __ldg(reinterpret_cast<const float4*>(a.e_shift + c0));
    if (a.e_center) ece = __ldg(reinterpret_cast<const float4*>(a.e_center + c0));
  }
  float f1[4] = {0.f, 0.f, 0.f, 0.f}, f2[4] = {0.f, 0.f, 0.f, 0.f};
  double d1[4] = {0.0, 0.0, 0.0, 0.0}, d2[4] = {0.0, 0.0, 0.0, 0.0};
  int since = 0;
  const int ppb = kThreads / TP;                        // pixels per block per iteration
  for (int m = blockIdx.x * ppb + tid / TP; m < M; m += gridDim.x * ppb) {
    const int qw = m % P.Wq;
    const int t = m / P.Wq;
    const int qh = t % P.Hq, n = t / P.Hq;
    float4 acc = bias;
    for (int tp = 0; tp < P.ntaps; ++tp) {
      const TapEntry tap = P.taps[tp];
      const int ih = qh * a.is + tap.dh, iw = qw * a.is + tap.dw;
      if (ih < 0 || ih >= a.Hs || iw < 0 || iw >= a.Ws) continue;
      float v = __ldg(a.src + ((size_t)n * a.Hs + ih) * a.Ws + iw);
      if (a.in_affine) v = fmaf(v - in_ce, in_sc, in_sh);
      if (a.in_act) v = lrelu(v, a.in_slope);
      const float4 w = *reinterpret_cast<const float4*>(&s_w[tap.widx * Cd + c0]);
      acc.x = fmaf(v, w.x, acc.x); acc.y = fmaf(v, w.y, acc.y); acc.z = fmaf(v, w.z, acc.z); acc.w = fmaf(v, w.w, acc.w);
    }
    const int oh = qh * a.os + P.ph, ow = qw * a.os + P.pw;
    const size_t off = (((size_t)n * a.Hd + oh) * a.Wd + ow) * Cd + c0;
    float o[4] = {acc.x, acc.y, acc.z, acc.w};
    if (a.epi == CVAE_EPI_STATS) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { f1[j] += o[j]; f2[j] = fmaf(o[j], o[j], f2[j]); }
    } else if (a.epi == CVAE_EPI_DACT) {
      const float4 r4 = __ldg(reinterpret_cast<const float4*>(a.epi_ref + off));
      float rf[4] = {r4.x - ece.x, r4.y - ece.y, r4.z - ece.z, r4.w - ece.w};
      const float sc[4] = {esc.x, esc.y, esc.z, esc.w}, sh[4] = {esh.x, esh.y, esh.z, esh.w};
      if (a.epi_add) {
        const float4 ad = __ldg(reinterpret_cast<const float4*>(a.epi_add + off));
        o[0] += ad.x; o[1] += ad.y; o[2] += ad.z; o[3] += ad.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float z = fmaf(rf[j], sc[j], sh[j]);
        o[j] = z > 0.f ? o[j] : o[j] * a.e_slope;
        f1[j] += o[j]; f2[j] = fmaf(o[j], rf[j], f2[j]);
      }
    }
    *reinterpret_cast<float4*>(a.dst + off) = make_float4(o[0], o[1], o[2], o[3]);
    if (a.epi != CVAE_EPI_PLAIN && ++since == 4) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { d1[j] += (double)f1[j]; d2[j] += (double)f2[j]; f1[j] = 0.f; f2[j] = 0.f; }
      since = 0;
    }
  }
  if (a.epi != CVAE_EPI_PLAIN && a.stats != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { s_red[tid][j] = d1[j] + (double)f1[j]; s_red[tid][4 + j] = d2[j] + (double)f2[j]; }
    __syncthreads();
    if (tid < TP * 8) {
      const int g = tid / 8, k = tid % 8;              // channel group, which of the 8 sums
      double t = 0.0;
      for (int r = g; r < kThreads; r += TP) t += s_red[r][k];
      atomicAdd(a.stats + (k < 4 ? 0 : Cd) + g * 4 + (k & 3), t);
    }
  }
}

// ---- Cs (multiple of 4, <= 64) -> Cd == 1, plain epilogue: image head forward ---------------------------
__global__ void __launch_bounds__(kThreads) conv_cd1_kernel(const __grid_constant__ GatherArgs a) {
  __shared__ __align__(16) float s_w[16 * 64];         // [widx][Cs]
  const int tid = threadIdx.x, Cs = a.Cs, TP = Cs >> 2;
  for (int i = tid; i < a.wtaps * Cs; i += kThreads) s_w[i] = __ldg(a.wt + i);
  __syncthreads();
  const PhaseGeom& P = a.phase[blockIdx.z];
  const int M = a.N * P.Hq * P.Wq;
  const int cg = tid % TP, c0 = cg * 4;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = zero, ce = zero;
  if (a.in_affine) {
    sc = __ldg(reinterpret_cast<const float4*>(a.in_scale + c0));
    sh = __ldg(reinterpret_cast<const float4*>(a.in_shift + c0));
    if (a.in_center) ce = __ldg(reinterpret_cast<const float4*>(a.in_center + c0));
  }
  const float bias = a.bias ? __ldg(a.bias) : 0.f;
  const int ppb = kThreads / TP;
  const int Mpad = (M + ppb - 1) / ppb * ppb;           // keep whole pixel groups converged for the shuffles
  for (int m = blockIdx.x * ppb + tid / TP; m < Mpad; m += gridDim.x * ppb) {
    const bool live = m < M;
    const int mm = live ? m : 0;
    const int qw = mm % P.Wq;
    const int t = mm / P.Wq;
    const int qh = t % P.Hq, n = t / P.Hq;
    float acc = 0.f;
    for (int tp = 0; tp < P.ntaps; ++tp) {
      const TapEntry tap = P.taps[tp];
      const int ih = qh * a.is + tap.dh, iw = qw * a.is + tap.dw;
      if (ih < 0 || ih >= a.Hs || iw < 0 || iw >= a.Ws) continue;
      float4 v = __ldg(reinterpret_cast<const float4*>(a.src + (((size_t)n * a.Hs + ih) * a.Ws + iw) * Cs + c0));
      v = xform4(v, sc, sh, ce, a.in_affine, a.in_act, a.in_slope);
      const float4 w = *reinterpret_cast<const float4*>(&s_w[tap.widx * Cs + c0]);
      acc = fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, fmaf(v.w, w.w, acc))));
    }
    for (int o = TP >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (live && cg == 0) {
      const int oh = qh * a.os + P.ph, ow = qw * a.os + P.pw;
      a.dst[((size_t)n * a.Hd + oh) * a.Wd + ow] = acc + bias;
    }
  }
}

// ---- weight gradients with a 1-channel operand ----------------------------------------------------------
// Cb == 1:  P[tap*Ca + ca] = sum_pix xa(ga[g(pix,tap)][ca]) * xb(db[pix])            (image head)
// Ca == 1:  P[tap][cb]     = sum_pix xa(ga[g(pix,tap)])     * xb(db[pix][cb])        (stem head)
template <int TAPS, bool CB1>
__global__ void __launch_bounds__(kThreads) wgrad_c1_kernel(const __grid_constant__ WgradArgs a) {
  __shared__ float s_red[kThreads][4];
  const int tid = threadIdx.x;
  const int C = CB1 ? a.Ca : a.Cb, TP = C >> 2;
  const int cg = tid % TP, c0 = cg * 4;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 vsc = make_float4(1.f, 1.f, 1.f, 1.f), vsh = zero, vce = zero;   // transform of the vector operand
  float ssc = 1.f, ssh = 0.f, sce = 0.f;                                  // transform of the scalar operand
  const bool v_aff = CB1 ? a.a_affine : a.b_affine, v_act = CB1 ? a.a_act : a.b_act;
  const bool s_aff = CB1 ? a.b_affine : a.a_affine, s_act = CB1 ? a.b_act : a.a_act;
  const float v_slope = CB1 ? a.a_slope : a.b_slope, s_slope = CB1 ? a.b_slope : a.a_slope;
  if (v_aff) {
    vsc = __ldg(reinterpret_cast<const float4*>((CB1 ? a.a_scale : a.b_scale) + c0));
    vsh = __ldg(reinterpret_cast<const float4*>((CB1 ? a.a_shift : a.b_shift) + c0));
    const float* cp = CB1 ? a.a_center : a.b_center;
    if (cp) vce = __ldg(reinterpret_cast<const float4*>(cp + c0));
  }
  if (s_aff) {
    ssc = __ldg(CB1 ? a.b_scale : a.a_scale); ssh = __ldg(CB1 ? a.b_shift : a.a_shift);
    const float* cp = CB1 ? a.b_center : a.a_center;
    if (cp) sce = __ldg(cp);
  }
  float acc[TAPS][4];
  int tdh[TAPS], tdw[TAPS];
#pragma unroll
  for (int t = 0; t < TAPS; ++t) {
    acc[t][0] = 0.f; acc[t][1] = 0.f; acc[t][2] = 0.f; acc[t][3] = 0.f;
    tdh[t] = t / a.kw - a.pad; tdw[t] = t % a.kw - a.pad;
  }
  const int ppb = kThreads / TP;
  for (int pix = blockIdx.x * ppb + tid / TP; pix < a.K; pix += gridDim.x * ppb) {
    const int qw = pix % a.Wq;
    const int t = pix / a.Wq;
    const int qh = t % a.Hq, n = t / a.Hq;
    float4 dv = zero;
    float ds = 0.f;
    if (CB1) {
      ds = __ldg(a.db + pix);
      if (s_aff) ds = fmaf(ds - sce, ssc, ssh);
      if (s_act) ds = lrelu(ds, s_slope);
    } else {
      dv = __ldg(reinterpret_cast<const float4*>(a.db + (size_t)pix * a.Cb + c0));
      dv = xform4(dv, vsc, vsh, vce, v_aff, v_act, v_slope);
    }
#pragma unroll
    for (int tp = 0; tp < TAPS; ++tp) {
      const int ih = qh * a.stride + tdh[tp], iw = qw * a.stride + tdw[tp];
      if (ih < 0 || ih >= a.Ha || iw < 0 || iw >= a.Wa) continue;
      const size_t gp = ((size_t)n * a.Ha + ih) * a.Wa + iw;
      if (CB1) {
        float4 g = __ldg(reinterpret_cast<const float4*>(a.ga + gp * a.Ca + c0));
        g = xform4(g, vsc, vsh, vce, v_aff, v_act, v_slope);
        acc[tp][0] = fmaf(g.x, ds, acc[tp][0]); acc[tp][1] = fmaf(g.y, ds, acc[tp][1]);
        acc[tp][2] = fmaf(g.z, ds, acc[tp][2]); acc[tp][3] = fmaf(g.w, ds, acc[tp][3]);
      } else {
        float g = __ldg(a.ga + gp);
        if (s_aff) g = fmaf(g - sce, ssc, ssh);
        if (s_act) g = lrelu(g, s_slope);
        acc[tp][0] = fmaf(g, dv.x, acc[tp][0]); acc[tp][1] = fmaf(g, dv.y, acc[tp][1]);
        acc[tp][2] = fmaf(g, dv.z, acc[tp][2]); acc[tp][3] = fmaf(g, dv.w, acc[tp][3]);
      }
    }
  }
  // block reduction, one tap at a time; threads [0, 4*TP) own (channel group, lane-in-vector)
#pragma unroll
  for (int tp = 0; tp < TAPS; ++tp) {
    __syncthreads();
    s_red[tid][0] = acc[tp][0]; s_red[tid][1] = acc[tp][1]; s_red[tid][2] = acc[tp][2]; s_red[tid][3] = acc[tp][3];
    __syncthreads();
    if (tid < TP * 4) {
      const int g = tid >> 2, u = tid & 3;
      float t = 0.f;
      for (int r = g; r < kThreads; r += TP) t += s_red[r][u];
      // CB1: row = tap*Ca + ca, col 0 ;  CA1: row = tap, col = cb
      atomicAdd(a.partial + (size_t)tp * C + g * 4 + u, t);
    }
  }
}

// ---- shared-memory tiled variants for the 3x3 image-sized layers ---------------------------------------
// The stream kernels above re-read every input pixel once per tap through L1 with full index arithmetic
// (measured 6-12x over their HBM time at 256x256, B = 64).  Below, a CTA stages an input patch (+ halo)
// in shared memory once, with the producer's BatchNorm + LeakyReLU applied while staging, and the taps
// become shared-memory reads at precomputed offsets.

// true when phase 0 is a single-phase 3x3 window with offsets in [-1, 1] (Conv2d k3 p1 s1|s2 forward,
// Conv2d k3 p1 s1 input-gradient)
static bool window3x3(const GatherArgs& g) {
  if (g.nphase != 1 || g.phase[0].ntaps != 9 || (g.is != 1 && g.is != 2) || g.os != 1) return false;
  for (int t = 0; t < 9; ++t) {
    const TapEntry& e = g.phase[0].taps[t];
    if (e.dh < -1 || e.dh > 1 || e.dw < -1 || e.dw > 1) return false;
  }
  return true;
}

// the same for a 4x4 window with offsets in [-1, 2] at stride 2 (Conv2d k4 s2 p1 forward, ConvTranspose2d k4 s2 p1
// input gradient: causal_cascade/models.py:9, :36)
static bool window4x4(const GatherArgs& g) {
  if (g.nphase != 1 || g.phase[0].ntaps != 16 || g.is != 2 || g.os != 1) return false;
  for (int t = 0; t < 16; ++t) {
    const TapEntry& e = g.phase[0].taps[t];
    if (e.dh < -1 || e.dh > 2 || e.dw < -1 || e.dw > 2) return false;
  }
  return true;
}

// Cs == 1 -> C channels (stem.0 forward: stride 2 + statistics; image-head input gradient: DACT + statistics)
template <int C, int IS, int KW = 3>     // KW = 4: the k4 / s2 / p1 windows of causal_cascade (window4x4 below)
__global__ void __launch_bounds__(kThreads) conv_cs1_tile_kernel(const __grid_constant__ GatherArgs a, const int patches,
                                                                 const int tiles_h, const int tiles_w) {
  constexpr int TH = 8, TW = 32, NB = C / 4, PPP = kThreads / NB, NT = KW * KW;
  constexpr int GR = (TH - 1) * IS + KW, GC = (TW - 1) * IS + KW;
  __shared__ float sG[GR * GC];
  __shared__ double s_red[kThreads][8];
  const int tid = threadIdx.x, c4 = tid % NB, c0 = c4 * 4, pp = tid / NB;
  const PhaseGeom& P = a.phase[0];
  float4 w[NT];
  int toff[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    w[t] = __ldg(reinterpret_cast<const float4*>(a.wt + P.taps[t].widx * C + c0));
    toff[t] = (P.taps[t].dh + 1) * GC + P.taps[t].dw + 1;
  }
  const float in_sc = a.in_affine ? __ldg(a.in_scale) : 1.f, in_sh = a.in_affine ? __ldg(a.in_shift) : 0.f;
  const float in_ce = (a.in_affine && a.in_center) ? __ldg(a.in_center) : 0.f;
  float4 bias = make_float4(0.f, 0.f, 0.f, 0.f), esc = make_float4(1.f, 1.f, 1.f, 1.f), esh = bias, ece = bias;
  if (a.bias) bias = __ldg(reinterpret_cast<const float4*>(a.bias + c0));
  if (a.e_affine) {
    esc = __ldg(reinterpret_cast<const float4*>(a.e_scale + c0));
    esh = __ldg(reinterpret_cast<const float4*>(a.e_shift + c0));
    if (a.e_center) ece = __ldg(reinterpret_cast<const float4*>(a.e_center + c0));
  }
  float f1[4] = {0.f, 0.f, 0.f, 0.f}, f2[4] = {0.f, 0.f, 0.f, 0.f};
  double d1[4] = {0.0, 0.0, 0.0, 0.0}, d2[4] = {0.0, 0.0, 0.0, 0.0};
  int since = 0;
  for (int patch = blockIdx.x; patch < patches; patch += gridDim.x) {
    const int tw = patch % tiles_w, tt = patch / tiles_w, th = tt % tiles_h, n = tt / tiles_h;
    const int h0 = th * TH, w0 = tw * TW;
    const int gh0 = h0 * IS - 1, gw0 = w0 * IS - 1;
    {
      // every load of the halo tile is issued before the first shared-memory store (one round trip per patch, not one
      // per loop iteration)
      constexpr int kIt = (GR * GC + kThreads - 1) / kThreads;
      float hv[kIt];
      bool hok[kIt];
#pragma unroll
      for (int u = 0; u < kIt; ++u) {
        const int idx = tid + u * kThreads;
        const int gi = idx / GC, gj = idx % GC, ih = gh0 + gi, iw = gw0 + gj;
        hok[u] = idx < GR * GC && (unsigned)ih < (unsigned)a.Hs && (unsigned)iw < (unsigned)a.Ws;
        hv[u] = 0.f;
        if (hok[u]) hv[u] = __ldg(a.src + ((size_t)n * a.Hs + ih) * a.Ws + iw);
      }
#pragma unroll
      for (int u = 0; u < kIt; ++u) {
        const int idx = tid + u * kThreads;
        float v = hv[u];
        if (hok[u]) {
          if (a.in_affine) v = fmaf(v - in_ce, in_sc, in_sh);
          if (a.in_act) v = lrelu(v, a.in_slope);
        }
        if (idx < GR * GC) sG[idx] = v;
      }
    }
    __syncthreads();
    for (int p = pp; p < TH * TW; p += PPP) {
      const int r = p / TW, c = p % TW, qh = h0 + r, qw = w0 + c;
      if (qh >= P.Hq || qw >= P.Wq) continue;
      const float* gp = sG + (r * IS) * GC + c * IS;
      float o[4] = {bias.x, bias.y, bias.z, bias.w};
#pragma unroll
      for (int t = 0; t < NT; ++t) fma4(o, gp[toff[t]], w[t]);          // two packed FMAs per tap (common.cuh)
      const size_t off = (((size_t)n * a.Hd + qh) * a.Wd + qw) * C + c0;
      if (a.epi == CVAE_EPI_STATS) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { f1[j] += o[j]; f2[j] = fmaf(o[j], o[j], f2[j]); }
      } else if (a.epi == CVAE_EPI_DACT) {
        const float4 r4 = __ldg(reinterpret_cast<const float4*>(a.epi_ref + off));
        float rf[4] = {r4.x - ece.x, r4.y - ece.y, r4.z - ece.z, r4.w - ece.w};
        const float sc[4] = {esc.x, esc.y, esc.z, esc.w}, sh[4] = {esh.x, esh.y, esh.z, esh.w};
        if (a.epi_add) {
          const float4 ad = __ldg(reinterpret_cast<const float4*>(a.epi_add + off));
          o[0] += ad.x; o[1] += ad.y; o[2] += ad.z; o[3] += ad.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float z = fmaf(rf[j], sc[j], sh[j]);
          o[j] = z > 0.f ? o[j] : o[j] * a.e_slope;
          f1[j] += o[j]; f2[j] = fmaf(o[j], rf[j], f2[j]);
        }
      }
      *reinterpret_cast<float4*>(a.dst + off) = make_float4(o[0], o[1], o[2], o[3]);
      if (a.epi != CVAE_EPI_PLAIN && ++since == 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { d1[j] += (double)f1[j]; d2[j] += (double)f2[j]; f1[j] = 0.f; f2[j] = 0.f; }
        since = 0;
      }
    }
    __syncthreads();
  }
  if (a.epi != CVAE_EPI_PLAIN && a.stats != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { s_red[tid][j] = d1[j] + (double)f1[j]; s_red[tid][4 + j] = d2[j] + (double)f2[j]; }
    __syncthreads();
    if (tid < NB * 8) {
      const int g = tid / 8, k = tid % 8;              // channel group, which of the 8 sums
      double t = 0.0;
      for (int r = g; r < kThreads; r += NB) t += s_red[r][k];
      atomicAdd(a.stats + (k < 4 ? 0 : C) + g * 4 + (k & 3), t);
    }
  }
}

// C channels -> 1 (image head forward), plain epilogue.  A thread owns two vertically adjacent output
// pixels: their 4 x 3 input window is read once, the channel-planar tile keeps a warp's reads contiguous.
template <int C>
__global__ void __launch_bounds__(kThreads) conv_cd1_tile_kernel(const __grid_constant__ GatherArgs a, const int patches,
                                                                 const int tiles_h, const int tiles_w) {
  constexpr int TH = 16, TW = 32, NA = C / 4;
  constexpr int GR = TH + 2, GC = TW + 2, GP = GR * GC;
  extern __shared__ __align__(16) float4 sT[];        // [NA][GP] planar tile, then [9][NA] weights
  float4* sW = sT + NA * GP;
  const int tid = threadIdx.x;
  const PhaseGeom& P = a.phase[0];
  for (int i = tid; i < 9 * NA; i += kThreads) {
    const int t = i / NA, g = i % NA;                   // weights stored by WINDOW position (dh+1)*3 + (dw+1)
    int wi = 0;
    for (int u = 0; u < 9; ++u)
      if ((P.taps[u].dh + 1) * 3 + P.taps[u].dw + 1 == t) wi = P.taps[u].widx;
    sW[i] = __ldg(reinterpret_cast<const float4*>(a.wt + wi * C + g * 4));
  }
  const int sg = tid % NA;                              // staging role: fixed channel group
  float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f), ce = sh;
  if (a.in_affine) {
    sc = __ldg(reinterpret_cast<const float4*>(a.in_scale + sg * 4));
    sh = __ldg(reinterpret_cast<const float4*>(a.in_shift + sg * 4));
    if (a.in_center) ce = __ldg(reinterpret_cast<const float4*>(a.in_center + sg * 4));
  }
  const float bias = a.bias ? __ldg(a.bias) : 0.f;
  const int col = tid % TW, rp = tid / TW;              // rows 2*rp, 2*rp + 1
  for (int patch = blockIdx.x; patch < patches; patch += gridDim.x) {
    const int tw = patch % tiles_w, tt = patch / tiles_w, th = tt % tiles_h, n = tt / tiles_h;
    const int h0 = th * TH, w0 = tw * TW;
    for (int idx = tid; idx < GP * NA; idx += kThreads) {
      const int pix = idx / NA, gi = pix / GC, gj = pix % GC, ih = h0 - 1 + gi, iw = w0 - 1 + gj;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if ((unsigned)ih < (unsigned)a.Hs && (unsigned)iw < (unsigned)a.Ws) {
        v = __ldg(reinterpret_cast<const float4*>(a.src + (((size_t)n * a.Hs + ih) * a.Ws + iw) * C + sg * 4));
        v = xform4(v, sc, sh, ce, a.in_affine, a.in_act, a.in_slope);
      }
      sT[sg * GP + pix] = v;
    }
    __syncthreads();
    float acc0 = bias, acc1 = bias;
#pragma unroll
    for (int g = 0; g < NA; ++g) {
      const float4* tp = sT + g * GP + (2 * rp) * GC + col;
#pragma unroll
      for (int dw = 0; dw < 3; ++dw) {
        const float4 x0 = tp[dw], x1 = tp[GC + dw], x2 = tp[2 * GC + dw], x3 = tp[3 * GC + dw];
        const float4 w0 = sW[(0 * 3 + dw) * NA + g], w1 = sW[(1 * 3 + dw) * NA + g], w2 = sW[(2 * 3 + dw) * NA + g];
        acc0 = fmaf(x0.x, w0.x, fmaf(x0.y, w0.y, fmaf(x0.z, w0.z, fmaf(x0.w, w0.w, acc0))));
        acc0 = fmaf(x1.x, w1.x, fmaf(x1.y, w1.y, fmaf(x1.z, w1.z, fmaf(x1.w, w1.w, acc0))));
        acc0 = fmaf(x2.x, w2.x, fmaf(x2.y, w2.y, fmaf(x2.z, w2.z, fmaf(x2.w, w2.w, acc0))));
        acc1 = fmaf(x1.x, w0.x, fmaf(x1.y, w0.y, fmaf(x1.z, w0.z, fmaf(x1.w, w0.w, acc1))));
        acc1 = fmaf(x2.x, w1.x, fmaf(x2.y, w1.y, fmaf(x2.z, w1.z, fmaf(x2.w, w1.w, acc1))));
        acc1 = fmaf(x3.x, w2.x, fmaf(x3.y, w2.y, fmaf(x3.z, w2.z, fmaf(x3.w, w2.w, acc1))));
      }
    }
    const int qh = h0 + 2 * rp, qw = w0 + col;
    if (qw < P.Wq) {
      if (qh < P.Hq) a.dst[((size_t)n * a.Hd + qh) * a.Wd + qw] = acc0;
      if (qh + 1 < P.Hq) a.dst[((size_t)n * a.Hd + qh + 1) * a.Wd + qw] = acc1;
    }
    __syncthreads();
  }
}

// ---- ConvTranspose2d(C -> 1, stride 2) forward with all four output phases in one pass ------------------
// (causal_cascade / mnist dec_conv tail, k4 s2 p1: causal_cascade/models.py:36).  conv_cd1_kernel ran one grid
// slice per phase, so every input vector was fetched four times by four different thread groups behind a runtime tap
// loop (114 us at batch 256 for 33.5 MB of input).  Here C / 4 threads own one input position q: the 3 x 3
// neighbourhood every phase draws its taps from is requested once, up front (nine 128-bit loads in flight per
// thread), the four phase sums are reduced over the channel groups with shuffles and leave as a 2 x 2 output block.
template <int TP>     // threads per position = C / 4 (power of two <= 16)
__global__ void __launch_bounds__(kThreads) convt_cd1_phases_kernel(const __grid_constant__ GatherArgs a) {
  __shared__ __align__(16) float s_w[16 * 64];         // [widx][Cs]
  __shared__ int s_sel[4][9];                          // weight index of (phase, neighbourhood position) or -1
  const int tid = threadIdx.x, Cs = a.Cs;
  for (int i = tid; i < a.wtaps * Cs; i += kThreads) s_w[i] = __ldg(a.wt + i);
  if (tid < 36) {
    const int ph = tid / 9, pos = tid % 9;
    int sel = -1;
    for (int t = 0; t < a.phase[ph].ntaps; ++t)
      if ((a.phase[ph].taps[t].dh + 1) * 3 + a.phase[ph].taps[t].dw + 1 == pos) sel = a.phase[ph].taps[t].widx;
    s_sel[ph][pos] = sel;
  }
  __syncthreads();
  const int Hq = a.phase[0].Hq, Wq = a.phase[0].Wq;
  const int M = a.N * Hq * Wq;
  const int cg = tid % TP, c0 = cg * 4;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = zero, ce = zero;
  if (a.in_affine) {
    sc = __ldg(reinterpret_cast<const float4*>(a.in_scale + c0));
    sh = __ldg(reinterpret_cast<const float4*>(a.in_shift + c0));
    if (a.in_center) ce = __ldg(reinterpret_cast<const float4*>(a.in_center + c0));
  }
  const float bias = a.bias ? __ldg(a.bias) : 0.f;
  constexpr int ppb = kThreads / TP;
  const int Mpad = (M + ppb - 1) / ppb * ppb;           // keep whole position groups converged for the shuffles
  for (int m = blockIdx.x * ppb + tid / TP; m < Mpad; m += gridDim.x * ppb) {
    const bool live = m < M;
    const int mm = live ? m : 0;
    const int qw = mm % Wq;
    const int t = mm / Wq;
    const int qh = t % Hq, n = t / Hq;
    float4 v[9];
    bool ok[9];
#pragma unroll
    for (int pos = 0; pos < 9; ++pos) {
      const int ih = qh + pos / 3 - 1, iw = qw + pos % 3 - 1;
      ok[pos] = live && (unsigned)ih < (unsigned)a.Hs && (unsigned)iw < (unsigned)a.Ws;
      v[pos] = zero;
      if (ok[pos]) v[pos] = __ldg(reinterpret_cast<const float4*>(a.src + (((size_t)n * a.Hs + ih) * a.Ws + iw) * Cs + c0));
    }
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int pos = 0; pos < 9; ++pos) {
      float4 x = v[pos];
      if (ok[pos]) x = xform4(x, sc, sh, ce, a.in_affine, a.in_act, a.in_slope);     // padding stays exactly 0
#pragma unroll
      for (int ph = 0; ph < 4; ++ph) {
        const int sel = s_sel[ph][pos];
        if (sel >= 0) {
          const float4 w = *reinterpret_cast<const float4*>(&s_w[sel * Cs + c0]);
          acc[ph] = fmaf(x.x, w.x, fmaf(x.y, w.y, fmaf(x.z, w.z, fmaf(x.w, w.w, acc[ph]))));
        }
      }
    }
#pragma unroll
    for (int ph = 0; ph < 4; ++ph)
#pragma unroll
      for (int o = TP >> 1; o > 0; o >>= 1) acc[ph] += __shfl_xor_sync(0xffffffffu, acc[ph], o);
    if (live && cg < 4) {                               // lane cg of the group writes phase cg
      const PhaseGeom& P = a.phase[cg];
      const int oh = qh * 2 + P.ph, ow = qw * 2 + P.pw;
      const float r = cg == 0 ? acc[0] : cg == 1 ? acc[1] : cg == 2 ? acc[2] : acc[3];
      if (oh < a.Hd && ow < a.Wd) a.dst[((size_t)n * a.Hd + oh) * a.Wd + ow] = r + bias;
    }
  }
}

// the four phases of a stride-2 transposed convolution whose taps all lie in the 3 x 3 neighbourhood of the input position
static bool convt_phases3x3(const GatherArgs& g) {
  if (g.nphase != 4 || g.os != 2 || g.is != 1) return false;
  for (int ph = 0; ph < 4; ++ph) {
    const PhaseGeom& P = g.phase[ph];
    if (P.Hq != g.phase[0].Hq || P.Wq != g.phase[0].Wq || P.Hq > g.Hs || P.Wq > g.Ws) return false;
    for (int t = 0; t < P.ntaps; ++t)
      if (P.taps[t].dh < -1 || P.taps[t].dh > 1 || P.taps[t].dw < -1 || P.taps[t].dw > 1) return false;
    for (int q = 0; q < ph; ++q)
      if (g.phase[q].ph == P.ph && g.phase[q].pw == P.pw) return false;
  }
  return true;
}

static inline int stream_grid(int64_t pixels, int tp) {
  const int64_t ppb = kThreads / tp;
  int64_t b = (pixels + ppb - 1) / ppb;
  const int64_t cap = (int64_t)kNumSMs * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

bool launch_conv_cs1(const GatherArgs& g, int maxM, cudaStream_t st) {
  if (g.Cs != 1 || (g.Cd & 3) || g.Cd > 64 || g.Cd < 4 || g.wtaps > 16) return false;
  const int tp = g.Cd / 4;
  if (kThreads % tp) return false;
  if (window3x3(g) && maxM >= 65536 && (g.Cd == 16 || g.Cd == 32)) {
    const PhaseGeom& P = g.phase[0];
    const int tiles_h = (P.Hq + 7) / 8, tiles_w = (P.Wq + 31) / 32;
    const int patches = g.N * tiles_h * tiles_w, grid = min(patches, kNumSMs * 6);
    if (g.Cd == 16 && g.is == 1) conv_cs1_tile_kernel<16, 1><<<grid, kThreads, 0, st>>>(g, patches, tiles_h, tiles_w);
    else if (g.Cd == 16) conv_cs1_tile_kernel<16, 2><<<grid, kThreads, 0, st>>>(g, patches, tiles_h, tiles_w);
    else if (g.is == 1) conv_cs1_tile_kernel<32, 1><<<grid, kThreads, 0, st>>>(g, patches, tiles_h, tiles_w);
    else conv_cs1_tile_kernel<32, 2><<<grid, kThreads, 0, st>>>(g, patches, tiles_h, tiles_w);
    return true;
  }
  if (window4x4(g) && maxM >= 65536 && g.Cd == 32) {     // causal_cascade at its training batch: 92 / 104 us on the streaming kernel
    const PhaseGeom& P = g.phase[0];
    const int tiles_h = (P.Hq + 7) / 8, tiles_w = (P.Wq + 31) / 32;
    const int patches = g.N * tiles_h * tiles_w, grid = min(patches, kNumSMs * 6);
    conv_cs1_tile_kernel<32, 2, 4><<<grid, kThreads, 0, st>>>(g, patches, tiles_h, tiles_w);
    return true;
  }
  conv_cs1_kernel<<<dim3(stream_grid(maxM, tp), 1, g.nphase), kThreads, 0, st>>>(g);
  return true;
}

bool launch_conv_cd1(const GatherArgs& g, int maxM, cudaStream_t st) {
  if (g.Cd != 1 || (g.Cs & 3) || g.Cs > 64 || g.Cs < 4 || g.wtaps > 16 || g.epi != CVAE_EPI_PLAIN) return false;
  const int tp = g.Cs / 4;
  if (tp & (tp - 1)) return false;                      // power of two: shuffle reduction inside a warp
  {
    static const bool on = [] { const char* e = getenv("CVAE_HEAD_FWD_NEW"); return !(e && e[0] == '0'); }();
    if (on && launch_conv16_head_fwd(g, st) == 1) return true;
  }
  if (window3x3(g) && g.is == 1 && maxM >= 65536 && g.Cs == 16) {
    const PhaseGeom& P = g.phase[0];
    const int tiles_h = (P.Hq + 15) / 16, tiles_w = (P.Wq + 31) / 32;
    const int patches = g.N * tiles_h * tiles_w;
    const size_t smem = sizeof(float4) * (4 * 18 * 34 + 9 * 4);
    conv_cd1_tile_kernel<16><<<min(patches, kNumSMs * 4), kThreads, smem, st>>>(g, patches, tiles_h, tiles_w);
    return true;
  }
  if (convt_phases3x3(g) && tp >= 4 && tp <= 16) {
    static const bool on = [] { const char* e = getenv("CVAE_CONVT_PHASES"); return !(e && e[0] == '0'); }();
    if (on) {
      const int grid = stream_grid(maxM, tp);
      if (tp == 4) convt_cd1_phases_kernel<4><<<grid, kThreads, 0, st>>>(g);
      else if (tp == 8) convt_cd1_phases_kernel<8><<<grid, kThreads, 0, st>>>(g);
      else convt_cd1_phases_kernel<16><<<grid, kThreads, 0, st>>>(g);
      return true;
    }
  }
  conv_cd1_kernel<<<dim3(stream_grid(maxM, tp), 1, g.nphase), kThreads, 0, st>>>(g);
  return true;
}

template <bool CB1>
static bool launch_wgrad_c1(const WgradArgs& a, int taps, cudaStream_t st) {
  const int C = CB1 ? a.Ca : a.Cb;
  if ((C & 3) || C > 64 || C < 4 || kThreads % (C / 4)) return false;
  if (cudaMemsetAsync(a.partial, 0, sizeof(float) * (size_t)taps * C, st) != cudaSuccess) return false;
  // every block ends with taps x C same-address atomics: 1184 blocks made 1184 serialised adds per output (the two 4x4
  // layers of the cascade model: 217 us each); two blocks per SM keep the loads in flight with a quarter of the atomics
  static const int bps = [] { const char* e = getenv("CVAE_WGC1_BPS"); return e ? atoi(e) : 2; }();
  const int grid = min(stream_grid(a.K, C / 4), kNumSMs * bps);
  if (taps == 9) wgrad_c1_kernel<9, CB1><<<grid, kThreads, 0, st>>>(a);
  else if (taps == 16) wgrad_c1_kernel<16, CB1><<<grid, kThreads, 0, st>>>(a);
  else return false;
  return true;
}
bool launch_wgrad_cb1(const WgradArgs& a, int taps, cudaStream_t st) { return a.Cb == 1 && launch_wgrad_c1<true>(a, taps, st); }
bool launch_wgrad_ca1(const WgradArgs& a, int taps, cudaStream_t st) { return a.Ca == 1 && launch_wgrad_c1<false>(a, taps, st); }

}  // namespace cvae
