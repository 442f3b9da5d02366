// Implicit-GEMM convolution / transposed convolution / linear family for sm_100a (fp32 SIMT core).
//
// One "gather" kernel serves Conv2d forward, ConvTranspose2d forward (phase-decomposed so no
// multiply hits an inserted zero), both input-gradients and nn.Linear (one tap).  The producer
// layer's BatchNorm + LeakyReLU is applied while the A operand is staged (training-mode BN forces a
// grid-wide barrier between a conv and its activation, so the normalise pass is folded into the
// consumer's load instead of costing a read+write of every activation), and the epilogue either
// accumulates the BN batch statistics of the output or applies the activation derivative and
// accumulates the BN-backward reductions.  A thread-per-pixel variant covers the skinny layers
// (Cin = 1 stem head, Cout = 1 image head) that are pure HBM streams.  The weight gradient is a
// pixels-contracted GEMM with deterministic split-K partials.
#include "common.cuh"
#include "conv_args.cuh"

namespace cvae {

// ------------------------------------------------------------------------------------------------
// tiled gather kernel: BM=128 output pixels x BN channels, BK=16 input channels of one tap per step
// ------------------------------------------------------------------------------------------------
template <int BN>
__global__ void __launch_bounds__(256) igemm_gather_kernel(const __grid_constant__ GatherArgs a) {
  constexpr int BM = 128, BK = 16, TM = 8, TN = BN / 16, BV = BN / 4;
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN];
  __shared__ int s_n[BM], s_qh[BM], s_qw[BM];

  // grid.z = phase (conv-transpose / dgrad) or, for single-phase problems with a long K and few
  // output tiles, the K-split index (partial sums are combined with atomics on a zeroed dst)
  const int kz = a.ksplit > 1 ? blockIdx.z : 0;
  const PhaseGeom& P = a.phase[a.ksplit > 1 ? 0 : blockIdx.z];
  const int tid = threadIdx.x;
  const int M = a.N * P.Hq * P.Wq;
  const int m0 = blockIdx.x * BM;
  if (m0 >= M) return;
  const int n0 = blockIdx.y * BN;

  if (tid < BM) {
    const int m = m0 + tid;
    if (m < M) {
      const int qw = m % P.Wq, t = m / P.Wq;
      s_qw[tid] = qw; s_qh[tid] = t % P.Hq; s_n[tid] = t / P.Hq;
    } else {
      s_n[tid] = -1; s_qh[tid] = 0; s_qw[tid] = 0;
    }
  }
  __syncthreads();

  const int ar = tid >> 2, kv = (tid & 3) * 4;
  int rn[2], rqh[2], rqw[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) { rn[i] = s_n[ar + 64 * i]; rqh[i] = s_qh[ar + 64 * i] * a.is; rqw[i] = s_qw[ar + 64 * i] * a.is; }
  const int bk = tid / BV, bn = (tid % BV) * 4;
  const bool b_active = bk < BK;

  const int kc = (a.Cs + BK - 1) / BK;
  const int Tall = P.ntaps * kc;
  const int Tper = (Tall + max(a.ksplit, 1) - 1) / max(a.ksplit, 1);
  const int t_beg = kz * Tper;
  const int T = min(Tall, t_beg + Tper);

  float4 ra[2], rb;
  auto load_tile = [&](int t) {
    const TapEntry tap = P.taps[t / kc];
    const int c0 = (t % kc) * BK;
    const int c = c0 + kv;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int ih = rqh[i] + tap.dh, iw = rqw[i] + tap.dw;
      if (rn[i] >= 0 && c < a.Cs && ih >= 0 && ih < a.Hs && iw >= 0 && iw < a.Ws) {
        v = __ldg(reinterpret_cast<const float4*>(a.src + (((size_t)rn[i] * a.Hs + ih) * a.Ws + iw) * a.Cs + c));
        if (a.in_affine) {
          const float4 sc = __ldg(reinterpret_cast<const float4*>(a.in_scale + c));
          const float4 sh = __ldg(reinterpret_cast<const float4*>(a.in_shift + c));
          if (a.in_center != nullptr) {
            const float4 ce = __ldg(reinterpret_cast<const float4*>(a.in_center + c));
            v.x -= ce.x; v.y -= ce.y; v.z -= ce.z; v.w -= ce.w;
          }
          v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
          v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
        }
        if (a.in_act) {
          v.x = lrelu(v.x, a.in_slope); v.y = lrelu(v.y, a.in_slope);
          v.z = lrelu(v.z, a.in_slope); v.w = lrelu(v.w, a.in_slope);
        }
      }
      ra[i] = v;
    }
    rb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b_active) {
      const int kk = c0 + bk, col = n0 + bn;
      if (kk < a.Cs && col < a.Cd)
        rb = __ldg(reinterpret_cast<const float4*>(a.wt + ((size_t)tap.widx * a.Cs + kk) * a.Cd + col));
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      As[buf][kv + 0][ar + 64 * i] = ra[i].x; As[buf][kv + 1][ar + 64 * i] = ra[i].y;
      As[buf][kv + 2][ar + 64 * i] = ra[i].z; As[buf][kv + 3][ar + 64 * i] = ra[i].w;
    }
    if (b_active) *reinterpret_cast<float4*>(&Bs[buf][bk][bn]) = rb;
  };

  const int tx = tid & 15, ty = tid >> 4;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  if (t_beg < T) {
    load_tile(t_beg);
    store_tile(t_beg & 1);
  }
  __syncthreads();
  for (int t = t_beg; t < T; ++t) {
    const int buf = t & 1;
    if (t + 1 < T) load_tile(t + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM + 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[TN];
      if constexpr (TN == 4) {
        const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
        bv[0] = b.x; bv[1] = b.y; bv[2] = b.z; bv[3] = b.w;
      } else if constexpr (TN == 2) {
        const float2 b = *reinterpret_cast<const float2*>(&Bs[buf][k][tx * 2]);
        bv[0] = b.x; bv[1] = b.y;
      } else {
        bv[0] = Bs[buf][k][tx];
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (t + 1 < T) store_tile(buf ^ 1);
    __syncthreads();
  }

  // ---------------- epilogue ----------------
  const int cbase = n0 + tx * TN;
  const bool col_ok = cbase < a.Cd;
  float bias[TN], esc[TN], esh[TN], ece[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    bias[j] = (a.bias != nullptr && col_ok && kz == 0) ? __ldg(a.bias + cbase + j) : 0.f;
    esc[j] = (a.e_affine && col_ok) ? __ldg(a.e_scale + cbase + j) : 1.f;
    esh[j] = (a.e_affine && col_ok) ? __ldg(a.e_shift + cbase + j) : 0.f;
    ece[j] = (a.e_affine && a.e_center != nullptr && col_ok) ? __ldg(a.e_center + cbase + j) : 0.f;
  }
  // per-thread statistics in double: fp32 products are exact in double, so E[y^2] - mean^2 keeps
  // two-pass accuracy even for channels with |mean| >> std
  double s1[TN], s2[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) { s1[j] = 0.0; s2[j] = 0.0; }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int r = ty * TM + i;
    const int n = s_n[r];
    if (n < 0 || !col_ok) continue;
    const int oh = s_qh[r] * a.os + P.ph, ow = s_qw[r] * a.os + P.pw;
    const size_t off = (((size_t)n * a.Hd + oh) * a.Wd + ow) * a.Cd + cbase;
    float v[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) v[j] = acc[i][j] + bias[j];
    if (a.epi == CVAE_EPI_STATS) {
#pragma unroll
      for (int j = 0; j < TN; ++j) { s1[j] += (double)v[j]; s2[j] += (double)v[j] * (double)v[j]; }
    } else if (a.epi == CVAE_EPI_DACT) {
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const float refc = __ldg(a.epi_ref + off + j) - ece[j];
        if (a.epi_add != nullptr) v[j] += __ldg(a.epi_add + off + j);
        const float z = fmaf(refc, esc[j], esh[j]);
        v[j] = z > 0.f ? v[j] : v[j] * a.e_slope;
        s1[j] += (double)v[j]; s2[j] += (double)v[j] * (double)refc;
      }
    }
    if (a.ksplit > 1) {
#pragma unroll
      for (int j = 0; j < TN; ++j) atomicAdd(a.dst + off + j, v[j]);
    } else if constexpr (TN == 4) {
      *reinterpret_cast<float4*>(a.dst + off) = make_float4(v[0], v[1], v[2], v[3]);
    } else if constexpr (TN == 2) {
      *reinterpret_cast<float2*>(a.dst + off) = make_float2(v[0], v[1]);
    } else {
      a.dst[off] = v[0];
    }
  }

  if (a.epi != CVAE_EPI_PLAIN && a.stats != nullptr) {
    static_assert(sizeof(As) >= 2 * 16 * BN * sizeof(double), "reduction scratch does not fit");
    double* red1 = reinterpret_cast<double*>(&As[0][0][0]);   // [16][BN]
    double* red2 = red1 + 16 * BN;                            // [16][BN]
    __syncthreads();
#pragma unroll
    for (int j = 0; j < TN; ++j) { red1[ty * BN + tx * TN + j] = s1[j]; red2[ty * BN + tx * TN + j] = s2[j]; }
    __syncthreads();
    if (tid < BN && n0 + tid < a.Cd) {
      double t1 = 0.0, t2 = 0.0;
#pragma unroll
      for (int y = 0; y < 16; ++y) { t1 += red1[y * BN + tid]; t2 += red2[y * BN + tid]; }
      atomicAdd(a.stats + n0 + tid, t1);
      atomicAdd(a.stats + a.Cd + n0 + tid, t2);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// thread-per-pixel gather kernel for skinny layers (Cs == 1 or Cd == 1 ...): weights in shared
// memory, grid-stride over pixels, statistics kept in registers until the end.
// ------------------------------------------------------------------------------------------------
template <int CD>
__global__ void __launch_bounds__(128) conv_pix_kernel(const __grid_constant__ GatherArgs a) {
  extern __shared__ float s_w[];  // [wtaps][Cs][CD]
  __shared__ double s_red[4][2 * CD];
  const int tid = threadIdx.x;
  const int wcount = a.wtaps * a.Cs * CD;
  for (int i = tid; i < wcount; i += blockDim.x) s_w[i] = __ldg(a.wt + i);
  __syncthreads();

  const PhaseGeom& P = a.phase[blockIdx.z];
  const int M = a.N * P.Hq * P.Wq;
  const bool vec = (a.Cs & 3) == 0;
  float s1[CD], s2[CD];          // fp32 partials over <= 8 pixels, flushed into the doubles below
  double d1[CD], d2[CD];
#pragma unroll
  for (int j = 0; j < CD; ++j) { s1[j] = 0.f; s2[j] = 0.f; d1[j] = 0.0; d2[j] = 0.0; }
  int since_flush = 0;

  for (int m = blockIdx.x * blockDim.x + tid; m < M; m += gridDim.x * blockDim.x) {
    const int qw = m % P.Wq, t = m / P.Wq, qh = t % P.Hq, n = t / P.Hq;
    float acc[CD];
#pragma unroll
    for (int j = 0; j < CD; ++j) acc[j] = a.bias != nullptr ? __ldg(a.bias + j) : 0.f;
    for (int tp = 0; tp < P.ntaps; ++tp) {
      const TapEntry tap = P.taps[tp];
      const int ih = qh * a.is + tap.dh, iw = qw * a.is + tap.dw;
      if (ih < 0 || ih >= a.Hs || iw < 0 || iw >= a.Ws) continue;
      const float* sp = a.src + (((size_t)n * a.Hs + ih) * a.Ws + iw) * a.Cs;
      const float* wp = s_w + (size_t)tap.widx * a.Cs * CD;
      if (vec) {
        for (int c = 0; c < a.Cs; c += 4) {
          const float4 v4 = __ldg(reinterpret_cast<const float4*>(sp + c));
          float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (a.in_affine) {
              if (a.in_center != nullptr) v[u] -= __ldg(a.in_center + c + u);
              v[u] = fmaf(v[u], __ldg(a.in_scale + c + u), __ldg(a.in_shift + c + u));
            }
            if (a.in_act) v[u] = lrelu(v[u], a.in_slope);
#pragma unroll
            for (int j = 0; j < CD; ++j) acc[j] = fmaf(v[u], wp[(c + u) * CD + j], acc[j]);
          }
        }
      } else {
        for (int c = 0; c < a.Cs; ++c) {
          float v = __ldg(sp + c);
          if (a.in_affine) {
            if (a.in_center != nullptr) v -= __ldg(a.in_center + c);
            v = fmaf(v, __ldg(a.in_scale + c), __ldg(a.in_shift + c));
          }
          if (a.in_act) v = lrelu(v, a.in_slope);
#pragma unroll
          for (int j = 0; j < CD; ++j) acc[j] = fmaf(v, wp[c * CD + j], acc[j]);
        }
      }
    }
    const int oh = qh * a.os + P.ph, ow = qw * a.os + P.pw;
    const size_t off = (((size_t)n * a.Hd + oh) * a.Wd + ow) * CD;
    if (a.epi == CVAE_EPI_STATS) {
#pragma unroll
      for (int j = 0; j < CD; ++j) { s1[j] += acc[j]; s2[j] = fmaf(acc[j], acc[j], s2[j]); }
    } else if (a.epi == CVAE_EPI_DACT) {
#pragma unroll
      for (int j = 0; j < CD; ++j) {
        float ref = __ldg(a.epi_ref + off + j);
        if (a.e_affine && a.e_center != nullptr) ref -= __ldg(a.e_center + j);
        if (a.epi_add != nullptr) acc[j] += __ldg(a.epi_add + off + j);
        const float z = a.e_affine ? fmaf(ref, __ldg(a.e_scale + j), __ldg(a.e_shift + j)) : ref;
        acc[j] = z > 0.f ? acc[j] : acc[j] * a.e_slope;
        s1[j] += acc[j]; s2[j] = fmaf(acc[j], ref, s2[j]);
      }
    }
    if (a.epi != CVAE_EPI_PLAIN && ++since_flush == 4) {
#pragma unroll
      for (int j = 0; j < CD; ++j) { d1[j] += (double)s1[j]; d2[j] += (double)s2[j]; s1[j] = 0.f; s2[j] = 0.f; }
      since_flush = 0;
    }
    if constexpr (CD % 4 == 0) {
#pragma unroll
      for (int j = 0; j < CD; j += 4)
        *reinterpret_cast<float4*>(a.dst + off + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < CD; ++j) a.dst[off + j] = acc[j];
    }
  }

  if (a.epi != CVAE_EPI_PLAIN && a.stats != nullptr) {
    const int lane = tid & 31, w = tid >> 5;
#pragma unroll
    for (int j = 0; j < CD; ++j) {
      const double t1 = warp_sum_d(d1[j] + (double)s1[j]), t2 = warp_sum_d(d2[j] + (double)s2[j]);
      if (lane == 0) { s_red[w][j] = t1; s_red[w][CD + j] = t2; }
    }
    __syncthreads();
    if (tid < 2 * CD) {
      const double t = s_red[0][tid] + s_red[1][tid] + s_red[2][tid] + s_red[3][tid];
      atomicAdd(a.stats + tid, t);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// weight gradient:  P[split][tap*Ca + ca][cb] = sum_{pix in split} xa(ga[g(pix,tap)][ca]) * xb(db[pix][cb])
// ------------------------------------------------------------------------------------------------
template <int BN, int VA, int VB>
__global__ void __launch_bounds__(256) wgrad_kernel(const __grid_constant__ WgradArgs a) {
  constexpr int BM = 64, BK = 16, TM = 4, TN = BN / 16, BV = BN / 4;
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kbeg = blockIdx.z * a.kchunk;
  const int kend = min(a.K, kbeg + a.kchunk);
  const int T = kend > kbeg ? (kend - kbeg + BK - 1) / BK : 0;

  // A-load mapping
  int a_k[VA == 4 ? 1 : 4], a_r, a_dh = 0, a_dw = 0, a_ca = 0;
  bool a_rowok;
  if constexpr (VA == 4) {
    a_k[0] = tid >> 4; a_r = (tid & 15) * 4;
  } else {
    a_r = tid & 63;
#pragma unroll
    for (int i = 0; i < 4; ++i) a_k[i] = (tid >> 6) + 4 * i;
  }
  {
    const int row = m0 + a_r;
    a_rowok = row < a.rows;
    const int tap = a_rowok ? row / a.Ca : 0;
    a_ca = a_rowok ? row % a.Ca : 0;
    a_dh = tap / a.kw - a.pad; a_dw = tap % a.kw - a.pad;
  }
  // B-load mapping
  int b_k, b_n;
  bool b_active;
  if constexpr (VB == 4) { b_k = tid / BV; b_n = (tid % BV) * 4; b_active = b_k < BK; }
  else { b_k = tid / BN; b_n = tid % BN; b_active = b_k < BK; }   // BN == 16 -> all 256 threads

  float ra[4], rb[4];
  auto gather_a = [&](int pix, float* out) {
#pragma unroll
    for (int u = 0; u < VA; ++u) out[u] = 0.f;
    if (!a_rowok || pix >= kend) return;
    const int qw = pix % a.Wq, t = pix / a.Wq, qh = t % a.Hq, n = t / a.Hq;
    const int ih = qh * a.stride + a_dh, iw = qw * a.stride + a_dw;
    if (ih < 0 || ih >= a.Ha || iw < 0 || iw >= a.Wa) return;
    const float* p = a.ga + (((size_t)n * a.Ha + ih) * a.Wa + iw) * a.Ca + a_ca;
    if constexpr (VA == 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p));
      out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
    } else {
      out[0] = __ldg(p);
    }
#pragma unroll
    for (int u = 0; u < VA; ++u) {
      if (a.a_affine) {
        if (a.a_center != nullptr) out[u] -= __ldg(a.a_center + a_ca + u);
        out[u] = fmaf(out[u], __ldg(a.a_scale + a_ca + u), __ldg(a.a_shift + a_ca + u));
      }
      if (a.a_act) out[u] = lrelu(out[u], a.a_slope);
    }
  };
  auto load_tile = [&](int t) {
    const int kb = kbeg + t * BK;
    if constexpr (VA == 4) {
      gather_a(kb + a_k[0], ra);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) gather_a(kb + a_k[i], &ra[i]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) rb[u] = 0.f;
    if (b_active) {
      const int pix = kb + b_k, col = n0 + b_n;
      if (pix < kend && col < a.Cb) {
        const float* p = a.db + (size_t)pix * a.Cb + col;
        if constexpr (VB == 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(p));
          rb[0] = v.x; rb[1] = v.y; rb[2] = v.z; rb[3] = v.w;
        } else {
          rb[0] = __ldg(p);
        }
#pragma unroll
        for (int u = 0; u < VB; ++u) {
          if (a.b_affine) {
            if (a.b_center != nullptr) rb[u] -= __ldg(a.b_center + col + u);
            rb[u] = fmaf(rb[u], __ldg(a.b_scale + col + u), __ldg(a.b_shift + col + u));
          }
          if (a.b_act) rb[u] = lrelu(rb[u], a.b_slope);
        }
      }
    }
  };
  auto store_tile = [&](int buf) {
    if constexpr (VA == 4) {
      *reinterpret_cast<float4*>(&As[buf][a_k[0]][a_r]) = make_float4(ra[0], ra[1], ra[2], ra[3]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) As[buf][a_k[i]][a_r] = ra[i];
    }
    if (b_active) {
      if constexpr (VB == 4) *reinterpret_cast<float4*>(&Bs[buf][b_k][b_n]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
      else Bs[buf][b_k][b_n] = rb[0];
    }
  };

  const int tx = tid & 15, ty = tid >> 4;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  if (T > 0) {
    load_tile(0);
    store_tile(0);
    __syncthreads();
    for (int t = 0; t < T; ++t) {
      const int buf = t & 1;
      if (t + 1 < T) load_tile(t + 1);
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM]);
        const float av[4] = {a4.x, a4.y, a4.z, a4.w};
        float bv[TN];
        if constexpr (TN == 4) {
          const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
          bv[0] = b.x; bv[1] = b.y; bv[2] = b.z; bv[3] = b.w;
        } else if constexpr (TN == 2) {
          const float2 b = *reinterpret_cast<const float2*>(&Bs[buf][k][tx * 2]);
          bv[0] = b.x; bv[1] = b.y;
        } else {
          bv[0] = Bs[buf][k][tx];
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      if (t + 1 < T) store_tile(buf ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int row = m0 + ty * TM + i;
    if (row >= a.rows) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int col = n0 + tx * TN + j;
      if (col < a.Cb) a.partial[((size_t)blockIdx.z * a.rows + row) * a.Cb + col] = acc[i][j];
    }
  }
}

// Sum the K-split partials.  A block owns 128 consecutive partial-order outputs j = (tap*ca + a)*cb + b
// (a warp reads 512 contiguous bytes per split) and spreads the splits over 8 thread groups; the result
// is written in torch layout dst[b][a][tap].  cb % 4 == 0 (vector path) or any cb (scalar path, V = 1).
template <int V>
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int taps, int ca,
                                                           int ca_real, int cb, float* __restrict__ dst, int accumulate) {
  __shared__ float red[8][32 * V + 1];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const size_t stride = (size_t)taps * ca * cb;
  const int total = taps * ca * cb;
  for (int j0 = blockIdx.x * 32 * V; j0 < total; j0 += gridDim.x * 32 * V) {
    const int j = j0 + lane * V;
    float s[V];
#pragma unroll
    for (int i = 0; i < V; ++i) s[i] = 0.f;
    if (j < total) {
      for (int k = grp; k < splits; k += 8) {
        if constexpr (V == 4) {
          const float4 t = *reinterpret_cast<const float4*>(partial + (size_t)k * stride + j);
          s[0] += t.x; s[1] += t.y; s[2] += t.z; s[3] += t.w;
        } else {
          s[0] += partial[(size_t)k * stride + j];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) red[grp][lane * V + i] = s[i];
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * V; e += 256) {
      const int jj = j0 + e;
      if (jj < total) {
        const float t = ((red[0][e] + red[1][e]) + (red[2][e] + red[3][e])) + ((red[4][e] + red[5][e]) + (red[6][e] + red[7][e]));
        const int b_ = jj % cb, q = jj / cb, a_ = q % ca, tap = q / ca;
        if (a_ < ca_real) {
          const size_t o = ((size_t)b_ * ca_real + a_) * taps + tap;
          dst[o] = accumulate ? dst[o] + t : t;
        }
      }
    }
    __syncthreads();
  }
}

// Large weight tensors (Linear 512 -> 16384: 8.4 M outputs): the torch layout dst[b][a][tap] is the
// transpose of the partial layout [tap][a][b], so a block sums the splits for a 32(a) x 32(b) tile of
// every tap with coalesced row reads, stages it in shared memory and writes contiguous runs of dst.
template <int TA>
__global__ void __launch_bounds__(256) wgrad_reduce_tile_kernel(const float* __restrict__ partial, int splits, int taps,
                                                                int ca, int ca_real, int cb, float* __restrict__ dst,
                                                                int accumulate) {
  extern __shared__ float tile[];                   // [32 b][TA a][taps] (+1 pad per b row)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int a0 = blockIdx.x * TA, b0 = blockIdx.y * 32;
  const size_t stride = (size_t)taps * ca * cb;
  const int ld = TA * taps + 1;
  for (int tap = 0; tap < taps; ++tap)
    for (int al = warp; al < TA; al += 8) {
      const int a_ = a0 + al, b_ = b0 + lane;
      float s = 0.f;
      if (a_ < ca && b_ < cb) {
        const float* p = partial + ((size_t)tap * ca + a_) * cb + b_;
#pragma unroll 8
        for (int k = 0; k < splits; ++k) s += __ldg(p + (size_t)k * stride);
      }
      tile[lane * ld + al * taps + tap] = s;
    }
  __syncthreads();
  const int na = min(TA, ca_real - a0);             // rows a >= ca_real are padding and dropped
  if (na <= 0) return;
  const int run = na * taps;
  for (int bl = warp; bl < 32; bl += 8) {
    const int b_ = b0 + bl;
    if (b_ >= cb) continue;
    float* o = dst + ((size_t)b_ * ca_real + a0) * taps;
    for (int i = lane; i < run; i += 32) o[i] = accumulate ? o[i] + tile[bl * ld + i] : tile[bl * ld + i];
  }
}

__global__ void pack_weight_kernel(const float* __restrict__ src, float* __restrict__ dst, int A,
                                   int A_pad, int B, int taps, int src_bat, int src_ld) {
  const size_t total = (size_t)taps * A_pad * B;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(i % B);
    const size_t r = i / B;
    const int a_ = (int)(r % A_pad), t = (int)(r / A_pad);
    float v = 0.f;
    if (a_ < A && (src_bat || b < src_ld))
      v = src_bat ? src[((size_t)b * src_ld + a_) * taps + t] : src[((size_t)a_ * src_ld + b) * taps + t];
    dst[i] = v;
  }
}

int build_geom(const cvae_conv_params_t* p, GatherArgs& g) {
  if (p->kh * p->kw > 16 || p->kh < 1 || p->kw < 1 || p->stride < 1 || p->stride > 2) return CVAE_ERR_UNSUPPORTED_SHAPE;
  g.wtaps = p->kh * p->kw;
  if (p->mode == CVAE_CONV_GATHER) {
    g.nphase = 1; g.os = 1; g.is = p->stride;
    PhaseGeom& P = g.phase[0];
    P.ph = 0; P.pw = 0; P.Hq = p->Hd; P.Wq = p->Wd; P.ntaps = 0;
    for (int kh = 0; kh < p->kh; ++kh)
      for (int kw = 0; kw < p->kw; ++kw) P.taps[P.ntaps++] = {kh - p->pad, kw - p->pad, kh * p->kw + kw};
    // shape check: Hd = (Hs + 2p - k)/s + 1
    if ((p->Hs + 2 * p->pad - p->kh) / p->stride + 1 != p->Hd || (p->Ws + 2 * p->pad - p->kw) / p->stride + 1 != p->Wd)
      return CVAE_ERR_BAD_ARG;
  } else if (p->mode == CVAE_CONV_SCATTER) {
    const int s = p->stride;
    g.os = s; g.is = 1; g.nphase = 0;
    // dst[oh] += src[ih] * w[kh]  with oh = ih*s - pad + kh ; Hd may include output_padding
    if ((p->Hs - 1) * s - 2 * p->pad + p->kh > p->Hd || (p->Hs - 1) * s - 2 * p->pad + p->kh + s <= p->Hd) return CVAE_ERR_BAD_ARG;
    if ((p->Ws - 1) * s - 2 * p->pad + p->kw > p->Wd || (p->Ws - 1) * s - 2 * p->pad + p->kw + s <= p->Wd) return CVAE_ERR_BAD_ARG;
    for (int ph = 0; ph < s; ++ph)
      for (int pw = 0; pw < s; ++pw) {
        PhaseGeom& P = g.phase[g.nphase++];
        P.ph = ph; P.pw = pw; P.ntaps = 0;
        P.Hq = (p->Hd - ph + s - 1) / s; P.Wq = (p->Wd - pw + s - 1) / s;
        for (int kh = 0; kh < p->kh; ++kh) {
          const int nh = ph + p->pad - kh;
          if (((nh % s) + s) % s != 0) continue;
          for (int kw = 0; kw < p->kw; ++kw) {
            const int nw = pw + p->pad - kw;
            if (((nw % s) + s) % s != 0) continue;
            const int dh = nh >= 0 ? nh / s : -((-nh) / s), dw = nw >= 0 ? nw / s : -((-nw) / s);
            P.taps[P.ntaps++] = {dh, dw, kh * p->kw + kw};
          }
        }
      }
  } else {
    return CVAE_ERR_BAD_ARG;
  }
  return CVAE_OK;
}

}  // namespace cvae

using namespace cvae;

extern "C" int cvae_conv_gather(const cvae_conv_params_t* p, cvae_stream_t s) {
  if (!p || !p->src || !p->wt || !p->dst || p->N <= 0) return CVAE_ERR_BAD_ARG;
  if (p->epi == CVAE_EPI_DACT && !p->epi_ref) return CVAE_ERR_BAD_ARG;
  GatherArgs g;
  g.a_image = nullptr;
  g.src = p->src; g.wt = p->wt; g.bias = p->bias; g.dst = p->dst;
  g.in_scale = p->in.scale; g.in_shift = p->in.shift; g.in_center = p->in.center; g.in_slope = p->in.slope;
  g.in_affine = p->in.scale != nullptr; g.in_act = p->in.slope != 1.0f;
  g.epi = p->epi; g.epi_ref = p->epi_ref; g.epi_add = p->epi_add;
  g.e_scale = p->epi_x.scale; g.e_shift = p->epi_x.shift; g.e_center = p->epi_x.center; g.e_slope = p->epi_x.slope;
  g.e_affine = p->epi_x.scale != nullptr;
  g.stats = p->stats;
  g.N = p->N; g.Hs = p->Hs; g.Ws = p->Ws; g.Cs = p->Cs; g.Hd = p->Hd; g.Wd = p->Wd; g.Cd = p->Cd;
  const int rc = build_geom(p, g);
  if (rc != CVAE_OK) return rc;
  int maxM = 0;
  for (int i = 0; i < g.nphase; ++i) maxM = max(maxM, p->N * g.phase[i].Hq * g.phase[i].Wq);
  cudaStream_t st = as_stream(s);

  g.ksplit = 1;
  {  // image-sized 16 -> 16 stride-2 layers: fp32 tile kernels (conv_few.cu)
    const int fr = launch_conv_few(p, g, st);
    if (fr < 0) return fr;
    if (fr == 1) { CVAE_LAUNCH_CHECK(); return CVAE_OK; }
  }
  {  // Linear layers with a handful of rows: N x K split kernels (linear_small.cu)
    const int lr = launch_linear_small(g, st);
    if (lr < 0) return lr;
    if (lr == 1) { CVAE_LAUNCH_CHECK(); return CVAE_OK; }
  }
  const bool tiled_ok = (p->Cs % 4 == 0) && (p->Cd % 4 == 0) && p->Cs >= 8 && p->Cd >= 8;
  if (tiled_ok) {
    const int gx = (maxM + 127) / 128;
    const int bn = p->Cd >= 64 ? 64 : p->Cd >= 32 ? 32 : 16;
    const int gy = (p->Cd + bn - 1) / bn;
    int gz = g.nphase;
    const int ksteps = g.wtaps * ((p->Cs + 15) / 16);
    if (g.nphase == 1 && p->epi == CVAE_EPI_PLAIN && gx * gy <= kNumSMs / 2 && ksteps >= 64) {
      int ks = (2 * kNumSMs + gx * gy - 1) / (gx * gy);
      ks = min(ks, ksteps / 16);
      if (ks > 1) {
        g.ksplit = ks; gz = ks;
        const size_t bytes = (size_t)p->N * p->Hd * p->Wd * p->Cd * sizeof(float);
        if (cudaMemsetAsync(p->dst, 0, bytes, st) != cudaSuccess) return CVAE_ERR_LAUNCH;
      }
    }
    if (bn == 64) {
      igemm_gather_kernel<64><<<dim3(gx, gy, gz), 256, 0, st>>>(g);
    } else if (bn == 32) {
      igemm_gather_kernel<32><<<dim3(gx, gy, gz), 256, 0, st>>>(g);
    } else {
      igemm_gather_kernel<16><<<dim3(gx, gy, gz), 256, 0, st>>>(g);
    }
  } else if (launch_conv_cs1(g, maxM, st) || launch_conv_cd1(g, maxM, st)) {
    // 1-channel layers: coalesced stream kernels (skinny.cu)
  } else {
    const size_t smem = (size_t)g.wtaps * p->Cs * p->Cd * sizeof(float);
    if (smem > 40 * 1024) return CVAE_ERR_UNSUPPORTED_SHAPE;
    int gx = (maxM + 127) / 128;
    gx = min(gx, kNumSMs * 8);
    dim3 grid(gx, 1, g.nphase);
#define CVAE_PIX_CASE(n) case n: conv_pix_kernel<n><<<grid, 128, smem, st>>>(g); break;
    switch (p->Cd) {   // narrow heads (logits, morphology vectors): any width up to 16, and 32
      CVAE_PIX_CASE(1) CVAE_PIX_CASE(2) CVAE_PIX_CASE(3) CVAE_PIX_CASE(4) CVAE_PIX_CASE(5) CVAE_PIX_CASE(6)
      CVAE_PIX_CASE(7) CVAE_PIX_CASE(8) CVAE_PIX_CASE(9) CVAE_PIX_CASE(10) CVAE_PIX_CASE(11) CVAE_PIX_CASE(12)
      CVAE_PIX_CASE(13) CVAE_PIX_CASE(14) CVAE_PIX_CASE(15) CVAE_PIX_CASE(16) CVAE_PIX_CASE(32)
      default: return CVAE_ERR_UNSUPPORTED_SHAPE;
    }
#undef CVAE_PIX_CASE
  }
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_wgrad_splits(int pixels, int rows, int cols) {
  if (cols == 1 || rows <= 16) return 1;   // 1-channel operands: single-pass stream kernels (skinny.cu)
  const int bn = cols >= 64 ? 64 : cols >= 32 ? 32 : 16;
  const int tiles = ((rows + 63) / 64) * ((cols + bn - 1) / bn);
  int splits = (2 * kNumSMs + tiles - 1) / tiles;
  const int max_by_k = (pixels + 255) / 256;
  if (splits > max_by_k) splits = max_by_k;
  if (splits > 512) splits = 512;
  if (splits < 1) splits = 1;
  return splits;
}

extern "C" int cvae_conv_wgrad(const cvae_wgrad_params_t* p, cvae_stream_t s) {
  if (!p || !p->ga || !p->db || !p->partial || p->splits < 1) return CVAE_ERR_BAD_ARG;
  if ((p->Ha + 2 * p->pad - p->kh) / p->stride + 1 != p->Hq || (p->Wa + 2 * p->pad - p->kw) / p->stride + 1 != p->Wq)
    return CVAE_ERR_BAD_ARG;
  WgradArgs a;
  a.ga = p->ga; a.db = p->db;
  a.a_scale = p->xa.scale; a.a_shift = p->xa.shift; a.a_center = p->xa.center; a.a_slope = p->xa.slope;
  a.a_affine = p->xa.scale != nullptr; a.a_act = p->xa.slope != 1.0f;
  a.b_scale = p->xb.scale; a.b_shift = p->xb.shift; a.b_center = p->xb.center; a.b_slope = p->xb.slope;
  a.b_affine = p->xb.scale != nullptr; a.b_act = p->xb.slope != 1.0f;
  a.partial = p->partial;
  a.N = p->N; a.Ha = p->Ha; a.Wa = p->Wa; a.Ca = p->Ca; a.Hq = p->Hq; a.Wq = p->Wq; a.Cb = p->Cb;
  a.kw = p->kw; a.stride = p->stride; a.pad = p->pad;
  a.rows = p->kh * p->kw * p->Ca;
  a.K = p->N * p->Hq * p->Wq;
  int chunk = (a.K + p->splits - 1) / p->splits;
  chunk = ((chunk + 15) / 16) * 16;
  a.kchunk = chunk;
  cudaStream_t st = as_stream(s);
  if (launch_wgrad_small(a, p->kh * p->kw, p->splits, st) == 1) {
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
  }
  if (p->splits == 1 && (launch_wgrad_cb1(a, p->kh * p->kw, st) || launch_wgrad_ca1(a, p->kh * p->kw, st))) {
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
  }
  const bool va4 = (p->Ca % 4) == 0, vb4 = (p->Cb % 4) == 0;
  const int gx = (a.rows + 63) / 64;
  if (vb4 && p->Cb >= 64) {
    dim3 grid(gx, (p->Cb + 63) / 64, p->splits);
    if (va4) wgrad_kernel<64, 4, 4><<<grid, 256, 0, st>>>(a); else wgrad_kernel<64, 1, 4><<<grid, 256, 0, st>>>(a);
  } else if (vb4 && p->Cb >= 32) {
    dim3 grid(gx, (p->Cb + 31) / 32, p->splits);
    if (va4) wgrad_kernel<32, 4, 4><<<grid, 256, 0, st>>>(a); else wgrad_kernel<32, 1, 4><<<grid, 256, 0, st>>>(a);
  } else if (vb4) {
    dim3 grid(gx, (p->Cb + 15) / 16, p->splits);
    if (va4) wgrad_kernel<16, 4, 4><<<grid, 256, 0, st>>>(a); else wgrad_kernel<16, 1, 4><<<grid, 256, 0, st>>>(a);
  } else {
    dim3 grid(gx, (p->Cb + 15) / 16, p->splits);
    if (va4) wgrad_kernel<16, 4, 1><<<grid, 256, 0, st>>>(a); else wgrad_kernel<16, 1, 1><<<grid, 256, 0, st>>>(a);
  }
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_wgrad_reduce(const float* partial, int splits, int taps, int ca, int ca_real, int cb,
                                 float* dst, int accumulate, cvae_stream_t s) {
  if (!partial || !dst || splits < 1 || ca_real > ca) return CVAE_ERR_BAD_ARG;
  const int total = cb * ca * taps;
  if (((ca + 31) / 32) * ((cb + 31) / 32) >= 128 && splits <= 32 && taps <= 16) {
    if (taps == 1 && ((ca + 31) / 32) * ((cb + 31) / 32) < 1024) {
      // mid-size Linear weights (256 x 512 ...): 32 x 32 tiles give < 1 block per SM and a serial chain of
      // `splits` loads per thread; 8-row tiles quadruple the blocks in flight
      const size_t smem = sizeof(float) * 32 * (8 * taps + 1);
      wgrad_reduce_tile_kernel<8><<<dim3((ca + 7) / 8, (cb + 31) / 32), 256, smem, as_stream(s)>>>(
          partial, splits, taps, ca, ca_real, cb, dst, accumulate);
    } else if (taps <= 9) {
      const size_t smem = sizeof(float) * 32 * (32 * taps + 1);
      wgrad_reduce_tile_kernel<32><<<dim3((ca + 31) / 32, (cb + 31) / 32), 256, smem, as_stream(s)>>>(
          partial, splits, taps, ca, ca_real, cb, dst, accumulate);
    } else {
      const size_t smem = sizeof(float) * 32 * (16 * taps + 1);
      wgrad_reduce_tile_kernel<16><<<dim3((ca + 15) / 16, (cb + 31) / 32), 256, smem, as_stream(s)>>>(
          partial, splits, taps, ca, ca_real, cb, dst, accumulate);
    }
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
  }
  if (cb % 4 == 0) {
    const int blocks = min((total + 127) / 128, kNumSMs * 16);
    wgrad_reduce_kernel<4><<<blocks, 256, 0, as_stream(s)>>>(partial, splits, taps, ca, ca_real, cb, dst, accumulate);
  } else {
    const int blocks = min((total + 31) / 32, kNumSMs * 16);
    wgrad_reduce_kernel<1><<<blocks, 256, 0, as_stream(s)>>>(partial, splits, taps, ca, ca_real, cb, dst, accumulate);
  }
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_pack_weight(const float* src, float* dst, int A, int A_pad, int B, int taps, int src_bat,
                                int src_ld, cvae_stream_t s) {
  if (!src || !dst || A_pad < A || (src_bat && src_ld < A)) return CVAE_ERR_BAD_ARG;
  const size_t total = (size_t)taps * A_pad * B;
  int blocks = (int)min((total + 255) / 256, (size_t)kNumSMs * 16);
  pack_weight_kernel<<<blocks, 256, 0, as_stream(s)>>>(src, dst, A, A_pad, B, taps, src_bat, src_ld);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}
