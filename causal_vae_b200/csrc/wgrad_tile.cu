// Shared-memory tiled fp32 weight gradient for the high-resolution, few-channel 3x3 layers (sm_100a).
//
//   P[split][tap*Ca + ca][cb] = sum_{pix in split} xa(ga[n, qh*S + dh - 1, qw*S + dw - 1, ca]) * xb(db[n, qh, qw, cb])
//
// The decoder tail (ConvT 16->16 @256^2, 32->16 @128^2), ResBlock(32) @64^2, stem.3 (32->64), the image
// head (16->1) and stem.0 (1->32) contract over 0.25-4 M pixels with only 16-32 channels per side.  On
// the tensor cores a kind::tf32 MMA costs ~51 clk (A in TMEM) whatever N <= 64 is (scripts/
// umma_rate.cu), so N = 16/32 leaves them 75-90 % idle and the per-(tap, channel) operand transposes
// dominate (wgrad_tc ran these layers 10-28x over their HBM time).  Here a CTA stages an 8 x 16 patch
// of db and the matching input halo of ga in shared memory ONCE (BatchNorm + LeakyReLU applied while
// staging), and every thread keeps a 3(dw) x 4(ca) x 4(cb) accumulator block in registers for its
// (dh, ca-group, cb-group); lanes that share an operand read it as a shared-memory broadcast.
// When the (dh, ca, cb) grid needs fewer than 256 threads, several thread groups ("pixel subsets") split
// the patch's pixels and are summed through shared memory at the end; every CTA writes one K-split
// slice of the partial buffer, reduced by cvae_wgrad_reduce.
#include "conv_args.cuh"

namespace cvae {

constexpr int kWtTH = 8, kWtTW = 16, kWtThreads = 256;
// resident CTAs per SM: the 1-channel variants are pure streams (12-24 accumulators, ~12 KB of tiles) and need
// more CTAs in flight to cover the stage -> barrier -> compute latency chain
__host__ __device__ constexpr int wt_ctas_per_sm(int ca, int cb) { return ca * cb <= 32 ? 4 : 2; }

template <int V> struct VecT { float v[V]; };

template <int V>
__device__ __forceinline__ VecT<V> ldv(const float* p) {
  VecT<V> r;
  if constexpr (V == 4) { const float4 t = *reinterpret_cast<const float4*>(p); r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; }
  else r.v[0] = *p;
  return r;
}
template <int V>
__device__ __forceinline__ VecT<V> ldgv(const float* p) {
  VecT<V> r;
  if constexpr (V == 4) { const float4 t = __ldg(reinterpret_cast<const float4*>(p)); r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; }
  else r.v[0] = __ldg(p);
  return r;
}
template <int V>
__device__ __forceinline__ void stv(float* p, const VecT<V>& r) {
  if constexpr (V == 4) *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  else *p = r.v[0];
}

template <int CA, int CB, int S, int KW = 3>       // KW: window size (3, or 4 for the Conv2d / ConvTranspose2d(k4, s2, p1) layers of causal_cascade)
__global__ void __launch_bounds__(kWtThreads, wt_ctas_per_sm(CA, CB))
wgrad_tile_kernel(const __grid_constant__ WgradArgs a, const int patches, const int tiles_h, const int tiles_w) {
  constexpr int VA = CA >= 4 ? 4 : 1, VB = CB >= 4 ? 4 : 1;
  constexpr int NA = CA / VA, NB = CB / VB;
  constexpr int T1 = KW * NA * NB;                 // threads covering every (dh, ca-group, cb-group)
  constexpr int PS = kWtThreads / T1;              // pixel subsets
  constexpr int GR = (kWtTH - 1) * S + KW, GC = (kWtTW - 1) * S + KW;
  constexpr int NPIX = kWtTH * kWtTW;
  static_assert(T1 <= kWtThreads, "tile does not fit the block");
  extern __shared__ __align__(16) float smem[];
  float* sG = smem;                                // [GR*GC][CA]
  float* sD = smem + ((GR * GC * CA + 3) & ~3);    // [NPIX][CB], 16-byte aligned

  const int tid = threadIdx.x;
  const int cb0 = blockIdx.y * CB;
  // staging roles: a thread always stages the same channel group of each operand
  const int ga_grp = tid % NA, db_grp = tid % NB;
  VecT<VA> asc, ash, ace;
  VecT<VB> bsc, bsh, bce;
#pragma unroll
  for (int i = 0; i < VA; ++i) {
    asc.v[i] = a.a_affine ? __ldg(a.a_scale + ga_grp * VA + i) : 1.f;
    ash.v[i] = a.a_affine ? __ldg(a.a_shift + ga_grp * VA + i) : 0.f;
    ace.v[i] = (a.a_affine && a.a_center) ? __ldg(a.a_center + ga_grp * VA + i) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < VB; ++i) {
    bsc.v[i] = a.b_affine ? __ldg(a.b_scale + cb0 + db_grp * VB + i) : 1.f;
    bsh.v[i] = a.b_affine ? __ldg(a.b_shift + cb0 + db_grp * VB + i) : 0.f;
    bce.v[i] = (a.b_affine && a.b_center) ? __ldg(a.b_center + cb0 + db_grp * VB + i) : 0.f;
  }
  // compute role
  const bool active = tid < PS * T1;
  const int ps = tid / T1, t1 = tid % T1;
  const int cb_i = t1 % NB, ca_i = (t1 / NB) % NA, dh = t1 / (NB * NA);
  float acc[KW][VA][VB];
#pragma unroll
  for (int w = 0; w < KW; ++w)
#pragma unroll
    for (int i = 0; i < VA; ++i)
#pragma unroll
      for (int j = 0; j < VB; ++j) acc[w][i][j] = 0.f;

  for (int patch = blockIdx.x; patch < patches; patch += gridDim.x) {
    const int tw = patch % tiles_w, tt = patch / tiles_w, th = tt % tiles_h, n = tt / tiles_h;
    const int h0 = th * kWtTH, w0 = tw * kWtTW;
    // ---- stage the input halo of ga and the db patch (transform applied; padding stays exactly 0) ----
    // Loads are issued in batches of up to kBatch per thread BEFORE the first shared-memory store of the batch: the
    // plain load -> transform -> store loop cannot be reordered by the compiler across the stores, so a thread paid one
    // full L2 / DRAM round trip per vector (nine per patch for the stride-2 halo) with nothing else in flight.
    const int gh0 = h0 * S - a.pad, gw0 = w0 * S - a.pad;
    constexpr int kBatch = 6;
    constexpr int kGaItems = GR * GC * NA, kDbItems = NPIX * NB;
    constexpr int kGaIters = (kGaItems + kWtThreads - 1) / kWtThreads, kDbIters = (kDbItems + kWtThreads - 1) / kWtThreads;
    {
      // the db patch first (its loads stay in flight while the halo batches are issued)
      VecT<VB> dv[kDbIters];
      int dpix[kDbIters];
#pragma unroll
      for (int u = 0; u < kDbIters; ++u) {
        const int idx = tid + u * kWtThreads;
        dpix[u] = -1;
#pragma unroll
        for (int i = 0; i < VB; ++i) dv[u].v[i] = 0.f;
        if (idx < kDbItems) {
          const int pix = idx / NB, r = pix / kWtTW, c = pix % kWtTW;
          const int qh = h0 + r, qw = w0 + c;
          dpix[u] = pix;
          if (qh < a.Hq && qw < a.Wq) {
            dv[u] = ldgv<VB>(a.db + (((size_t)n * a.Hq + qh) * a.Wq + qw) * a.Cb + cb0 + db_grp * VB);
            dpix[u] |= 1 << 30;
          }
        }
      }
#pragma unroll
      for (int b0 = 0; b0 < kGaIters; b0 += kBatch) {
        VecT<VA> gv[kBatch];
        int gpix[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          gpix[u] = -1;
#pragma unroll
          for (int i = 0; i < VA; ++i) gv[u].v[i] = 0.f;
          if (b0 + u < kGaIters) {
            const int idx = tid + (b0 + u) * kWtThreads;
            if (idx < kGaItems) {
              const int pix = idx / NA, gi = pix / GC, gj = pix % GC;
              const int ih = gh0 + gi, iw = gw0 + gj;
              gpix[u] = pix;
              if ((unsigned)ih < (unsigned)a.Ha && (unsigned)iw < (unsigned)a.Wa) {
                gv[u] = ldgv<VA>(a.ga + (((size_t)n * a.Ha + ih) * a.Wa + iw) * CA + ga_grp * VA);
                gpix[u] |= 1 << 30;
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          if (gpix[u] < 0) continue;
          if (gpix[u] & (1 << 30)) {
#pragma unroll
            for (int i = 0; i < VA; ++i) {
              if (a.a_affine) gv[u].v[i] = fmaf(gv[u].v[i] - ace.v[i], asc.v[i], ash.v[i]);
              if (a.a_act) gv[u].v[i] = lrelu(gv[u].v[i], a.a_slope);
            }
          }
          stv<VA>(sG + (gpix[u] & ~(1 << 30)) * CA + ga_grp * VA, gv[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < kDbIters; ++u) {
        if (dpix[u] < 0) continue;
        if (dpix[u] & (1 << 30)) {
#pragma unroll
          for (int i = 0; i < VB; ++i) {
            if (a.b_affine) dv[u].v[i] = fmaf(dv[u].v[i] - bce.v[i], bsc.v[i], bsh.v[i]);
            if (a.b_act) dv[u].v[i] = lrelu(dv[u].v[i], a.b_slope);
          }
        }
        stv<VB>(sD + (dpix[u] & ~(1 << 30)) * CB + db_grp * VB, dv[u]);
      }
    }
    __syncthreads();
    // ---- accumulate ----
    if (active) {
#pragma unroll 2
      for (int pix = ps; pix < NPIX; pix += PS) {
        const int r = pix / kWtTW, c = pix % kWtTW;
        const VecT<VB> d = ldv<VB>(sD + pix * CB + cb_i * VB);
        const float* gp = sG + ((r * S + dh) * GC + c * S) * CA + ca_i * VA;
#pragma unroll
        for (int w = 0; w < KW; ++w) {
          const VecT<VA> g = ldv<VA>(gp + w * CA);
          if constexpr (VB == 4) {                     // four columns per row value: two packed FMAs (common.cuh)
            const float4 d4 = make_float4(d.v[0], d.v[1], d.v[2], d.v[3]);
#pragma unroll
            for (int i = 0; i < VA; ++i) fma4(acc[w][i], g.v[i], d4);
          } else {
#pragma unroll
            for (int i = 0; i < VA; ++i)
#pragma unroll
              for (int j = 0; j < VB; ++j) acc[w][i][j] = fmaf(g.v[i], d.v[j], acc[w][i][j]);
          }
        }
      }
    }
    __syncthreads();
  }
  // ---- pixel subsets are summed through shared memory; every CTA is one K-split slice of the partial buffer ----
  float* out = a.partial + (size_t)blockIdx.x * a.rows * a.Cb;
  if constexpr (PS == 1) {
    if (active) {
#pragma unroll
      for (int w = 0; w < KW; ++w)
#pragma unroll
        for (int i = 0; i < VA; ++i) {
          VecT<VB> v;
#pragma unroll
          for (int j = 0; j < VB; ++j) v.v[j] = acc[w][i][j];
          stv<VB>(out + (size_t)((dh * KW + w) * CA + ca_i * VA + i) * a.Cb + cb0 + cb_i * VB, v);
        }
    }
  } else {
    constexpr int OUT = KW * KW * CA * CB;
    float* sR = smem;                              // [PS][9*CA][CB]  (the tiles are dead: last loop barrier passed)
    if (active) {
#pragma unroll
      for (int w = 0; w < KW; ++w)
#pragma unroll
        for (int i = 0; i < VA; ++i) {
          VecT<VB> v;
#pragma unroll
          for (int j = 0; j < VB; ++j) v.v[j] = acc[w][i][j];
          stv<VB>(sR + ps * OUT + ((dh * KW + w) * CA + ca_i * VA + i) * CB + cb_i * VB, v);
        }
    }
    __syncthreads();
    for (int o = tid; o < OUT / VB; o += kWtThreads) {
      VecT<VB> v = ldv<VB>(sR + o * VB);
      for (int q = 1; q < PS; ++q) {
        const VecT<VB> u = ldv<VB>(sR + q * OUT + o * VB);
#pragma unroll
        for (int j = 0; j < VB; ++j) v.v[j] += u.v[j];
      }
      const int row = (o * VB) / CB, col = (o * VB) % CB;
      stv<VB>(out + (size_t)row * a.Cb + cb0 + col, v);
    }
  }
}

template <int CA, int CB, int S, int KW = 3>
static int launch_tile(const WgradArgs& a, cudaStream_t st) {
  constexpr int GR = (kWtTH - 1) * S + KW, GC = (kWtTW - 1) * S + KW;
  constexpr int VA = CA >= 4 ? 4 : 1, VB = CB >= 4 ? 4 : 1, PS = kWtThreads / (KW * (CA / VA) * (CB / VB));
  size_t floats = (size_t)((GR * GC * CA + 3) & ~3) + (size_t)kWtTH * kWtTW * CB;
  if (PS > 1) floats = max(floats, (size_t)PS * KW * KW * CA * CB);
  const size_t smem = sizeof(float) * floats;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(wgrad_tile_kernel<CA, CB, S, KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess)
      return CVAE_ERR_LAUNCH;
    attr_set = true;
  }
  const int tiles_h = (a.Hq + kWtTH - 1) / kWtTH, tiles_w = (a.Wq + kWtTW - 1) / kWtTW;
  const int patches = a.N * tiles_h * tiles_w;
  wgrad_tile_kernel<CA, CB, S, KW><<<dim3(kNumSMs * wt_ctas_per_sm(CA, CB), a.Cb / CB), kWtThreads, smem, st>>>(a, patches, tiles_h, tiles_w);
  return CVAE_OK;
}

static inline int tile_cb_slice(int Cb) { return Cb > 32 ? 32 : Cb; }

}  // namespace cvae

using namespace cvae;

// > 0: the K-split count the tiled kernel needs (size of the partial buffer) when it covers the shape; 0 otherwise.
extern "C" int cvae_wgrad_tile_splits(int pixels, int Ca, int Cb, int k, int stride, int pad) {
  // 4x4 / stride 2 / pad 1 with one channel on the image side: first Conv2d and last ConvTranspose2d of causal_cascade
  // (causal_cascade/models.py:9, :36); were 160-176 us each on the streaming fallback at batch 256
  if (k == 4) return (stride == 2 && pad == 1 && Ca == 1 && Cb == 32 && pixels >= 32768) ? kNumSMs * wt_ctas_per_sm(1, 32) : 0;
  if (k != 3 || pad != 1 || (stride != 1 && stride != 2)) return 0;
  if (!(Ca == 1 || Ca == 16 || Ca == 32) || !(Cb == 1 || Cb == 16 || Cb == 32 || Cb == 64)) return 0;
  if (Ca == 1 && Cb == 1) return 0;
  // 64 output columns fill a tensor-core tile well enough: the pipelined wgrad_tc kernel is 1.8x faster
  // there (stem.3 32 -> 64: 257 us vs 470 us; decoder.8 64 -> 32 transposed: 77 us vs 138 us, B = 64)
  if (Cb == 64 && Ca >= 16) return 0;
  if (pixels < 32768) return 0;
  return kNumSMs * wt_ctas_per_sm(Ca, tile_cb_slice(Cb));
}

extern "C" int cvae_conv_wgrad_tile(const cvae_wgrad_params_t* p, cvae_stream_t s) {
  if (!p || !p->ga || !p->db || !p->partial) return CVAE_ERR_BAD_ARG;
  if (p->kh != p->kw || (p->kh != 3 && p->kh != 4)) return CVAE_ERR_UNSUPPORTED_SHAPE;
  const int need = cvae_wgrad_tile_splits(p->N * p->Hq * p->Wq, p->Ca, p->Cb, p->kh, p->stride, p->pad);
  if (need == 0) return CVAE_ERR_UNSUPPORTED_SHAPE;
  if (p->splits != need) return CVAE_ERR_BAD_ARG;
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
  a.K = p->N * p->Hq * p->Wq; a.kchunk = 0;
  cudaStream_t st = as_stream(s);
  const int cbs = tile_cb_slice(p->Cb);
  int rc = CVAE_ERR_UNSUPPORTED_SHAPE;
  if (p->kh == 4) {
    rc = launch_tile<1, 32, 2, 4>(a, st);
    if (rc != CVAE_OK) return rc;
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
  }
#define CVAE_WT(ca, cb)                                                                      \
  if (p->Ca == ca && cbs == cb) rc = p->stride == 1 ? launch_tile<ca, cb, 1>(a, st) : launch_tile<ca, cb, 2>(a, st);
  CVAE_WT(1, 16) CVAE_WT(1, 32) CVAE_WT(16, 1) CVAE_WT(16, 16) CVAE_WT(16, 32) CVAE_WT(32, 1) CVAE_WT(32, 16) CVAE_WT(32, 32)
#undef CVAE_WT
  if (rc != CVAE_OK) return rc;
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}
