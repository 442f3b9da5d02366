// tcgen05 / TMEM weight-gradient kernel for sm_100a (3xTF32).
//
//   P[split][tap*Ca + ca][cb] = sum_{pix in split} xa(ga[g(pix, tap)][ca]) * xb(db[pix][cb])
//
// The contraction runs over PIXELS, so both operands are channel-contiguous ("MN-major").  That
// makes tensor memory the natural home of the A operand: a TMEM lane is one (tap, ca) output row
// and its columns are consecutive pixels, which is exactly what a warp produces when lane = channel
// loads one coalesced 128-byte channel vector per pixel.  A never touches shared memory:
//   * A warps (8 = 4 lane quarters x 2 stage parities): 32 pixels x 1 channel per thread, apply the
//     producer layer's BatchNorm + LeakyReLU in registers, split into tf32 hi + lo, tcgen05.st both.
//   * B warps (4): 32 pixels x BN channels per stage through registers (transform, split) into
//     MN-major shared tiles in the 128B / 32-byte-base swizzle (the only one tf32 MN-major accepts).
//   * one thread issues, per 8 pixels, main += Ahi*Bhi and cross += Alo*Bhi + Ahi*Blo (separate
//     accumulators: the tensor core truncates on accumulate, so the 2^-11-scaled cross terms must
//     not share the long main chain); the epilogue adds them and writes the split-K partial tile.
#include "common.cuh"
#include "conv_args.cuh"
#include "tc_common.cuh"

namespace cvae {

using namespace tc;

constexpr int kWgThreads = 13 * 32;   // 8 A warps + 4 B warps + 1 MMA warp
constexpr int kWgStages = 4;
constexpr int kWgKPix = 32;           // pixels (K) per stage

// A-operand loaders: 32 consecutive pixels of one (tap, channel) row into registers; bit p of the
// returned mask is set where the value is real data (padding / out-of-range pixels stay exactly 0).
// wg_load_seg: the stage is made of 32/SEG whole output-row segments (Wq % SEG == 0, stages start on
// multiples of 32 pixels), so row / image coordinates and the vertical bound are computed once per
// segment instead of once per pixel (the per-pixel form, ~25 instructions per value, made this
// producer the bottleneck of the kernel).  wg_load_lin: 1x1 / Linear (pixel = row of the matrix).
template <int SEG>
__device__ __forceinline__ uint32_t wg_load_seg(const WgradArgs& a, int kb, int kend, bool row_ok, int dh, int dw, int ca,
                                                float (&v)[32]) {
  uint32_t okm = 0;
#pragma unroll
  for (int sg = 0; sg < 32 / SEG; ++sg) {
    const int pixs = kb + sg * SEG;
    const int qw0 = pixs % a.Wq, t = pixs / a.Wq, qh = t % a.Hq, n = t / a.Hq;
    const int ih = qh * a.stride + dh;
    const bool hok = row_ok && pixs < kend && (unsigned)ih < (unsigned)a.Ha;
    const float* rowp = a.ga + (((size_t)n * a.Ha + (hok ? ih : 0)) * a.Wa) * a.Ca + ca;
    int iw = qw0 * a.stride + dw;
#pragma unroll
    for (int j = 0; j < SEG; ++j) {
      const bool ok = hok && (unsigned)iw < (unsigned)a.Wa;
      v[sg * SEG + j] = 0.f;
      if (ok) { v[sg * SEG + j] = __ldg(rowp + (size_t)iw * a.Ca); okm |= 1u << (sg * SEG + j); }
      iw += a.stride;
    }
  }
  return okm;
}
__device__ __forceinline__ uint32_t wg_load_lin(const WgradArgs& a, int kb, int kend, bool row_ok, int ca, float (&v)[32]) {
  uint32_t okm = 0;
  const float* src = a.ga + (size_t)kb * a.Ca + ca;
#pragma unroll
  for (int p = 0; p < 32; ++p) {
    v[p] = 0.f;
    if (row_ok && kb + p < kend) { v[p] = __ldg(src + (size_t)p * a.Ca); okm |= 1u << p; }
  }
  return okm;
}

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(const __grid_constant__ WgradArgs a, const int BN) {
  extern __shared__ uint8_t dsm_raw[];
  __shared__ __align__(8) uint64_t s_full[kWgStages];
  __shared__ __align__(8) uint64_t s_empty[kWgStages];
  __shared__ __align__(8) uint64_t s_accum;
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * BN;
  const int kbeg = blockIdx.z * a.kchunk;
  const int kend = min(a.K, kbeg + a.kchunk);
  const int nst = kend > kbeg ? (kend - kbeg + kWgKPix - 1) / kWgKPix : 0;
  const uint32_t stage_bytes = 2u * kWgKPix * 128u * (uint32_t)((BN + 31) >> 5);   // hi + lo, [blocks of 32 ch][32 pix][128 B]
  const uint32_t half_bytes = stage_bytes / 2;
  uint8_t* dsm_gen = dsm_raw + ((1024u - (smem_u32(dsm_raw) & 1023u)) & 1023u);
  const uint32_t dsm = smem_u32(dsm_gen);

  if (warp == 12) {
    if (lane == 0) {
      for (int i = 0; i < kWgStages; ++i) { mbar_init(smem_u32(&s_full[i]), 8); mbar_init(smem_u32(&s_empty[i]), 1); }
      mbar_init(smem_u32(&s_accum), 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(&s_tmem), 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t t_main = tmem, t_cross = tmem + (uint32_t)BN, t_a = tmem + 2u * BN;   // A slot s at t_a + 64 s

  if (warp < 8) {
    // ============================== A producers: registers -> TMEM ==============================
    const int q = warp & 3, par = warp >> 2;
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < a.rows;
    const int tap = row_ok ? row / a.Ca : 0;
    const int ca = row_ok ? row % a.Ca : 0;
    const int dh = tap / a.kw - a.pad, dw = tap % a.kw - a.pad;
    float sc = 1.f, sh = 0.f, ce = 0.f;
    if (a.a_affine && row_ok) {
      sc = __ldg(a.a_scale + ca); sh = __ldg(a.a_shift + ca);
      if (a.a_center != nullptr) ce = __ldg(a.a_center + ca);
    }
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    for (int st = par; st < nst; st += 2) {
      const int slot = st % kWgStages;
      const int kb = kbeg + st * kWgKPix;
      float v[kWgKPix];
      uint32_t okm = 0;                          // padding / out-of-range pixels stay exactly 0
      if (a.seg == 32) okm = wg_load_seg<32>(a, kb, kend, row_ok, dh, dw, ca, v);
      else if (a.seg == 16) okm = wg_load_seg<16>(a, kb, kend, row_ok, dh, dw, ca, v);
      else if (a.seg == 8) okm = wg_load_seg<8>(a, kb, kend, row_ok, dh, dw, ca, v);
      else if (a.seg == -1) okm = wg_load_lin(a, kb, kend, row_ok, ca, v);
      else {
        int qw = kb % a.Wq, t = kb / a.Wq, qh = t % a.Hq, n = t / a.Hq;
#pragma unroll
        for (int p = 0; p < kWgKPix; ++p) {
          const int ih = qh * a.stride + dh, iw = qw * a.stride + dw;
          const bool ok = row_ok && (kb + p) < kend && (unsigned)ih < (unsigned)a.Ha && (unsigned)iw < (unsigned)a.Wa;
          v[p] = 0.f;
          if (ok) {
            v[p] = __ldg(a.ga + (((size_t)n * a.Ha + ih) * a.Wa + iw) * a.Ca + ca);
            okm |= 1u << p;
          }
          if (++qw == a.Wq) { qw = 0; if (++qh == a.Hq) { qh = 0; ++n; } }
        }
      }
      uint32_t hi[kWgKPix];
#pragma unroll
      for (int p = 0; p < kWgKPix; ++p) {
        float x = v[p];
        if ((okm >> p) & 1u) {
          if (a.a_affine) x = fmaf(x - ce, sc, sh);
          if (a.a_act) x = lrelu(x, a.a_slope);
        }
        const float h = tf32_rn(x);
        hi[p] = __float_as_uint(h);
        v[p] = tf32_rn(x - h);
      }
      mbar_wait(smem_u32(&s_empty[slot]), (uint32_t)(((st / kWgStages) & 1) ^ 1));
      tc_fence_after();
      tmem_st32(t_a + lane_sel + (uint32_t)(slot * 64), hi);
      tmem_st32(t_a + lane_sel + (uint32_t)(slot * 64 + 32), reinterpret_cast<const uint32_t*>(v));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_full[slot]));
    }

    // ============================== epilogue: main + cross -> partial tile ==============================
    mbar_wait(smem_u32(&s_accum), 0);
    tc_fence_after();
    if (nst > 0) {
      const int half = BN >> 1;                  // BN >= 32: each parity takes half of the columns
      const int c_beg = BN >= 32 ? par * half : 0, c_end = BN >= 32 ? c_beg + half : (par == 0 ? BN : 0);
      for (int c = c_beg; c < c_end; c += 16) {
        float mn[16], cr[16];
        tmem_ld16(t_main + lane_sel + (uint32_t)c, mn);
        tmem_ld16(t_cross + lane_sel + (uint32_t)c, cr);
        if (row_ok) {
          float* dst = a.partial + ((size_t)blockIdx.z * a.rows + row) * a.Cb + n0 + c;
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(dst + i) = make_float4(mn[i] + cr[i], mn[i + 1] + cr[i + 1], mn[i + 2] + cr[i + 2], mn[i + 3] + cr[i + 3]);
        }
      }
    } else if (row_ok && par == 0) {
      float* dst = a.partial + ((size_t)blockIdx.z * a.rows + row) * a.Cb + n0;
      for (int c = 0; c < BN; ++c) dst[c] = 0.f;
    }
  } else if (warp < 12) {
    // ============================== B producers: registers -> swizzled shared tiles ==============================
    const int bt = tid - 256;                    // 0..127
    const int chunk = bt & 7, prow = bt >> 3;    // 16-byte chunk of a 128-byte row; pixel rows prow, prow + 16
    const int nblk = (BN + 31) >> 5;
    for (int st = 0; st < nst; ++st) {
      const int slot = st % kWgStages;
      const int kb = kbeg + st * kWgKPix;
      float4 v[8];
      // nblk * 2 float4 per thread (BN = 128 -> 8)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int pix = kb + prow + 16 * i, col = n0 + b * 32 + chunk * 4;
          v[b * 2 + i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (b < nblk && pix < kend && b * 32 + chunk * 4 < BN)
            v[b * 2 + i] = __ldg(reinterpret_cast<const float4*>(a.db + (size_t)pix * a.Cb + col));
        }
      }
      mbar_wait(smem_u32(&s_empty[slot]), (uint32_t)(((st / kWgStages) & 1) ^ 1));
      uint8_t* sB = dsm_gen + (size_t)slot * stage_bytes;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        if (b >= nblk) break;
        const int col = n0 + b * 32 + chunk * 4;
        const bool cok = b * 32 + chunk * 4 < BN;
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f), ce = sh;
        if (a.b_affine && cok) {
          sc = __ldg(reinterpret_cast<const float4*>(a.b_scale + col));
          sh = __ldg(reinterpret_cast<const float4*>(a.b_shift + col));
          if (a.b_center != nullptr) ce = __ldg(reinterpret_cast<const float4*>(a.b_center + col));
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int p = prow + 16 * i;
          float4 x = v[b * 2 + i];
          if (cok && kb + p < kend) {
            if (a.b_affine) {
              x.x = fmaf(x.x - ce.x, sc.x, sh.x); x.y = fmaf(x.y - ce.y, sc.y, sh.y);
              x.z = fmaf(x.z - ce.z, sc.z, sh.z); x.w = fmaf(x.w - ce.w, sc.w, sh.w);
            }
            if (a.b_act) {
              x.x = lrelu(x.x, a.b_slope); x.y = lrelu(x.y, a.b_slope);
              x.z = lrelu(x.z, a.b_slope); x.w = lrelu(x.w, a.b_slope);
            }
          }
          float4 hi, lo;
          split4(x, hi, lo);
          const uint32_t off = (uint32_t)b * (kWgKPix * 128) + sw128b32_off(p, chunk);
          *reinterpret_cast<float4*>(sB + off) = hi;
          *reinterpret_cast<float4*>(sB + half_bytes + off) = lo;
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_full[slot]));
    }
  } else if (lane == 0) {
    // ============================== MMA issuer ==============================
    const uint32_t idesc = make_idesc_tf32(128, BN, 0, 1);   // A: TMEM (K along columns); B: MN-major shared
    const uint32_t blk_stride = kWgKPix * 128;               // bytes between 32-channel blocks of B
    uint32_t acc_m = 0, acc_c = 0;
    for (int st = 0; st < nst; ++st) {
      const int slot = st % kWgStages;
      mbar_wait(smem_u32(&s_full[slot]), (uint32_t)((st / kWgStages) & 1));
      tc_fence_after();
      const uint32_t b_hi = dsm + (uint32_t)slot * stage_bytes, b_lo = b_hi + half_bytes;
      const uint32_t a_hi = t_a + (uint32_t)(slot * 64), a_lo = a_hi + 32;
#pragma unroll
      for (int k = 0; k < kWgKPix / 8; ++k) {
        // MN-major descriptor: leading byte offset = stride between 32-channel blocks, stride byte
        // offset = stride between 4-pixel swizzle atoms; one MMA consumes 8 pixel rows (1024 B)
        const uint64_t dbh = make_smem_desc(b_hi + k * 1024, blk_stride, 512, kLayoutSw128Base32);
        const uint64_t dbl = make_smem_desc(b_lo + k * 1024, blk_stride, 512, kLayoutSw128Base32);
        mma_tf32_ts(t_cross, a_lo + k * 8, dbh, idesc, acc_c);
        mma_tf32_ts(t_cross, a_hi + k * 8, dbl, idesc, 1u);
        mma_tf32_ts(t_main, a_hi + k * 8, dbh, idesc, acc_m);
        acc_m = 1u; acc_c = 1u;
      }
      mma_commit(smem_u32(&s_empty[slot]));
    }
    mma_commit(smem_u32(&s_accum));
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 12) tmem_dealloc(tmem, 512);
}

static int wg_pick_bn(int Cb) {
  for (int bn = 128; bn >= 16; bn >>= 1)
    if (Cb % bn == 0) return bn;
  return 0;
}

}  // namespace cvae

using namespace cvae;

extern "C" int cvae_wgrad_tc_eligible(int pixels, int rows, int Cb) {
  return (Cb % 16 == 0) && pixels >= 512 && rows >= 32;
}

// K-split for the tensor-core weight gradient: enough CTAs to fill the GPU, at most 4096 pixels per
// CTA (bounds the length of one fp32 tensor-core accumulation chain), at least 256.
extern "C" int cvae_wgrad_tc_splits(int pixels, int rows, int Cb) {
  const int bn = wg_pick_bn(Cb);
  if (bn == 0) return 1;
  const int tiles = ((rows + 127) / 128) * (Cb / bn);
  int splits = (2 * kNumSMs + tiles - 1) / tiles;
  splits = max(splits, (pixels + 4095) / 4096);
  splits = min(splits, max(1, pixels / 256));
  return max(1, min(splits, 4096));
}

extern "C" int cvae_conv_wgrad_tc(const cvae_wgrad_params_t* p, cvae_stream_t s) {
  if (!p || !p->ga || !p->db || !p->partial || p->splits < 1) return CVAE_ERR_BAD_ARG;
  if ((p->Ha + 2 * p->pad - p->kh) / p->stride + 1 != p->Hq || (p->Wa + 2 * p->pad - p->kw) / p->stride + 1 != p->Wq)
    return CVAE_ERR_BAD_ARG;
  const int BN = wg_pick_bn(p->Cb);
  if (BN == 0) return CVAE_ERR_UNSUPPORTED_SHAPE;
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
  chunk = ((chunk + kWgKPix - 1) / kWgKPix) * kWgKPix;
  a.kchunk = chunk;
  // fast A-operand loaders (see wg_load_seg / wg_load_lin)
  a.seg = 0;
  if (p->kh == 1 && p->kw == 1 && p->Ha == 1 && p->Wa == 1 && p->stride == 1 && p->pad == 0) a.seg = -1;
  else if (p->Wq % 32 == 0) a.seg = 32;
  else if (p->Wq == 16) a.seg = 16;
  else if (p->Wq == 8) a.seg = 8;
  const size_t smem = (size_t)kWgStages * 2 * kWgKPix * 128 * ((BN + 31) / 32) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) != cudaSuccess)
      return CVAE_ERR_LAUNCH;
    attr_set = true;
  }
  dim3 grid((a.rows + 127) / 128, p->Cb / BN, p->splits);
  wgrad_tc_kernel<<<grid, kWgThreads, smem, as_stream(s)>>>(a, BN);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}
