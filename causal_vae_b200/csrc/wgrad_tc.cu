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
#include <cstdlib>
#include "common.cuh"
#include "conv_args.cuh"
#include "tc_common.cuh"

namespace cvae {

using namespace tc;

constexpr int kWgThreads = 13 * 32;   // 8 A warps + 4 B warps + 1 MMA warp
constexpr int kWgStages = 4;
constexpr int kWgKPix = 32;           // pixels (K) per stage

// A-operand loaders: NP consecutive pixels of one (tap, channel) row into registers; bit p of the
// returned mask is set where the value is real data (padding / out-of-range pixels stay exactly 0).
// wg_load_seg: the run is made of whole output-row segments of min(SEG, NP) pixels (Wq % SEG == 0, runs
// start on multiples of NP pixels), so row / image coordinates and the vertical bound are computed once
// per segment instead of once per pixel (the per-pixel form, ~25 instructions per value, made this
// producer the bottleneck of the kernel).  wg_load_lin: 1x1 / Linear (pixel = row of the matrix).
template <int SEG, int NP>
__device__ __forceinline__ uint32_t wg_load_seg(const WgradArgs& a, int kb, int kend, bool row_ok, int dh, int dw, int ca,
                                                float (&v)[NP]) {
  constexpr int L = SEG < NP ? SEG : NP;
  uint32_t okm = 0;
#pragma unroll
  for (int sg = 0; sg < NP / L; ++sg) {
    const int pixs = kb + sg * L;
    const int qw0 = pixs % a.Wq, t = pixs / a.Wq, qh = t % a.Hq, n = t / a.Hq;
    const int ih = qh * a.stride + dh;
    const bool hok = row_ok && pixs < kend && (unsigned)ih < (unsigned)a.Ha;
    const float* rowp = a.ga + (((size_t)n * a.Ha + (hok ? ih : 0)) * a.Wa) * a.Ca + ca;
    int iw = qw0 * a.stride + dw;
#pragma unroll
    for (int j = 0; j < L; ++j) {
      const bool ok = hok && (unsigned)iw < (unsigned)a.Wa;
      v[sg * L + j] = 0.f;
      if (ok) { v[sg * L + j] = __ldg(rowp + (size_t)iw * a.Ca); okm |= 1u << (sg * L + j); }
      iw += a.stride;
    }
  }
  return okm;
}
template <int NP>
__device__ __forceinline__ uint32_t wg_load_lin(const WgradArgs& a, int kb, int kend, bool row_ok, int ca, float (&v)[NP]) {
  uint32_t okm = 0;
  const float* src = a.ga + (size_t)kb * a.Ca + ca;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    v[p] = 0.f;
    if (row_ok && kb + p < kend) { v[p] = __ldg(src + (size_t)p * a.Ca); okm |= 1u << p; }
  }
  return okm;
}
template <int NP>
__device__ __forceinline__ uint32_t wg_load_any(const WgradArgs& a, int kb, int kend, bool row_ok, int dh, int dw, int ca,
                                                float (&v)[NP]) {
  if (a.seg == 32) return wg_load_seg<32, NP>(a, kb, kend, row_ok, dh, dw, ca, v);
  if (a.seg == 16) return wg_load_seg<16, NP>(a, kb, kend, row_ok, dh, dw, ca, v);
  if (a.seg == 8) return wg_load_seg<8, NP>(a, kb, kend, row_ok, dh, dw, ca, v);
  if (a.seg == -1) return wg_load_lin<NP>(a, kb, kend, row_ok, ca, v);
  uint32_t okm = 0;
  int qw = kb % a.Wq, t = kb / a.Wq, qh = t % a.Hq, n = t / a.Hq;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const int ih = qh * a.stride + dh, iw = qw * a.stride + dw;
    const bool ok = row_ok && (kb + p) < kend && (unsigned)ih < (unsigned)a.Ha && (unsigned)iw < (unsigned)a.Wa;
    v[p] = 0.f;
    if (ok) {
      v[p] = __ldg(a.ga + (((size_t)n * a.Ha + ih) * a.Wa + iw) * a.Ca + ca);
      okm |= 1u << p;
    }
    if (++qw == a.Wq) { qw = 0; if (++qh == a.Hq) { qh = 0; ++n; } }
  }
  return okm;
}

// Stateful loader for the segment modes: the position (output column / row / image) of the NEXT run is
// carried along and advanced incrementally (the div / mod form cost ~100 instructions per 16 pixels), and a
// run that lies inside the input row is loaded with unpredicated strided loads (the predicated form spent
// ~15 instructions per value on bounds, mask and 64-bit address arithmetic - ncu, profiles/).
struct WgPos { int pix, qw, qh, n; };
__device__ __forceinline__ void wg_advance(WgPos& p, int d, int Wq, int Hq) {
  p.pix += d; p.qw += d;
  while (p.qw >= Wq) { p.qw -= Wq; if (++p.qh == Hq) { p.qh = 0; ++p.n; } }
}
template <int L>
__device__ __forceinline__ uint32_t wg_load_run(const WgradArgs& a, const WgPos& p, int kend, bool row_ok, int dh, int dw,
                                                int ca, float* v) {
  const int ih = p.qh * a.stride + dh;
  const bool hok = row_ok && p.pix < kend && (unsigned)ih < (unsigned)a.Ha;
  const int iw0 = p.qw * a.stride + dw;
  if (!hok) {
#pragma unroll
    for (int j = 0; j < L; ++j) v[j] = 0.f;
    return 0u;
  }
  const float* ptr = a.ga + (((long long)p.n * a.Ha + ih) * a.Wa + iw0) * a.Ca + ca;   // dereferenced only where valid
  const long long step = (long long)a.stride * a.Ca;
  if (iw0 >= 0 && iw0 + (L - 1) * a.stride < a.Wa) {
#pragma unroll
    for (int j = 0; j < L; ++j) v[j] = __ldg(ptr + j * step);
    return L == 32 ? 0xFFFFFFFFu : ((1u << L) - 1u);
  }
  uint32_t okm = 0;
  int iw = iw0;
#pragma unroll
  for (int j = 0; j < L; ++j) {
    v[j] = 0.f;
    if ((unsigned)iw < (unsigned)a.Wa) { v[j] = __ldg(ptr + j * step); okm |= 1u << j; }
    iw += a.stride;
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
    // Software pipeline over HALF stages (16 pixels): while one half is transformed, split into tf32
    // hi / lo and stored to tensor memory, the loads of the next two halves are in flight (ncu: 42 % of
    // the stall samples of the unpipelined loop sat on the first use of the loaded values).  Three
    // 16-value buffers + one 16-value store staging array = the register budget of the old 32 + 32.
    constexpr int kH = kWgKPix / 2;
    const int nmine = nst > par ? (nst - par + 1) / 2 : 0;      // stages par, par + 2, ...
    const int nh = 2 * nmine;
    WgPos pos;                                   // position of the next half to load (segment modes)
    pos.pix = kbeg + par * kWgKPix;
    pos.qw = pos.pix % a.Wq; { const int t = pos.pix / a.Wq; pos.qh = t % a.Hq; pos.n = t / a.Hq; }
    auto load_half = [&](int u, float (&v)[kH]) -> uint32_t {
      uint32_t m = 0;
      if (a.seg >= 16) {                         // one 16-pixel run inside an output row
        m = wg_load_run<16>(a, pos, kend, row_ok, dh, dw, ca, v);
        wg_advance(pos, (u & 1) ? 16 + kWgKPix : 16, a.Wq, a.Hq);
      } else if (a.seg == 8) {                   // two runs of 8 (8-wide feature maps)
        m = wg_load_run<8>(a, pos, kend, row_ok, dh, dw, ca, v);
        wg_advance(pos, 8, a.Wq, a.Hq);
        m |= wg_load_run<8>(a, pos, kend, row_ok, dh, dw, ca, v + 8) << 8;
        wg_advance(pos, (u & 1) ? 8 + kWgKPix : 8, a.Wq, a.Hq);
      } else {
        const int kb = kbeg + (par + 2 * (u >> 1)) * kWgKPix + (u & 1) * kH;
        m = wg_load_any<kH>(a, kb, kend, row_ok, dh, dw, ca, v);
      }
      return m;
    };
    float b0[kH], b1[kH], b2[kH];
    uint32_t m0 = 0, m1 = 0, m2 = 0;
    if (nh > 0) m0 = load_half(0, b0);
    if (nh > 1) m1 = load_half(1, b1);
    for (int u = 0; u < nh; ++u) {
      if (u + 2 < nh) m2 = load_half(u + 2, b2);
      const int st = par + 2 * (u >> 1), slot = st % kWgStages, h = (u & 1) * kH;
      if (m0 == 0xFFFFu) {                       // interior run (warp-uniform when Ca % 32 == 0)
#pragma unroll
        for (int p = 0; p < kH; ++p) {
          if (a.a_affine) b0[p] = fmaf(b0[p] - ce, sc, sh);
          if (a.a_act) b0[p] = fmaxf(b0[p], b0[p] * a.a_slope);      // slope in [0, 1] (launcher)
        }
      } else if (m0 != 0u) {
#pragma unroll
        for (int p = 0; p < kH; ++p) {
          if ((m0 >> p) & 1u) {                  // padding stays exactly 0
            if (a.a_affine) b0[p] = fmaf(b0[p] - ce, sc, sh);
            if (a.a_act) b0[p] = fmaxf(b0[p], b0[p] * a.a_slope);
          }
        }
      }
      if ((u & 1) == 0) {
        mbar_wait(smem_u32(&s_empty[slot]), (uint32_t)(((st / kWgStages) & 1) ^ 1));
        tc_fence_after();
      }
      uint32_t t16[kH];
#pragma unroll
      for (int p = 0; p < kH; ++p) {
        const float hi = tf32_rn(b0[p]);
        t16[p] = __float_as_uint(hi);
        b0[p] = tf32_rn(b0[p] - hi);
      }
      tmem_st16(t_a + lane_sel + (uint32_t)(slot * 64 + h), t16);
      tmem_st16(t_a + lane_sel + (uint32_t)(slot * 64 + 32 + h), reinterpret_cast<const uint32_t*>(b0));
      if (u & 1) {
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s_full[slot]));
      }
#pragma unroll
      for (int p = 0; p < kH; ++p) { b0[p] = b1[p]; b1[p] = b2[p]; }
      m0 = m1; m1 = m2;
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
        if (a.direct != nullptr) {
          // direct mode: the K splits of this tile meet in the zeroed gradient (torch layout [cb][ca][tap]).  Lanes are
          // consecutive input channels: stride taps * 4 B (Linear: one 128-byte line per warp and column)
          if (row_ok && ca < a.ca_real) {
            const size_t cstride = (size_t)a.ca_real * a.taps;
            float* dst = a.direct + (size_t)(n0 + c) * cstride + (size_t)ca * a.taps + tap;
#pragma unroll
            for (int i = 0; i < 16; ++i) atomicAdd(dst + (size_t)i * cstride, mn[i] + cr[i]);
          }
        } else if (row_ok) {
          float* dst = a.partial + ((size_t)blockIdx.z * a.rows + row) * a.Cb + n0 + c;
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(dst + i) = make_float4(mn[i] + cr[i], mn[i + 1] + cr[i + 1], mn[i + 2] + cr[i + 2], mn[i + 3] + cr[i + 3]);
        }
      }
    } else if (row_ok && par == 0 && a.direct == nullptr) {
      float* dst = a.partial + ((size_t)blockIdx.z * a.rows + row) * a.Cb + n0;
      for (int c = 0; c < BN; ++c) dst[c] = 0.f;
    }
  } else if (warp < 12) {
    // ============================== B producers: registers -> swizzled shared tiles ==============================
    const int bt = tid - 256;                    // 0..127
    const int chunk = bt & 7, prow = bt >> 3;    // 16-byte chunk of a 128-byte row; pixel rows prow, prow + 16
    const int nblk = (BN + 31) >> 5;
    // nblk * 2 float4 per thread and stage (BN = 128 -> 8).  The loads of stage st + 1 are issued before stage st is
    // transformed and stored: with one stage of loads in flight the four B warps paid a full DRAM round trip per stage
    // and were the critical path of the kernel (ncu, profiles/r2_ncu_wgrad_tc_raw.txt: the first use of the loaded vectors
    // carried the stall samples; 3.2 kclk per 32-pixel stage against 0.8 kclk of MMA time).
    auto load_stage = [&](int st, float4 (&v)[8]) {
      const int kb = kbeg + st * kWgKPix;
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
    };
    float4 v[8], vn[8];
    if (nst > 0) load_stage(0, v);
    for (int st = 0; st < nst; ++st) {
      const int slot = st % kWgStages;
      const int kb = kbeg + st * kWgKPix;
      if (st + 1 < nst) load_stage(st + 1, vn);
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
              x.x = fmaxf(x.x, x.x * a.b_slope); x.y = fmaxf(x.y, x.y * a.b_slope);      // slope in [0, 1] (launcher)
              x.z = fmaxf(x.z, x.z * a.b_slope); x.w = fmaxf(x.w, x.w * a.b_slope);
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
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = vn[i];
    }
  } else {
    // ============================== MMA issuer: converged warp, one elected lane issues (see conv_halo_tc.cu) ==============================
    const bool leader = elect_one();
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
        if (leader) mma_tf32_ts(t_cross, a_lo + k * 8, dbh, idesc, acc_c);
        if (leader) mma_tf32_ts(t_cross, a_hi + k * 8, dbl, idesc, 1u);
        if (leader) mma_tf32_ts(t_main, a_hi + k * 8, dbh, idesc, acc_m);
        acc_m = 1u; acc_c = 1u;
      }
      if (leader) mma_commit(smem_u32(&s_empty[slot]));
    }
    if (leader) mma_commit(smem_u32(&s_accum));
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
  // One CTA per SM: the grid runs in waves of 148, so 300 CTAs cost three waves (ncu: grid (5,1,60) ran
  // 2 full waves + 4 CTAs).  Pick the split count that minimises waves x (stages per CTA + fixed
  // prologue / epilogue cost) within [pixels / 4096, pixels / 256].
  const int smin = max(1, (pixels + 4095) / 4096), smax = max(smin, min(4096, pixels / 256));
  if (const char* e = getenv("CVAE_WG_SPLITS")) {            // tuning hook (scripts/bench_layers.py sweeps)
    const int v = atoi(e);
    if (v > 0) return max(smin, min(smax, v));
  }
  int best = smin;
  long long best_cost = -1;
  for (int sp = smin; sp <= smax; ++sp) {
    const long long waves = ((long long)tiles * sp + kNumSMs - 1) / kNumSMs;
    const long long stages = ((pixels + sp - 1) / sp + kWgKPix - 1) / kWgKPix + 8;   // + ~8 stages of fixed cost
    const long long cost = waves * stages;
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = sp; }
  }
  return best;
}

static int wgrad_tc_launch(const cvae_wgrad_params_t* p, float* direct, int ca_real, cvae_stream_t s) {
  if (!p || !p->ga || !p->db || (!p->partial && !direct) || p->splits < 1) return CVAE_ERR_BAD_ARG;
  if (direct && (ca_real < 1 || ca_real > p->Ca)) return CVAE_ERR_BAD_ARG;
  if ((p->Ha + 2 * p->pad - p->kh) / p->stride + 1 != p->Hq || (p->Wa + 2 * p->pad - p->kw) / p->stride + 1 != p->Wq)
    return CVAE_ERR_BAD_ARG;
  const int BN = wg_pick_bn(p->Cb);
  if (BN == 0) return CVAE_ERR_UNSUPPORTED_SHAPE;
  if ((p->xa.slope != 1.0f && !(p->xa.slope >= 0.f && p->xa.slope <= 1.f)) ||
      (p->xb.slope != 1.0f && !(p->xb.slope >= 0.f && p->xb.slope <= 1.f)))
    return CVAE_ERR_UNSUPPORTED_SHAPE;          // the producers evaluate the leaky ReLU as max(v, slope * v)
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
  a.direct = direct; a.ca_real = ca_real; a.taps = p->kh * p->kw;
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

extern "C" int cvae_conv_wgrad_tc(const cvae_wgrad_params_t* p, cvae_stream_t s) { return wgrad_tc_launch(p, nullptr, 0, s); }

extern "C" int cvae_conv_wgrad_tc_direct(const cvae_wgrad_params_t* p, float* grad, int ca_real, cvae_stream_t s) {
  if (!grad) return CVAE_ERR_BAD_ARG;
  return wgrad_tc_launch(p, grad, ca_real, s);
}
