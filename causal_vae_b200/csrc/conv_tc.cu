// tcgen05 / TMEM implicit-GEMM convolution family for sm_100a (3xTF32: fp32-grade accuracy on the
// 5th-generation tensor cores).
//
//   D[pixel, cout] = sum_{tap, cin} xform(src[gather(pixel, tap), cin]) * W[tap][cin][cout]
//
// One CTA computes a 128-pixel x BN-channel tile with the accumulator in tensor memory.
//   * A operand (activations): 8 producer warps gather 128 pixels x 32 channels of one tap per
//     stage with 128-bit loads, apply the producer layer's BatchNorm + LeakyReLU in registers
//     (training-mode BN cannot be folded: it needs the whole batch first), split every value into
//     tf32 hi + lo and store both as K-major, 128B-swizzled UMMA tiles.  There is no register stage
//     between TMA and the MMA, which is why this operand is staged by warps and not by TMA.
//   * B operand (weights): pre-split / pre-swizzled by tc_pack_weight_kernel into the exact shared
//     memory image, so a stage is two 1-D bulk copies (TMA) completing on the stage's mbarrier.
//   * one elected thread issues hi*hi + lo*hi + hi*lo tcgen05.mma (kind::tf32, M=128, N=BN, K=8)
//     per 8 channels; tcgen05.commit releases the stage / publishes the accumulator.
//   * epilogue: TMEM -> registers -> padded shared tile -> coalesced 128-bit stores with the fused
//     bias, BatchNorm statistics (double) or activation-derivative + BN-backward sums.
#include "common.cuh"
#include "conv_args.cuh"
#include "tc_common.cuh"
#include <cstdlib>

namespace cvae {

using namespace tc;

constexpr int kTcProdWarps = 8;                 // 2 producer groups of 4 warps
constexpr int kTcThreads = (kTcProdWarps + 1 + 4) * 32;   // + 1 MMA warp + 4 epilogue warps
constexpr int kTcAStage = 2 * 128 * 128;        // hi + lo tiles of 128 rows x 128 bytes
constexpr int kEpiLd = 36;                      // padded row stride (floats) of the epilogue staging tile
constexpr int kEpiBytes = 128 * kEpiLd * 4;

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// Static persistent schedule: CTA b owns tiles b, b + grid, ...; a tile is (m-tile fastest, n-tile,
// phase).  Every warp role walks the same list and skips the same out-of-range tiles.
struct TcTile { int ph, n0, m0, M, T; };
__device__ __forceinline__ bool tc_decode(const GatherArgs& a, int t, int tiles_m, int tiles_n, int BN, int KB, TcTile& o) {
  const int mt = t % tiles_m, r = t / tiles_m;
  o.ph = r / tiles_n;
  o.n0 = (r % tiles_n) * BN;
  o.m0 = mt * 128;
  const PhaseGeom& P = a.phase[o.ph];
  o.M = a.N * P.Hq * P.Wq;
  o.T = P.ntaps * KB;
  return o.m0 < o.M;
}

__global__ void __launch_bounds__(kTcThreads, 1)
igemm_tc_kernel(const __grid_constant__ GatherArgs a, const int BN, const int NS, const int tiles_m, const int tiles_n) {
  extern __shared__ uint8_t dsm_raw[];
  __shared__ __align__(8) uint64_t s_full[8];
  __shared__ __align__(8) uint64_t s_empty[8];
  __shared__ __align__(8) uint64_t s_tfull[2];
  __shared__ __align__(8) uint64_t s_tempty[2];
  __shared__ uint32_t s_tmem;
  __shared__ int s_out[128];
  __shared__ double s_part[4][64];
  __shared__ double s_stat[512];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KB = (a.Cs + 31) >> 5;
  const int total = tiles_m * tiles_n * a.nphase;
  const uint32_t stage_bytes = kTcAStage + 256u * BN;
  uint32_t tmem_cols = 32;
  // cross terms (lo*hi' + hi*lo') in their own accumulator columns [BN, 2BN) of each buffer, added to the main chain in
  // the epilogue: the tensor core truncates on accumulate, so the 2^-11-scaled terms must not ride the long main chain
  // (as in conv_halo_tc.cu / wgrad_tc.cu; the launcher keeps BN <= 128 so that 2 buffers x 2 BN columns fit)
  const uint32_t accw = 2u * BN;
  while (tmem_cols < 2u * accw) tmem_cols <<= 1;
  uint8_t* dsm_gen = dsm_raw + ((1024u - (smem_u32(dsm_raw) & 1023u)) & 1023u);
  const uint32_t dsm = smem_u32(dsm_gen);

  for (int i = tid; i < 512; i += kTcThreads) s_stat[i] = 0.0;
  if (warp == kTcProdWarps) {
    if (lane == 0) {
      for (int i = 0; i < NS; ++i) { mbar_init(smem_u32(&s_full[i]), 5); mbar_init(smem_u32(&s_empty[i]), 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&s_tfull[i]), 1); mbar_init(smem_u32(&s_tempty[i]), 4); }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(&s_tmem), tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  if (warp < kTcProdWarps) {
    // ============================== A / B producers ==============================
    // Group g fills the k-blocks with (running index % 2 == g).  The global loads of a group's
    // NEXT k-block are issued before its current one is transformed and stored, so a full memory
    // latency is always covered by work.
    const int grp = warp >> 2, wq = warp & 3;
    const int chunk = lane & 7, rsub = lane >> 3;

    // ---- load cursor (one k-block ahead of the store cursor) ----
    int l_t = blockIdx.x, l_kb = grp;          // tile index, k-block within the tile
    TcTile lt;
    bool l_live = false;
    int pix[8], hw[8];
    auto l_enter = [&]() {                     // position on the first valid tile at / after l_t holding l_kb
      while (l_t < total) {
        if (tc_decode(a, l_t, tiles_m, tiles_n, BN, KB, lt)) {
          if (l_kb < lt.T) break;
          l_kb -= lt.T;
        }
        l_t += gridDim.x;
      }
      l_live = l_t < total;
      if (l_live) {
        const PhaseGeom& P = a.phase[lt.ph];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int m = lt.m0 + wq * 32 + j * 4 + rsub;
          if (m < lt.M) {
            const int qw = m % P.Wq, t = m / P.Wq, qh = t % P.Hq, n = t / P.Hq;
            pix[j] = (n * a.Hs + qh * a.is) * a.Ws + qw * a.is;
            hw[j] = (qh * a.is) | ((qw * a.is) << 16);
          } else {
            pix[j] = -1; hw[j] = 0;
          }
        }
      }
    };
    float4 v[8];
    uint32_t okmask = 0;
    auto l_issue = [&]() {
      const TapEntry tap = a.phase[lt.ph].taps[l_kb / KB];
      const int c = (l_kb % KB) * 32 + chunk * 4;
      const bool cok = c < a.Cs;
      const int doff = tap.dh * a.Ws + tap.dw;
      okmask = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ih = (hw[j] & 0xffff) + tap.dh, iw = (hw[j] >> 16) + tap.dw;
        const bool ok = cok && pix[j] >= 0 && (unsigned)ih < (unsigned)a.Hs && (unsigned)iw < (unsigned)a.Ws;
        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) {
          v[j] = __ldg(reinterpret_cast<const float4*>(a.src + (size_t)(pix[j] + doff) * a.Cs + c));
          okmask |= 1u << j;
        }
      }
    };

    // ---- store cursor ----
    int s_kb = grp;
    uint32_t it = grp;                          // running k-block index of this CTA
    TcTile st;
    l_enter();
    if (l_live) l_issue();
    while (l_live) {
      // the store cursor takes over the load cursor's position and data
      s_kb = l_kb; st = lt;
      float4 cur[8];
      const uint32_t cur_ok = okmask;
#pragma unroll
      for (int j = 0; j < 8; ++j) cur[j] = v[j];
      // advance the load cursor by this group's stride (2 k-blocks) and issue its loads
      l_kb += 2;
      if (l_kb >= lt.T) { l_kb -= lt.T; l_t += gridDim.x; l_enter(); }
      if (l_live) l_issue();

      const int slot = it % NS;
      const uint32_t full = smem_u32(&s_full[slot]);
      mbar_wait(smem_u32(&s_empty[slot]), (uint32_t)(((it / NS) & 1) ^ 1));
      const TapEntry tap = a.phase[st.ph].taps[s_kb / KB];
      const int kbk = s_kb % KB;
      if (wq == 0 && lane == 0) {
        const uint32_t bbytes = 128u * BN;
        const uint32_t sB = dsm + (uint32_t)slot * stage_bytes + kTcAStage;
        mbar_arrive_expect_tx(full, 2u * bbytes);
        const float* wsrc = a.wt + (((size_t)tap.widx * KB + kbk) * 2) * (size_t)a.Cd * 32 + (size_t)st.n0 * 32;
        bulk_g2s(sB, wsrc, bbytes, full);
        bulk_g2s(sB + bbytes, wsrc + (size_t)a.Cd * 32, bbytes, full);
      }
      const int c = kbk * 32 + chunk * 4;
      float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f), ce = sh;
      if (a.in_affine && c < a.Cs) {
        sc = __ldg(reinterpret_cast<const float4*>(a.in_scale + c));
        sh = __ldg(reinterpret_cast<const float4*>(a.in_shift + c));
        if (a.in_center != nullptr) ce = __ldg(reinterpret_cast<const float4*>(a.in_center + c));
      }
      uint8_t* sA = dsm_gen + (size_t)slot * stage_bytes;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 x = cur[j];
        if ((cur_ok >> j) & 1u) {   // padding stays exactly 0
          if (a.in_affine) {
            x.x = fmaf(x.x - ce.x, sc.x, sh.x); x.y = fmaf(x.y - ce.y, sc.y, sh.y);
            x.z = fmaf(x.z - ce.z, sc.z, sh.z); x.w = fmaf(x.w - ce.w, sc.w, sh.w);
          }
          if (a.in_act) {
            x.x = lrelu(x.x, a.in_slope); x.y = lrelu(x.y, a.in_slope);
            x.z = lrelu(x.z, a.in_slope); x.w = lrelu(x.w, a.in_slope);
          }
        }
        float4 hi, lo;
        split4(x, hi, lo);
        const uint32_t off = sw128_off(wq * 32 + j * 4 + rsub, chunk);
        *reinterpret_cast<float4*>(sA + off) = hi;
        *reinterpret_cast<float4*>(sA + 16384 + off) = lo;
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(full);
      it += 2;
    }
  } else if (warp == kTcProdWarps) {
    // ============================== MMA issuer: converged warp, one elected lane issues (see conv_halo_tc.cu) ==============================
    {
      const bool leader = elect_one();
      const uint32_t idesc = make_idesc_tf32(128, BN, 0, 0);
      uint32_t it = 0, tcount = 0;
      TcTile tl;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        if (!tc_decode(a, t, tiles_m, tiles_n, BN, KB, tl)) continue;
        const uint32_t acc = tcount & 1u;
        mbar_wait(smem_u32(&s_tempty[acc]), ((tcount >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem + acc * accw;
        uint32_t accumulate = 0;
        for (int kb = 0; kb < tl.T; ++kb, ++it) {
          const int slot = it % NS;
          mbar_wait(smem_u32(&s_full[slot]), (it / NS) & 1u);
          tc_fence_after();
          const int c0 = (kb % KB) * 32;
          const int ksteps = min(32, a.Cs - c0) >> 3;
          const uint32_t a_hi = dsm + (uint32_t)slot * stage_bytes, a_lo = a_hi + 16384;
          const uint32_t b_hi = a_hi + kTcAStage, b_lo = b_hi + 128u * BN;
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t dah = make_smem_desc(a_hi + k * 32, 16, 1024, kLayoutSw128);
            const uint64_t dal = make_smem_desc(a_lo + k * 32, 16, 1024, kLayoutSw128);
            const uint64_t dbh = make_smem_desc(b_hi + k * 32, 16, 1024, kLayoutSw128);
            const uint64_t dbl = make_smem_desc(b_lo + k * 32, 16, 1024, kLayoutSw128);
            if (leader) mma_tf32(d_tmem + (uint32_t)BN, dal, dbh, idesc, accumulate);
            if (leader) mma_tf32(d_tmem + (uint32_t)BN, dah, dbl, idesc, 1u);
            if (leader) mma_tf32(d_tmem, dah, dbh, idesc, accumulate);
            accumulate = 1u;
          }
          if (leader) mma_commit(smem_u32(&s_empty[slot]));
        }
        if (leader) mma_commit(smem_u32(&s_tfull[acc]));
        ++tcount;
      }
    }
  } else {
    // ============================== epilogue warps ==============================
    const int q = warp & 3;                         // TMEM lane quarter this warp may read
    const int gt = (warp - (kTcProdWarps + 1)) * 32 + lane;
    const int ew = warp - (kTcProdWarps + 1);
    float* ebuf = reinterpret_cast<float*>(dsm_gen + (size_t)NS * stage_bytes);
    const int nchunks = (BN + 31) >> 5;
    const bool want_stats = a.epi != CVAE_EPI_PLAIN && a.stats != nullptr;
    uint32_t tcount = 0;
    int cur_n0 = -1;
    TcTile tl;
    auto flush_stats = [&]() {
      named_bar_sync(1, 128);
      for (int i = gt; i < 2 * BN; i += 128) {
        const double sv = s_stat[i];
        if (sv != 0.0) atomicAdd(a.stats + (i < BN ? cur_n0 + i : a.Cd + cur_n0 + i - BN), sv);
        s_stat[i] = 0.0;
      }
      named_bar_sync(1, 128);
    };
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      if (!tc_decode(a, t, tiles_m, tiles_n, BN, KB, tl)) continue;
      if (want_stats && cur_n0 >= 0 && cur_n0 != tl.n0) flush_stats();
      cur_n0 = tl.n0;
      {
        const PhaseGeom& P = a.phase[tl.ph];
        const int m = tl.m0 + q * 32 + lane;
        int o = -1;
        if (m < tl.M) {
          const int qw = m % P.Wq, tt = m / P.Wq, qh = tt % P.Hq, n = tt / P.Hq;
          o = (n * a.Hd + qh * a.os + P.ph) * a.Wd + qw * a.os + P.pw;
        }
        s_out[q * 32 + lane] = o;
      }
      const uint32_t acc = tcount & 1u;
      mbar_wait(smem_u32(&s_tfull[acc]), (tcount >> 1) & 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem + acc * accw + ((uint32_t)(q * 32) << 16);
      for (int ch = 0; ch < nchunks; ++ch) {
        const int cw = min(32, BN - ch * 32);
        {  // TMEM -> padded shared tile (thread = accumulator row)
          const int row = q * 32 + lane;
          for (int h = 0; h < cw; h += 16) {
            float r16[16], c16[16];
            tmem_ld16(d_tmem + (uint32_t)(ch * 32 + h), r16);
            tmem_ld16(d_tmem + (uint32_t)(BN + ch * 32 + h), c16);
#pragma unroll
            for (int i = 0; i < 16; ++i) r16[i] += c16[i];
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              *reinterpret_cast<float4*>(ebuf + row * kEpiLd + h + i) = make_float4(r16[i], r16[i + 1], r16[i + 2], r16[i + 3]);
          }
        }
        if (ch == nchunks - 1) {   // accumulator drained: hand the TMEM buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&s_tempty[acc]));
        }
        named_bar_sync(1, 128);
        const int cgs = cw >> 2, rstep = 128 / cgs;
        const int cg = gt % cgs, r0 = gt / cgs;
        const int col = tl.n0 + ch * 32 + cg * 4;
        float4 bias = make_float4(0.f, 0.f, 0.f, 0.f), esc = make_float4(1.f, 1.f, 1.f, 1.f), esh = bias, ece = bias;
        if (a.bias != nullptr) bias = __ldg(reinterpret_cast<const float4*>(a.bias + col));
        if (a.e_affine) {
          esc = __ldg(reinterpret_cast<const float4*>(a.e_scale + col));
          esh = __ldg(reinterpret_cast<const float4*>(a.e_shift + col));
          if (a.e_center != nullptr) ece = __ldg(reinterpret_cast<const float4*>(a.e_center + col));
        }
        double s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
        for (int p0 = 0; p0 < cgs; p0 += 4) {     // 4 rows per batch: their global loads overlap
          int o[4];
          float4 r4[4], d4[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            o[u] = s_out[r0 + (p0 + u) * rstep];
            r4[u] = make_float4(0.f, 0.f, 0.f, 0.f); d4[u] = r4[u];
            if (a.epi == CVAE_EPI_DACT && o[u] >= 0) {
              const size_t goff = (size_t)o[u] * a.Cd + col;
              r4[u] = __ldg(reinterpret_cast<const float4*>(a.epi_ref + goff));
              if (a.epi_add != nullptr) d4[u] = __ldg(reinterpret_cast<const float4*>(a.epi_add + goff));
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (o[u] < 0) continue;
            const float4 t4 = *reinterpret_cast<const float4*>(ebuf + (r0 + (p0 + u) * rstep) * kEpiLd + cg * 4);
            float x[4] = {t4.x + bias.x, t4.y + bias.y, t4.z + bias.z, t4.w + bias.w};
            if (a.epi == CVAE_EPI_STATS) {
#pragma unroll
              for (int j = 0; j < 4; ++j) { s1[j] += (double)x[j]; s2[j] += (double)x[j] * (double)x[j]; }
            } else if (a.epi == CVAE_EPI_DACT) {
              const float refc[4] = {r4[u].x - ece.x, r4[u].y - ece.y, r4[u].z - ece.z, r4[u].w - ece.w};
              x[0] += d4[u].x; x[1] += d4[u].y; x[2] += d4[u].z; x[3] += d4[u].w;
              const float z[4] = {fmaf(refc[0], esc.x, esh.x), fmaf(refc[1], esc.y, esh.y), fmaf(refc[2], esc.z, esh.z),
                                  fmaf(refc[3], esc.w, esh.w)};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                x[j] = z[j] > 0.f ? x[j] : x[j] * a.e_slope;
                s1[j] += (double)x[j]; s2[j] += (double)x[j] * (double)refc[j];
              }
            }
            *reinterpret_cast<float4*>(a.dst + (size_t)o[u] * a.Cd + col) = make_float4(x[0], x[1], x[2], x[3]);
          }
        }
        if (want_stats) {
          // lanes sharing a column group are lane, lane + cgs, ...: xor-reduce over offsets >= cgs
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            for (int off = 16; off >= cgs; off >>= 1) {
              s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], off);
              s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], off);
            }
          }
          if (lane < cgs) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { s_part[ew][cg * 4 + j] = s1[j]; s_part[ew][32 + cg * 4 + j] = s2[j]; }
          }
          named_bar_sync(1, 128);
          if (gt < 2 * cw) {
            const int which = gt / cw, cc = gt % cw;
            const double sv = s_part[0][which * 32 + cc] + s_part[1][which * 32 + cc] + s_part[2][which * 32 + cc] +
                              s_part[3][which * 32 + cc];
            s_stat[which * BN + ch * 32 + cc] += sv;
          }
        }
        named_bar_sync(1, 128);   // staging tile / s_part / s_out free for reuse
      }
      ++tcount;
    }
    if (want_stats && cur_n0 >= 0) flush_stats();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kTcProdWarps) tmem_dealloc(tmem, tmem_cols);
}

// weights -> [tap][k-block of 32][hi | lo][Cd rows][32 floats, 128B-swizzled by (row & 7)]
__global__ void tc_pack_weight_kernel(const float* __restrict__ src, float* __restrict__ dst, int A, int A_pad, int B,
                                      int taps, int src_bat, int src_ld) {
  const int KB = (A_pad + 31) >> 5;
  const size_t total = (size_t)taps * KB * 2 * B * 32;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int kk = (int)(i & 31);
    size_t r = i >> 5;
    const int n = (int)(r % B); r /= B;
    const int h = (int)(r & 1); r >>= 1;
    const int kb = (int)(r % KB);
    const int tap = (int)(r / KB);
    const int lc = (kk >> 2) ^ (n & 7);
    const int k = kb * 32 + lc * 4 + (kk & 3);
    float v = 0.f;
    if (k < A && (src_bat || n < src_ld))
      v = src_bat ? src[((size_t)n * src_ld + k) * taps + tap] : src[((size_t)k * src_ld + n) * taps + tap];
    float hi, lo;
    split_tf32(v, hi, lo);
    dst[i] = h ? lo : hi;
  }
}

// Rows of a plain [M][K] matrix -> the tensor-core A-operand image: per (128-row tile, 32-channel k-block) one 32 KB
// block = tf32 hi plane then lo plane, each 128 rows x 128 bytes in the 128B swizzle -- exactly what a stage of the
// halo kernel's A ring holds, so the kernel fetches it with two bulk copies and needs no producer warps.
__global__ void __launch_bounds__(256) tc_pack_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, long long M,
                                                           int K, int KB) {
  const int mt = blockIdx.x / KB, kb = blockIdx.x - mt * KB;
  uint8_t* tile = reinterpret_cast<uint8_t*>(dst) + ((size_t)mt * KB + kb) * 32768;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int item = threadIdx.x + i * 256, r = item >> 3, chunk = item & 7;
    const long long row = (long long)mt * 128 + r;
    const int c = kb * 32 + chunk * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < M && c < K) v = __ldg(reinterpret_cast<const float4*>(src + row * K + c));
    float4 hi, lo;
    tc::split4(v, hi, lo);
    const uint32_t off = tc::sw128_off(r, chunk);
    *reinterpret_cast<float4*>(tile + off) = hi;
    *reinterpret_cast<float4*>(tile + 16384 + off) = lo;
  }
}

int build_geom(const cvae_conv_params_t* p, GatherArgs& g);  // conv.cu
int launch_conv_halo_tc(const GatherArgs& g, cudaStream_t st);   // conv_halo_tc.cu (conv-shaped layers, k > 1)

static int tc_pick_bn(int Cd) {
  for (int bn = 128; bn >= 16; bn >>= 1)
    if (Cd % bn == 0) return bn;
  return 0;
}

}  // namespace cvae

using namespace cvae;

extern "C" int cvae_tc_eligible(int Cs, int Cd, int64_t M) {
  return (Cs % 16 == 0) && (Cd % 16 == 0) && Cs >= 16 && M >= 1;
}

extern "C" int64_t cvae_tc_pack_floats(int A_pad, int B, int taps) {
  return (int64_t)taps * ((A_pad + 31) / 32) * 2 * B * 32;
}

extern "C" int cvae_tc_pack_weight(const float* src, float* dst, int A, int A_pad, int B, int taps, int src_bat,
                                   int src_ld, cvae_stream_t s) {
  if (!src || !dst || A < 1 || A_pad < A || B < 1 || (src_bat && src_ld < A)) return CVAE_ERR_BAD_ARG;
  const size_t total = (size_t)cvae_tc_pack_floats(A_pad, B, taps);
  const int blocks = (int)min((total + 255) / 256, (size_t)kNumSMs * 16);
  tc_pack_weight_kernel<<<blocks, 256, 0, as_stream(s)>>>(src, dst, A, A_pad, B, taps, src_bat, src_ld);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int64_t cvae_tc_pack_rows_floats(int64_t M, int K) {
  return ((M + 127) / 128) * ((K + 31) / 32) * 8192;
}

extern "C" int cvae_tc_pack_rows(const float* src, float* dst, int64_t M, int K, cvae_stream_t s) {
  if (!src || !dst || M < 1 || K < 4 || (K & 3)) return CVAE_ERR_BAD_ARG;
  const int KB = (K + 31) / 32;
  const long long blocks = ((M + 127) / 128) * KB;
  if (blocks >= (1ll << 31)) return CVAE_ERR_UNSUPPORTED_SHAPE;
  tc_pack_rows_kernel<<<(int)blocks, 256, 0, as_stream(s)>>>(src, dst, M, K, KB);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

static int conv_gather_tc_impl(const cvae_conv_params_t* p, const float* a_image, cvae_stream_t s) {
  if (!p || !p->src || !p->wt || !p->dst || p->N <= 0) return CVAE_ERR_BAD_ARG;
  if (p->epi == CVAE_EPI_DACT && !p->epi_ref) return CVAE_ERR_BAD_ARG;
  if (p->Cs % 16 != 0 || p->Cd % 16 != 0) return CVAE_ERR_UNSUPPORTED_SHAPE;
  GatherArgs g;
  g.a_image = a_image;
  g.src = p->src; g.wt = p->wt; g.bias = p->bias; g.dst = p->dst;
  g.in_scale = p->in.scale; g.in_shift = p->in.shift; g.in_center = p->in.center; g.in_slope = p->in.slope;
  g.in_affine = p->in.scale != nullptr; g.in_act = p->in.slope != 1.0f;
  g.epi = p->epi; g.epi_ref = p->epi_ref; g.epi_add = p->epi_add;
  g.e_scale = p->epi_x.scale; g.e_shift = p->epi_x.shift; g.e_center = p->epi_x.center; g.e_slope = p->epi_x.slope;
  g.e_affine = p->epi_x.scale != nullptr;
  g.stats = p->stats;
  g.N = p->N; g.Hs = p->Hs; g.Ws = p->Ws; g.Cs = p->Cs; g.Hd = p->Hd; g.Wd = p->Wd; g.Cd = p->Cd;
  g.ksplit = 1;
  const int rc = build_geom(p, g);
  if (rc != CVAE_OK) return rc;
  if (p->Hs >= 32768 || p->Ws >= 32768) return CVAE_ERR_UNSUPPORTED_SHAPE;
  if ((int64_t)p->N * p->Hs * p->Ws >= (1ll << 31) || (int64_t)p->N * p->Hd * p->Wd >= (1ll << 31))
    return CVAE_ERR_UNSUPPORTED_SHAPE;
  int maxM = 0;
  for (int i = 0; i < g.nphase; ++i) {
    if (g.phase[i].ntaps < 1) return CVAE_ERR_UNSUPPORTED_SHAPE;
    maxM = max(maxM, p->N * g.phase[i].Hq * g.phase[i].Wq);
  }
  {
    // conv-shaped layers: the halo-tile kernel stages the input once for all taps (CVAE_HALO=0 forces
    // the per-tap gather kernel below, kept for 1x1 / Linear layers and shapes the halo plan rejects)
    static const bool halo_on = [] { const char* e = getenv("CVAE_HALO"); return !(e && e[0] == '0'); }();
    if (halo_on || a_image != nullptr) {
      const int hr = launch_conv_halo_tc(g, as_stream(s));
      if (hr <= 0) return hr;
    }
    if (a_image != nullptr) return CVAE_ERR_UNSUPPORTED_SHAPE;     // the packed operand exists for the halo kernel only
  }
  const int BN = tc_pick_bn(p->Cd);
  if (BN == 0) return CVAE_ERR_UNSUPPORTED_SHAPE;
  const int stage_bytes = kTcAStage + 256 * BN;
  int NS = (198 * 1024) / stage_bytes;
  NS = max(2, min(NS, 6));
  const size_t smem = (size_t)NS * stage_bytes + kEpiBytes + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(igemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess)
      return CVAE_ERR_LAUNCH;
    attr_set = true;
  }
  const int tiles_m = (maxM + 127) / 128, tiles_n = p->Cd / BN;
  const int total = tiles_m * tiles_n * g.nphase;
  igemm_tc_kernel<<<min(total, kNumSMs), kTcThreads, smem, as_stream(s)>>>(g, BN, NS, tiles_m, tiles_n);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_conv_gather_tc(const cvae_conv_params_t* p, cvae_stream_t s) { return conv_gather_tc_impl(p, nullptr, s); }

// Linear layer (kh = kw = 1, identity input transform) whose A operand was packed by cvae_tc_pack_rows: p->src is the
// plain matrix the image was made from (kept for the argument checks), a_image the packed copy the kernel reads.
extern "C" int cvae_linear_tc_packed(const cvae_conv_params_t* p, const float* a_image, cvae_stream_t s) {
  if (!p || !a_image) return CVAE_ERR_BAD_ARG;
  if (p->kh != 1 || p->kw != 1 || p->stride != 1 || p->pad != 0 || p->in.scale != nullptr || p->in.slope != 1.0f)
    return CVAE_ERR_UNSUPPORTED_SHAPE;
  return conv_gather_tc_impl(p, a_image, s);
}
