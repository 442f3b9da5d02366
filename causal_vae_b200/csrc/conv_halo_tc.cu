// Halo-tile tcgen05 / TMEM implicit-GEMM convolution for sm_100a (3xTF32) -- second generation of
// conv_tc.cu for conv-shaped layers (k > 1).
//
// conv_tc.cu gathers the A operand once per TAP (9x for a 3x3 conv): ncu showed it bound by the
// producers' instruction stream (~100 warp instructions per 16-byte vector, tensor pipe 15 % busy).
// Here an input HALO TILE is staged once per 32-channel block and every tap reads it through its own
// shared-memory descriptor:
//
//   * an M tile is a 16 x 8 patch of output positions ("q-space") of one image; MMA row r = 8*rh + rw;
//   * the halo tile is stored slot-major (slot = one pixel x 32 channels = 128 bytes, hi and lo tf32
//     planes), row pitch C slots, with the 128B swizzle applied on ADDRESS bits (chunk ^= bits [7,10));
//   * tap (di, dj) is the K-major SWIZZLE_128B operand whose start address is shifted by
//     (di*C + dj) slots and whose 8-row-group pitch (stride byte offset) is C*128 bytes.  The tensor
//     core applies the swizzle on absolute address bits (measured: scripts/umma_probe.cu -- any
//     128-byte-aligned start and any 128-byte-multiple group pitch read back exactly, descriptor
//     base_offset = 0), so all taps share one staged copy: producer work drops 6.4x for a 3x3 conv;
//   * stride-2 gathers (Conv2d s2 forward, ConvTranspose2d input-gradient) stage the four input
//     parity planes one after the other (a stage = one k-block of one plane), each tap reading the
//     plane of its parity; scatter mode (ConvTranspose2d forward, Conv2d s2 input-gradient) keeps one
//     TMEM accumulator per output phase, all fed from the same staged tile;
//   * weights (pre-split, pre-swizzled image from tc_pack_weight_kernel) stream through their own
//     TMA ring, one (tap, k-block) per stage;
//   * epilogue as in conv_tc.cu (TMEM -> registers -> padded smem -> coalesced 128-bit stores with
//     fused bias / BatchNorm statistics / activation-derivative + BN-backward sums).
#include <cstdlib>
#include "common.cuh"
#include "conv_args.cuh"
#include "tc_common.cuh"

namespace cvae {

using namespace tc;

constexpr int kHTH = 16, kHTW = 8;                 // q-space tile: 16 rows x 8 columns = 128 MMA rows
constexpr int kHMaxSlots = 184;                    // (16+2)*(8+2) = 180, rounded so a half is 23 KiB
constexpr int kHHalf = kHMaxSlots * 128;           // bytes of one tf32 plane (hi or lo) of an A stage
constexpr int kHAStage = 2 * kHHalf;
constexpr int kHNA = 3;                            // largest A ring depth in shared memory (plan: HaloPlan::na)
constexpr size_t kHMaxDyn = 226 * 1024;            // dynamic shared memory cap: 227 KB per CTA minus the static barriers (~200 B)
constexpr int kHRawStages = 4;                     // Linear mode: cp.async ring of raw row chunks (k-blocks in flight + 1)
constexpr int kHNAT = 4;                           // A ring depth in tensor-memory mode (64 columns per stage)
// Producer warps are a template parameter of the kernel (PW): 8 (18 warps -> 96 registers per thread) for forward-type
// launches, whose operand transform is the limiter and scales with the number of warps staging it; 6 (16 warps -> 128
// registers, software-pipelined stages) for input-gradient launches, whose activation-derivative epilogue wants the registers.
constexpr int kHEpiWarps = 8;                      // two groups of four (a warp reads only its own TMEM lane quarter)
__host__ __device__ constexpr int h_threads(int pw) { return (pw + 2 + kHEpiWarps) * 32; }   // + MMA warp + weight-loader warp + epilogue warps
__host__ __device__ constexpr int h_items(int pw) { return (kHMaxSlots * 8 + pw * 32 - 1) / (pw * 32); }   // 16-byte vectors per producer thread

// A tap GROUP: the taps (of different output phases) that read the same staged window.  They run as ONE
// MMA whose B operand stacks their weight tiles along N and whose D spans their (adjacent) accumulators:
// a kind::tf32 MMA costs ~96 clk whatever N <= 128 is (scripts/umma_rate.cu), so the 9 taps of a
// stride-2 transposed conv take 4 MMAs per k-step instead of 9.  Gather mode: one tap per group.
struct HaloTap { int off; int nsub; int pos0; int widx[4]; };   // off: slot offset of the window; pos0: first accumulator
struct HaloPlane { int pr, pc, imin, jmin, ntaps; HaloTap taps[16]; };
struct HaloPlan {
  int nplanes, R, C;          // staged rows / columns (uniform over planes)
  int tiles_h, tiles_w, tiles_n, BN, NB;
  int Hq, Wq;                 // q-space extent per image
  int ph[4], pw[4];           // output offset of each phase
  int pos[4];                 // accumulator position (TMEM column block) of each phase
  int bslot_bytes;            // weight ring slot: 256 * BN * (largest group)
  int a_pre;                  // 1: Linear mode with a pre-packed A image (cvae_tc_pack_rows): the loader warp fetches 32 KB stages
  int a_stage, a_half;        // A ring geometry in shared memory: bytes per stage, offset of the lo plane
  int a_tmem;                 // 1: Linear / 1x1 mode with the A operand staged in tensor memory (see the producer)
  int stat_off;               // byte offset (dynamic shared memory) of the epilogue warps' fp64 statistic slices
  int fence_mode;             // 0: producers fence their stores; 1: the MMA issuer fences once per stage (see the producers)
  int prefetch;               // 1: the loader warp requests epilogue reference rows / next halo tiles into L2
  int na;                     // A ring depth (stages of kHAStage bytes in shared memory, or 64-column stages in tensor memory)
  int xacc;                   // 1: cross terms (hi*lo' + lo*hi') in their own accumulator columns [BN, 2BN) -- see the MMA issuer
  int ksplit;                 // Linear mode: CTAs per output tile, each taking KB / ksplit k-blocks and adding its part into the
                              // pre-zeroed output (1: plain stores) -- see launch_conv_halo_tc
  HaloPlane plane[4];
};

#ifdef CVAE_TIMING
// role-level wait accounting (debug builds only): clock64 deltas summed per role, read by cvae_debug_read
__device__ unsigned long long g_halo_dbg[16];
#define T_DECL unsigned long long t_acc[4] = {0, 0, 0, 0}; unsigned long long t_0 = clock64(), t_1;
#define T_WAIT(i, stmt) { t_1 = clock64(); stmt; t_acc[i] += clock64() - t_1; }
#define T_FLUSH(base, n) { for (int i_ = 0; i_ < n; ++i_) atomicAdd(&g_halo_dbg[base + i_], t_acc[i_]); atomicAdd(&g_halo_dbg[base + n], clock64() - t_0); }
#else
#define T_DECL
#define T_WAIT(i, stmt) { stmt; }
#define T_FLUSH(base, n) {}
#endif

__device__ __forceinline__ void hbar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

struct HTile { int n, h0, w0, n0, ks; };
template <bool KS>
__device__ __forceinline__ HTile h_decode(const HaloPlan& p, int t) {
  HTile o;
  o.ks = 0;
  if constexpr (KS) { o.ks = t % p.ksplit; t /= p.ksplit; }       // work item = (tile, K slice)
  o.n0 = (t % p.tiles_n) * p.BN;
  int sp = t / p.tiles_n;
  o.w0 = (sp % p.tiles_w) * kHTW; sp /= p.tiles_w;
  o.h0 = (sp % p.tiles_h) * kHTH;
  o.n = sp / p.tiles_h;
  return o;
}

// KS: split-K instantiation (Linear launches with few tiles, see launch_conv_halo_tc).  A template parameter, not a run-time
// branch: with `if (p.ksplit > 1) atomicAdd ... else store` in the epilogue the compiler predicated the two atomics into
// EVERY launch (ncu on the stem.3 input gradient: 2.2 M issued, predicated-off instructions; 148 -> 158 us).
template <int PW, bool KS = false>
__global__ void __launch_bounds__(h_threads(PW), 1)
conv_halo_tc_kernel(const __grid_constant__ GatherArgs a, const __grid_constant__ HaloPlan p, const int total) {
  extern __shared__ uint8_t dsm_raw[];
  __shared__ __align__(8) uint64_t s_afull[kHNAT], s_aempty[kHNAT];
  __shared__ __align__(8) uint64_t s_bfull[4], s_bempty[4];
  __shared__ __align__(8) uint64_t s_tfull[2], s_tempty[2];
  __shared__ uint32_t s_tmem;

  constexpr int kHProdWarps = PW, kHItems = h_items(PW);
  constexpr bool kPipe = PW <= 6;                  // register room for a second set of stage loads
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int BN = p.BN, NB = p.NB;
  const int KB = (a.Cs + 31) >> 5;
  const int ksplit = KS ? p.ksplit : 1;             // compile-time 1 in the ordinary instantiations
  const uint32_t bstage = (uint32_t)p.bslot_bytes;
  const uint32_t acc_cols = (uint32_t)(a.nphase * BN * (p.xacc ? 2 : 1));
  const uint32_t a_cols = p.a_tmem ? (uint32_t)(kHNAT * 64) : 0u;     // A ring in tensor memory: hi | lo, 32 columns each
  const int na = p.na;
  uint32_t tmem_cols = 32;
  while (tmem_cols < 2u * acc_cols + a_cols) tmem_cols <<= 1;
  uint8_t* dsm_gen = dsm_raw + ((1024u - (smem_u32(dsm_raw) & 1023u)) & 1023u);
  const uint32_t dsm = smem_u32(dsm_gen);
  const uint32_t b_base = dsm + (p.a_tmem ? 0u : (uint32_t)(na * p.a_stage));

  if (warp == kHProdWarps) {
    if (lane == 0) {
      for (int i = 0; i < kHNAT; ++i) { mbar_init(smem_u32(&s_afull[i]), p.a_pre ? 1 : (p.a_tmem ? (PW >= 8 ? 8 : 4) : kHProdWarps)); mbar_init(smem_u32(&s_aempty[i]), 1); }
      for (int i = 0; i < NB; ++i) { mbar_init(smem_u32(&s_bfull[i]), 1); mbar_init(smem_u32(&s_bempty[i]), 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&s_tfull[i]), 1); mbar_init(smem_u32(&s_tempty[i]), kHEpiWarps); }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(&s_tmem), tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  if (warp < kHProdWarps && p.a_pre) {
    // A operand pre-packed (cvae_tc_pack_rows): nothing to transform or split -- the loader warp streams it
  } else if (warp < kHProdWarps && p.a_tmem) {
    // ============================== A producers, Linear / 1x1 mode: rows -> tensor memory ==============================
    // With both operands in shared memory a kind::tf32 MMA costs ~100-120 clk whatever N <= 256 is, with A in tensor
    // memory 51-64 clk (scripts/umma_rate*.cu) - and a Linear layer has no halo to share between taps, so nothing is
    // lost by leaving shared memory out: thread = one row of the 128-row tile (TMEM lane; warps 0-3, the other producer
    // warps idle), hi and lo planes of a 32-channel k-block go to columns [0,32) / [32,64) of the stage with tcgen05.st.
    constexpr int kLinWarps = PW >= 8 ? 8 : 4;       // 8: warps w and w + 4 take channels 0-15 / 16-31 of a k-block
    constexpr int kHpt = kLinWarps == 8 ? 1 : 2;      // 16-channel halves per thread
    constexpr int kPitch = kHpt == 1 ? 80 : 144;      // bytes per thread and stage (+16: conflict-free 128-bit reads)
    if (warp < kLinWarps) {
      const int q = warp & 3, h0 = kLinWarps == 8 ? (warp >> 2) : 0;
      const int row = q * 32 + lane;
      const uint32_t t_a = tmem + 2u * acc_cols + ((uint32_t)(q * 32) << 16);
      // A thread's channels of a k-block (64 or 128 bytes of its own row) travel global -> shared with cp.async,
      // kHRawStages - 1 k-blocks ahead and across tile boundaries, into a slot only this thread reads: no block barrier,
      // no registers held while the loads fly.  The copies of a warp's 32 rows are issued COOPERATIVELY: kCh consecutive
      // lanes fetch the kCh 16-byte chunks of one row, so one instruction touches 32 / kCh rows with full sectors.  (In the
      // thread-per-row form every instruction touched 32 different 128-byte lines, 16 of 32 sector bytes used: 1024 tag
      // lookups per k-block and SM - scripts/probe_linear_fixed.py: ~0.9 us per k-block whatever the MMA work, 17.6 us
      // for ONE 128 x 256 x 256 tile.  First form of all: load -> transform -> tcgen05.st -> next load, one dependent
      // round trip per k-block.)
      constexpr int kCh = 4 * kHpt, kRpi = 32 / kCh;   // chunks per row part, rows per instruction
      uint8_t* raw = dsm_gen + (size_t)NB * bstage + (size_t)tid * kPitch;
      const uint32_t raw_warp = smem_u32(dsm_gen + (size_t)NB * bstage + (size_t)(tid - lane) * kPitch);
      const int crow = lane / kCh, cch = lane % kCh;   // this lane's row (within an instruction) and chunk
      const long long Mrows = (long long)a.Hs * 8;     // Linear view: ONE image of M / 8 x 8 positions, a tile = 128 consecutive rows
      const int kbper = KB / ksplit;                   // k-blocks of one work item (host: KB % ksplit == 0)
      int ti = blockIdx.x, kbi = 0, kbend = 0;         // next (tile, k-block) to request; end of the item's K slice
      const float* wbase = nullptr;                    // first row of this warp quarter in tile ti
      int nvalid = 0;                                  // rows of the quarter inside the matrix
      auto set_tile = [&](int t) {
        wbase = nullptr; nvalid = 0;
        if (t >= total) return;
        const HTile tl = h_decode<KS>(p, t);
        kbi = tl.ks * kbper; kbend = kbi + kbper;
        const long long r0 = (long long)tl.h0 * 8 + q * 32;
        wbase = a.src + (size_t)r0 * a.Cs;
        const long long left = Mrows - r0;
        nvalid = left >= 32 ? 32 : (left > 0 ? (int)left : 0);
      };
      set_tile(ti);
      auto issue = [&](int slot) {
        if (ti < total) {
          const int c = kbi * 32 + h0 * 16 + 4 * cch;
          const uint32_t dstw = raw_warp + (uint32_t)slot * (uint32_t)(kLinWarps * 32 * kPitch) + 16u * cch;
#pragma unroll
          for (int j = 0; j < kCh; ++j) {
            const int rl = j * kRpi + crow;
            const uint32_t dst = dstw + (uint32_t)(rl * kPitch);
            if (rl < nvalid && c < a.Cs)
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(wbase + (size_t)rl * a.Cs + c) : "memory");
            else
              asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "f"(0.f) : "memory");
          }
          if (++kbi == kbend) { ti += gridDim.x; set_tile(ti); }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");     // always: one group per iteration keeps the count uniform
      };
#pragma unroll
      for (int sidx = 0; sidx < kHRawStages - 1; ++sidx) issue(sidx);
      uint32_t it = 0;
      T_DECL
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const HTile tl = h_decode<KS>(p, t);
        const int qh = tl.h0 + (row >> 3), qw = tl.w0 + (row & 7);
        const bool rok = qh < a.Hs && qw < a.Ws;
        for (int kb = tl.ks * kbper; kb < (tl.ks + 1) * kbper; ++kb, ++it) {
          issue((int)((it + kHRawStages - 1) % kHRawStages));
          T_WAIT(1, asm volatile("cp.async.wait_group %0;" ::"n"(kHRawStages - 1) : "memory"))
          __syncwarp();                                  // the row was copied by other lanes of this warp
          const uint8_t* rs = raw + (size_t)(it % kHRawStages) * (kLinWarps * 32 * kPitch);
          const int slot = it % kHNAT;
#pragma unroll
          for (int hh = 0; hh < kHpt; ++hh) {
            const int half = h0 + hh;
            const int c0 = kb * 32 + half * 16;
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float4 x = *reinterpret_cast<const float4*>(rs + 64 * hh + 16 * j);
              const int c = c0 + 4 * j;
              if (rok && c < a.Cs) {   // padding stays exactly 0
                if (a.in_affine) {
                  const float4 sc = __ldg(reinterpret_cast<const float4*>(a.in_scale + c));
                  const float4 sh = __ldg(reinterpret_cast<const float4*>(a.in_shift + c));
                  float4 ce = make_float4(0.f, 0.f, 0.f, 0.f);
                  if (a.in_center != nullptr) ce = __ldg(reinterpret_cast<const float4*>(a.in_center + c));
                  x.x = fmaf(x.x - ce.x, sc.x, sh.x); x.y = fmaf(x.y - ce.y, sc.y, sh.y);
                  x.z = fmaf(x.z - ce.z, sc.z, sh.z); x.w = fmaf(x.w - ce.w, sc.w, sh.w);
                }
                if (a.in_act) {
                  x.x = fmaxf(x.x, x.x * a.in_slope); x.y = fmaxf(x.y, x.y * a.in_slope);
                  x.z = fmaxf(x.z, x.z * a.in_slope); x.w = fmaxf(x.w, x.w * a.in_slope);
                }
              }
              float4 h4, l4;
              split4(x, h4, l4);
              hi[4 * j] = __float_as_uint(h4.x); hi[4 * j + 1] = __float_as_uint(h4.y);
              hi[4 * j + 2] = __float_as_uint(h4.z); hi[4 * j + 3] = __float_as_uint(h4.w);
              lo[4 * j] = __float_as_uint(l4.x); lo[4 * j + 1] = __float_as_uint(l4.y);
              lo[4 * j + 2] = __float_as_uint(l4.z); lo[4 * j + 3] = __float_as_uint(l4.w);
            }
            if (hh == 0) {
              T_WAIT(0, mbar_wait(smem_u32(&s_aempty[slot]), ((it / kHNAT) & 1u) ^ 1u))
              tc_fence_after();
            }
            tmem_st16(t_a + (uint32_t)(slot * 64 + half * 16), hi);
            tmem_st16(t_a + (uint32_t)(slot * 64 + 32 + half * 16), lo);
          }
          T_WAIT(2, tmem_st_wait())
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&s_afull[slot]));
        }
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      if (tid == 0) T_FLUSH(0, 1)
#ifdef CVAE_TIMING
      if (tid == 0) { atomicAdd(&g_halo_dbg[10], t_acc[1]); atomicAdd(&g_halo_dbg[11], t_acc[2]); }   // Linear producer: cp.async wait, tcgen05.wait::st
#endif
    }
  } else if (warp < kHProdWarps) {
    // ============================== A producers: halo tile -> swizzled hi / lo planes ==============================
    const int chunk = tid & 7;
    const int nslots = p.R * p.C;
    // per item, fixed for the whole kernel: where its 16-byte vector lands in the swizzled stage (-1: no such slot) and
    // where it comes from relative to the stage's first pixel -- an interior stage is then one 64-bit add, one load
    // per item
    int soff[kHItems], goff[kHItems];
#pragma unroll
    for (int k = 0; k < kHItems; ++k) {
      const int s = (tid >> 3) + k * (kHProdWarps * 4);
      const int i = s / p.C, j = s - i * p.C;
      soff[k] = s < nslots ? s * 128 + (((chunk ^ s) & 7) << 4) : -1;
      goff[k] = (i * a.is * a.Ws + j * a.is) * a.Cs;
    }
    const int span_h = (p.R - 1) * a.is, span_w = (p.C - 1) * a.is;
    // Software pipeline over stages: the loads of stage i + 1 are issued BEFORE stage i is transformed and stored, so a
    // stage costs max(memory latency, transform) instead of their sum (role timers: the producers were busy 90-95 %
    // of the forward stride-2 launches at 0.12 instructions per cycle and warp -- one dependent round trip per stage).
    struct Stage { int t, kb, pl; };
    auto advance = [&](Stage st) -> Stage {
      if (++st.pl == p.nplanes) { st.pl = 0; if (++st.kb == KB) { st.kb = 0; st.t += gridDim.x; } }
      return st;
    };
    int ld_t = -1;
    HTile ld_tl{0, 0, 0, 0, 0};
    auto load_stage = [&](const Stage& st, float4 (&v)[kHItems]) -> uint32_t {
      uint32_t okm = 0;
      if (st.t != ld_t) { ld_tl = h_decode<KS>(p, st.t); ld_t = st.t; }
      const int c = st.kb * 32 + chunk * 4;
      const bool cok = c < a.Cs;
      const HaloPlane& P = p.plane[st.pl];
      const int bh = (ld_tl.h0 + P.imin) * a.is + P.pr, bw = (ld_tl.w0 + P.jmin) * a.is + P.pc;
      const long long base = (((long long)ld_tl.n * a.Hs + bh) * a.Ws + bw) * a.Cs + c;     // float index of the stage's first pixel
      if (cok && bh >= 0 && bw >= 0 && bh + span_h < a.Hs && bw + span_w < a.Ws) {   // interior stage (block-uniform)
#pragma unroll
        for (int k = 0; k < kHItems; ++k) {
          v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (soff[k] >= 0) { v[k] = __ldg(reinterpret_cast<const float4*>(a.src + (base + goff[k]))); okm |= 1u << k; }
        }
      } else {
#pragma unroll
        for (int k = 0; k < kHItems; ++k) {
          const int sl = (tid >> 3) + k * (kHProdWarps * 4);
          const int i = sl / p.C, j = sl - i * p.C;
          const int ih = bh + i * a.is, iw = bw + j * a.is;
          const bool ok = cok && soff[k] >= 0 && (unsigned)ih < (unsigned)a.Hs && (unsigned)iw < (unsigned)a.Ws;
          v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok) { v[k] = __ldg(reinterpret_cast<const float4*>(a.src + (base + goff[k]))); okm |= 1u << k; }
        }
      }
      return okm;
    };
    uint32_t it = 0;
    T_DECL
    Stage cur{(int)blockIdx.x, 0, 0};
    float4 v[kHItems], vn[kHItems];
    uint32_t okm = 0, okn = 0;
    if (cur.t < total) okm = load_stage(cur, v);
    int tr_kb = -1;
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f), ce = sh;
    while (cur.t < total) {
      const Stage nxt = advance(cur);
      if constexpr (kPipe) { if (nxt.t < total) okn = load_stage(nxt, vn); }
      if (cur.kb != tr_kb) {                         // BatchNorm coefficients of this k-block's channels
        tr_kb = cur.kb;
        const int c = cur.kb * 32 + chunk * 4;
        sc = make_float4(1.f, 1.f, 1.f, 1.f); sh = make_float4(0.f, 0.f, 0.f, 0.f); ce = sh;
        if (a.in_affine && c < a.Cs) {
          sc = __ldg(reinterpret_cast<const float4*>(a.in_scale + c));
          sh = __ldg(reinterpret_cast<const float4*>(a.in_shift + c));
          if (a.in_center != nullptr) ce = __ldg(reinterpret_cast<const float4*>(a.in_center + c));
        }
      }
      const int slot = it % na;
      T_WAIT(0, mbar_wait(smem_u32(&s_aempty[slot]), ((it / na) & 1u) ^ 1u))
      uint8_t* sA = dsm_gen + (size_t)slot * p.a_stage;
      // Straight-line transform: no per-item branches, so the compiler interleaves the 32 independent value chains of the
      // 8 items (the per-item `if`s of the first form left one dependent chain per basic block: ncu showed the producer
      // warps issuing one instruction per ~7.7 clk, stall reason "wait").  Without BatchNorm the coefficients are
      // (1, 0, 0) and without activation the slope is 1: fma(x - 0, 1, 0) and max(x, x) reproduce x exactly; padding is
      // forced back to exactly 0 by the select.
      const float slope = a.in_act ? a.in_slope : 1.f;
#pragma unroll
      for (int k = 0; k < kHItems; ++k) {
        float4 x = v[k];
        x.x = fmaf(x.x - ce.x, sc.x, sh.x); x.y = fmaf(x.y - ce.y, sc.y, sh.y);
        x.z = fmaf(x.z - ce.z, sc.z, sh.z); x.w = fmaf(x.w - ce.w, sc.w, sh.w);
        x.x = fmaxf(x.x, x.x * slope); x.y = fmaxf(x.y, x.y * slope);
        x.z = fmaxf(x.z, x.z * slope); x.w = fmaxf(x.w, x.w * slope);
        const bool real = (okm >> k) & 1u;
        x.x = real ? x.x : 0.f; x.y = real ? x.y : 0.f; x.z = real ? x.z : 0.f; x.w = real ? x.w : 0.f;
        float4 hi, lo;
        split4(x, hi, lo);
        if (soff[k] >= 0) {
          *reinterpret_cast<float4*>(sA + soff[k]) = hi;
          *reinterpret_cast<float4*>(sA + kHHalf + soff[k]) = lo;
        }
      }
      // generic-proxy stores -> async-proxy reads (tcgen05.mma operands) need a proxy fence somewhere on the causality
      // path.  On the writers' side (fence_mode 0) every producer thread fences with its 128-bit stores still in flight:
      // ncu attributed 29 % of ALL stall samples of the kernel to that MEMBAR (~45 % of the producers' time).
      // fence_mode 1: the writers only release through the mbarrier and the MMA issuer fences once per stage after
      // its acquire (PTX memory model: the proxy fence may sit anywhere along the base causality order).
      if (p.fence_mode == 0) fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_afull[slot]));
      if constexpr (kPipe) {
#pragma unroll
        for (int k = 0; k < kHItems; ++k) v[k] = vn[k];
        okm = okn;
      } else {
        if (nxt.t < total) okm = load_stage(nxt, v);
      }
      cur = nxt;
      ++it;
    }
    if (tid == 0) T_FLUSH(0, 1)
  } else if (warp == kHProdWarps) {
    // ============================== MMA issuer (one elected lane of a CONVERGED warp) ==============================
    // The whole warp runs the loop (barrier waits, descriptor arithmetic: all warp-uniform values) and only the issue
    // itself is predicated on the elected lane.  Inside `if (lane == 0)` the compiler must treat every operand as
    // per-thread data and wraps each tcgen05.mma in an ELECT / BRA.U.ANY loop with predicate chains (6 dependent
    // instructions per MMA, SASS); that loop, not the tensor core, was the "96.6 clk per kind::tf32 MMA whatever N"
    // of scripts/umma_rate.cu -- converged, the descriptors live in uniform registers and the MMAs issue back to back.
    {
      const bool leader = elect_one();
      const uint32_t sbo = (uint32_t)p.C * 128u;
      const uint64_t a_desc_hi = make_smem_desc(0, 16, sbo, kLayoutSw128);     // everything but the address
      const uint64_t b_desc_hi = make_smem_desc(0, 16, 1024, kLayoutSw128);
      uint32_t ita = 0, itb = 0, tcount = 0;
      T_DECL
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++tcount) {
        const uint32_t acc = tcount & 1u;
        T_WAIT(0, mbar_wait(smem_u32(&s_tempty[acc]), ((tcount >> 1) & 1u) ^ 1u))
        tc_fence_after();
        const uint32_t d_base = tmem + acc * acc_cols;
        uint32_t started = 0, startedx = 0;
        const int kbper = KB / ksplit, kb0 = KS ? (t % ksplit) * kbper : 0;
        for (int kb = kb0; kb < kb0 + kbper; ++kb) {
          const int ksteps = min(32, a.Cs - kb * 32) >> 3;
          for (int pl = 0; pl < p.nplanes; ++pl, ++ita) {
            const HaloPlane& P = p.plane[pl];
            const int aslot = ita % na;
            T_WAIT(1, mbar_wait(smem_u32(&s_afull[aslot]), (ita / na) & 1u))
            if (p.fence_mode == 1 && !p.a_tmem) fence_async_smem();
            tc_fence_after();
            const uint32_t a_hi0 = dsm + (uint32_t)(aslot * p.a_stage);
            if (p.a_tmem) {          // one tap, A operand in tensor memory (columns: hi [0,32), lo [32,64) of the stage)
              const int bslot = itb % NB;
              T_WAIT(2, mbar_wait(smem_u32(&s_bfull[bslot]), (itb / NB) & 1u))
              tc_fence_after();
              const uint32_t idesc = make_idesc_tf32(128, BN, 0, 0);
              const uint32_t b_hi = b_base + (uint32_t)bslot * bstage;
              uint64_t dbh = b_desc_hi | (uint64_t)((b_hi & 0x3FFFFu) >> 4);
              uint64_t dbl = b_desc_hi | (uint64_t)(((b_hi + 128u * (uint32_t)BN) & 0x3FFFFu) >> 4);
              uint32_t ta_hi = tmem + 2u * acc_cols + (uint32_t)(aslot * 64), ta_lo = ta_hi + 32u;
              uint32_t accum = started ? 1u : 0u;
              if (p.xacc) {
                // [W_hi | W_lo] are adjacent in the slot: ONE MMA of N = 2*BN gives A_hi*W_hi in columns [0, BN)
                // and A_hi*W_lo in [BN, 2BN); A_lo*W_hi joins the cross columns.  Two MMAs per product instead of
                // three, and the main chain sees one truncating accumulate per k-step instead of three.
                const uint32_t idesc2 = make_idesc_tf32(128, 2 * BN, 0, 0);
#pragma unroll 4
                for (int k = 0; k < ksteps; ++k) {
                  if (leader) mma_tf32_ts(d_base, ta_hi, dbh, idesc2, accum);
                  if (leader) mma_tf32_ts(d_base + (uint32_t)BN, ta_lo, dbh, idesc, 1u);
                  accum = 1u;
                  ta_hi += 8; ta_lo += 8; dbh += 2;
                }
              } else {
#pragma unroll 4
                for (int k = 0; k < ksteps; ++k) {
                  if (leader) mma_tf32_ts(d_base, ta_lo, dbh, idesc, accum);
                  if (leader) mma_tf32_ts(d_base, ta_hi, dbl, idesc, 1u);
                  if (leader) mma_tf32_ts(d_base, ta_hi, dbh, idesc, 1u);
                  accum = 1u;
                  ta_hi += 8; ta_lo += 8; dbh += 2; dbl += 2;
                }
              }
              started = 1u;
              if (leader) mma_commit(smem_u32(&s_bempty[bslot]));
              ++itb;
              if (leader) mma_commit(smem_u32(&s_aempty[aslot]));
              continue;
            }
            for (int tp = 0; tp < P.ntaps; ++tp, ++itb) {
              const HaloTap tap = P.taps[tp];
              const int bslot = itb % NB;
              T_WAIT(2, mbar_wait(smem_u32(&s_bfull[bslot]), (itb / NB) & 1u))
              tc_fence_after();
              const uint32_t ng = (uint32_t)(tap.nsub * BN);                 // N of this group's MMA
              const uint32_t idesc = make_idesc_tf32(128, (int)ng, 0, 0);
              const uint32_t a_hi = a_hi0 + (uint32_t)tap.off * 128u;
              const uint32_t b_hi = b_base + (uint32_t)bslot * bstage;
              const uint32_t d_tmem = d_base + (uint32_t)(tap.pos0 * BN);
              const uint32_t gmask = ((1u << tap.nsub) - 1u) << tap.pos0;
              uint32_t accum = (started & gmask) ? 1u : 0u;                  // host plan: never mixed within a group
              // descriptors differ only in their 14-bit address field: one 64-bit add per k-step
              // (the issuing thread is latency-bound: rebuilding four descriptors per k-step capped
              // the issue rate at ~1 MMA / 100 clk -- measured with the CVAE_TIMING build)
              uint64_t dah = a_desc_hi | (uint64_t)((a_hi & 0x3FFFFu) >> 4);
              uint64_t dal = a_desc_hi | (uint64_t)(((a_hi + (uint32_t)p.a_half) & 0x3FFFFu) >> 4);
              uint64_t dbh = b_desc_hi | (uint64_t)((b_hi & 0x3FFFFu) >> 4);
              uint64_t dbl = b_desc_hi | (uint64_t)(((b_hi + 128u * ng) & 0x3FFFFu) >> 4);
              if (p.xacc && a.nphase > 1) {
                // scatter plans: the stacked groups of different phases cannot all be followed by their own cross
                // columns, so the cross accumulators live nphase*BN columns further on and take two MMAs of their own
                const uint32_t dx_tmem = d_tmem + (uint32_t)(a.nphase * BN);
                uint32_t accx = (startedx & gmask) ? 1u : 0u;
#pragma unroll 4
                for (int k = 0; k < ksteps; ++k) {
                  if (leader) mma_tf32(dx_tmem, dal, dbh, idesc, accx);
                  if (leader) mma_tf32(dx_tmem, dah, dbl, idesc, 1u);
                  if (leader) mma_tf32(d_tmem, dah, dbh, idesc, accum);
                  accum = 1u; accx = 1u;
                  dah += 2; dal += 2; dbh += 2; dbl += 2;
                }
                startedx |= gmask;
              } else if (p.xacc) {   // one tap per group, one phase: [main | cross] accumulators (see the a_tmem branch)
                const uint32_t idesc2 = make_idesc_tf32(128, 2 * BN, 0, 0);
#pragma unroll 4
                for (int k = 0; k < ksteps; ++k) {
                  if (leader) mma_tf32(d_tmem, dah, dbh, idesc2, accum);
                  if (leader) mma_tf32(d_tmem + (uint32_t)BN, dal, dbh, idesc, 1u);
                  accum = 1u;
                  dah += 2; dal += 2; dbh += 2;
                }
              } else {
#pragma unroll 4
                for (int k = 0; k < ksteps; ++k) {
                  if (leader) mma_tf32(d_tmem, dal, dbh, idesc, accum);
                  if (leader) mma_tf32(d_tmem, dah, dbl, idesc, 1u);
                  if (leader) mma_tf32(d_tmem, dah, dbh, idesc, 1u);
                  accum = 1u;
                  dah += 2; dal += 2; dbh += 2; dbl += 2;      // + 32 bytes (8 tf32) along K
                }
              }
              started |= gmask;
              if (leader) mma_commit(smem_u32(&s_bempty[bslot]));
            }
            if (leader) mma_commit(smem_u32(&s_aempty[aslot]));
          }
        }
        if (leader) mma_commit(smem_u32(&s_tfull[acc]));
      }
      if (leader) T_FLUSH(2, 3)
    }
  } else if (warp == kHProdWarps + 1) {
    // ============================== weight loader (TMA, one thread) + L2 prefetcher (the warp) ==============================
    // The warp first requests into L2 what the OTHER roles will read from global memory about one tile later: the
    // producer's raw output rows that the activation-derivative epilogue re-reads for THIS tile (the epilogue keeps only
    // one work unit of loads in flight -- 16 KB per SM against the ~90 KB that 44 B/ns x 2 us of DRAM latency needs;
    // role timers before: epilogue busy 275 of 294 kclk on the stem.3 input gradient) and the input halo rows of the
    // NEXT tile.  One lane per image row segment; the n-tile-0 CTA of a spatial tile does it.  Off by default
    // (CVAE_HALO_PREFETCH=1): measured neutral to harmful, see DESIGN.md.
    const bool pf_on = p.prefetch != 0;
    const bool ld_leader = elect_one();
    uint32_t itb = 0, ita = 0;
    T_DECL
    const uint32_t bbytes = 128u * BN;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      if (pf_on) {
        __syncwarp();
        if (a.epi == CVAE_EPI_DACT && t % p.tiles_n == 0) {
          const HTile tl = h_decode<KS>(p, t);
          const int oh0 = tl.h0 * a.os, ow0 = tl.w0 * a.os;
          const int npx = min(kHTW * a.os, a.Wd - ow0);
          const uint32_t bytes = (uint32_t)(npx * a.Cd) * 4u;
          for (int r = lane; r < kHTH * a.os; r += 32) {
            if (oh0 + r >= a.Hd || npx <= 0) break;
            const size_t o = ((size_t)(tl.n * a.Hd + oh0 + r) * a.Wd + ow0) * a.Cd;
            prefetch_l2_lines(a.epi_ref + o, bytes);
            if (a.epi_add != nullptr) prefetch_l2_lines(a.epi_add + o, bytes);
          }
        }
        const int tn = t + gridDim.x;
        if (tn < total && !p.a_tmem && tn % p.tiles_n == 0) {
          const HTile tl = h_decode<KS>(p, tn);
          // union of the planes' staged windows (plane rows / columns are `is` apart)
          int r_lo = 1 << 30, r_hi = -(1 << 30), c_lo = 1 << 30, c_hi = -(1 << 30);
          for (int pl = 0; pl < p.nplanes; ++pl) {
            const HaloPlane& P = p.plane[pl];
            const int bh = (tl.h0 + P.imin) * a.is + P.pr, bw = (tl.w0 + P.jmin) * a.is + P.pc;
            r_lo = min(r_lo, bh); r_hi = max(r_hi, bh + (p.R - 1) * a.is);
            c_lo = min(c_lo, bw); c_hi = max(c_hi, bw + (p.C - 1) * a.is);
          }
          r_lo = max(r_lo, 0); r_hi = min(r_hi, a.Hs - 1); c_lo = max(c_lo, 0); c_hi = min(c_hi, a.Ws - 1);
          if (c_hi >= c_lo) {
            const uint32_t bytes = (uint32_t)((c_hi - c_lo + 1) * a.Cs) * 4u;
            for (int r = r_lo + lane; r <= r_hi; r += 32)
              prefetch_l2_lines(a.src + ((size_t)(tl.n * a.Hs + r) * a.Ws + c_lo) * a.Cs, bytes);
          }
        }
        __syncwarp();
      }
      if (ld_leader) {
        const int tt = t / ksplit, kbper = KB / ksplit, kb0 = (t % ksplit) * kbper;
        const int n0 = (tt % p.tiles_n) * BN;
        const int mt = tt / p.tiles_n;                 // a_pre: 128-row tile of the packed A image
        for (int kb = kb0; kb < kb0 + kbper; ++kb)
          for (int pl = 0; pl < p.nplanes; ++pl) {
            if (p.a_pre) {                             // this k-block's A stage: hi and lo planes, 16 KB each
              const int aslot = ita % na;
              T_WAIT(0, mbar_wait(smem_u32(&s_aempty[aslot]), ((ita / na) & 1u) ^ 1u))
              const uint32_t afull = smem_u32(&s_afull[aslot]);
              const uint32_t sA = dsm + (uint32_t)(aslot * p.a_stage);
              const float* asrc = a.a_image + ((size_t)mt * KB + kb) * 8192;
              mbar_arrive_expect_tx(afull, 32768u);
              bulk_g2s(sA, asrc, 16384u, afull);
              bulk_g2s(sA + (uint32_t)p.a_half, asrc + 4096, 16384u, afull);
              ++ita;
            }
            const HaloPlane& P = p.plane[pl];
            for (int tp = 0; tp < P.ntaps; ++tp, ++itb) {
              const int bslot = itb % NB;
              T_WAIT(0, mbar_wait(smem_u32(&s_bempty[bslot]), ((itb / NB) & 1u) ^ 1u))
              const uint32_t full = smem_u32(&s_bfull[bslot]);
              const uint32_t sB = b_base + (uint32_t)bslot * bstage;
              const HaloTap tap = P.taps[tp];
              const uint32_t ng = (uint32_t)(tap.nsub * BN);
              mbar_arrive_expect_tx(full, 2u * 128u * ng);
              for (int sb = 0; sb < tap.nsub; ++sb) {        // stack the sub-taps' weight tiles along N
                const float* wsrc = a.wt + (((size_t)tap.widx[sb] * KB + kb) * 2) * (size_t)a.Cd * 32 + (size_t)n0 * 32;
                bulk_g2s(sB + (uint32_t)sb * bbytes, wsrc, bbytes, full);
                bulk_g2s(sB + 128u * ng + (uint32_t)sb * bbytes, wsrc + (size_t)a.Cd * 32, bbytes, full);
              }
            }
          }
      }
    }
    if (ld_leader) T_FLUSH(6, 1)
  } else {
    // ============================== epilogue warps (two groups of four) ==============================
    // The accumulators leave tensor memory in the 16-lane x 256-bit shape: thread t of a warp receives rows t/4 and
    // t/4 + 8 of the 16 lanes, columns 2*(t%4), +1 of each 8-column group -- i.e. four consecutive lanes hold 32
    // contiguous bytes of one output pixel, so every global access below is a full 32-byte sector per lane quad and
    // NO shared-memory transposition and NO barrier is needed (the first two generations staged a 128 x 32 tile
    // through padded shared memory behind a named barrier; with four warps that made the epilogue the limiter of
    // the HBM-shaped launches: role timers 327 of 345 kclk busy, profiles/r1_ncu_halo_stem3_raw.txt).
    // Work unit = 16 accumulator columns of one phase; the two groups take alternate units of the SAME tile, so a
    // tile's drain is shared by eight warps (short tail for the layers with one or two tiles per CTA).
    // MMA row r = 8*rh + rw is TMEM lane r: warp quarter q holds rh = 4q .. 4q+3, the two 16-lane loads give
    // (rh, rh+1) and (rh+2, rh+3) for the fixed rw = t/4 of the thread.
    const int ew = warp - (kHProdWarps + 2);
    const int grp = ew >> 2, q = warp & 3;
    const int rw = lane >> 2, cpair = (lane & 3) << 1;
    const int upp = BN >> 4;                          // units per phase (1, 2, 4 or 8)
    const int upp_sh = 31 - __clz(upp);
    const int nunits = a.nphase * upp;
    const bool want_stats = a.epi != CVAE_EPI_PLAIN && a.stats != nullptr;
    // BatchNorm sums of the CTA: one private [2][256] fp64 slice per epilogue warp (its lanes 0-3 own distinct channels
    // after the shuffle reduction), summed over the eight warps at the end.  The shared-memory fp64 atomics this replaces
    // (a compare-and-swap loop per add) were 13 % of the kernel's stall samples on the stem.3 forward.
    double* wstat = reinterpret_cast<double*>(dsm_gen + p.stat_off);
    double* wst = wstat + ew * 512;
    if (want_stats) {
      for (int i = lane; i < 512; i += 32) wst[i] = 0.0;
      __syncwarp();
    }
    // a thread sees the same four channels in every unit when the tile spans all channels and a group always gets
    // the same column half: the sums then stay in registers for the whole kernel
    const bool reg_stats = want_stats && BN <= 32 && p.tiles_n == 1;
    const bool dact = a.epi == CVAE_EPI_DACT;
    const bool has_add = dact && a.epi_add != nullptr;
    const int ostep = a.os * a.Wd;                    // pixel distance between a thread's consecutive rows
    double s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
    struct Unit { int obase, vmask, col, t, ui, ks; };
    HTile ctl = h_decode<KS>(p, blockIdx.x);
    int ct = blockIdx.x;
    auto unit_of = [&](int t, int ui) -> Unit {
      Unit u{0, 0, 0, t, ui, 0};
      if (t >= total || ui >= nunits) return u;
      if (t != ct) { ctl = h_decode<KS>(p, t); ct = t; }
      u.ks = ctl.ks;
      const int phs = ui >> upp_sh, hf = ui - (phs << upp_sh);
      u.col = ctl.n0 + hf * 16 + cpair;
      const int qh = ctl.h0 + 4 * q, qw = ctl.w0 + rw;
      const int oh = qh * a.os + p.ph[phs], ow = qw * a.os + p.pw[phs];
      if (qw >= p.Wq || ow >= a.Wd) return u;
      u.obase = (ctl.n * a.Hd + oh) * a.Wd + ow;
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (qh + r < p.Hq && oh + r * a.os < a.Hd) u.vmask |= 1 << r;
      return u;
    };
    float2 pre[4][2];                                 // reference rows of the unit about to be processed (DACT)
    auto prefetch = [&](const Unit& u) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (dact && ((u.vmask >> r) & 1)) {
          const float* rp = a.epi_ref + (size_t)(u.obase + r * ostep) * a.Cd + u.col;
          pre[r][0] = __ldg(reinterpret_cast<const float2*>(rp));
          pre[r][1] = __ldg(reinterpret_cast<const float2*>(rp + 8));
        }
    };
#pragma unroll
    for (int r = 0; r < 4; ++r) pre[r][0] = pre[r][1] = make_float2(0.f, 0.f);
    Unit cur = unit_of(blockIdx.x, grp);
    prefetch(cur);
    uint32_t tcount = 0;
    T_DECL
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++tcount) {
      const uint32_t acc = tcount & 1u;
      T_WAIT(0, mbar_wait(smem_u32(&s_tfull[acc]), (tcount >> 1) & 1u))
      tc_fence_after();
      for (int ui = grp; ui < nunits; ui += 2) {
        const int phs = ui >> upp_sh, hf = ui - (phs << upp_sh);
        const uint32_t taddr = tmem + acc * acc_cols + (uint32_t)(p.pos[phs] * BN + hf * 16) + ((uint32_t)(q * 32) << 16);
        uint32_t v[2][8];
        tmem_ld_16x256b_x2(taddr, v[0]);
        tmem_ld_16x256b_x2(taddr + (16u << 16), v[1]);
        if (p.xacc) {                                 // + the cross-term accumulator (rounded fp32 add)
          uint32_t c[2][8];
          const uint32_t xaddr = taddr + (uint32_t)(a.nphase * BN);
          tmem_ld_16x256b_x2(xaddr, c[0]);
          tmem_ld_16x256b_x2(xaddr + (16u << 16), c[1]);
          tmem_ld_wait();
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < 8; ++i) v[h][i] = __float_as_uint(__uint_as_float(v[h][i]) + __uint_as_float(c[h][i]));
        } else {
          tmem_ld_wait();
        }
        const bool last = ui + 2 >= nunits;
        if (last) {                                   // this warp's share of the accumulator is in registers
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&s_tempty[acc]));
        }
        const Unit nxt = last ? unit_of(t + gridDim.x, grp) : unit_of(t, ui + 2);
        const int col = cur.col;
        float f1[4] = {0.f, 0.f, 0.f, 0.f}, f2[4] = {0.f, 0.f, 0.f, 0.f};   // fp32 partial sums over the unit's 4 rows
#pragma unroll
        for (int g = 0; g < 2; ++g) {                 // the thread's two channel pairs: col + 8g, col + 8g + 1
          float2 bias = make_float2(0.f, 0.f), esc = make_float2(1.f, 1.f), esh = bias, ece = bias;
          if (a.bias != nullptr && (!KS || cur.ks == 0)) bias = __ldg(reinterpret_cast<const float2*>(a.bias + col + 8 * g));   // once per tile
          if (a.e_affine) {
            esc = __ldg(reinterpret_cast<const float2*>(a.e_scale + col + 8 * g));
            esh = __ldg(reinterpret_cast<const float2*>(a.e_shift + col + 8 * g));
            if (a.e_center != nullptr) ece = __ldg(reinterpret_cast<const float2*>(a.e_center + col + 8 * g));
          }
          float2 d2[4];                               // skip-path gradient joining here (ResBlock input gradients)
#pragma unroll
          for (int r = 0; r < 4; ++r) d2[r] = make_float2(0.f, 0.f);
          if (has_add) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
              if ((cur.vmask >> r) & 1)
                d2[r] = __ldg(reinterpret_cast<const float2*>(a.epi_add + (size_t)(cur.obase + r * ostep) * a.Cd + col + 8 * g));
          }
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            if ((cur.vmask >> r) & 1) {
              const int vi = 4 * g + 2 * (r & 1);
              float x0 = __uint_as_float(v[r >> 1][vi]) + bias.x, x1 = __uint_as_float(v[r >> 1][vi + 1]) + bias.y;
              if (a.epi == CVAE_EPI_STATS) {
                f1[2 * g] += x0; f2[2 * g] = fmaf(x0, x0, f2[2 * g]);
                f1[2 * g + 1] += x1; f2[2 * g + 1] = fmaf(x1, x1, f2[2 * g + 1]);
              } else if (dact) {
                const float rc0 = pre[r][g].x - ece.x, rc1 = pre[r][g].y - ece.y;
                x0 += d2[r].x; x1 += d2[r].y;
                const float z0 = fmaf(rc0, esc.x, esh.x), z1 = fmaf(rc1, esc.y, esh.y);
                x0 = z0 > 0.f ? x0 : x0 * a.e_slope;
                x1 = z1 > 0.f ? x1 : x1 * a.e_slope;
                f1[2 * g] += x0; f2[2 * g] = fmaf(x0, rc0, f2[2 * g]);
                f1[2 * g + 1] += x1; f2[2 * g + 1] = fmaf(x1, rc1, f2[2 * g + 1]);
              }
              float* op = a.dst + (size_t)(cur.obase + r * ostep) * a.Cd + col + 8 * g;
              if constexpr (KS) { atomicAdd(op, x0); atomicAdd(op + 1, x1); }       // K slices meet in the pre-zeroed output
              else *reinterpret_cast<float2*>(op) = make_float2(x0, x1);
            }
          }
        }
        prefetch(nxt);                                // in flight while the next accumulator is awaited / loaded
        if (want_stats) {
          if (reg_stats) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { s1[j] += (double)f1[j]; s2[j] += (double)f2[j]; }
          } else {
            // the unit's channels differ from unit to unit: reduce the 8 row-owners of each channel pair with shuffles
            // (lanes t, t^4, t^8, t^16 share t%4) and fold the warp's 32-row sums into its fp64 slice in shared memory.
            // The sums go to fp64 BEFORE the shuffles: the gradient sums cancel heavily, and fp32 partials over 32 rows
            // (instead of this thread's 4) showed up as 1.2 x the tolerance in the batch-of-4 CNN-variant gradient test
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              double a1 = (double)f1[j], a2 = (double)f2[j];
#pragma unroll
              for (int off = 4; off < 32; off <<= 1) {
                a1 += __shfl_xor_sync(0xffffffffu, a1, off);
                a2 += __shfl_xor_sync(0xffffffffu, a2, off);
              }
              if (lane < 4) {                         // this warp's own slots: plain read-modify-write, no atomics
                const int ch = col + (j & 1) + 8 * (j >> 1);
                wst[ch] += a1;
                wst[256 + ch] += a2;
              }
            }
          }
        }
        cur = nxt;
      }
      if (grp >= nunits) {                            // a single-unit tile: the second group only keeps the barrier phase
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s_tempty[acc]));
      }
    }
    if (ew == 0 && lane == 0) T_FLUSH(8, 1)
    if (want_stats) {
      if (reg_stats && grp < nunits) {
        const int c0 = (BN == 32 ? grp * 16 : 0) + cpair;       // tiles_n == 1: n0 = 0
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int off = 4; off < 32; off <<= 1) {
            s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], off);
            s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], off);
          }
        }
        if (lane < 4) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int ch = c0 + (j & 1) + 8 * (j >> 1);
            wst[ch] += s1[j];
            wst[256 + ch] += s2[j];
          }
        }
      }
      hbar_sync(1, kHEpiWarps * 32);
      for (int i = ew * 32 + lane; i < 2 * a.Cd; i += kHEpiWarps * 32) {
        const int which = i / a.Cd, cc = i % a.Cd;
        double sv = 0.0;
#pragma unroll
        for (int w8 = 0; w8 < kHEpiWarps; ++w8) sv += wstat[w8 * 512 + which * 256 + cc];
        if (sv != 0.0) atomicAdd(a.stats + i, sv);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kHProdWarps) tmem_dealloc(tmem, tmem_cols);
}

int build_geom(const cvae_conv_params_t* p, GatherArgs& g);  // conv.cu

static inline int floordiv2(int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); }

// Build the staging plan from the phase / tap geometry.  Returns false when the shape is not covered.
// max_sub: how many taps may be stacked along N in one MMA (1 = no grouping).
static bool build_halo_plan(const GatherArgs& g, HaloPlan& hp, int max_sub) {
  const int is = g.is;
  if (is != 1 && is != 2) return false;
  if (is == 2 && g.nphase != 1) return false;
  hp.nplanes = is * is;
  for (int pl = 0; pl < 4; ++pl) { hp.plane[pl].ntaps = 0; hp.plane[pl].imin = 1 << 20; hp.plane[pl].jmin = 1 << 20; }
  int imax[4] = {-(1 << 20), -(1 << 20), -(1 << 20), -(1 << 20)}, jmax[4] = {imax[0], imax[0], imax[0], imax[0]};
  auto plane_of = [&](const TapEntry& te, int& di, int& dj) {
    const int pr = is == 2 ? (te.dh & 1) : 0, pc = is == 2 ? (te.dw & 1) : 0;
    di = is == 2 ? floordiv2(te.dh) : te.dh; dj = is == 2 ? floordiv2(te.dw) : te.dw;
    hp.plane[pr * is + pc].pr = pr; hp.plane[pr * is + pc].pc = pc;
    return pr * is + pc;
  };
  // pass 1: plane membership and extents
  for (int ph = 0; ph < g.nphase; ++ph)
    for (int t = 0; t < g.phase[ph].ntaps; ++t) {
      int di, dj;
      const int pl = plane_of(g.phase[ph].taps[t], di, dj);
      HaloPlane& P = hp.plane[pl];
      P.imin = min(P.imin, di); P.jmin = min(P.jmin, dj);
      imax[pl] = max(imax[pl], di); jmax[pl] = max(jmax[pl], dj);
    }
  int eh = 0, ew = 0;
  for (int pl = 0; pl < hp.nplanes; ++pl) {
    if (hp.plane[pl].imin > imax[pl]) return false;      // a parity plane without taps (e.g. k = 1, s = 2)
    eh = max(eh, imax[pl] - hp.plane[pl].imin); ew = max(ew, jmax[pl] - hp.plane[pl].jmin);
  }
  hp.R = kHTH + eh; hp.C = kHTW + ew;
  if (hp.R * hp.C > kHMaxSlots) return false;
  for (int ph = 0; ph < g.nphase; ++ph) { hp.ph[ph] = g.phase[ph].ph; hp.pw[ph] = g.phase[ph].pw; hp.pos[ph] = ph; }

  // pass 2: windows (plane, off) with the (phase, widx) pairs that read them
  struct Win { int pl, off, n, phase[4], widx[4]; };
  Win win[64];
  int nwin = 0;
  for (int ph = 0; ph < g.nphase; ++ph)
    for (int t = 0; t < g.phase[ph].ntaps; ++t) {
      int di, dj;
      const TapEntry& te = g.phase[ph].taps[t];
      const int pl = plane_of(te, di, dj);
      const int off = (di - hp.plane[pl].imin) * hp.C + (dj - hp.plane[pl].jmin);
      int w = 0;
      while (w < nwin && !(win[w].pl == pl && win[w].off == off)) ++w;
      if (w == nwin) { if (nwin == 64) return false; win[nwin++] = {pl, off, 0, {0, 0, 0, 0}, {0, 0, 0, 0}}; }
      if (win[w].n == 4) return false;                   // one window is read at most once per phase
      win[w].phase[win[w].n] = ph; win[w].widx[win[w].n] = te.widx; ++win[w].n;
    }
  // pass 3: choose accumulator positions so that every window's phases are adjacent (try all orders),
  // split windows wider than max_sub, and order the groups so that a group never mixes first-touch and
  // accumulate positions (largest windows first)
  int perm[4] = {0, 1, 2, 3}, best[4] = {0, 1, 2, 3};
  bool found = max_sub <= 1 || g.nphase == 1;
  if (!found) {
    int idx[4] = {0, 1, 2, 3};
    auto contiguous = [&](const int* pos) {
      for (int w = 0; w < nwin; ++w) {
        int lo = 4, hi = -1;
        for (int i = 0; i < win[w].n; ++i) { lo = min(lo, pos[win[w].phase[i]]); hi = max(hi, pos[win[w].phase[i]]); }
        if (hi - lo + 1 != win[w].n) return false;
      }
      return true;
    };
    // all permutations of up to 4 phases (Heap-free brute force)
    for (int a0 = 0; a0 < g.nphase && !found; ++a0)
      for (int a1 = 0; a1 < g.nphase && !found; ++a1)
        for (int a2 = 0; a2 < g.nphase && !found; ++a2)
          for (int a3 = 0; a3 < g.nphase && !found; ++a3) {
            idx[0] = a0; idx[1] = a1; idx[2] = a2; idx[3] = a3;
            bool ok = true;
            for (int i = 0; i < g.nphase && ok; ++i)
              for (int j = i + 1; j < g.nphase; ++j)
                if (idx[i] == idx[j]) { ok = false; break; }
            if (!ok) continue;
            for (int i = 0; i < g.nphase; ++i) perm[i] = idx[i];     // perm[phase] = position
            if (contiguous(perm)) { found = true; for (int i = 0; i < 4; ++i) best[i] = perm[i]; }
          }
  }
  const bool grouped = found && max_sub > 1 && g.nphase > 1;
  if (grouped) for (int ph = 0; ph < g.nphase; ++ph) hp.pos[ph] = best[ph];
  // emit groups: windows sorted by size (descending) within each plane
  uint32_t started = 0;
  int max_group = 1;
  for (int pass = 4; pass >= 1; --pass)
    for (int w = 0; w < nwin; ++w) {
      if (win[w].n != pass) continue;
      HaloPlane& P = hp.plane[win[w].pl];
      // sub-taps in accumulator-position order
      int order[4] = {0, 1, 2, 3};
      for (int i = 0; i < win[w].n; ++i)
        for (int j = i + 1; j < win[w].n; ++j)
          if (hp.pos[win[w].phase[order[j]]] < hp.pos[win[w].phase[order[i]]]) { const int tmp = order[i]; order[i] = order[j]; order[j] = tmp; }
      const int chunk = grouped ? max_sub : 1;
      for (int i0 = 0; i0 < win[w].n; i0 += chunk) {
        if (P.ntaps >= 16) return false;
        HaloTap& T = P.taps[P.ntaps++];
        T.off = win[w].off; T.nsub = min(chunk, win[w].n - i0); T.pos0 = hp.pos[win[w].phase[order[i0]]];
        for (int i = 0; i < 4; ++i) T.widx[i] = i < T.nsub ? win[w].widx[order[i0 + i]] : 0;
        const uint32_t gmask = ((1u << T.nsub) - 1u) << T.pos0;
        max_group = max(max_group, T.nsub);
        (void)started; (void)gmask;
      }
    }
  // the kernel derives one accumulate flag per group from the positions already written: verify that
  // no group mixes written and unwritten positions in ISSUE order (k-block major, plane, tap)
  for (int pl = 0; pl < hp.nplanes; ++pl)
    for (int t = 0; t < hp.plane[pl].ntaps; ++t) {
      const HaloTap& T = hp.plane[pl].taps[t];
      const uint32_t gmask = ((1u << T.nsub) - 1u) << T.pos0;
      if ((started & gmask) != 0 && (started & gmask) != gmask) return false;
      started |= gmask;
    }
  hp.bslot_bytes = max_group;     // finished by the caller (x 256 * BN)
  return true;
}

// exported to conv_tc.cu: returns CVAE_OK when launched, 1 when the shape is not covered (caller falls
// back to the per-tap gather kernel), < 0 on error.
int launch_conv_halo_tc(const GatherArgs& g_in, cudaStream_t st) {
  GatherArgs g = g_in;
  if (g.wtaps < 2) {
    // 1x1 / Linear: a plain GEMM over M = N*H*W rows.  Viewed as ONE image of M/8 x 8 positions, a
    // 16 x 8 tile is 128 consecutive rows and the single "tap" is the tile itself (no halo).
    const long long M = (long long)g.N * g.Hs * g.Ws;
    if (g.nphase != 1 || g.is != 1 || g.os != 1 || g.Hs != g.Hd || g.Ws != g.Wd || M % 8 != 0 || M >= (1ll << 31)) return 1;
    if (g.phase[0].ntaps != 1 || g.phase[0].taps[0].dh != 0 || g.phase[0].taps[0].dw != 0) return 1;
    g.N = 1; g.Hs = g.Hd = (int)(M / 8); g.Ws = g.Wd = 8;
    g.phase[0].Hq = g.Hs; g.phase[0].Wq = 8;
  }
  if (g.Cs % 16 != 0 || g.Cd % 16 != 0) return 1;
  if (g.in_act && !(g.in_slope >= 0.f && g.in_slope <= 1.f)) return 1;      // the producers use max(v, slope * v)
  if (g.Cd > 256 && g.epi != CVAE_EPI_PLAIN && g.stats != nullptr) return 1;   // s_stat holds 256 channels
  // cross-term accumulators (xacc, see the MMA issuer) double the accumulator columns.  Gather plans keep their tile
  // width (N <= 128: 2 * 2 * 128 = 512 columns); scatter plans (4 phases) fit with BN <= 32, which costs the wide
  // layers more n-tiles (measured: stem.9 / stem.12 input gradients +25 %).  Forward accuracy is what decides which side
  // of a LeakyReLU kink a unit takes, so the wide layers pay that price only in forward-type launches (no
  // activation-derivative epilogue); input gradients with Cd >= 64 keep the single accumulator (their 3-8e-6 is far
  // inside the 1e-4 gradient tolerance).  CVAE_XACC_SCATTER=0 / 2: never / always.
  static const bool xacc_on = [] { const char* e = getenv("CVAE_XACC"); return !(e && e[0] == '0'); }();
  static const int xacc_sc = [] { const char* e = getenv("CVAE_XACC_SCATTER"); return e ? atoi(e) : 1; }();
  const bool want_xs = xacc_on && g.nphase > 1 &&
                       (xacc_sc == 2 || (xacc_sc == 1 && (g.Cd % 64 != 0 || g.epi != CVAE_EPI_DACT)));
  int bn = 0;
  for (int c : {128, 64, 32, 16})
    if (g.Cd % c == 0 && 2 * g.nphase * c * (want_xs ? 2 : 1) <= 512) { bn = c; break; }
  if (bn == 0) return 1;
  HaloPlan hp;
  // stack up to 128 output columns per MMA (weight-ring slot <= 32 KiB); fall back to one tap per MMA
  const int max_sub = g.nphase > 1 ? max(1, min(4, 128 / bn)) : 1;
  if (!build_halo_plan(g, hp, max_sub) && !build_halo_plan(g, hp, 1)) return 1;
  // q-space extent: identical for every phase of the layers covered here (even output sizes)
  hp.Hq = g.phase[0].Hq; hp.Wq = g.phase[0].Wq;
  for (int i = 1; i < g.nphase; ++i)
    if (g.phase[i].Hq != hp.Hq || g.phase[i].Wq != hp.Wq) return 1;
  hp.BN = bn;
  {  // Linear / 1x1 layers: A operand through tensor memory (CVAE_LIN_TMEM=0 keeps it in shared memory)
    static const bool lin_tmem = [] { const char* e = getenv("CVAE_LIN_TMEM"); return !(e && e[0] == '0'); }();
    hp.a_pre = (g.a_image != nullptr && g.wtaps < 2 && hp.nplanes == 1 && g.nphase == 1 && hp.plane[0].ntaps == 1 &&
                !g.in_affine && !g.in_act) ? 1 : 0;
    if (g.a_image != nullptr && !hp.a_pre) return 1;
    hp.a_tmem = (!hp.a_pre && lin_tmem && g.wtaps < 2 && hp.nplanes == 1 && g.nphase == 1 && hp.plane[0].ntaps == 1 &&
                 2 * bn + kHNAT * 64 <= 512) ? 1 : 0;
    // separate cross-term accumulators: gather-type plans (one phase, one tap per MMA) whose doubled accumulators
    // still fit tensor memory twice (CVAE_XACC=0 restores the single-accumulator 3-MMA form)
    hp.xacc = (xacc_on && g.nphase == 1 && hp.bslot_bytes == 1 &&
               2 * 2 * bn + (hp.a_tmem ? kHNAT * 64 : 0) <= 512) ? 1 : 0;
    if (want_xs && 2 * 2 * g.nphase * bn <= 512) hp.xacc = 1;
  }
  hp.bslot_bytes *= 256 * bn;
  hp.NB = hp.bslot_bytes >= 32768 ? (bn >= 128 ? 3 : 2) : 4;
  if (hp.bslot_bytes == 32768 && bn < 128) hp.NB = 2;
  hp.tiles_h = (hp.Hq + kHTH - 1) / kHTH; hp.tiles_w = (hp.Wq + kHTW - 1) / kHTW; hp.tiles_n = g.Cd / bn;
  {
    // A tile is a 16 x 8 patch of ONE image: feature maps smaller than that (4 x 4 in the causal_cascade stack: 16 of
    // 128 MMA rows in use, its ConvTranspose2d 256->128 ran 631 us; 8 x 8: half of the rows, and a scatter plan with
    // cross-term accumulators then also splits 64 output channels into two n-tiles -- cascade dec_conv.2 forward 172 us
    // for the FLOPs its 4 x 4 neighbour does in 46 us) go to the per-tap gather kernel, whose 128-row tiles run across
    // images (CVAE_HALO_MIN_UTIL: least percentage of tile rows in use, default 60; measured 40 -> 60: cascade step
    // 2.27 -> 2.12 ms, vessel step 7.610 -> 7.596 ms, latent_translator and mnist unchanged)
    static const int min_util = [] { const char* e = getenv("CVAE_HALO_MIN_UTIL"); return e ? atoi(e) : 60; }();
    if (g.wtaps >= 2 && 100ll * hp.Hq * hp.Wq < (long long)min_util * hp.tiles_h * hp.tiles_w * kHTH * kHTW) return 1;
  }
  long long total = (long long)g.N * hp.tiles_h * hp.tiles_w * hp.tiles_n;
  if (total >= (1ll << 31)) return 1;
  // Linear launches whose tiles leave three quarters of the SMs idle and whose K loop is long (latent_translator fc2 and the
  // input gradient of fc1: K = 1024, 34 tiles at M = 2176): two CTAs per tile, each with half of the k-blocks, adding into
  // the pre-zeroed output.  Two addends meet a zero, so the sum does not depend on their order (bitwise reproducible).
  // Measured (scripts/ab_small.py / ab_step.py, CVAE_LIN_KSPLIT=1 turns it off): latent_translator step 5.88 -> 5.77 ms; the
  // vessel step's 66-tile launches (M = 4160) LOSE with it (7.58 -> 7.63 ms: the zeroing launch and the atomics cost more
  // than the shorter K loop returns, and in backward the idle SMs were already taken by the weight-gradient stream), hence
  // the quarter-of-the-machine threshold.
  hp.ksplit = 1;
  {
    static const int ks_env = [] { const char* e = getenv("CVAE_LIN_KSPLIT"); return e ? atoi(e) : 2; }();
    const int KBh = (g.Cs + 31) / 32;
    if (ks_env == 2 && hp.a_tmem && g.epi == CVAE_EPI_PLAIN && 4 * total <= kNumSMs && KBh >= 16 && KBh % 2 == 0 && g.Cs % 32 == 0) {
      if (cudaMemsetAsync(g.dst, 0, sizeof(float) * (size_t)g.Hd * 8 * g.Cd, st) != cudaSuccess) return CVAE_ERR_LAUNCH;
      hp.ksplit = 2;
      total *= 2;
    }
  }
  // A ring: tensor memory in Linear mode; else as many shared-memory stages (2 or 3) as fit beside the weight ring --
  // the producers issue a stage's global loads before they wait for its slot, so a deeper ring is more DRAM latency hidden
  static const bool pf = [] { const char* e = getenv("CVAE_HALO_PREFETCH"); return e && e[0] == '1'; }();
  hp.prefetch = pf ? 1 : 0;
  static const int fmode = [] { const char* e = getenv("CVAE_HALO_FENCE"); return e ? atoi(e) : 1; }();
  hp.fence_mode = fmode;
  static const int na_max = [] { const char* e = getenv("CVAE_HALO_NA"); return e ? max(2, min(kHNA, atoi(e))) : 2; }();   // measured: 3 stages no faster
  hp.na = hp.a_tmem ? kHNAT : (hp.a_pre ? 3 : 2);
  hp.a_stage = hp.a_pre ? 32768 : kHAStage;
  hp.a_half = hp.a_pre ? 16384 : kHHalf;
  if (!hp.a_tmem && !hp.a_pre && na_max >= 3 && (size_t)3 * kHAStage + (size_t)hp.NB * hp.bslot_bytes + 1024 <= kHMaxDyn) hp.na = 3;
  const size_t stat_bytes = (g.epi != CVAE_EPI_PLAIN && g.stats != nullptr) ? (size_t)kHEpiWarps * 512 * sizeof(double) : 0;
  // producer warps (template parameter): 6 + pipelined stages for the activation-derivative launches, 8 otherwise
  static const int pw_env = [] { const char* e = getenv("CVAE_HALO_PW"); return e ? atoi(e) : 0; }();
  // measured on the vessel step (B = 64): 8 everywhere 8.23 ms, 6 for every activation-derivative launch 8.38 ms, 6 everywhere
  // 8.69 ms -- the 6-warp form only pays on the HBM-shaped scatter launches with few channels (stem.3 input gradient
  // 175 -> 152 us), so that is where it is used
  static const int pw10 = [] { const char* e = getenv("CVAE_HALO_PW10"); return e ? atoi(e) : 0; }();   // 1: forward-type conv launches, 2: + DACT ones
  int pw = pw_env == 6 || pw_env == 8 || pw_env == 10 ? pw_env
           : ((g.epi == CVAE_EPI_DACT && g.nphase > 1 && g.Cd <= 32 && !hp.a_tmem && !hp.a_pre) ? 6 : 8);
  if (pw_env == 0 && pw == 8 && !hp.a_tmem && !hp.a_pre && (pw10 == 2 || (pw10 == 1 && g.epi != CVAE_EPI_DACT))) pw = 10;
  const size_t a_bytes = hp.a_tmem ? (size_t)kHRawStages * (pw >= 8 ? 256 * 80 : 128 * 144) : (size_t)hp.na * hp.a_stage;
  while (hp.NB > 2 && a_bytes + (size_t)hp.NB * hp.bslot_bytes + stat_bytes + 1024 > kHMaxDyn) --hp.NB;   // shallower weight ring
  size_t smem = a_bytes + (size_t)hp.NB * hp.bslot_bytes;
  hp.stat_off = (int)smem;
  smem += stat_bytes + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(conv_halo_tc_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHMaxDyn) != cudaSuccess ||
        cudaFuncSetAttribute(conv_halo_tc_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHMaxDyn) != cudaSuccess ||
        cudaFuncSetAttribute(conv_halo_tc_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHMaxDyn) != cudaSuccess ||
        cudaFuncSetAttribute(conv_halo_tc_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHMaxDyn) != cudaSuccess)
      return CVAE_ERR_LAUNCH;
    attr_set = true;
  }
  if (smem > kHMaxDyn) return 1;
  const int grid = (int)min(total, (long long)kNumSMs);
  if (hp.ksplit > 1) conv_halo_tc_kernel<8, true><<<grid, h_threads(8), smem, st>>>(g, hp, (int)total);     // Linear mode: pw == 8
  else if (pw == 6) conv_halo_tc_kernel<6><<<grid, h_threads(6), smem, st>>>(g, hp, (int)total);
  else if (pw == 10) conv_halo_tc_kernel<10><<<grid, h_threads(10), smem, st>>>(g, hp, (int)total);
  else conv_halo_tc_kernel<8><<<grid, h_threads(8), smem, st>>>(g, hp, (int)total);
  if (cudaPeekAtLastError() != cudaSuccess) { cudaGetLastError(); return CVAE_ERR_LAUNCH; }
  return CVAE_OK;
}

}  // namespace cvae

// Debug hook (timing builds): [0] producer wait-empty, [1] producer total, [2] mma wait-tmem-empty, [3] mma wait-A,
// [4] mma wait-B, [5] mma total, [6] B-loader wait-empty, [7] B-loader total, [8] epilogue wait-full, [9] epilogue total
// (clock cycles summed over CTAs).  Returns 0 when the library was built without CVAE_TIMING.
extern "C" int cvae_debug_read(unsigned long long* out16, int reset) {
#ifdef CVAE_TIMING
  unsigned long long z[16] = {0};
  if (out16 && cudaMemcpyFromSymbol(out16, cvae::g_halo_dbg, sizeof(z)) != cudaSuccess) return -1;
  if (reset && cudaMemcpyToSymbol(cvae::g_halo_dbg, z, sizeof(z)) != cudaSuccess) return -1;
  return 1;
#else
  (void)out16; (void)reset;
  return 0;
#endif
}
