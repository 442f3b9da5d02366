// Linear layers with a handful of rows (M = batch <= 512; 128 rows per CTA): the adapters, the morphology head and
// decoder_input of CausalViTVAE (vessel_analysis/00_core/models.py:225-250, vit_backbone.py:186-188), and
// every Linear of the MNIST / cascade models.
//
// As [128-row tile] x [64-column tile] implicit GEMMs these ran on 1-8 CTAs with a serial K loop: 15-70 us
// per launch for a few MFLOP (ncu launch list, profiles/), ~0.7 ms of the vessel step.  Here the work is cut
// along N *and* K so that even a 64 x 512 x 256 product fills the GPU:
//
//   forward / input gradient:  y[m][n] (+)= sum_{k in chunk} xf(x[m][k]) * w[k][n]
//     CTA = 32 columns x one 64-deep K chunk x all rows; x chunk staged in shared memory (transform applied),
//     a lane owns one column (coalesced weight reads), a warp owns M/4 rows (x values are 128-bit shared
//     broadcasts); K chunks meet in the pre-zeroed output with fp32 atomics.  Bias rides on chunk 0; the
//     BatchNorm-statistics / activation-derivative epilogues run as a second tiny kernel over the finished
//     [M, N] matrix.
//   weight gradient:  P[ca][cb] = sum_m xa(ga[m][ca]) * xb(db[m][cb])   (one pass, no K split).
#include "conv_args.cuh"

namespace cvae {

constexpr int kLsKC = 64;        // K chunk per CTA
constexpr int kLsMaxM = 512;     // rows beyond 128 go to further row blocks (grid z): the cascade model trains at batch 256

template <int R>   // rows per warp; a CTA covers rows [4 R z, 4 R (z + 1)) of the M
__global__ void __launch_bounds__(128) linear_small_kernel(const __grid_constant__ GatherArgs a, const int M, const int cpc) {
  __shared__ __align__(16) float xs[4 * R * kLsKC];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mb = blockIdx.z * 4 * R;                 // first row of this CTA's row block
  const int n = blockIdx.x * 32 + lane;
  const bool nok = n < a.Cd;
  float acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.f;
  const float b = (nok && blockIdx.y == 0 && a.bias != nullptr) ? __ldg(a.bias + n) : 0.f;   // requested with the weights
  // `cpc` consecutive K chunks per CTA (launch_linear_small: as many as still leave ~4 CTAs per SM), so a wide layer
  // (decoder_input, 512 -> 16384: 512 column blocks) runs without K split -- no pre-zeroing, no atomics: its 8.4 M
  // fp32 atomics were ~30 us of a 105 us launch -- and the 16384 -> 512 input gradient meets in 43 instead of 256 adds
  for (int kc = 0; kc < cpc; ++kc) {
  const int k0 = (blockIdx.y * cpc + kc) * kLsKC;
  if (k0 >= a.Cs) break;
  if (kc > 0) __syncthreads();
  // ---- this lane's 64 weights of the chunk: ALL requested before anything waits on them ----
  // (the K loop used to fetch four per iteration: 16 dependent round trips to weights that left L2 a step ago, which
  // was the whole life of these CTAs -- 14-18 us for a 64 x 64 x 12 layer; the loads now overlap the x staging too)
  const int kmax = min(kLsKC, a.Cs - k0);            // Cs % 4 == 0
  float wv[kLsKC];
  {
    const float* wp = a.wt + (size_t)k0 * a.Cd + (nok ? n : 0);
#pragma unroll
    for (int kk = 0; kk < kLsKC; ++kk) wv[kk] = (nok && kk < kmax) ? __ldg(wp + (size_t)kk * a.Cd) : 0.f;
  }
  // ---- stage x[:, k0 : k0 + 64] with the producer's BatchNorm + activation applied ----
  // A thread always stages the same four channels (128 threads = 8 rows x 16 vectors per pass), so their coefficients are
  // loaded once, and the row loads go out in batches BEFORE the first shared-memory store of the batch: the plain
  // load -> transform -> store loop is not reordered across the stores, i.e. one dependent round trip per vector
  // (R / 2 of them: the other half of these CTAs' 14-18 us).
  {
    constexpr int kIt = R / 2, kB = kIt < 8 ? kIt : 8;
    const int c4 = tid & 15, k = k0 + c4 * 4, mrow = tid >> 4;
    const bool kok = k < a.Cs;
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f), ce = sh;
    if (a.in_affine && kok) {
      sc = __ldg(reinterpret_cast<const float4*>(a.in_scale + k));
      sh = __ldg(reinterpret_cast<const float4*>(a.in_shift + k));
      if (a.in_center != nullptr) ce = __ldg(reinterpret_cast<const float4*>(a.in_center + k));
    }
#pragma unroll
    for (int b0 = 0; b0 < kIt; b0 += kB) {
      float4 v[kB];
      bool ok[kB];
#pragma unroll
      for (int u = 0; u < kB; ++u) {
        const int m = mrow + (b0 + u) * 8;
        ok[u] = kok && mb + m < M;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok[u]) v[u] = __ldg(reinterpret_cast<const float4*>(a.src + (size_t)(mb + m) * a.Cs + k));
      }
#pragma unroll
      for (int u = 0; u < kB; ++u) {
        const int m = mrow + (b0 + u) * 8;
        float4 x = v[u];
        if (ok[u]) {
          if (a.in_affine) {
            x.x = fmaf(x.x - ce.x, sc.x, sh.x); x.y = fmaf(x.y - ce.y, sc.y, sh.y);
            x.z = fmaf(x.z - ce.z, sc.z, sh.z); x.w = fmaf(x.w - ce.w, sc.w, sh.w);
          }
          if (a.in_act) { x.x = lrelu(x.x, a.in_slope); x.y = lrelu(x.y, a.in_slope); x.z = lrelu(x.z, a.in_slope); x.w = lrelu(x.w, a.in_slope); }
        }
        *reinterpret_cast<float4*>(xs + m * kLsKC + c4 * 4) = x;
      }
    }
  }
  __syncthreads();
  const float* xr = xs + warp * R * kLsKC;
  float part[R];                                     // this chunk's sums: chains stay 64 long whatever cpc is
#pragma unroll
  for (int r = 0; r < R; ++r) part[r] = 0.f;
#pragma unroll
  for (int kk = 0; kk < kLsKC; kk += 4) {            // x beyond Cs is staged as 0 and its weights are 0: no tail case
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float4 xv = *reinterpret_cast<const float4*>(xr + r * kLsKC + kk);
      part[r] = fmaf(xv.x, wv[kk], fmaf(xv.y, wv[kk + 1], fmaf(xv.z, wv[kk + 2], fmaf(xv.w, wv[kk + 3], part[r]))));
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] += part[r];
  }
  if (!nok) return;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int m = mb + warp * R + r;
    if (m < M) {
      if (gridDim.y == 1) a.dst[(size_t)m * a.Cd + n] = acc[r] + b;      // single K chunk: plain store, no pre-zeroing
      else atomicAdd(a.dst + (size_t)m * a.Cd + n, acc[r] + b);
    }
  }
}

// second pass over the finished [M, Cd] matrix: BatchNorm statistics, or activation derivative (+ residual
// add) with the BN-backward sums.  Block = 32 columns x 4 row groups.
__global__ void __launch_bounds__(128) linear_small_epi_kernel(const __grid_constant__ GatherArgs a, const int M) {
  __shared__ double red[2][4][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 32 + lane;
  double s1 = 0.0, s2 = 0.0;
  if (n < a.Cd) {
    float esc = 1.f, esh = 0.f, ece = 0.f;
    if (a.e_affine) { esc = __ldg(a.e_scale + n); esh = __ldg(a.e_shift + n); if (a.e_center) ece = __ldg(a.e_center + n); }
    // rows in batches of 8 with every load of a batch issued before its first store: the matrix is read and (DACT)
    // rewritten in place, so the plain loop was one dependent round trip per row (10-12 us at 64 rows, 40 us at 256)
    const bool dact = a.epi != CVAE_EPI_STATS, has_add = dact && a.epi_add != nullptr;
    for (int mbase = warp; mbase < M; mbase += 32) {
      float x[8], rf[8], ad[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int m = mbase + 4 * u;
        x[u] = 0.f; rf[u] = 0.f; ad[u] = 0.f;
        if (m < M) {
          const size_t o = (size_t)m * a.Cd + n;
          x[u] = a.dst[o];
          if (dact) rf[u] = __ldg(a.epi_ref + o);
          if (has_add) ad[u] = __ldg(a.epi_add + o);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int m = mbase + 4 * u;
        if (m >= M) continue;
        float xv = x[u];
        if (!dact) {
          s1 += (double)xv; s2 += (double)xv * (double)xv;
        } else {   // CVAE_EPI_DACT
          const float refc = rf[u] - ece;
          if (has_add) xv += ad[u];
          const float z = fmaf(refc, esc, esh);
          xv = z > 0.f ? xv : xv * a.e_slope;
          a.dst[(size_t)m * a.Cd + n] = xv;
          s1 += (double)xv; s2 += (double)xv * (double)refc;
        }
      }
    }
  }
  red[0][warp][lane] = s1; red[1][warp][lane] = s2;
  __syncthreads();
  if (warp == 0 && n < a.Cd && a.stats != nullptr) {
    atomicAdd(a.stats + n, red[0][0][lane] + red[0][1][lane] + red[0][2][lane] + red[0][3][lane]);
    atomicAdd(a.stats + a.Cd + n, red[1][0][lane] + red[1][1][lane] + red[1][2][lane] + red[1][3][lane]);
  }
}

// ---- wide layers with a handful of rows: decoder_input (Linear 512 -> 16384 at M = batch, vit_backbone.py:186-188), its
// input gradient (K = 16384) and weight gradient (512 x 16384 outputs) ----------------------------------------------
// The 32-column / 64-deep CTAs above make 4096 CTAs with 8.4 M fp32 atomics out of such a layer (105 / 110 / 102 us for
// 0.5 GFMA and 33.5 MB of weights).  Here a CTA owns a 64 x 128 output tile, a thread a 4 x 8 register block, and the
// reduction dimension streams through shared memory in chunks of 16 (weights by cp.async, double-buffered):
// 32 FMA per 3 shared-memory loads.  Split K (grid.y) only where the output has too few tiles to fill the GPU.
constexpr int kLrK = 16;

__device__ __forceinline__ void lr_fma_chunk(const float (*xs)[68], const float (*ws)[128], int ty, int tx, float (&acc)[4][8]) {
#pragma unroll
  for (int kk = 0; kk < kLrK; ++kk) {
    const float4 x4 = *reinterpret_cast<const float4*>(&xs[kk][4 * ty]);
    const float4 w0 = *reinterpret_cast<const float4*>(&ws[kk][4 * tx]);
    const float4 w1 = *reinterpret_cast<const float4*>(&ws[kk][64 + 4 * tx]);
    const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {                     // 4 x 8 block as eight packed FMAs per row pair (common.cuh)
      fma4(*reinterpret_cast<float(*)[4]>(&acc[i][0]), xv[i], w0);
      fma4(*reinterpret_cast<float(*)[4]>(&acc[i][4]), xv[i], w1);
    }
  }
}

// rows [r0, r0 + 16) x columns [n0, n0 + 128) of a row-major [R][ld] matrix -> ws[16][128] (16-byte cp.async, zero fill)
__device__ __forceinline__ void lr_load_w(float (*ws)[128], const float* __restrict__ src, int r0, int R, int n0, int ld, int tid) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int idx = tid + u * 256, kk = idx >> 5, c4 = (idx & 31) << 2;
    float* d = &ws[kk][c4];
    if (r0 + kk < R && n0 + c4 < ld) {
      const uint32_t da = static_cast<uint32_t>(__cvta_generic_to_shared(d));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(da), "l"(src + (size_t)(r0 + kk) * ld + n0 + c4) : "memory");
    } else {
      *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// y[m][n] (+)= sum_k xf(x[m][k]) * w[k][n];  grid (Cd / 128, K splits, row blocks of 64)
__global__ void __launch_bounds__(256) linear_rows_kernel(const __grid_constant__ GatherArgs a, const int M, const int kper) {
  // weights: a ring of kLrWS chunks by cp.async, kLrWS - 1 of them in flight (with two buffers a CTA had ONE 8 KB chunk in
  // flight: 128 CTAs x 8 KB against the ~6 MB that the HBM rate x latency asks for -- decoder_input's 33.5 MB of weights
  // took 40 us); x (a few rows, L2-resident) stays double-buffered through registers
  constexpr int kLrWS = 4;
  __shared__ __align__(16) float xs[2][kLrK][68];
  __shared__ __align__(16) float ws[kLrWS][kLrK][128];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int n0 = blockIdx.x * 128, m0 = blockIdx.z * 64;
  const int kbeg = blockIdx.y * kper, kend = min(a.Cs, kbeg + kper);
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const int xm = tid >> 2, xk = (tid & 3) << 2;           // this thread's x vector of a chunk: row xm, k offset xk
  auto load_x = [&](int k0) -> float4 {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const int k = k0 + xk;
    if (m0 + xm < M && k < kend) {
      v = __ldg(reinterpret_cast<const float4*>(a.src + (size_t)(m0 + xm) * a.Cs + k));
      if (a.in_affine) {
        const float4 sc = __ldg(reinterpret_cast<const float4*>(a.in_scale + k));
        const float4 sh = __ldg(reinterpret_cast<const float4*>(a.in_shift + k));
        float4 ce = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.in_center != nullptr) ce = __ldg(reinterpret_cast<const float4*>(a.in_center + k));
        v.x = fmaf(v.x - ce.x, sc.x, sh.x); v.y = fmaf(v.y - ce.y, sc.y, sh.y);
        v.z = fmaf(v.z - ce.z, sc.z, sh.z); v.w = fmaf(v.w - ce.w, sc.w, sh.w);
      }
      if (a.in_act) { v.x = lrelu(v.x, a.in_slope); v.y = lrelu(v.y, a.in_slope); v.z = lrelu(v.z, a.in_slope); v.w = lrelu(v.w, a.in_slope); }
    }
    return v;
  };
  auto store_x = [&](int b, const float4 v) { xs[b][xk][xm] = v.x; xs[b][xk + 1][xm] = v.y; xs[b][xk + 2][xm] = v.z; xs[b][xk + 3][xm] = v.w; };
  const int nch = (kend - kbeg + kLrK - 1) / kLrK;
  if (nch > 0) store_x(0, load_x(kbeg));
#pragma unroll
  for (int s = 0; s < kLrWS - 1; ++s) {              // one commit group per chunk slot, empty past the end: uniform counts
    if (s < nch) lr_load_w(ws[s], a.wt, kbeg + s * kLrK, kend, n0, a.Cd, tid);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int c = 0; c < nch; ++c) {
    const int b = c & 1;
    float4 xn = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c + 1 < nch) xn = load_x(kbeg + (c + 1) * kLrK);
    if (c + kLrWS - 1 < nch)                         // into the slot chunk c - 1 left (its reads ended before the last barrier)
      lr_load_w(ws[(c + kLrWS - 1) % kLrWS], a.wt, kbeg + (c + kLrWS - 1) * kLrK, kend, n0, a.Cd, tid);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(kLrWS - 1) : "memory");
    __syncthreads();
    lr_fma_chunk(xs[b], ws[c % kLrWS], ty, tx, acc);
    if (c + 1 < nch) store_x(b ^ 1, xn);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + 4 * ty + i;
    if (m >= M) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + 64 * h + 4 * tx;
      if (n >= a.Cd) continue;
      float4 o = make_float4(acc[i][4 * h], acc[i][4 * h + 1], acc[i][4 * h + 2], acc[i][4 * h + 3]);
      if (blockIdx.y == 0 && a.bias != nullptr) {
        const float4 bv = __ldg(reinterpret_cast<const float4*>(a.bias + n));
        o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
      }
      float* d = a.dst + (size_t)m * a.Cd + n;
      if (gridDim.y == 1) *reinterpret_cast<float4*>(d) = o;
      else { atomicAdd(d, o.x); atomicAdd(d + 1, o.y); atomicAdd(d + 2, o.z); atomicAdd(d + 3, o.w); }
    }
  }
}

// 1: launched, 0: not covered
int launch_linear_rows(const GatherArgs& g, int M, cudaStream_t st) {
  if (M > kLsMaxM || (g.Cs & 3) || (g.Cd & 3) || (long long)g.Cs * g.Cd < (1ll << 21)) return 0;
  if (((reinterpret_cast<uintptr_t>(g.src) | reinterpret_cast<uintptr_t>(g.wt) | reinterpret_cast<uintptr_t>(g.dst)) & 15) != 0) return 0;
  if (g.bias && (reinterpret_cast<uintptr_t>(g.bias) & 15) != 0) return 0;
  if (g.in_affine && (((reinterpret_cast<uintptr_t>(g.in_scale) | reinterpret_cast<uintptr_t>(g.in_shift)) & 15) != 0 ||
                      (g.in_center && (reinterpret_cast<uintptr_t>(g.in_center) & 15) != 0))) return 0;
  const int tiles = ((g.Cd + 127) / 128) * ((M + 63) / 64);
  int splits = max(1, min((g.Cs + 255) / 256, (2 * kNumSMs) / max(tiles, 1)));    // >= 256 deep per CTA, ~2 CTAs per SM
  if (4 * tiles >= 3 * kNumSMs) splits = 1;                                      // enough tiles: plain stores, no atomics
  int kper = (g.Cs + splits - 1) / splits;
  kper = ((kper + kLrK - 1) / kLrK) * kLrK;
  splits = (g.Cs + kper - 1) / kper;
  if (splits > 1 && cudaMemsetAsync(g.dst, 0, sizeof(float) * (size_t)M * g.Cd, st) != cudaSuccess) return CVAE_ERR_LAUNCH;
  linear_rows_kernel<<<dim3((g.Cd + 127) / 128, splits, (M + 63) / 64), 256, 0, st>>>(g, M, kper);
  if (g.epi != CVAE_EPI_PLAIN) linear_small_epi_kernel<<<(g.Cd + 31) / 32, 128, 0, st>>>(g, M);
  return 1;
}

// weight gradient of such a layer: P[ca][cb] = sum_m xa(ga[m][ca]) * xb(db[m][cb]), m < M: the same 64 x 128 tile with
// the batch as the streamed dimension (both operands are read row-wise as they lie: no transposition)
__global__ void __launch_bounds__(256) wgrad_rows_kernel(const __grid_constant__ WgradArgs a, const int M) {
  __shared__ __align__(16) float as[2][kLrK][68];
  __shared__ __align__(16) float bs[2][kLrK][128];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int ca0 = blockIdx.y * 64, cb0 = blockIdx.x * 128;
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  auto load_a = [&](int b, int m0) {        // 16 rows x 64 channels of ga (transform applied): one float4 per thread
    const int kk = tid >> 4, c = ca0 + ((tid & 15) << 2);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m0 + kk < M && c < a.Ca) {
      v = __ldg(reinterpret_cast<const float4*>(a.ga + (size_t)(m0 + kk) * a.Ca + c));
      if (a.a_affine) {
        const float4 sc = __ldg(reinterpret_cast<const float4*>(a.a_scale + c));
        const float4 sh = __ldg(reinterpret_cast<const float4*>(a.a_shift + c));
        float4 ce = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.a_center != nullptr) ce = __ldg(reinterpret_cast<const float4*>(a.a_center + c));
        v.x = fmaf(v.x - ce.x, sc.x, sh.x); v.y = fmaf(v.y - ce.y, sc.y, sh.y);
        v.z = fmaf(v.z - ce.z, sc.z, sh.z); v.w = fmaf(v.w - ce.w, sc.w, sh.w);
      }
      if (a.a_act) { v.x = lrelu(v.x, a.a_slope); v.y = lrelu(v.y, a.a_slope); v.z = lrelu(v.z, a.a_slope); v.w = lrelu(v.w, a.a_slope); }
    }
    *reinterpret_cast<float4*>(&as[b][kk][(tid & 15) << 2]) = v;
  };
  const int nch = (M + kLrK - 1) / kLrK;
  load_a(0, 0);
  lr_load_w(bs[0], a.db, 0, M, cb0, a.Cb, tid);
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int c = 0; c < nch; ++c) {
    const int b = c & 1;
    if (c + 1 < nch) {
      load_a(b ^ 1, (c + 1) * kLrK);
      lr_load_w(bs[b ^ 1], a.db, (c + 1) * kLrK, M, cb0, a.Cb, tid);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    lr_fma_chunk(as[b], bs[b], ty, tx, acc);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ca = ca0 + 4 * ty + i;
    if (ca >= a.Ca) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int cb = cb0 + 64 * h + 4 * tx;
      if (cb < a.Cb)
        *reinterpret_cast<float4*>(a.partial + (size_t)ca * a.Cb + cb) = make_float4(acc[i][4 * h], acc[i][4 * h + 1], acc[i][4 * h + 2], acc[i][4 * h + 3]);
    }
  }
}

// 1: launched, 0: not covered
int launch_linear_small(const GatherArgs& g, cudaStream_t st) {
  if (g.wtaps != 1 || g.nphase != 1 || g.is != 1 || g.os != 1 || g.Hs != g.Hd || g.Ws != g.Wd) return 0;
  if (g.phase[0].ntaps != 1 || g.phase[0].taps[0].dh != 0 || g.phase[0].taps[0].dw != 0) return 0;
  const long long M = (long long)g.N * g.Hs * g.Ws;
  if (M > kLsMaxM || (g.Cs & 3) || g.Cs < 4) return 0;
  if ((reinterpret_cast<uintptr_t>(g.src) & 15) != 0) return 0;
  if (g.in_affine && (((reinterpret_cast<uintptr_t>(g.in_scale) | reinterpret_cast<uintptr_t>(g.in_shift)) & 15) != 0 ||
                      (g.in_center && (reinterpret_cast<uintptr_t>(g.in_center) & 15) != 0))) return 0;
  if (launch_linear_rows(g, (int)M, st)) return 1;      // wide layers (decoder_input): register-blocked tile kernel below
  const int ncol = (g.Cd + 31) / 32, nk = (g.Cs + kLsKC - 1) / kLsKC;
  const int cpc = 1;     // K chunks per CTA: > 1 measured slower (a CTA's life is latency; serial chunks add to it)
  const dim3 grid(ncol, (nk + cpc - 1) / cpc);
  if (grid.y > 1 && cudaMemsetAsync(g.dst, 0, sizeof(float) * (size_t)M * g.Cd, st) != cudaSuccess) return CVAE_ERR_LAUNCH;
  if (M <= 16) linear_small_kernel<4><<<grid, 128, 0, st>>>(g, (int)M, cpc);
  else if (M <= 32) linear_small_kernel<8><<<grid, 128, 0, st>>>(g, (int)M, cpc);
  else if (M <= 64) linear_small_kernel<16><<<grid, 128, 0, st>>>(g, (int)M, cpc);
  else linear_small_kernel<32><<<dim3(grid.x, grid.y, (unsigned)((M + 127) / 128)), 128, 0, st>>>(g, (int)M, cpc);
  if (g.epi != CVAE_EPI_PLAIN) linear_small_epi_kernel<<<(g.Cd + 31) / 32, 128, 0, st>>>(g, (int)M);
  return 1;
}

// ---- weight gradient: P[ca][cb] = sum_m xa(ga[m][ca]) * xb(db[m][cb]),  m < M <= 512 (128 rows staged at a time) ----
// CTA tile 32 (ca) x 64 (cb); thread = 2 ca x 4 cb.
__global__ void __launch_bounds__(256) wgrad_small_kernel(const __grid_constant__ WgradArgs a, const int M) {
  extern __shared__ __align__(16) float ws_sm[];      // min(M, 128) x (32 + 64) floats: sized by the launch, so M = 64 leaves 9 CTAs per SM
  const int MC = min(M, 128);                         // rows staged at a time (M <= 128: one pass, as before)
  float* As = ws_sm;
  float* Bs = ws_sm + MC * 32;
  const int tid = threadIdx.x;
  const int ca0 = blockIdx.x * 32, cb0 = blockIdx.y * 64;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  for (int m0 = 0; m0 < M; m0 += MC) {
    const int mc = min(MC, M - m0);
    if (m0 > 0) __syncthreads();
    // a thread always stages the same channel of each operand (256 threads = 8 rows x 32 / 4 rows x 64 per pass): the
    // coefficients are loaded once and the rows go out in batches of 8 loads before the first shared-memory store
    // (the plain loop paid one dependent round trip per row pass: 15 us at 64 rows, 55 us at 256)
    {
      const int c = ca0 + (tid & 31), mrow = tid >> 5;
      const bool cok = c < a.Ca;
      float sc = 1.f, sh = 0.f, ce = 0.f;
      if (a.a_affine && cok) { sc = __ldg(a.a_scale + c); sh = __ldg(a.a_shift + c); ce = a.a_center ? __ldg(a.a_center + c) : 0.f; }
      for (int r0 = 0; r0 < mc; r0 += 64) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int r = r0 + mrow + 8 * u;
          v[u] = 0.f;
          if (cok && r < mc) v[u] = __ldg(a.ga + (size_t)(m0 + r) * a.Ca + c);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int r = r0 + mrow + 8 * u;
          if (r >= mc) continue;
          float x = v[u];
          if (cok) {
            if (a.a_affine) x = fmaf(x - ce, sc, sh);
            if (a.a_act) x = lrelu(x, a.a_slope);
          }
          As[r * 32 + (tid & 31)] = x;
        }
      }
    }
    {
      const int c = cb0 + (tid & 63), mrow = tid >> 6;
      const bool cok = c < a.Cb;
      float sc = 1.f, sh = 0.f, ce = 0.f;
      if (a.b_affine && cok) { sc = __ldg(a.b_scale + c); sh = __ldg(a.b_shift + c); ce = a.b_center ? __ldg(a.b_center + c) : 0.f; }
      for (int r0 = 0; r0 < mc; r0 += 32) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int r = r0 + mrow + 4 * u;
          v[u] = 0.f;
          if (cok && r < mc) v[u] = __ldg(a.db + (size_t)(m0 + r) * a.Cb + c);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int r = r0 + mrow + 4 * u;
          if (r >= mc) continue;
          float x = v[u];
          if (cok) {
            if (a.b_affine) x = fmaf(x - ce, sc, sh);
            if (a.b_act) x = lrelu(x, a.b_slope);
          }
          Bs[r * 64 + (tid & 63)] = x;
        }
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int m = 0; m < mc; ++m) {
      const float2 av = *reinterpret_cast<const float2*>(As + m * 32 + ty * 2);
      const float4 bv = *reinterpret_cast<const float4*>(Bs + m * 64 + tx * 4);
      acc[0][0] = fmaf(av.x, bv.x, acc[0][0]); acc[0][1] = fmaf(av.x, bv.y, acc[0][1]);
      acc[0][2] = fmaf(av.x, bv.z, acc[0][2]); acc[0][3] = fmaf(av.x, bv.w, acc[0][3]);
      acc[1][0] = fmaf(av.y, bv.x, acc[1][0]); acc[1][1] = fmaf(av.y, bv.y, acc[1][1]);
      acc[1][2] = fmaf(av.y, bv.z, acc[1][2]); acc[1][3] = fmaf(av.y, bv.w, acc[1][3]);
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int ca = ca0 + ty * 2 + i;
    if (ca >= a.Ca) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cb = cb0 + tx * 4 + j;
      if (cb < a.Cb) a.partial[(size_t)ca * a.Cb + cb] = acc[i][j];
    }
  }
}

// 1: launched, 0: not covered.  Requires splits == 1 (the caller's partial buffer is [rows][Cb]).
int launch_wgrad_small(const WgradArgs& a, int taps, int splits, cudaStream_t st) {
  if (taps != 1 || splits != 1 || a.K > kLsMaxM || a.K < 1) return 0;
  if ((long long)a.Ca * a.Cb >= (1ll << 21) && (a.Ca & 3) == 0 && (a.Cb & 3) == 0 && !a.b_affine && !a.b_act &&
      ((reinterpret_cast<uintptr_t>(a.ga) | reinterpret_cast<uintptr_t>(a.db) | reinterpret_cast<uintptr_t>(a.partial)) & 15) == 0) {
    wgrad_rows_kernel<<<dim3((a.Cb + 127) / 128, (a.Ca + 63) / 64), 256, 0, st>>>(a, a.K);
    return 1;
  }
  wgrad_small_kernel<<<dim3((a.Ca + 31) / 32, (a.Cb + 63) / 64), 256, (size_t)min(a.K, 128) * 96 * sizeof(float), st>>>(a, a.K);
  return 1;
}

}  // namespace cvae
