// Image-sized 16 -> 16 channel stride-2 layers as fp32 SIMT tile kernels (sm_100a).
//
// The last up-stage of the decoder (ConvTranspose2d 16 -> 16, 128^2 -> 256^2; vit_backbone.py:119-156)
// moves the largest tensors of the step (268 MB at B = 64) through the smallest GEMM: K = N = 16.  On
// the tcgen05 path a 128 x 16 x 8 MMA costs the same ~96 clk as a 128 x 128 x 8 one (operand fetch,
// scripts/umma_rate.cu) and three are needed per product (3xTF32), so the tensor cores deliver less than
// the fp32 FMA pipes here, and the warp-specialised kernel was bound by its 4 epilogue warps (forward)
// or its 8 producer warps (input gradient): 293 us / 387 us per launch against an HBM time of ~50 us.
//
// Here every thread of a full-occupancy CTA does FMAs: the input tile (+ halo) is staged once in shared
// memory channel-planar (BatchNorm + LeakyReLU applied while staging), a thread owns 4 output channels of
// 4 (forward: x 4 output phases) or 8 positions, x operands are conflict-free scalar shared loads,
// weights are 128-bit shared loads shared by every thread with the same channel quarter, and the
// per-channel BatchNorm sums stay in registers across the persistent tile loop.
#include "conv_args.cuh"

namespace cvae {

constexpr int kFewThreads = 256;

struct FewEpi {    // per-thread statistics of 4 channels (fp32 within a tile, fp64 across tiles)
  float f1[4], f2[4];
  double d1[4], d2[4];
};
struct FewEpiC {   // epilogue constants of 4 channels, loaded per tile so they are not live in the FMA loop
  float4 bias, esc, esh, ece;
};

__device__ __forceinline__ void few_epi_init(FewEpi& e) {
#pragma unroll
  for (int j = 0; j < 4; ++j) { e.f1[j] = 0.f; e.f2[j] = 0.f; e.d1[j] = 0.0; e.d2[j] = 0.0; }
}
__device__ __forceinline__ FewEpiC few_epi_consts(const GatherArgs& a, int c0) {
  FewEpiC k;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  k.bias = zero; k.esc = make_float4(1.f, 1.f, 1.f, 1.f); k.esh = zero; k.ece = zero;
  if (a.bias) k.bias = __ldg(reinterpret_cast<const float4*>(a.bias + c0));
  if (a.e_affine) {
    k.esc = __ldg(reinterpret_cast<const float4*>(a.e_scale + c0));
    k.esh = __ldg(reinterpret_cast<const float4*>(a.e_shift + c0));
    if (a.e_center) k.ece = __ldg(reinterpret_cast<const float4*>(a.e_center + c0));
  }
  return k;
}

// one output vector: bias, statistics / activation-derivative epilogue, 128-bit store
__device__ __forceinline__ void few_epi_store(FewEpi& e, const FewEpiC& k, const GatherArgs& a, const float (&acc)[4],
                                              size_t off, const float4& ref, const float4& add) {
  float o[4] = {acc[0] + k.bias.x, acc[1] + k.bias.y, acc[2] + k.bias.z, acc[3] + k.bias.w};
  if (a.epi == CVAE_EPI_STATS) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { e.f1[j] += o[j]; e.f2[j] = fmaf(o[j], o[j], e.f2[j]); }
  } else if (a.epi == CVAE_EPI_DACT) {
    const float rf[4] = {ref.x - k.ece.x, ref.y - k.ece.y, ref.z - k.ece.z, ref.w - k.ece.w};
    const float sc[4] = {k.esc.x, k.esc.y, k.esc.z, k.esc.w}, sh[4] = {k.esh.x, k.esh.y, k.esh.z, k.esh.w};
    o[0] += add.x; o[1] += add.y; o[2] += add.z; o[3] += add.w;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float z = fmaf(rf[j], sc[j], sh[j]);
      o[j] = z > 0.f ? o[j] : o[j] * a.e_slope;
      e.f1[j] += o[j]; e.f2[j] = fmaf(o[j], rf[j], e.f2[j]);
    }
  }
  *reinterpret_cast<float4*>(a.dst + off) = make_float4(o[0], o[1], o[2], o[3]);
}

__device__ __forceinline__ void few_epi_fold(FewEpi& e) {   // fp32 tile sums -> fp64 running sums
#pragma unroll
  for (int j = 0; j < 4; ++j) { e.d1[j] += (double)e.f1[j]; e.d2[j] += (double)e.f2[j]; e.f1[j] = 0.f; e.f2[j] = 0.f; }
}

// block reduction of the per-thread channel sums (threads with the same tid & 3 own the same 4 channels)
__device__ __forceinline__ void few_epi_flush(FewEpi& e, const GatherArgs& a, double* s_red) {
  if (a.epi == CVAE_EPI_PLAIN || a.stats == nullptr) return;
  few_epi_fold(e);
  const int tid = threadIdx.x;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) { s_red[tid * 8 + j] = e.d1[j]; s_red[tid * 8 + 4 + j] = e.d2[j]; }
  __syncthreads();
  if (tid < 32) {
    const int q = tid >> 3, k = tid & 7;
    double t = 0.0;
    for (int r = q; r < kFewThreads; r += 4) t += s_red[r * 8 + k];
    atomicAdd(a.stats + (k < 4 ? 0 : 16) + q * 4 + (k & 3), t);
  }
}

// ============================== ConvTranspose2d 16 -> 16, k3 s2 p1 op1: forward ==============================
// q-space tile: 8 x 32 input positions -> 16 x 64 output pixels.  thread = (row, 4 adjacent columns,
// channel quarter): 4 phases x 4 positions x 4 channels = 64 accumulators.
constexpr int kUpTH = 8, kUpTW = 32, kUpGC = kUpTW + 1;
constexpr int kUpPix = (kUpTH + 1) * kUpGC;          // 297 staged positions
constexpr int kUpPlane = 298;                        // 4 * plane % 32 == 8: the staging stores of a warp hit 32 banks

struct UpTaps { int w[4][4]; };                      // [phase = ph*2 + pw][dh*2 + dw] -> weight index (unused: 0)

__global__ void __launch_bounds__(kFewThreads, 2)
convt16_up_kernel(const __grid_constant__ GatherArgs a, const __grid_constant__ UpTaps T, const int tiles_h,
                  const int tiles_w, const int total) {
  __shared__ __align__(16) float sX[16 * kUpPlane];
  __shared__ __align__(16) float4 sW[9 * 16 * 4];    // [widx][ci][channel quarter]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 9 * 16 * 4; i += kFewThreads) sW[i] = __ldg(reinterpret_cast<const float4*>(a.wt) + i);
  // staging role: fixed channel quarter (tid & 3)
  const int s4 = tid & 3;
  // compute role
  const int g = lane >> 2, cq = lane & 3;
  const int p0 = warp * kUpGC + 4 * g;
  FewEpi E;
  few_epi_init(E);
  const float4* wq = sW + cq;
  // weight-tile offsets straight from the parameter bank (uniform registers, not per-thread ones)
#define o00 (T.w[0][0] * 64)
#define o01a (T.w[1][0] * 64)
#define o01b (T.w[1][1] * 64)
#define o10a (T.w[2][0] * 64)
#define o10b (T.w[2][2] * 64)
#define o11a (T.w[3][0] * 64)
#define o11b (T.w[3][1] * 64)
#define o11c (T.w[3][2] * 64)
#define o11d (T.w[3][3] * 64)

  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int tw = tile % tiles_w, tt = tile / tiles_w, th = tt % tiles_h, n = tt / tiles_h;
    const int h0 = th * kUpTH, w0 = tw * kUpTW;
    __syncthreads();                                 // previous tile's readers are done (also orders the sW fill)
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f), ce = sh;
    if (a.in_affine) {                               // (re)loaded per tile: not live across the FMA loop
      sc = __ldg(reinterpret_cast<const float4*>(a.in_scale + s4 * 4));
      sh = __ldg(reinterpret_cast<const float4*>(a.in_shift + s4 * 4));
      if (a.in_center) ce = __ldg(reinterpret_cast<const float4*>(a.in_center + s4 * 4));
    }
    {   // 297 x 4 vectors = at most 5 per thread: all loads issued before the first store
      float4 v[5];
      int pixs[5];
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        const int idx = tid + u * kFewThreads;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        pixs[u] = -1;
        if (idx < kUpPix * 4) {
          const int pix = idx >> 2, gi = pix / kUpGC, gj = pix - gi * kUpGC, ih = h0 + gi, iw = w0 + gj;
          pixs[u] = pix;
          if (ih < a.Hs && iw < a.Ws) {
            v[u] = __ldg(reinterpret_cast<const float4*>(a.src + (((size_t)n * a.Hs + ih) * a.Ws + iw) * 16 + s4 * 4));
            pixs[u] |= 1 << 30;                       // valid pixel: the transform applies (padding stays exactly 0)
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        if (pixs[u] < 0) continue;
        float4 x = v[u];
        if (pixs[u] & (1 << 30)) {
          if (a.in_affine) {
            x.x = fmaf(x.x - ce.x, sc.x, sh.x); x.y = fmaf(x.y - ce.y, sc.y, sh.y);
            x.z = fmaf(x.z - ce.z, sc.z, sh.z); x.w = fmaf(x.w - ce.w, sc.w, sh.w);
          }
          if (a.in_act) { x.x = lrelu(x.x, a.in_slope); x.y = lrelu(x.y, a.in_slope); x.z = lrelu(x.z, a.in_slope); x.w = lrelu(x.w, a.in_slope); }
        }
        float* d = sX + (s4 * 4) * kUpPlane + (pixs[u] & ~(1 << 30));
        d[0] = x.x; d[kUpPlane] = x.y; d[2 * kUpPlane] = x.z; d[3 * kUpPlane] = x.w;
      }
    }
    __syncthreads();
    float acc[4][4][4];
#pragma unroll
    for (int f = 0; f < 4; ++f)
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[f][p][c] = 0.f;
#pragma unroll 2
    for (int ci = 0; ci < 16; ++ci) {
      const float* xp = sX + ci * kUpPlane + p0;
      float x0[5], x1[5];
#pragma unroll
      for (int t = 0; t < 5; ++t) { x0[t] = xp[t]; x1[t] = xp[kUpGC + t]; }
      const float4* wp = wq + ci * 4;
      float4 w;
      w = wp[o00];
#pragma unroll
      for (int p = 0; p < 4; ++p) fma4(acc[0][p], x0[p], w);
      w = wp[o01a];
#pragma unroll
      for (int p = 0; p < 4; ++p) fma4(acc[1][p], x0[p], w);
      w = wp[o01b];
#pragma unroll
      for (int p = 0; p < 4; ++p) fma4(acc[1][p], x0[p + 1], w);
      w = wp[o10a];
#pragma unroll
      for (int p = 0; p < 4; ++p) fma4(acc[2][p], x0[p], w);
      w = wp[o10b];
#pragma unroll
      for (int p = 0; p < 4; ++p) fma4(acc[2][p], x1[p], w);
      w = wp[o11a];
#pragma unroll
      for (int p = 0; p < 4; ++p) fma4(acc[3][p], x0[p], w);
      w = wp[o11b];
#pragma unroll
      for (int p = 0; p < 4; ++p) fma4(acc[3][p], x0[p + 1], w);
      w = wp[o11c];
#pragma unroll
      for (int p = 0; p < 4; ++p) fma4(acc[3][p], x1[p], w);
      w = wp[o11d];
#pragma unroll
      for (int p = 0; p < 4; ++p) fma4(acc[3][p], x1[p + 1], w);
    }
    const int qh = h0 + warp;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const FewEpiC K = few_epi_consts(a, cq * 4);
    if (qh < a.Hs) {
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int qw = w0 + 4 * g + p;
        if (qw >= a.Ws) continue;
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          const int oh = 2 * qh + (f >> 1), ow = 2 * qw + (f & 1);
          const size_t off = (((size_t)n * a.Hd + oh) * a.Wd + ow) * 16 + cq * 4;
          few_epi_store(E, K, a, acc[f][p], off, zero, zero);
        }
      }
    }
    if (a.epi != CVAE_EPI_PLAIN) few_epi_fold(E);
  }
  few_epi_flush(E, a, reinterpret_cast<double*>(sX));
#undef o00
#undef o01a
#undef o01b
#undef o10a
#undef o10b
#undef o11a
#undef o11b
#undef o11c
#undef o11d
}

// ============================== Conv2d 16 -> 16, k3 s2 p1 (gather): forward of a stride-2 conv, =================
// ============================== input gradient of the transposed conv above ====================================
// q-space tile: 16 x 32 outputs from a 33 x 65 input window, staged 8 channels at a time (two passes).
// thread = (2 rows, 4 adjacent columns, channel quarter): 32 accumulators.
constexpr int kDnTH = 16, kDnTW = 32, kDnGR = 2 * kDnTH + 1, kDnGC = 2 * kDnTW + 1;
constexpr int kDnPitch = 73;                         // 65 columns + 4 words of skew per 32 columns (see dn_col)
constexpr int kDnPlane = kDnGR * kDnPitch;           // 2409
constexpr int kDnSmem = (8 * kDnPlane + 9 * 16 * 16) * 4;

// Column b of the staged window lives at b + 4 * (b / 32): threads of a warp read columns 8 g + t
// (g = 0..7), a stride of 8 words that would put g and g + 4 on the same bank.
__device__ __forceinline__ int dn_col(int b) { return b + ((b >> 5) << 2); }

struct DnTaps { int w[3][3]; };                      // [kh][kw] -> weight index

__global__ void __launch_bounds__(kFewThreads, 2)
conv16_dn_kernel(const __grid_constant__ GatherArgs a, const __grid_constant__ DnTaps T, const int tiles_h,
                 const int tiles_w, const int total) {
  extern __shared__ __align__(16) float dsm[];
  float* sX = dsm;                                                     // [8][33][73]
  float4* sW = reinterpret_cast<float4*>(dsm + 8 * kDnPlane);          // [widx][c_in][channel quarter]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 9 * 16 * 4; i += kFewThreads) sW[i] = __ldg(reinterpret_cast<const float4*>(a.wt) + i);
  const int g = lane >> 2, cq = lane & 3;
  // columns 8 g + t, t = 0..7, never cross a 32-column skew boundary inside a thread; t = 8 may
  const int cb = dn_col(8 * g), c8 = dn_col(8 * g + 8);
  FewEpi E;
  few_epi_init(E);

  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int tw = tile % tiles_w, tt = tile / tiles_w, th = tt % tiles_h, n = tt / tiles_h;
    const int h0 = th * kDnTH, w0 = tw * kDnTW;
    float acc[2][4][4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][p][c] = 0.f;
    for (int hc = 0; hc < 2; ++hc) {
      __syncthreads();                               // previous pass's readers are done (also orders the sW fill)
      // Batches of 4 pixels per thread with all 8 loads issued before the first store: the one-pixel-at-a-time
      // loop paid a full global round trip per iteration (26 % of the stall samples in profiles/r1_ncu_few_raw.txt).
      for (int idx0 = tid; idx0 < kDnGR * kDnGC; idx0 += 4 * kFewThreads) {
        float4 v0[4], v1[4];
        int dofs[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int idx = idx0 + u * kFewThreads;
          v0[u] = make_float4(0.f, 0.f, 0.f, 0.f); v1[u] = v0[u];
          dofs[u] = -1;
          if (idx < kDnGR * kDnGC) {
            const int gi = idx / kDnGC, gj = idx - gi * kDnGC, ih = 2 * h0 - 1 + gi, iw = 2 * w0 - 1 + gj;
            dofs[u] = gi * kDnPitch + dn_col(gj);
            if ((unsigned)ih < (unsigned)a.Hs && (unsigned)iw < (unsigned)a.Ws) {
              const float4* sp = reinterpret_cast<const float4*>(a.src + (((size_t)n * a.Hs + ih) * a.Ws + iw) * 16 + hc * 8);
              v0[u] = __ldg(sp); v1[u] = __ldg(sp + 1);
              dofs[u] |= 1 << 30;                     // valid pixel: the transform applies (padding stays exactly 0)
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (dofs[u] < 0) continue;
          float v[8] = {v0[u].x, v0[u].y, v0[u].z, v0[u].w, v1[u].x, v1[u].y, v1[u].z, v1[u].w};
          if (dofs[u] & (1 << 30)) {
            if (a.in_affine) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const int c = hc * 8 + k;
                const float cen = a.in_center ? __ldg(a.in_center + c) : 0.f;
                v[k] = fmaf(v[k] - cen, __ldg(a.in_scale + c), __ldg(a.in_shift + c));
              }
            }
            if (a.in_act) {
#pragma unroll
              for (int k = 0; k < 8; ++k) v[k] = lrelu(v[k], a.in_slope);
            }
          }
          float* d = sX + (dofs[u] & ~(1 << 30));
#pragma unroll
          for (int k = 0; k < 8; ++k) d[k * kDnPlane] = v[k];
        }
      }
      __syncthreads();
#pragma unroll 1
      for (int co = 0; co < 8; ++co) {
        const float* xp = sX + co * kDnPlane + (4 * warp) * kDnPitch;
        const float4* wp = sW + (hc * 8 + co) * 4 + cq;
        float x[5][9];
#pragma unroll
        for (int r = 0; r < 5; ++r) {
#pragma unroll
          for (int t = 0; t < 8; ++t) x[r][t] = xp[r * kDnPitch + cb + t];
          x[r][8] = xp[r * kDnPitch + c8];
        }
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const float4 w = wp[T.w[kh][kw] * 64];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
              for (int p = 0; p < 4; ++p) fma4(acc[r][p], x[2 * r + kh][2 * p + kw], w);
          }
      }
    }
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const FewEpiC K = few_epi_consts(a, cq * 4);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int qh = h0 + 2 * warp + r;
      if (qh >= a.Hd) continue;
      float4 ref[4], add[4];
      size_t off[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) {                  // the four reference loads of a row overlap
        const int qw = w0 + 4 * g + p;
        off[p] = (((size_t)n * a.Hd + qh) * a.Wd + qw) * 16 + cq * 4;
        ref[p] = zero; add[p] = zero;
        if (a.epi == CVAE_EPI_DACT && qw < a.Wd) {
          ref[p] = __ldg(reinterpret_cast<const float4*>(a.epi_ref + off[p]));
          if (a.epi_add) add[p] = __ldg(reinterpret_cast<const float4*>(a.epi_add + off[p]));
        }
      }
#pragma unroll
      for (int p = 0; p < 4; ++p)
        if (w0 + 4 * g + p < a.Wd) few_epi_store(E, K, a, acc[r][p], off[p], ref[p], add[p]);
    }
    if (a.epi != CVAE_EPI_PLAIN) few_epi_fold(E);
  }
  few_epi_flush(E, a, reinterpret_cast<double*>(dsm));
}

// ============================== Conv2d 16 -> 1, k3 s1 p1 (image head forward), plain epilogue ==============================
// Output tile 32 x 32; the 34 x 34 x 16 input window is staged channel-planar (BatchNorm + LeakyReLU applied while
// staging).  thread = 4 vertically adjacent outputs of one column: the 32 lanes of a warp read 32 consecutive words
// of a plane (conflict-free, one wavefront per load), a 6 x 3 window per channel feeds 36 FMAs, and the 9 weights of
// the channel come as three 128-bit broadcasts.  The previous kernel (skinny.cu conv_cd1_tile) read the window as
// 128-bit vectors, four wavefronts each: 168 shared-memory wavefronts per output against 84 here.
constexpr int kHdT = 32, kHdG = kHdT + 2, kHdPix = kHdG * kHdG;     // 1156 staged pixels
constexpr int kHdPlane = 1162;                                       // 4 * plane % 32 == 8 (conflict-free staging stores)
constexpr int kHdSmem = (16 * kHdPlane + 16 * 12) * 4;

__global__ void __launch_bounds__(kFewThreads, 3)
conv16_head_fwd_kernel(const __grid_constant__ GatherArgs a, const int tiles_h, const int tiles_w, const int total) {
  extern __shared__ __align__(16) float dsm[];
  float* sX = dsm;                                                   // [16][34 x 34]
  float4* sW = reinterpret_cast<float4*>(dsm + 16 * kHdPlane);       // [ci][3]: taps (kh, kw) row-major, padded to 12
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const PhaseGeom& P = a.phase[0];
  if (tid < 16 * 12) {
    const int ci = tid / 12, t = tid % 12;
    float w = 0.f;
    if (t < 9)
      for (int u = 0; u < 9; ++u)
        if ((P.taps[u].dh + 1) * 3 + P.taps[u].dw + 1 == t) w = __ldg(a.wt + P.taps[u].widx * 16 + ci);
    reinterpret_cast<float*>(sW)[tid] = w;
  }
  const int s4 = tid & 3;
  const float bias = a.bias ? __ldg(a.bias) : 0.f;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int tw = tile % tiles_w, tt = tile / tiles_w, th = tt % tiles_h, n = tt / tiles_h;
    const int h0 = th * kHdT, w0 = tw * kHdT;
    __syncthreads();
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f), ce = sh;
    if (a.in_affine) {
      sc = __ldg(reinterpret_cast<const float4*>(a.in_scale + s4 * 4));
      sh = __ldg(reinterpret_cast<const float4*>(a.in_shift + s4 * 4));
      if (a.in_center) ce = __ldg(reinterpret_cast<const float4*>(a.in_center + s4 * 4));
    }
    // Staging by ROWS: a window row is 34 pixels x 16 channels = 2176 contiguous bytes (NHWC), so a warp takes whole rows
    // (warp, warp + 8, ...) and a lane the vectors lane + 32 k of the row: the channel quarter (lane & 3) is a per-thread
    // constant, the addresses are row pointer + immediate, and the five loads of a row are issued before its stores (the
    // flat index form spent ~50 instructions per vector on divisions by 34, bounds and 64-bit address arithmetic).
    for (int gi = warp; gi < kHdG; gi += kFewThreads / 32) {
      const int ih = h0 - 1 + gi;
      const bool rok = (unsigned)ih < (unsigned)a.Hs;
      const float* rowp = a.src + (((size_t)n * a.Hs + (rok ? ih : 0)) * a.Ws + (w0 - 1)) * 16 + lane * 4;   // dereferenced only where valid
      float4 v[5];
      bool ok[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const int gj = (lane >> 2) + 8 * k, iw = w0 - 1 + gj;
        ok[k] = rok && gj < kHdG && (unsigned)iw < (unsigned)a.Ws;
        v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok[k]) v[k] = __ldg(reinterpret_cast<const float4*>(rowp + k * 128));
      }
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const int gj = (lane >> 2) + 8 * k;
        if (gj >= kHdG) continue;
        float4 x = v[k];
        if (ok[k]) {
          if (a.in_affine) {
            x.x = fmaf(x.x - ce.x, sc.x, sh.x); x.y = fmaf(x.y - ce.y, sc.y, sh.y);
            x.z = fmaf(x.z - ce.z, sc.z, sh.z); x.w = fmaf(x.w - ce.w, sc.w, sh.w);
          }
          if (a.in_act) { x.x = lrelu(x.x, a.in_slope); x.y = lrelu(x.y, a.in_slope); x.z = lrelu(x.z, a.in_slope); x.w = lrelu(x.w, a.in_slope); }
        }
        float* d = sX + (s4 * 4) * kHdPlane + gi * kHdG + gj;
        d[0] = x.x; d[kHdPlane] = x.y; d[2 * kHdPlane] = x.z; d[3 * kHdPlane] = x.w;
      }
    }
    __syncthreads();
    float acc[4] = {bias, bias, bias, bias};
    const float* xp0 = sX + (4 * warp) * kHdG + lane;                // window rows 4 warp .. 4 warp + 5, columns lane .. lane + 2
#pragma unroll 4
    for (int ci = 0; ci < 16; ++ci) {
      const float* xp = xp0 + ci * kHdPlane;
      const float4 wa = sW[ci * 3], wb = sW[ci * 3 + 1], wc = sW[ci * 3 + 2];
      const float w[9] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, wc.x};
      float x[6][3];
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) x[r][c] = xp[r * kHdG + c];
#pragma unroll
      for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) acc[o] = fmaf(x[o + kh][kw], w[kh * 3 + kw], acc[o]);
    }
    const int qw = w0 + lane;
    if (qw < a.Wd) {
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        const int qh = h0 + 4 * warp + o;
        if (qh < a.Hd) a.dst[((size_t)n * a.Hd + qh) * a.Wd + qw] = acc[o];
      }
    }
  }
}

// 1: launched, 0: not covered
int launch_conv16_head_fwd(const GatherArgs& g, cudaStream_t st) {
  if (g.Cs != 16 || g.Cd != 1 || g.nphase != 1 || g.is != 1 || g.os != 1 || g.phase[0].ntaps != 9 || g.epi != CVAE_EPI_PLAIN) return 0;
  if (g.Hs != g.Hd || g.Ws != g.Wd || (long long)g.N * g.Hd * g.Wd < 65536) return 0;
  for (int t = 0; t < 9; ++t) {
    const TapEntry& e = g.phase[0].taps[t];
    if (e.dh < -1 || e.dh > 1 || e.dw < -1 || e.dw > 1) return 0;
  }
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(conv16_head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHdSmem) != cudaSuccess) return CVAE_ERR_LAUNCH;
    attr_set = true;
  }
  const int tiles_h = (g.Hd + kHdT - 1) / kHdT, tiles_w = (g.Wd + kHdT - 1) / kHdT;
  const long long total = (long long)g.N * tiles_h * tiles_w;
  if (total >= (1ll << 31)) return 0;
  conv16_head_fwd_kernel<<<(int)min(total, (long long)kNumSMs * 3), kFewThreads, kHdSmem, st>>>(g, tiles_h, tiles_w, (int)total);
  return 1;
}

// ---- host side ------------------------------------------------------------------------------------------
static bool few_shape(int Cs, int Cd, int kh, int kw, int stride, int pad, int mode, int N, int Hs, int Ws, int Hd,
                      int Wd, int epi) {
  if (Cs != 16 || Cd != 16 || kh != 3 || kw != 3 || stride != 2 || pad != 1) return false;
  if (mode == CVAE_CONV_SCATTER) {
    if (Hd != 2 * Hs || Wd != 2 * Ws || epi == CVAE_EPI_DACT) return false;
    return (long long)N * Hs * Ws >= 65536;
  }
  if (mode == CVAE_CONV_GATHER) {
    if (Hs != 2 * Hd || Ws != 2 * Wd) return false;
    return (long long)N * Hd * Wd >= 65536;
  }
  return false;
}

// 1: launched; 0: shape not covered (caller continues with its generic kernels); < 0: error
int launch_conv_few(const cvae_conv_params_t* p, const GatherArgs& g, cudaStream_t st) {
  if (!few_shape(p->Cs, p->Cd, p->kh, p->kw, p->stride, p->pad, p->mode, p->N, p->Hs, p->Ws, p->Hd, p->Wd, p->epi)) return 0;
  if (p->mode == CVAE_CONV_SCATTER) {
    if (g.nphase != 4) return 0;
    UpTaps T;
    for (int f = 0; f < 4; ++f) for (int k = 0; k < 4; ++k) T.w[f][k] = -1;
    for (int i = 0; i < 4; ++i) {
      const PhaseGeom& P = g.phase[i];
      if (P.ph < 0 || P.ph > 1 || P.pw < 0 || P.pw > 1 || P.ntaps != (P.ph + 1) * (P.pw + 1)) return 0;
      for (int t = 0; t < P.ntaps; ++t) {
        const TapEntry& e = P.taps[t];
        if (e.dh < 0 || e.dh > P.ph || e.dw < 0 || e.dw > P.pw) return 0;
        T.w[P.ph * 2 + P.pw][e.dh * 2 + e.dw] = e.widx;
      }
    }
    // every tap the kernel reads must have been assigned
    const int need[4][4] = {{1, 0, 0, 0}, {1, 1, 0, 0}, {1, 0, 1, 0}, {1, 1, 1, 1}};
    for (int f = 0; f < 4; ++f)
      for (int k = 0; k < 4; ++k) {
        if (need[f][k] && T.w[f][k] < 0) return 0;
        if (!need[f][k]) T.w[f][k] = 0;
      }
    const int tiles_h = (p->Hs + kUpTH - 1) / kUpTH, tiles_w = (p->Ws + kUpTW - 1) / kUpTW;
    const long long total = (long long)p->N * tiles_h * tiles_w;
    if (total >= (1ll << 31)) return 0;
    convt16_up_kernel<<<(int)min(total, (long long)kNumSMs * 2), kFewThreads, 0, st>>>(g, T, tiles_h, tiles_w, (int)total);
    return 1;
  }
  if (g.nphase != 1 || g.phase[0].ntaps != 9) return 0;
  DnTaps T;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) T.w[i][j] = -1;
  for (int t = 0; t < 9; ++t) {
    const TapEntry& e = g.phase[0].taps[t];
    if (e.dh < -1 || e.dh > 1 || e.dw < -1 || e.dw > 1) return 0;
    T.w[e.dh + 1][e.dw + 1] = e.widx;
  }
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) if (T.w[i][j] < 0) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(conv16_dn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDnSmem) != cudaSuccess) return CVAE_ERR_LAUNCH;
    attr_set = true;
  }
  const int tiles_h = (p->Hd + kDnTH - 1) / kDnTH, tiles_w = (p->Wd + kDnTW - 1) / kDnTW;
  const long long total = (long long)p->N * tiles_h * tiles_w;
  if (total >= (1ll << 31)) return 0;
  conv16_dn_kernel<<<(int)min(total, (long long)kNumSMs * 2), kFewThreads, kDnSmem, st>>>(g, T, tiles_h, tiles_w, (int)total);
  return 1;
}

}  // namespace cvae

// 1 when cvae_conv_gather runs this layer on the few-channel SIMT kernels (the caller then packs the weights in
// the fp32 [tap][Cin][Cout] layout and does not take the tensor-core entry point).
extern "C" int cvae_conv_few_eligible(int Cs, int Cd, int k, int stride, int pad, int mode, int N, int Hs, int Ws, int Hd,
                                      int Wd, int epi) {
  return cvae::few_shape(Cs, Cd, k, k, stride, pad, mode, N, Hs, Ws, Hd, Wd, epi) ? 1 : 0;
}
