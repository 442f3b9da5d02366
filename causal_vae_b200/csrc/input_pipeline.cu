// Vessel input pipeline on the device: what the reference's dataset does to every raw image on
// DataLoader worker CPUs (vessel_analysis/00_core/dataset.py:186,216-249) — antialiased bilinear resize, the
// deterministic flip by idx % 4, per-image min-max, threshold at the image mean -> {0,1} mask — as three launches
// over a whole batch, bit-for-bit the arithmetic of the CPU path:
//
//   * the resize follows ATen's separable CPU kernel (aten/src/ATen/native/cpu/UpSampleKernel.cpp): width pass
//     first, every intermediate rounded to fp32, taps accumulated sequentially, and — because the compiled loop
//     rounds the product in groups of four taps and fuses only the remainder (established against the live kernel,
//     tests/golden/make_input_golden.py --probe) —
//     with explicit __fmul_rn / __fadd_rn / __fmaf_rn so nvcc cannot contract differently;
//   * min / max are order-free (atomics on an order-preserving integer image of the float);
//   * the mean is the correctly rounded one: fp64 sum of the fp32 normalised values, rounded once.
//
// HBM-shaped: the raw batch is read once (66 MB at B = 64, 512^2), the resized image (L2-sized per chunk) is written
// once and re-read twice, the mask is written once.  One CTA = a 64-column x TH-row output tile: the horizontally
// resized rows it needs live in shared memory, so the intermediate [Hin, W] image of the CPU path never exists.
#include <cstdlib>
#include "common.cuh"

namespace cvae {

struct PreStats {          // per image, zero-initialised by the entry point
  unsigned int max_enc;    // atomicMax of enc(v)
  unsigned int min_inv;    // atomicMax of ~enc(v)
  double sum;              // fp64 sum of the normalised image
};

// order-preserving map float -> uint32 (and back)
__device__ __forceinline__ unsigned int f2ord(float f) {
  const unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int e) {
  return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e);
}

// One axis of ATen's antialiased bilinear resize for float input
// (HelperInterpBase::_compute_indices_min_size_weights_aa): scale, support, center, invscale, each weight and the
// running total are float; i + 0.5, (center -/+ support) + 0.5 and the filter argument are formed in double.
__global__ void aa_weights_kernel(int in_size, int out_size, int max_interp, int* __restrict__ xmin,
                                  int* __restrict__ xsize, float* __restrict__ w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= out_size) return;
  float* wi = w + (size_t)i * max_interp;
  if (in_size == out_size) {           // ATen skips an axis whose size does not change
    xmin[i] = i;
    xsize[i] = 1;
    wi[0] = 1.0f;
    for (int j = 1; j < max_interp; ++j) wi[j] = 0.f;
    return;
  }
  const float scale = __fdiv_rn((float)in_size, (float)out_size);
  const float support = scale >= 1.0f ? scale : 1.0f;
  const float invscale = scale >= 1.0f ? (float)__ddiv_rn(1.0, (double)scale) : 1.0f;
  const float center = (float)__dmul_rn((double)scale, __dadd_rn((double)i, 0.5));
  long long lo = (long long)__dadd_rn((double)__fsub_rn(center, support), 0.5);
  if (lo < 0) lo = 0;
  long long hi = (long long)__dadd_rn((double)__fadd_rn(center, support), 0.5);
  if (hi > in_size) hi = in_size;
  long long n = hi - lo;
  n = n < 0 ? 0 : (n > max_interp ? max_interp : n);
  float total = 0.f;
  for (int j = 0; j < (int)n; ++j) {
    const float d = __fsub_rn((float)(j + lo), center);
    const float arg = (float)__dmul_rn(__dadd_rn((double)d, 0.5), (double)invscale);
    const float a = fabsf(arg);
    const float wj = a < 1.0f ? __fsub_rn(1.0f, a) : 0.f;
    wi[j] = wj;
    total = __fadd_rn(total, wj);
  }
  if (total != 0.f)
    for (int j = 0; j < (int)n; ++j) wi[j] = __fdiv_rn(wi[j], total);
  for (int j = (int)n; j < max_interp; ++j) wi[j] = 0.f;
  xmin[i] = (int)lo;
  xsize[i] = (int)n;
}

// t = s0*w0; t += sj*wj sequentially; the first 4*floor((n-1)/4) steps round the product, the rest are fused.
template <typename LoadF>
__device__ __forceinline__ float aa_taps(LoadF src, const float* __restrict__ w, int n) {
  float t = __fmul_rn(src(0), w[0]);
  const int unfused = 1 + ((n - 1) & ~3);
  int j = 1;
  for (; j < unfused; ++j) t = __fadd_rn(t, __fmul_rn(src(j), w[j]));
  for (; j < n; ++j) t = __fmaf_rn(src(j), w[j], t);
  return t;
}
// the same with a compile-time tap count: the runtime loop above costs ~12 instructions per tap (ncu: the kernel
// issue-bound at 81 %), the unrolled form 2-3 (immediate-offset shared-memory load + FMUL/FADD or FFMA).
template <int N, typename LoadF>
__device__ __forceinline__ float aa_taps_n(LoadF src, const float* w) {
  constexpr int unfused = 1 + ((N - 1) & ~3);
  float t = __fmul_rn(src(0), w[0]);
#pragma unroll
  for (int j = 1; j < N; ++j) t = j < unfused ? __fadd_rn(t, __fmul_rn(src(j), w[j])) : __fmaf_rn(src(j), w[j], t);
  return t;
}
constexpr int kMaxUnrolledTaps = 12;
#define CVAE_TAP_SWITCH(n, CALL, FALLBACK)                                                              \
  switch (n) {                                                                                          \
    case 1: CALL(1); break;  case 2: CALL(2); break;  case 3: CALL(3); break;  case 4: CALL(4); break;   \
    case 5: CALL(5); break;  case 6: CALL(6); break;  case 7: CALL(7); break;  case 8: CALL(8); break;   \
    case 9: CALL(9); break;  case 10: CALL(10); break; case 11: CALL(11); break; case 12: CALL(12); break; \
    default: FALLBACK; break;                                                                           \
  }

constexpr int kTileW = 64;
constexpr int kPreThreads = 256;

// width pass of one thread: its column's N weights live in registers across the rows of the chunk
template <int N>
__device__ __forceinline__ void width_rows(const float* __restrict__ src, int pitch, float* __restrict__ dst, int rr0,
                                           int rc, const float* __restrict__ wc) {
  float w[N];
#pragma unroll
  for (int j = 0; j < N; ++j) w[j] = wc[j];
  // three rows per trip: their shared-memory loads are independent, and the loop / address arithmetic is amortised
  constexpr int kStep = kPreThreads / kTileW;
  int rr = rr0;
  for (; rr + 2 * kStep < rc; rr += 3 * kStep) {
    const float* r0 = src + rr * pitch;
    const float* r1 = r0 + kStep * pitch;
    const float* r2 = r1 + kStep * pitch;
    const float a = aa_taps_n<N>([&](int j) { return r0[j]; }, w);
    const float b = aa_taps_n<N>([&](int j) { return r1[j]; }, w);
    const float c = aa_taps_n<N>([&](int j) { return r2[j]; }, w);
    dst[rr * kTileW] = a;
    dst[(rr + kStep) * kTileW] = b;
    dst[(rr + 2 * kStep) * kTileW] = c;
  }
  for (; rr < rc; rr += kStep) {
    const float* row = src + rr * pitch;
    dst[rr * kTileW] = aa_taps_n<N>([&](int j) { return row[j]; }, w);
  }
}

// grid (ceil(W/64), ceil(H/TH), B).  Dynamic smem: rows_max*64 floats (horizontally resized rows) +
// 64*mx floats (width-axis weights of the tile) + TH*my floats (height-axis weights) + stage_floats (raw rows).
// The raw rows a tile needs are staged through shared memory in chunks with coalesced 128-bit loads, four per
// thread in flight (the direct strided gather kept one scalar load per thread in flight and ran at 0.16 of the
// copy bandwidth); the taps are then shared-memory reads.
__global__ void __launch_bounds__(kPreThreads)
resize_aa_kernel(const float* __restrict__ raw, float* __restrict__ resized, PreStats* __restrict__ stats,
                 const int* __restrict__ aug_mode, const int* __restrict__ xmin, const int* __restrict__ xsize,
                 const float* __restrict__ xw, int mx, const int* __restrict__ ymin, const int* __restrict__ ysize,
                 const float* __restrict__ yw, int my, int Hin, int Win, int H, int W, int TH, int rows_max,
                 int stage_floats) {
  extern __shared__ __align__(16) float smem[];
  float* sraw = smem;                                 // [rows per chunk][pitch]
  float* tmp = sraw + stage_floats;                   // [rows_max][64]
  float* wxs = tmp + (size_t)rows_max * kTileW;       // [64][mx]  (mx is odd: conflict-free)
  float* wys = wxs + kTileW * mx;                     // [TH][my]
  __shared__ unsigned int red_max[kPreThreads / 32], red_min[kPreThreads / 32];

  const int b = blockIdx.z;
  const int ox0 = blockIdx.x * kTileW, oy0 = blockIdx.y * TH;
  const int tw = min(kTileW, W - ox0), th = min(TH, H - oy0);
  const int tid = threadIdx.x;
  for (int i = tid; i < tw * mx; i += kPreThreads) wxs[i] = xw[(size_t)ox0 * mx + i];
  for (int i = tid; i < th * my; i += kPreThreads) wys[i] = yw[(size_t)oy0 * my + i];
  const int r0 = ymin[oy0];
  const int r1 = ymin[oy0 + th - 1] + ysize[oy0 + th - 1];
  const int nrows = r1 - r0;
  const int c0 = xmin[ox0] & ~3;                                        // staged columns [c0, c1)
  const int c1 = xmin[ox0 + tw - 1] + xsize[ox0 + tw - 1];
  const int pitch = (c1 - c0 + 3) & ~3;
  if (nrows > rows_max || pitch > stage_floats) __trap();               // host bound violated: never silently wrong
  const int rows_chunk = stage_floats / pitch;
  const bool vec = (Win & 3) == 0 && (reinterpret_cast<uintptr_t>(raw) & 15) == 0;

  const int c = tid & (kTileW - 1);
  const int lo = c < tw ? xmin[ox0 + c] - c0 : 0, n = c < tw ? xsize[ox0 + c] : 0;
  const float* wc = wxs + c * mx;
  const float* img = raw + ((size_t)b * Hin + r0) * Win + c0;
  for (int rbase = 0; rbase < nrows; rbase += rows_chunk) {
    const int rc = min(rows_chunk, nrows - rbase);
    __syncthreads();                                                    // previous chunk consumed, weights visible
    if (vec) {
      // thread = (float4 column q, row ty + 4k): no index division; four independent 128-bit loads in flight per thread
      const int p4 = pitch >> 2, w4 = Win >> 2;
      const int tx = tid & (kTileW - 1), ty = tid / kTileW;
      constexpr int kStep = kPreThreads / kTileW;
      for (int q = tx; q < p4; q += kTileW) {
        if (c0 + 4 * q >= Win) continue;                                // right of the image: never read by a tap
        const float4* g = reinterpret_cast<const float4*>(img + (size_t)rbase * Win) + q;
        float4* d = reinterpret_cast<float4*>(sraw) + q;
        int rr = ty;
        for (; rr + 3 * kStep < rc; rr += 4 * kStep) {
          const float4 v0 = __ldg(g + (size_t)rr * w4), v1 = __ldg(g + (size_t)(rr + kStep) * w4);
          const float4 v2 = __ldg(g + (size_t)(rr + 2 * kStep) * w4), v3 = __ldg(g + (size_t)(rr + 3 * kStep) * w4);
          d[rr * p4] = v0; d[(rr + kStep) * p4] = v1; d[(rr + 2 * kStep) * p4] = v2; d[(rr + 3 * kStep) * p4] = v3;
        }
        for (; rr < rc; rr += kStep) d[rr * p4] = __ldg(g + (size_t)rr * w4);
      }
    } else {
      const int total = rc * pitch;
      for (int i = tid; i < total; i += kPreThreads) {
        const int rr = i / pitch, q = i - rr * pitch;
        sraw[i] = c0 + q < Win ? __ldg(img + (size_t)(rbase + rr) * Win + q) : 0.f;
      }
    }
    __syncthreads();
    // ---- width pass: staged raw rows -> tmp (fp32, rounded exactly as the CPU path's intermediate image)
    if (c < tw) {
      const float* src = sraw + lo;
      float* dst = tmp + rbase * kTileW + c;
      const int rr0 = tid / kTileW;
#define CVAE_W(N) width_rows<N>(src, pitch, dst, rr0, rc, wc)
      CVAE_TAP_SWITCH(n, CVAE_W, {
        for (int rr = rr0; rr < rc; rr += kPreThreads / kTileW) {
          const float* row = src + rr * pitch;
          dst[rr * kTileW] = aa_taps([&](int j) { return row[j]; }, wc, n);
        }
      })
#undef CVAE_W
    }
  }
  __syncthreads();

  // ---- height pass + flip on store + min / max
  const int aug = aug_mode ? aug_mode[b] : 0;
  float vmax = -INFINITY, vmin = INFINITY;
  if (c < tw) {
    const int ox = ox0 + c;
    const int fx = (aug & 1) ? W - 1 - ox : ox;
    for (int orow = tid / kTileW; orow < th; orow += kPreThreads / kTileW) {
      const int oy = oy0 + orow;
      const float* col = tmp + (ymin[oy] - r0) * kTileW + c;
      const float* wr = wys + orow * my;
      const int ny = ysize[oy];
      float v;
#define CVAE_H(N) v = aa_taps_n<N>([&](int j) { return col[j * kTileW]; }, wr)
      CVAE_TAP_SWITCH(ny, CVAE_H, v = aa_taps([&](int j) { return col[j * kTileW]; }, wr, ny))
#undef CVAE_H
      const int fy = (aug & 2) ? H - 1 - oy : oy;
      resized[((size_t)b * H + fy) * W + fx] = v;
      vmax = fmaxf(vmax, v);
      vmin = fminf(vmin, v);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
  }
  if ((tid & 31) == 0) { red_max[tid >> 5] = f2ord(vmax); red_min[tid >> 5] = ~f2ord(vmin); }
  __syncthreads();
  if (tid == 0) {
    unsigned int a = red_max[0], i = red_min[0];
    for (int k = 1; k < kPreThreads / 32; ++k) { a = max(a, red_max[k]); i = max(i, red_min[k]); }
    atomicMax(&stats[b].max_enc, a);
    atomicMax(&stats[b].min_inv, i);
  }
}

// (v - min) / (max - min) in fp32, all zeros when max == min (dataset.py:229-232)
__device__ __forceinline__ float norm01(float v, float lo, float range, bool flat) {
  return flat ? 0.f : __fdiv_rn(__fsub_rn(v, lo), range);
}

// grid (chunks, B): fp64 sum of the normalised image
__global__ void __launch_bounds__(256) norm_sum_kernel(const float* __restrict__ resized, PreStats* __restrict__ stats,
                                                       int64_t n) {
  __shared__ double red[32];
  const int b = blockIdx.y;
  const float hi = ord2f(stats[b].max_enc), lo = ord2f(~stats[b].min_inv);
  const bool flat = !(hi > lo);
  const float range = __fsub_rn(hi, lo);
  const float* img = resized + (size_t)b * n;
  double acc = 0.0;
  if ((n & 3) == 0) {
    const float4* p = reinterpret_cast<const float4*>(img);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (n >> 2); i += (int64_t)gridDim.x * blockDim.x) {
      const float4 v = p[i];
      acc += (double)norm01(v.x, lo, range, flat) + (double)norm01(v.y, lo, range, flat) +
             (double)norm01(v.z, lo, range, flat) + (double)norm01(v.w, lo, range, flat);
    }
  } else {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
      acc += (double)norm01(img[i], lo, range, flat);
  }
  const double s = block_sum_d(acc, red);
  if (threadIdx.x == 0) atomicAdd(&stats[b].sum, s);
}

// grid (chunks, B): mask = normalised > mean (dataset.py:236-237)
__global__ void __launch_bounds__(256) threshold_kernel(const float* __restrict__ resized,
                                                        const PreStats* __restrict__ stats, float* __restrict__ mask,
                                                        float* __restrict__ thr_out, int64_t n) {
  const int b = blockIdx.y;
  const float hi = ord2f(stats[b].max_enc), lo = ord2f(~stats[b].min_inv);
  const bool flat = !(hi > lo);
  const float range = __fsub_rn(hi, lo);
  const float thr = (float)__ddiv_rn(stats[b].sum, (double)n);
  if (thr_out && blockIdx.x == 0 && threadIdx.x == 0) thr_out[b] = thr;
  const float* img = resized + (size_t)b * n;
  float* out = mask + (size_t)b * n;
  if ((n & 3) == 0) {
    const float4* p = reinterpret_cast<const float4*>(img);
    float4* q = reinterpret_cast<float4*>(out);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (n >> 2); i += (int64_t)gridDim.x * blockDim.x) {
      const float4 v = p[i];
      float4 m;
      m.x = norm01(v.x, lo, range, flat) > thr ? 1.f : 0.f;
      m.y = norm01(v.y, lo, range, flat) > thr ? 1.f : 0.f;
      m.z = norm01(v.z, lo, range, flat) > thr ? 1.f : 0.f;
      m.w = norm01(v.w, lo, range, flat) > thr ? 1.f : 0.f;
      q[i] = m;
    }
  } else {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
      out[i] = norm01(img[i], lo, range, flat) > thr ? 1.f : 0.f;
  }
}

// StandardScaler.transform in fp64, stored as fp32 (dataset.py:116,240)
__global__ void scaler_transform_kernel(const double* __restrict__ m, const double* __restrict__ mean,
                                        const double* __restrict__ scale, float* __restrict__ out, int64_t n, int cols) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cols);
    out[i] = (float)__ddiv_rn(__dsub_rn(m[i], mean[c]), scale[c]);
  }
}

static int host_max_interp(int in_size, int out_size) {
  if (in_size == out_size) return 1;
  const float scale = (float)in_size / (float)out_size;
  const float support = scale >= 1.0f ? scale : 1.0f;
  return (int)ceilf(support) * 2 + 1;
}

}  // namespace cvae

using namespace cvae;

extern "C" int cvae_aa_max_interp(int in_size, int out_size) {
  if (in_size <= 0 || out_size <= 0) return CVAE_ERR_BAD_ARG;
  return host_max_interp(in_size, out_size);
}

extern "C" int cvae_aa_weights(int in_size, int out_size, int* xmin, int* xsize, float* w, cvae_stream_t s) {
  if (in_size <= 0 || out_size <= 0 || !xmin || !xsize || !w) return CVAE_ERR_BAD_ARG;
  aa_weights_kernel<<<(out_size + 127) / 128, 128, 0, as_stream(s)>>>(in_size, out_size,
                                                                      host_max_interp(in_size, out_size), xmin, xsize, w);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_vessel_preprocess(const cvae_preproc_t* p, cvae_stream_t s) {
  if (!p || p->B < 0 || p->Hin <= 0 || p->Win <= 0 || p->H <= 0 || p->W <= 0) return CVAE_ERR_BAD_ARG;
  if (p->B == 0) return CVAE_OK;
  if (!p->raw || !p->resized || !p->stats || !p->mask || !p->xmin || !p->xsize || !p->xw || !p->ymin || !p->ysize ||
      !p->yw)
    return CVAE_ERR_BAD_ARG;
  if (p->B > 65535) return CVAE_ERR_UNSUPPORTED_SHAPE;
  if (((uintptr_t)p->resized | (uintptr_t)p->mask | (uintptr_t)p->stats) & 15) return CVAE_ERR_ALIGNMENT;
  const int mx = host_max_interp(p->Win, p->W), my = host_max_interp(p->Hin, p->H);
  // rows of the width-pass image one TH-row output tile can need: TH*scale + the two half windows (+ rounding)
  const float sy = p->Hin == p->H ? 1.0f : (float)p->Hin / (float)p->H;
  const float sx = p->Win == p->W ? 1.0f : (float)p->Win / (float)p->W;
  // raw columns a 64-column tile can need (+ 3 for the 16-byte align-down, rounded up to a multiple of 4)
  const int pitch_max = ((int)ceilf(kTileW * sx) + mx + 2 + 3 + 3) & ~3;
  const int budget = 47 * 1024 / (int)sizeof(float);   // dynamic + 64 B static must stay under the 48 KB default
  int TH = 16, rows_max = 0, stage_floats = 0;
  if (const char* e = getenv("CVAE_PRE_TH")) {          // tile-height experiment switch (power of two)
    const int v = atoi(e);
    if (v >= 1 && v <= 128 && (v & (v - 1)) == 0) TH = v;
  }
  for (; TH >= 1; TH >>= 1) {
    rows_max = (int)ceilf(TH * sy) + my + 2;
    const long long fixed = (long long)rows_max * kTileW + (long long)kTileW * mx + (long long)TH * my;
    const long long left = budget - fixed;
    const int want_rows = TH == 1 ? 1 : (rows_max < 8 ? rows_max : 8);   // rows per staging chunk worth having
    if (left >= (long long)pitch_max * want_rows) {
      const long long all = (long long)rows_max * pitch_max;
      stage_floats = (int)(left < all ? left : all) & ~3;
      break;
    }
  }
  if (TH < 1) return CVAE_ERR_UNSUPPORTED_SHAPE;
  const size_t smem = ((size_t)stage_floats + (size_t)rows_max * kTileW + (size_t)kTileW * mx + (size_t)TH * my) *
                      sizeof(float);
  cudaStream_t st = as_stream(s);
  if (cudaMemsetAsync(p->stats, 0, (size_t)p->B * sizeof(PreStats), st) != cudaSuccess) return CVAE_ERR_LAUNCH;
  const dim3 grid((p->W + kTileW - 1) / kTileW, (p->H + TH - 1) / TH, p->B);
  if (grid.y > 65535) return CVAE_ERR_UNSUPPORTED_SHAPE;
  resize_aa_kernel<<<grid, kPreThreads, smem, st>>>(p->raw, p->resized, (PreStats*)p->stats, p->aug_mode, p->xmin,
                                                    p->xsize, p->xw, mx, p->ymin, p->ysize, p->yw, my, p->Hin, p->Win,
                                                    p->H, p->W, TH, rows_max, stage_floats);
  CVAE_LAUNCH_CHECK();
  const int64_t n = (int64_t)p->H * p->W;
  // two 128-bit loads per thread: these passes read L2-resident data, their cost is latency, so many short CTAs
  int chunks = (int)((n / 4 + 511) / 512);
  chunks = max(1, min(chunks, 1024));
  const dim3 g2(chunks, p->B);
  norm_sum_kernel<<<g2, 256, 0, st>>>(p->resized, (PreStats*)p->stats, n);
  CVAE_LAUNCH_CHECK();
  threshold_kernel<<<g2, 256, 0, st>>>(p->resized, (const PreStats*)p->stats, p->mask, p->thr, n);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}

extern "C" int cvae_scaler_transform(const double* m, const double* mean, const double* scale, float* out, int64_t rows,
                                     int cols, cvae_stream_t s) {
  if (!m || !mean || !scale || !out || rows < 0 || cols <= 0) return CVAE_ERR_BAD_ARG;
  const int64_t n = rows * cols;
  if (n == 0) return CVAE_OK;
  const int64_t want = (n + 255) / 256;
  const int blocks = (int)(want < 4 * kNumSMs ? want : 4 * kNumSMs);
  scaler_transform_kernel<<<blocks, 256, 0, as_stream(s)>>>(m, mean, scale, out, n, cols);
  CVAE_LAUNCH_CHECK();
  return CVAE_OK;
}
