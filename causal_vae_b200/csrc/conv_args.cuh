// Kernel argument blocks shared by conv.cu (tiled implicit GEMM) and skinny.cu (1-channel layers).
#pragma once
#include "common.cuh"

namespace cvae {

struct TapEntry { int dh, dw, widx; };
struct PhaseGeom { int ph, pw, Hq, Wq, ntaps; TapEntry taps[16]; };

struct GatherArgs {
  const float* src; const float* wt; const float* bias; float* dst;
  const float* in_scale; const float* in_shift; const float* in_center; float in_slope; int in_affine; int in_act;
  int epi; const float* epi_ref; const float* epi_add;
  const float* e_scale; const float* e_shift; const float* e_center; float e_slope; int e_affine;
  double* stats;
  int N, Hs, Ws, Cs, Hd, Wd, Cd;
  int os, is, wtaps, nphase, ksplit;
  PhaseGeom phase[4];
  const float* a_image;   // Linear layers only: the A operand already split and swizzled by cvae_tc_pack_rows (else nullptr)
};

struct WgradArgs {
  const float* ga; const float* db;
  const float* a_scale; const float* a_shift; const float* a_center; float a_slope; int a_affine; int a_act;
  const float* b_scale; const float* b_shift; const float* b_center; float b_slope; int b_affine; int b_act;
  float* partial;
  int N, Ha, Wa, Ca, Hq, Wq, Cb;
  int kw, stride, pad, rows, kchunk, K;
  int seg;   // wgrad_tc A-loader variant: 32 / 16 / 8 = whole output-row segments per stage, -1 = Linear, 0 = general
  // wgrad_tc direct mode (cvae_conv_wgrad_tc_direct): split-K tiles meet in the PRE-ZEROED torch-layout gradient
  // direct[cb][ca < ca_real][tap] through fp32 reductions (red.global.add) - no partial buffer, no reduce launch
  float* direct = nullptr; int ca_real = 0; int taps = 0;
};

// skinny.cu: 1-channel layers as pure HBM streams.  Each returns false when the shape is not covered.
bool launch_conv_cs1(const GatherArgs& g, int maxM, cudaStream_t st);   // Cs == 1  -> Cd % 4 == 0, Cd <= 64
bool launch_conv_cd1(const GatherArgs& g, int maxM, cudaStream_t st);   // Cs % 4 == 0, Cs <= 64 -> Cd == 1, plain epilogue
bool launch_wgrad_cb1(const WgradArgs& a, int taps, cudaStream_t st);   // Cb == 1, Ca % 4 == 0, Ca <= 64
bool launch_wgrad_ca1(const WgradArgs& a, int taps, cudaStream_t st);   // Ca == 1, Cb % 4 == 0, Cb <= 64

// conv_few.cu: image-sized 16 -> 16 stride-2 layers.  1: launched, 0: not covered, < 0: error.
int launch_conv_few(const cvae_conv_params_t* p, const GatherArgs& g, cudaStream_t st);
int launch_conv16_head_fwd(const GatherArgs& g, cudaStream_t st);   // Conv 16 -> 1 3x3 s1 (image head forward)

// linear_small.cu: Linear layers / weight gradients with M <= 512 rows.  1: launched, 0: not covered, < 0: error.
int launch_linear_small(const GatherArgs& g, cudaStream_t st);
int launch_wgrad_small(const WgradArgs& a, int taps, int splits, cudaStream_t st);

}  // namespace cvae
