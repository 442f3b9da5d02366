/* cvae_b200.h — C ABI of the B200-native CausalVAE hot path (libcvae_b200.so).
 *
 * Plain C: POD structs, raw device pointers, explicit sizes, a cudaStream_t passed as void*.
 * No torch types, no allocation or free inside the library (PyTorch owns every buffer), no host
 * synchronisation: every entry point only enqueues work on the given stream and returns
 * CVAE_OK (0) or a negative error code for an unsupported shape / alignment / bad argument.
 *
 * The reference (bjo5029/causal-vae) has no native interface: its device code is reached through
 * torch.nn modules.  Each entry point therefore cites the reference call site whose ATen / cuDNN /
 * cuBLAS work it replaces (paths relative to the reference root).  The Python host side
 * (causal_vae_b200/) mirrors the reference's models.py module API on top of these calls;
 * INTEGRATION.md shows the ctypes binding.
 *
 * Layout conventions
 *   - all tensors fp32, contiguous; image activations are NHWC ("channels_last" physical layout,
 *     logical NCHW on the torch side); matrices are row-major [rows, cols];
 *   - conv weights are consumed in a packed tap-major form [tap][Cin_gather][Cout] produced by
 *     cvae_pack_weight from the torch layouts (Conv2d: [Cout][Cin][kh][kw]; ConvTranspose2d:
 *     [Cin][Cout][kh][kw]; Linear: [out][in] with one tap);
 *   - per-channel statistics buffers are double[2*C] (sum, sum of squares / sum, sum*ref).
 */
#ifndef CVAE_B200_H
#define CVAE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CVAE_OK 0
#define CVAE_ERR_BAD_ARG (-1)
#define CVAE_ERR_UNSUPPORTED_SHAPE (-2)
#define CVAE_ERR_ALIGNMENT (-3)
#define CVAE_ERR_LAUNCH (-4)

typedef void* cvae_stream_t; /* cudaStream_t */

/* library version (major*100 + minor) and the SM architecture the kernels were built for (100). */
int cvae_version(void);
int cvae_built_arch(void);

/* ---- per-channel input transform applied while an operand is loaded ----------------------
 * v' = (v - center[c]) * scale[c] + shift[c]  (skipped when scale == NULL; center == NULL means 0),
 * then leaky-relu with `slope` (slope == 1 -> identity, 0 -> ReLU).  The centred form is the
 * reference's BatchNorm arithmetic ((x - mean) * rstd * gamma + beta): it does not cancel for
 * channels whose |mean| >> std.  This is how training-mode BatchNorm + activation of the
 * PRODUCER layer is applied on the CONSUMER's operand load instead of in a separate pass
 * (vit_backbone.py:74-90,124-156: Conv -> BatchNorm2d -> LeakyReLU chains). */
typedef struct {
  const float* scale; /* [C] or NULL */
  const float* shift; /* [C] or NULL (must be non-NULL when scale is) */
  const float* center; /* [C] or NULL */
  float slope;
} cvae_xform_t;

/* Epilogue of the gather (conv / conv-transpose / linear) kernel. */
enum {
  CVAE_EPI_PLAIN = 0,  /* dst = acc + bias */
  CVAE_EPI_STATS = 1,  /* dst = acc + bias; stats[c] += sum(dst), stats[C+c] += sum(dst^2)  (BN fwd) */
  CVAE_EPI_DACT = 2    /* g = acc (+ add); z = xform(ref) pre-activation; dst = g * act'(z);
                          stats[c] += sum(dst), stats[C+c] += sum(dst*(ref - center))   (act + BN backward) */
};

enum { CVAE_CONV_GATHER = 0, /* Conv2d forward; ConvTranspose2d input-gradient */
       CVAE_CONV_SCATTER = 1 /* ConvTranspose2d forward; Conv2d input-gradient (phase-decomposed) */ };

typedef struct {
  const float* src;   /* [N, Hs, Ws, Cs] NHWC */
  const float* wt;    /* packed [kh*kw][Cs][Cd] */
  const float* bias;  /* [Cd] or NULL */
  float* dst;         /* [N, Hd, Wd, Cd] NHWC */
  cvae_xform_t in;    /* transform of src on load (padding stays exactly 0) */
  int epi;            /* CVAE_EPI_* */
  const float* epi_ref; /* [N,Hd,Wd,Cd] pre-BN output of the layer whose activation is differentiated */
  const float* epi_add; /* optional [N,Hd,Wd,Cd] added to acc before the derivative (residual grad) */
  cvae_xform_t epi_x; /* transform giving the pre-activation z from epi_ref */
  double* stats;      /* [2*Cd] accumulated with atomics (caller zeroes) or NULL */
  int N, Hs, Ws, Cs, Hd, Wd, Cd;
  int kh, kw, stride, pad;
  int mode;           /* CVAE_CONV_GATHER / CVAE_CONV_SCATTER */
} cvae_conv_params_t;

/* Implicit-GEMM convolution family.  Replaces cuDNN conv fwd / dgrad reached from nn.Conv2d,
 * nn.ConvTranspose2d (vit_backbone.py:74-90,124-156; causal_cascade/models.py:12-17,51-54;
 * mnist_test/01_baseline_causal_vae/models.py:19-22,46-47) and cuBLAS sgemm reached from nn.Linear
 * (1x1 "conv" with N = rows, H = W = 1). */
int cvae_conv_gather(const cvae_conv_params_t* p, cvae_stream_t s);

typedef struct {
  const float* ga;   /* gathered operand [N, Ha, Wa, Ca] (read at q*stride - pad + k) */
  const float* db;   /* direct operand   [N, Hq, Wq, Cb] */
  cvae_xform_t xa;   /* transform of ga on load */
  cvae_xform_t xb;   /* transform of db on load */
  float* partial;    /* workspace [splits][kh*kw*Ca][Cb] */
  int splits;        /* K-split count chosen by the caller via cvae_wgrad_splits */
  int N, Ha, Wa, Ca, Hq, Wq, Cb;
  int kh, kw, stride, pad;
} cvae_wgrad_params_t;

/* Weight gradient as a pixels-contracted GEMM:  P[tap][ca][cb] = sum_pix xa(ga[g(pix,tap)][ca]) * xb(db[pix][cb]).
 * Conv2d:  ga = layer input, db = dL/dout.  ConvTranspose2d: ga = dL/dout, db = layer input.
 * Linear: one tap.  Replaces cuDNN wgrad / cuBLAS sgemm-TN in the reference's loss.backward()
 * (vessel_analysis/01_train/train.py:84). */
int cvae_wgrad_splits(int pixels, int rows, int cols); /* recommended K-split for a problem size */
int cvae_conv_wgrad(const cvae_wgrad_params_t* p, cvae_stream_t s);
/* Sum the K-split partials and write the gradient in torch layout dW[cb][ca_real][tap]
 * (ca >= ca_real are padding rows and dropped); accumulate != 0 adds into dst. */
int cvae_wgrad_reduce(const float* partial, int splits, int taps, int ca, int ca_real, int cb,
                      float* dst, int accumulate, cvae_stream_t s);

/* Weight packing: torch layout -> [tap][A_pad][B].  src_bat != 0: src is [B][src_ld][tap] (Conv2d
 * forward, ConvTranspose2d input-grad, Linear forward); else src is [A][src_ld][tap].  src_ld is
 * the size of the source's middle dimension (>= A resp. B: a leading sub-block can be packed).
 * Rows a >= A (up to A_pad) and, for src_bat == 0, columns b >= src_ld are zero-filled. */
int cvae_pack_weight(const float* src, float* dst, int A, int A_pad, int B, int taps, int src_bat,
                     int src_ld, cvae_stream_t s);

/* 1 when cvae_conv_gather runs the layer on the few-channel fp32 tile kernels (image-sized 16 -> 16
 * channel 3x3 stride-2 layers: the decoder's last ConvTranspose2d, vit_backbone.py:146-150, and its
 * input gradient).  Callers that would otherwise take the tensor-core entry point must not: at
 * K = N = 16 the fp32 FMA pipes out-run 3xTF32 tcgen05 (csrc/conv_few.cu).  `wt` is the fp32
 * cvae_pack_weight layout. */
int cvae_conv_few_eligible(int Cs, int Cd, int k, int stride, int pad, int mode, int N, int Hs, int Ws,
                           int Hd, int Wd, int epi);

/* Backward of the image head (nn.Conv2d(C, 1, 3, padding=1), vit_backbone.py:152-156) in one pass over its
 * input: weight gradient and input gradient together, replacing cvae_conv_wgrad + cvae_wgrad_reduce +
 * cvae_conv_gather(EPI_DACT) for this layer (cuDNN wgrad + dgrad in the reference's loss.backward()).
 *   g  [N,H,W,1] dL/dout;  y [N,H,W,C] raw producer output, `x` its BatchNorm + LeakyReLU transform;
 *   w  [1][C][3][3] torch layout;  dz [N,H,W,C] = conv^T(g) * act'(z);  stats [2C] += (sum dz, sum dz*(y-center));
 *   dw [1][C][3][3] torch layout, overwritten. */
int cvae_head_bwd_eligible(int N, int H, int W, int C);
int cvae_head_bwd(const float* g, const float* y, cvae_xform_t x, const float* w, float* dz, double* stats,
                  float* dw, int N, int H, int W, int C, cvae_stream_t s);

/* Batched weight packing: every cvae_pack_weight / cvae_tc_pack_weight call of a training step as one
 * launch.  `jobs_dev` is a DEVICE array sorted by block0 (job i owns blocks [block0_i, block0_{i+1}),
 * cvae_pack_batch_blocks(...) blocks each); tc != 0 selects the tensor-core layout.  The reference
 * has no counterpart: its weights are consumed in torch layout by cuDNN / cuBLAS. */
typedef struct {
  const float* src; float* dst;
  int32_t A, A_pad, B, taps, src_bat, src_ld, tc, block0;
} cvae_pack_job_t;
int cvae_pack_batch_blocks(int A_pad, int B, int taps, int tc);
int cvae_pack_batch(const cvae_pack_job_t* jobs_dev, int njobs, int nblocks, cvae_stream_t s);

/* ---- tensor-core (tcgen05 / TMEM, 3xTF32) variant of the gather family ----------------------------
 * Same contract and parameter block as cvae_conv_gather, for layers with Cs % 16 == 0 and
 * Cd % 16 == 0 (the GEMM-shaped ones: vit_backbone.py:74-90 stem.3..12, :124-156 decoder.0..15 and
 * ResBlocks, the ViT / adapter nn.Linear layers).  `wt` must have been produced by
 * cvae_tc_pack_weight: [tap][ceil(A_pad/32)][hi|lo][B][32] floats, each value split into
 * tf32 hi + lo and each 128-byte row swizzled, i.e. the shared-memory image of the UMMA B operand.
 * Accuracy: hi*hi + lo*hi + hi*lo with fp32 accumulation, ~2^-21 relative per product. */
int cvae_tc_eligible(int Cs, int Cd, int64_t M);          /* 1 when cvae_conv_gather_tc accepts the shape */
int64_t cvae_tc_pack_floats(int A_pad, int B, int taps);  /* size of the packed weight buffer, in floats */
int cvae_tc_pack_weight(const float* src, float* dst, int A, int A_pad, int B, int taps, int src_bat,
                        int src_ld, cvae_stream_t s);     /* arguments as cvae_pack_weight */
int cvae_conv_gather_tc(const cvae_conv_params_t* p, cvae_stream_t s);
/* Linear layers whose input needs no transform (the ViT blocks' Linears, nn.MultiheadAttention's projections and their
 * input gradients; vit_backbone.py:13-41): cvae_tc_pack_rows splits the [M][K] matrix into tf32 hi / lo planes in the
 * tensor core's swizzled shared-memory tile order (cvae_tc_pack_rows_floats floats), and cvae_linear_tc_packed runs the
 * GEMM with BOTH operands streamed by bulk copies -- no thread touches the A operand.  p as for cvae_conv_gather_tc with
 * kh = kw = 1 and an identity input transform; p->src is the matrix the image was packed from. */
int64_t cvae_tc_pack_rows_floats(int64_t M, int K);
int cvae_tc_pack_rows(const float* src, float* dst, int64_t M, int K, cvae_stream_t s);
int cvae_linear_tc_packed(const cvae_conv_params_t* p, const float* a_image, cvae_stream_t s);
/* Tensor-core weight gradient: same contract, parameter block and partial layout as cvae_conv_wgrad
 * (followed by cvae_wgrad_reduce), for Cb % 16 == 0.  The gathered operand is staged straight into
 * tensor memory (lane = (tap, ca) row, columns = pixels); the split count must come from
 * cvae_wgrad_tc_splits (it also bounds the length of one accumulation chain). */
int cvae_wgrad_tc_eligible(int pixels, int rows, int Cb);
int cvae_wgrad_tc_splits(int pixels, int rows, int Cb);
int cvae_conv_wgrad_tc(const cvae_wgrad_params_t* p, cvae_stream_t s);
/* The same kernel without the partial buffer and without cvae_wgrad_reduce: every K split adds its tile into `grad`
 * (torch layout [Cb][ca_real][kh*kw], rows ca >= ca_real of a padded operand are dropped) with fp32 reductions, so
 * `grad` must hold zeros (or the value to accumulate onto) when the launch starts -- the state the trainers' flat
 * gradient buffer is in after zero_grad() (train.py:78 optimizer.zero_grad() in the reference).  p->partial is unused. */
int cvae_conv_wgrad_tc_direct(const cvae_wgrad_params_t* p, float* grad, int ca_real, cvae_stream_t s);

/* Shared-memory tiled fp32 weight gradient for 3x3 (pad 1, stride 1 | 2) layers with few channels and
 * many pixels (Ca in {1,16,32}, Cb in {1,16,32,64}: vit_backbone.py:74-78 stem.0/stem.3, :136-156
 * decoder.8..18): same contract and partial layout as cvae_conv_wgrad, followed by cvae_wgrad_reduce.
 * cvae_wgrad_tile_splits returns the K-split count to allocate (> 0) when the shape is covered, else 0. */
int cvae_wgrad_tile_splits(int pixels, int Ca, int Cb, int k, int stride, int pad);
int cvae_conv_wgrad_tile(const cvae_wgrad_params_t* p, cvae_stream_t s);

/* ---- BatchNorm (training mode: batch statistics; eval: running statistics) -------------------
 * nn.BatchNorm2d / nn.BatchNorm1d, eps 1e-5, momentum 0.1 (vit_backbone.py:76-89;
 * vessel_analysis/00_core/models.py:227,237; causal_cascade/models.py:36). */
/* stats (sum, sumsq over `count` elements per channel) -> scale/shift for the consumer, saved
 * mean/rstd for backward, running-stat EMA update (unbiased variance), num_batches_tracked += 1. */
int cvae_bn_finalize(const double* stats, int C, double count, const float* gamma, const float* beta,
                     float eps, float momentum, float* running_mean, float* running_var,
                     int64_t* num_batches_tracked, float* scale, float* shift, float* mean,
                     float* rstd, cvae_stream_t s);
/* eval mode: scale/shift from running statistics. */
int cvae_bn_eval_coeffs(const float* running_mean, const float* running_var, const float* gamma,
                        const float* beta, float eps, int C, float* scale, float* shift,
                        cvae_stream_t s);
/* sum / sum-of-squares of a [rows, C] matrix into double stats[2C] (for producers without a fused
 * statistics epilogue). */
int cvae_col_stats(const float* y, int64_t rows, int C, double* stats, cvae_stream_t s);
/* backward coefficients from stats = (sum dz, sum dz*(y-mean)):  dy = ca*dz + cb*(y-mean) + cc;
 * dgamma, dbeta, and the gradient of a conv bias feeding the BN (sum dy, analytically zero). */
int cvae_bn_bwd_finalize(const double* stats, int C, double count, const float* gamma,
                         const float* mean, const float* rstd, float* ca, float* cb, float* cc,
                         float* dgamma, float* dbeta, float* dbias_pre, cvae_stream_t s);
/* out[r][c] = xform(a[r][c]) (+ xform2(b[r][c]) when b != NULL): materialise an activation /
 * residual sum (vit_backbone.py:18-19 ResBlock `x + conv(x)`). */
int cvae_affine_act(const float* a, cvae_xform_t xa, const float* b, cvae_xform_t xb, float* out,
                    int64_t rows, int C, cvae_stream_t s);
/* out = ca[c]*dz + cb[c]*(y - mean[c]) + cc[c]  (BN input gradient) */
int cvae_bn_bwd_apply(const float* dz, const float* y, const float* ca, const float* cb,
                      const float* cc, const float* mean, float* out, int64_t rows, int C,
                      cvae_stream_t s);
/* cvae_bn_bwd_finalize + cvae_bn_bwd_apply as one launch (C <= 512, C % 4 == 0): out = BatchNorm-backward of dz
 * given the producer's raw output y and the sums in `stats`; dgamma / dbeta / dbias_pre [C] may be NULL. */
int cvae_bn_bwd_fused_ok(int C);
int cvae_bn_bwd(const float* dz, const float* y, const double* stats, double count, const float* gamma,
                const float* mean, const float* rstd, float* out, float* dgamma, float* dbeta,
                float* dbias_pre, int64_t rows, int C, cvae_stream_t s);
/* dz = g * act'(xform(ref)); stats += (sum dz, sum dz*ref)  — the CVAE_EPI_DACT epilogue as a
 * standalone pass. */
int cvae_dact_stats(const float* g, const float* ref, cvae_xform_t x, float* dz, double* stats,
                    int64_t rows, int C, cvae_stream_t s);

/* ---- LayerNorm over the last dim (nn.LayerNorm, eps 1e-5; vit_backbone.py:26,31,111) ---------- */
int cvae_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean,
                       float* rstd, int64_t rows, int D, int64_t x_row_stride, float eps,
                       cvae_stream_t s);
/* dx (row stride dx_row_stride, accumulate_dx != 0 adds), dgamma/dbeta accumulated with atomics
 * (caller zeroes). */
int cvae_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean,
                       const float* rstd, float* dx, float* dgamma, float* dbeta, int64_t rows, int D,
                       int64_t x_row_stride, int64_t dx_row_stride, int accumulate_dx,
                       cvae_stream_t s);
/* Residual add fused with the LayerNorm that follows it (x = x + attn(...); mlp(norm2(x)), vit_backbone.py:44-46):
 * sum = a + b, y = LayerNorm(sum); backward: dx = LayerNorm-backward(dy) + dadd (dadd = gradient of `sum` from its
 * other consumer), dgamma / dbeta accumulated.  D % 32 == 0, D <= 256 for the backward form. */
int cvae_add_layernorm_fwd(const float* a, const float* b, const float* gamma, const float* beta, float* sum,
                           float* y, float* mean, float* rstd, int64_t rows, int D, float eps, cvae_stream_t s);
int cvae_layernorm_bwd_add(const float* dy, const float* x, const float* gamma, const float* mean,
                           const float* rstd, const float* dadd, float* dx, float* dgamma, float* dbeta,
                           int64_t rows, int D, cvae_stream_t s);

/* ---- multi-head self-attention core (nn.MultiheadAttention; vit_backbone.py:28-30,43) --------
 * qkv: [B, S, 3*D] packed projections; out: [B, S, D]; probs: [B, H, S, S] saved softmax.
 * dropout on the probabilities uses the counter-based generator (seed, offset); `counter`
 * (device int64, may be NULL) is mixed into the seed so a replayed CUDA graph draws fresh masks
 * every step.  S <= 128: one CTA per (batch, head), everything in shared memory (csrc/attention.cu); S > 128
 * (961 tokens at the reference's default 768x1280 image, vessel_analysis/00_core/config.py:10-11): strips of 32
 * rows with the other operand streamed through shared memory (csrc/attention_long.cu), forward through the same
 * entry point, backward through cvae_attention_bwd_ws with a workspace of cvae_attention_ws_bytes()
 * (cvae_attention_bwd itself returns CVAE_ERR_UNSUPPORTED_SHAPE above 128). */
int cvae_attention_fwd(const float* qkv, float* out, float* probs, int B, int S, int H, int d,
                       float dropout_p, uint64_t seed, uint64_t offset, const int64_t* counter,
                       cvae_stream_t s);
int cvae_attention_bwd(const float* qkv, const float* probs, const float* dout, float* dqkv, int B,
                       int S, int H, int d, float dropout_p, uint64_t seed, uint64_t offset,
                       const int64_t* counter, cvae_stream_t s);
int64_t cvae_attention_ws_bytes(int B, int S, int H, int d);
int cvae_attention_bwd_ws(const float* qkv, const float* probs, const float* dout, float* dqkv, float* ws,
                          int64_t ws_bytes, int B, int S, int H, int d, float dropout_p, cvae_stream_t s);

/* ---- elementwise / layout ---------------------------------------------------------------------- */
enum { CVAE_ACT_LRELU = 0, CVAE_ACT_GELU = 1, CVAE_ACT_SIGMOID = 2 };
int cvae_act_fwd(const float* x, float* y, int64_t n, int act, float slope, cvae_stream_t s);
/* dx = dy * act'(x)  (sigmoid takes y = sigmoid(x) in `x`) */
int cvae_act_bwd(const float* dy, const float* x, float* dx, int64_t n, int act, float slope,
                 cvae_stream_t s);
int cvae_add(const float* a, const float* b, float* out, int64_t n, cvae_stream_t s);
/* y = x * mask / (1-p), mask from the counter-based generator; p == 0 copies. Same call on the
 * gradient in backward (nn.Dropout(0.1); vit_backbone.py:35-37,104). */
int cvae_dropout(const float* x, float* y, int64_t n, float p, uint64_t seed, uint64_t offset,
                 const int64_t* counter, cvae_stream_t s);
/* Fused forms with the SAME mask cvae_dropout draws for (seed, offset, counter), 0 < p < 1:
 * mode 0: y = act(x) * mask/(1-p)  (nn.GELU -> nn.Dropout, vit_backbone.py:33-35);  mode 1 (its backward):
 * y = aux * act'(x) * mask/(1-p), aux = dL/dy;  mode 2: y = aux + x * mask/(1-p)  (Dropout -> residual add). */
int cvae_dropout_fused(const float* x, const float* aux, float* y, int64_t n, int mode, int act, float slope,
                       float p, uint64_t seed, uint64_t offset, const int64_t* counter, cvae_stream_t s);
/* y = clamp(x, lo, hi); dx = dy where lo <= x <= hi (torch.clamp; models.py:285-286,294) */
int cvae_clamp_fwd(const float* x, float* y, int64_t n, float lo, float hi, cvae_stream_t s);
int cvae_clamp_bwd(const float* dy, const float* x, float* dx, int64_t n, float lo, float hi,
                   cvae_stream_t s);
/* tokens[b,0] = cls + pos[0]; tokens[b,1+i] = feat[b,i] + pos[1+i]  (vit_backbone.py:164-170) */
int cvae_tokens_fwd(const float* feat, const float* cls, const float* pos, float* tok, int B, int n,
                    int D, cvae_stream_t s);
int cvae_tokens_bwd(const float* dtok, float* dfeat, float* dcls, float* dpos, int B, int n, int D,
                    cvae_stream_t s);
/* [B, C, HW] <-> [B, HW, C] (decoder_input .view(-1,256,gh,gw), vit_backbone.py:188) */
int cvae_transpose_bc(const float* src, float* dst, int B, int rows, int cols, cvae_stream_t s);
/* copy a [rows, w] block into dst[:, col0:col0+w] of a [rows, ld] matrix (torch.cat along dim 1);
 * accumulate != 0 adds. src_ld is the source row stride, src_col0 the first source column. */
int cvae_copy_cols(const float* src, int64_t src_ld, int src_col0, float* dst, int64_t dst_ld,
                   int dst_col0, int64_t rows, int w, int accumulate, cvae_stream_t s);
int cvae_fill(float* dst, int64_t n, float v, cvae_stream_t s);
/* column sums of a [rows, C] matrix: out[c] (+)= sum_r x[r][c] */
int cvae_col_sum(const float* x, int64_t rows, int C, float* out, int accumulate, cvae_stream_t s);

/* ---- latent: clamp + reparameterise + KL (vessel_analysis/00_core/models.py:281-288,252-255;
 * train.py:49).  h: [B, 2Z] adapter output (mu | logvar).  mu = clamp(h[:, :Z], +-mu_clamp),
 * logvar = clamp(h[:, Z:], +-lv_clamp) (clamp <= 0 disables), z = mu + eps*exp(0.5*logvar);
 * kld_sum accumulates -0.5*sum(1 + logvar - mu^2 - exp(logvar)) in double (caller zeroes). */
int cvae_latent_fwd(const float* h, const float* eps, float* mu, float* logvar, float* z,
                    double* kld_sum, int B, int Z, float mu_clamp, float lv_clamp, cvae_stream_t s);
/* dh from dz, dmu, dlogvar (any may be NULL) with the clamp masks recomputed from h. */
int cvae_latent_bwd(const float* h, const float* eps, const float* dz, const float* dmu,
                    const float* dlogvar, float* dh, int B, int Z, float mu_clamp, float lv_clamp,
                    cvae_stream_t s);
/* Gaussian NLL 0.5*sum(lv + (m-mu)^2/exp(lv)) with lv = clamp(raw_lv) (models.py:294, train.py:55-58),
 * KL term gradient helpers: d/dmu = g*mu, d/dlogvar = g*0.5*(exp(lv)-1). */
int cvae_gauss_nll_fwd(const float* m, const float* m_mu, const float* raw_lv, float* lv_out,
                       double* nll_sum, int64_t n, float lv_clamp, cvae_stream_t s);
int cvae_gauss_nll_bwd(const float* m, const float* m_mu, const float* raw_lv, const float* gscale,
                       float gmul, float* d_mu, float* d_raw_lv, int64_t n, float lv_clamp,
                       cvae_stream_t s);
/* sum += -0.5 * sum(1 + logvar - mu^2 - exp(logvar)) */
int cvae_kld_fwd(const float* mu, const float* logvar, int64_t n, double* sum, cvae_stream_t s);
int cvae_kld_bwd(const float* mu, const float* logvar, const float* gscale, float gmul, float* dmu,
                 float* dlogvar, int64_t n, int accumulate, cvae_stream_t s);

/* ---- reconstruction losses -------------------------------------------------------------------- */
/* vessel loss_function (vessel_analysis/01_train/train.py:18-46): pass 1 = sum(x) -> sums[0];
 * pass 2 = weighted MSE -> sums[1], sparsity L1 -> sums[2], with
 * pos_weight = clamp((1-pf)/(pf+1e-6), 1, 50), pf = sum(x)/(n+1e-6) computed on device. */
int cvae_vessel_xsum(const float* x, int64_t n, double* sums, cvae_stream_t s);
int cvae_vessel_recon_fwd(const float* recon, const float* x, int64_t n, double* sums, cvae_stream_t s);
/* d_recon = g_recon*2*(r-x)*w + g_sparsity*sign(r)*[x<0.1]; g_* are device scalars (upstream
 * gradients of the two loss terms). */
int cvae_vessel_recon_bwd(const float* recon, const float* x, int64_t n, const double* sums,
                          const float* g_recon, const float* g_sparsity, float* d_recon,
                          cvae_stream_t s);
/* sum-reduced MSE (causal_cascade/train.py:7,10; latent_translator/engine.py:25 with scale 1/n) and
 * sum-reduced BCE with log clamped at -100 (mnist_test/01_baseline_causal_vae/train.py:70). */
int cvae_mse_fwd(const float* a, const float* b, int64_t n, double* sum, cvae_stream_t s);
int cvae_mse_bwd(const float* a, const float* b, int64_t n, const float* g, float gmul, float* da,
                 cvae_stream_t s);
int cvae_bce_fwd(const float* p, const float* y, int64_t n, double* sum, cvae_stream_t s);
int cvae_bce_bwd(const float* p, const float* y, int64_t n, const float* g, float gmul, float* dp,
                 cvae_stream_t s);
/* double accumulator -> float scalar (optionally scaled) */
int cvae_finish_scalar(const double* acc, float mul, float* out, cvae_stream_t s);
/* out = wa*a + wb*b + wc*c + wd*d over device scalars (b, c, d may be NULL): the weighted total loss
 * (train.py:82); scale4: out4[i] = w_i * g, its backward. */
int cvae_scalar_combine(const float* a, const float* b, const float* c, const float* d, float wa, float wb,
                        float wc, float wd, float* out, cvae_stream_t s);
int cvae_scalar_scale4(const float* g, float wa, float wb, float wc, float wd, float* out4, cvae_stream_t s);

/* ---- treatment labels: argmax / one-hot (integer, bit-exact) and softmax losses on [rows, T] logits -----
 * torch.argmax(t, dim=1) (first maximum; mnist_test/01_baseline_causal_vae/train.py:38),
 * F.one_hot(t, T).float() (causal_cascade/models.py:71),
 * F.cross_entropy(logits, idx) summed over rows (train.py:55; the caller scales by 1/rows), and
 * F.kl_div(log_softmax(logits), U(1/T), 'batchmean') * rows (train.py:78-82).  Backward writes
 * dlogits = (softmax - onehot) * g*gmul  resp.  (softmax - 1/T) * g*gmul; g is a device scalar or NULL. */
int cvae_argmax_rows(const float* t, int64_t rows, int T, int64_t* out, cvae_stream_t s);
int cvae_one_hot(const int64_t* idx, int64_t rows, int T, float* out, cvae_stream_t s);
int cvae_softmax_ce_fwd(const float* logits, const int64_t* target, int64_t rows, int T, double* sum,
                        cvae_stream_t s);
int cvae_softmax_ce_bwd(const float* logits, const int64_t* target, int64_t rows, int T, const float* g,
                        float gmul, float* dlogits, cvae_stream_t s);
int cvae_uniform_kl_fwd(const float* logits, int64_t rows, int T, double* sum, cvae_stream_t s);
int cvae_uniform_kl_bwd(const float* logits, int64_t rows, int T, const float* g, float gmul,
                        float* dlogits, cvae_stream_t s);

/* ---- counterfactual: do(M_k := M_k + delta | M_k := value) over all K concepts ------------------
 * (generate_counterfactual.py:86-88; analyze_vessel.py:101-104).  Builds the decoder-adapter input
 * rows [m' | z] for S sources x K concepts: row (s*K + k) = cat(do_k(m[s]), z[s]). */
int cvae_do_expand(const float* m, const float* z, float* out, int S, int K, int Z, int set_value,
                   float v, cvae_stream_t s);
/* rows (i*N + j) = cat(m[i], scale * z[j]): the M x Z cross-product grid of check_mechanism_z_perm.py:100-118 */
int cvae_pair_expand(const float* m, const float* z, float* out, int N, int K, int Z, float scale, cvae_stream_t s);
/* elementwise mean and unbiased std over the reconstructions of n_models <= 8 fold models (`preds`: HOST array of device
 * pointers; torch.stack(..).mean(0) / .std(0), ensemble_reconstruction.py:80-86); stdv may be NULL */
int cvae_ensemble_mean_std(const float* const* preds, int n_models, float* mean, float* stdv, int64_t n, cvae_stream_t s);
/* per-row L2 norm of (a - b[row / group]) (analyze_vessel.py:115) */
int cvae_rowdiff_l2(const float* a, const float* b, float* out, int64_t rows, int64_t rowlen,
                    int group, cvae_stream_t s);

/* ---- optimizer: global-norm clip + Adam over flat fp32 buffers ----------------------------------
 * torch.nn.utils.clip_grad_norm_(max_norm) + torch.optim.Adam defaults
 * (vessel_analysis/01_train/train.py:85-86,152). */
int cvae_sumsq(const float* g, int64_t n, double* acc, cvae_stream_t s);
/* step_count is a device int64 incremented by the kernel (graph-replay safe); max_norm <= 0
 * disables clipping; grad_scale multiplies the gradient first (1/world for mean-reduced losses). */
int cvae_clip_adam(float* p, const float* g, float* m, float* v, int64_t n, const double* sumsq,
                   float max_norm, float lr, float beta1, float beta2, float eps, float grad_scale,
                   int64_t* step_count, cvae_stream_t s);

/* nn.Upsample(scale_factor=2, mode='nearest') of the CNN vessel decoder (vessel_analysis/00_core/models.py:123-145)
 * on NHWC: y[n, 2h+a, 2w+b, c] = x[n, h, w, c] (H, W are the INPUT sizes); backward sums each 2x2 block. */
int cvae_upsample2x_fwd(const float* x, float* y, int N, int H, int W, int C, cvae_stream_t s);
int cvae_upsample2x_bwd(const float* dy, float* dx, int N, int H, int W, int C, cvae_stream_t s);

/* ---- vessel input pipeline on the device ---------------------------------------
 * What VesselDataset.__getitem__ does per raw image on DataLoader worker CPUs
 * (vessel_analysis/00_core/dataset.py:186,216-237): transforms.Resize((H, W), antialias=True) (ATen's separable
 * antialiased bilinear kernel, width pass first, fp32), hflip / vflip / both by aug_mode = idx % 4
 * (dataset.py:219-226), per-image min-max to [0,1] (all zeros when max == min, :229-232), threshold at the image
 * mean -> {0,1} (:236-237) — for a whole batch of same-sized raw images.  Bit-for-bit the CPU arithmetic; the mean
 * is the correctly rounded one (fp64 sum of the fp32 normalised values), see oracle/input_oracle.py.
 *
 * cvae_aa_max_interp: taps per output index of one axis (host-side, no launch; 1 when in == out).
 * cvae_aa_weights: fills xmin[out], xsize[out], w[out][max_interp] for one axis (cache per (in, out)). */
int cvae_aa_max_interp(int in_size, int out_size);
int cvae_aa_weights(int in_size, int out_size, int* xmin, int* xsize, float* w, cvae_stream_t s);
typedef struct {
  const float* raw;    /* [B, Hin, Win] raw images */
  float* resized;      /* [B, H, W] workspace: the resized + flipped fp32 image (16-byte aligned) */
  void* stats;         /* workspace, 16 bytes per image (zeroed by the call) */
  float* mask;         /* [B, H, W] output {0,1} == x[B,1,H,W] */
  float* thr;          /* [B] optional: the fp32 threshold each image was cut at */
  const int* aug_mode; /* [B] 0..3 or NULL (no flips: validation) */
  const int* xmin; const int* xsize; const float* xw; /* width axis, cvae_aa_weights(Win, W) */
  const int* ymin; const int* ysize; const float* yw; /* height axis, cvae_aa_weights(Hin, H) */
  int B, Hin, Win, H, W;
} cvae_preproc_t;
int cvae_vessel_preprocess(const cvae_preproc_t* p, cvae_stream_t s);
/* StandardScaler.transform of the morphology features in fp64, stored fp32 (dataset.py:113-116,240):
 * out[r][c] = (float)((m[r][c] - mean[c]) / scale[c]). */
int cvae_scaler_transform(const double* m, const double* mean, const double* scale, float* out, int64_t rows,
                          int cols, cvae_stream_t s);

/* ---- replay of a captured step with per-node priorities ------------------------------------------------------------
 * The reference's step (train.py:77-86) is one stream-ordered sequence; here it is captured once into a CUDA graph whose
 * weight-gradient family runs as a parallel branch.  cvae_graph_set_priorities marks every kernel node of the captured
 * cudaGraph_t: `prio_side` for the weight-gradient family (kernel names containing "wgrad" / "col_reduce"), `prio_main`
 * for everything else (CUDA convention: lower value = higher priority); counts[3] = {kernel nodes, side nodes, nodes
 * whose name could not be resolved}.  cvae_graph_instantiate_prio instantiates with
 * cudaGraphInstantiateFlagUseNodePriority (a default instantiation ignores node priorities); cvae_graph_launch replays
 * it on `s`.  The caller keeps the cudaGraph_t and its memory pool alive. */
int cvae_graph_set_priorities(void* graph, int prio_main, int prio_side, int* counts);
int cvae_graph_instantiate_prio(void* graph, void** exec_out);
int cvae_graph_launch(void* exec, cvae_stream_t s);
int cvae_graph_exec_destroy(void* exec);

/* ---- debug: role-level wait accounting of the tensor-core kernels (builds with -DCVAE_TIMING only;
 * returns 0 and leaves out16 untouched otherwise).  Not part of the reference-facing surface. */
int cvae_debug_read(unsigned long long* out16, int reset);

#ifdef __cplusplus
}
#endif
#endif /* CVAE_B200_H */
