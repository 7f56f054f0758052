/*
 * stemb200 — C ABI of the B200-native STEM P-frame hot path (libstemb200.so).
 *
 * The reference (mmSir/SpatioTemporalEntropyModel, a CompressAI 1.1.1 fork) has no FFI layer on this path:
 * its boundary is the Python nn.Module API, whose arithmetic is ATen library calls. Each entry point below
 * therefore cites the reference *call site* (file:line under /root/reference) whose arithmetic it replaces.
 * The Python package `spatiotemporalentropymodel_b200` binds these with ctypes and mirrors the module API.
 *
 * Conventions
 *   - plain pointers and sizes only; device pointers unless the name says `host`;
 *   - every function returns 0 on success or a negative STEMB200_E_* code (never throws, never allocates
 *     device memory; outputs and workspaces are caller-allocated);
 *   - `stream` is a cudaStream_t passed as void*; functions are stateless and thread-safe per stream;
 *   - activations between kernels are NHWC ("pixels x channels"), element type given by STEMB200_DT_*;
 *     tensors crossing the reference API are NCHW fp32 and are converted by the nchw/nhwc entry points.
 */
#ifndef STEMB200_H_
#define STEMB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STEMB200_OK 0
#define STEMB200_E_INVALID -1   /* bad argument / unsupported geometry */
#define STEMB200_E_CUDA -2      /* a CUDA runtime / driver call failed (see stemb200_last_error) */
#define STEMB200_E_NODEVICE -3  /* no sm_100 device */

/* epilogues of stemb200_conv2d_fwd */
#define STEMB200_EPI_LINEAR 0 /* out = LeakyReLU(acc + bias) */
#define STEMB200_EPI_SFT 1    /* spatial feature transform (stem_utils.py:36-43): the conv computes gamma and beta at
                                 once (c_out = 2 x channels, per 128 accumulator columns [gamma(64) | beta(64)]),
                                 out = LeakyReLU(aux * gamma + beta) with c_out/2 channels; fold the "+1" of
                                 (1 + gamma) into gamma's bias */
#define STEMB200_EPI_ADD 2    /* out = LeakyReLU(acc + bias) + aux (SFTResblk skip connection, stem_utils.py:55-60) */

#define STEMB200_DT_F16 0 /* IEEE half operands, fp32 accumulate (tcgen05 kind::f16) */
#define STEMB200_DT_F32 1 /* fp32 storage (epilogue outputs that feed quantisation) */

const char* stemb200_version(void);
const char* stemb200_last_error(void);
/* number of kernels this library has launched since process start (bench.py: gpu_launches) */
uint64_t stemb200_launch_count(void);

/* ---------------------------------------------------------------------------------------------------
 * Dense contractions: nn.Conv2d / nn.ConvTranspose2d / MaskedConv2d / GDN's gamma conv as implicit GEMM
 * on tcgen05 (TMEM accumulators, TMA-fed). Replaces F.conv2d / F.conv_transpose2d at
 *   compressai/models/utils.py:112-130 (conv/deconv factories, k5 s2),
 *   compressai/models/spatiotemporalpriors.py:523-554 (TPM, HE, HD, context_prediction, EPM),
 *   compressai/layers/layers.py:44-47 (MaskedConv2d), compressai/layers/gdn.py:58 (gamma 1x1 conv).
 * ------------------------------------------------------------------------------------------------- */
typedef struct stemb200_conv_desc {
  int32_t batch;        /* N */
  int32_t h_in, w_in;   /* input spatial size (all sources share it) */
  int32_t n_src;        /* 1..3 inputs concatenated along channels (torch.cat(..., 1) folded into K) */
  int32_t c_in[3];      /* channels per source, multiples of 8 (the last 64-channel K chunk of a source is
                           zero-filled by TMA when the count is not a multiple of 64) */
  int32_t c_out;        /* output channels (multiple of 16) */
  int32_t kh, kw;       /* kernel size (1, 3 or 5; square padding k/2) */
  int32_t stride;       /* 1 or 2 */
  int32_t transposed;   /* 1: ConvTranspose2d(k, stride 2, padding k/2, output_padding 1) */
  uint32_t tap_mask;    /* bit (r*kw+s) set = tap used; 0 = all taps (mask 'A' 5x5 = 0x00000FFF) */
  int32_t epilogue;     /* STEMB200_EPI_* */
  float lrelu_slope;    /* LeakyReLU negative slope; 1.0f = no activation, 0.0f = ReLU */
  int32_t out_dtype;    /* STEMB200_DT_F16 or STEMB200_DT_F32 (NHWC) */
  float sq_scale;       /* conv2d_gdn_fwd only: x is prescaled by this before squaring so that x^2 stays inside
                           the fp16 range; the normaliser is rescaled by 1/sq_scale^2 */
  int32_t tile_h, tile_w; /* output patch per 128-row MMA tile, tile_h*tile_w <= 128; 0 = choose */
  int32_t direct_store; /* 1: epilogue stores straight from registers (debug / c_out < 32) */
  int32_t row_taps;     /* 1: few-channel first layer (conv k x k, stride 2, c_in[0] == 8, k <= 8; priors.py:422).
                           in[0] is the zero-bordered canvas written by stemb200_frame_to_nhwc8:
                           [batch][h_in + 2*(k/2)][w_in + 2*(k/2)][8] fp16 (+ 64 elements of slack), h_in / w_in the
                           logical (even) input size. One K step per kernel ROW: the k taps of a row are 8k
                           contiguous fp16 of the canvas, fetched as one 64-element TMA box through a tensor map
                           whose pixel stride (2 pixels) is smaller than its box (overlapping windows), so no
                           im2col buffer exists. Packed weight: [c_out][k*64], K index = r*64 + s*8 + ch. */
} stemb200_conv_desc;

/* K extent of the packed weight matrix [c_out][K] for this geometry */
int64_t stemb200_conv2d_packed_k(const stemb200_conv_desc* d);
/* Repack a PyTorch weight (Conv2d: [c_out][c_in][kh][kw]; ConvTranspose2d: [c_in][c_out][kh][kw]), fp32 on
 * device, into the K-major fp16 matrix the kernel's TMA descriptor expects. Masked taps are dropped. */
int stemb200_conv2d_pack_weight(const stemb200_conv_desc* d, const float* weight_f32, void* packed_f16,
                                void* stream);
/* in[s]: NHWC fp16 [batch][h_in][w_in][c_in[s]]; out: NHWC [batch][h_out][w_out][c_out] (c_out/2 channels for
 * EPI_SFT); bias: fp32 [c_out]; aux: NHWC fp16 with the output's shape (EPI_SFT: x, EPI_ADD: residual) or NULL. */
int stemb200_conv2d_fwd(const stemb200_conv_desc* d, const void* const* in, const void* packed_weight,
                        const float* bias, const void* aux, void* out, void* stream);

/* Convolution / transposed convolution with GDN (inverse = 0) or IGDN (inverse = 1) fused into the epilogue:
 *   x = conv(in) + bias;  out = x * rsqrt(beta + gamma . x^2)   |   x * sqrt(beta + gamma . x^2)
 * (priors.py:421-439 conv/deconv followed by layers/gdn.py:52-67). c_out must be 128 or 192; packed_gamma is the
 * [c_out][c_out] matrix produced by stemb200_conv2d_pack_weight for a 1x1 conv of the *re-parametrised* gamma
 * (ops/parametrizers.py:42-45), beta the re-parametrised beta (fp32). d->sq_scale prescales x before squaring
 * (fp16 range); out is NHWC fp16. x and x^2 never leave the SM. */
int stemb200_conv2d_gdn_fwd(const stemb200_conv_desc* d, const void* const* in, const void* packed_weight,
                            const float* bias, const void* packed_gamma, const float* beta, int32_t inverse,
                            void* out, void* stream);

/* The last two synthesis layers in one pass (priors.py:436-438: deconv(N, N) + IGDN, then deconv(N, 3)): as
 * stemb200_conv2d_gdn_fwd with inverse = 1 and c_out = 192, plus, while a tile of the layer's output is still in
 * shared memory, its product with packed_w6 = the [96][192] fp16 matrix
 * W6[stemb200_synthesis_col_index(r, s, c)][ci] = w[ci][c][r][s] of the final ConvTranspose2d(192, 3, 5, stride 2)
 * (the other 21 rows zero; pack it with stemb200_conv2d_pack_weight as a 1x1 conv 192 -> 96). col_out: NHWC fp16 [batch][2 h_in][2 w_in][96], the per-pixel tap contributions that
 * stemb200_synthesis_col2im sums; act_out: the layer's own activation (NHWC fp16) or NULL to skip writing it
 * (it is then never in HBM). */
int stemb200_conv2d_gdn_last_fwd(const stemb200_conv_desc* d, const void* const* in, const void* packed_weight,
                                 const float* bias, const void* packed_gamma, const float* beta,
                                 const void* packed_w6, void* col_out, void* act_out, void* stream);
/* Column of the col rows that holds tap (r, s) of output channel c (0 <= r, s <= 4, 0 <= c <= 2), in [0, 96), or a
 * negative error code. The order groups the taps by the 2x2 output quad they land in, so that the col2im kernel sums
 * a quad from aligned vector loads: taps r in {0,1} / {2,3} / {4} reach the quad one row above / in / below the
 * input pixel, likewise s for columns; inside a group the order is [r & 1][s & 1][c]. */
int stemb200_synthesis_col_index(int32_t r, int32_t s, int32_t c);
/* col2im + bias + clamp + squared error of the final deconv (k5, s2, p2, op1): x_hat[c][oh][ow] = bias[c] +
 * sum over taps (r, s) with oh = 2 i - 2 + r, ow = 2 j - 2 + s of col[i][j][stemb200_synthesis_col_index(r, s, c)];
 * col: NHWC fp16 [n][h2][w2][96], 16-byte aligned; x_hat: NCHW fp32 [n][3][2 h2][2 w2], 8-byte aligned; x_ref /
 * sq_err as in stemb200_synthesis_tail. */
int stemb200_synthesis_col2im(const void* col_f16, const float* bias3, float* x_hat_nchw, int32_t n, int32_t h2,
                              int32_t w2, const float* x_ref, int32_t h_ref, int32_t w_ref, int32_t pad_top,
                              int32_t pad_left, double* sq_err, int32_t clamp01, void* stream);
/* Same with the reference frame as 8-bit samples [n][3][h_ref][w_ref]: the pixel value is float(v) / 255 (IEEE
 * division), i.e. what torchvision's ToTensor hands to the model for the PNG frames of stem/evalSTEM.py:185 -
 * results are bit-identical to stemb200_synthesis_col2im on the converted frame, with a quarter of the bytes
 * crossing PCIe and HBM. */
int stemb200_synthesis_col2im_u8(const void* col_f16, const float* bias3, float* x_hat_nchw, int32_t n, int32_t h2,
                                 int32_t w2, const uint8_t* x_ref, int32_t h_ref, int32_t w_ref, int32_t pad_top,
                                 int32_t pad_left, double* sq_err, int32_t clamp01, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Layout / staging kernels at the API boundary
 * ------------------------------------------------------------------------------------------------- */
/* NCHW fp32 -> NHWC fp16 of (in - sub) (sub may be NULL), optionally rounded to nearest-even integer first:
 * y_hat = round(y) / round(y_cur - y_conditioned) (entropy_models.py:141, spatiotemporalpriors.py:852-856) */
int stemb200_nchw_f32_to_nhwc_f16(const float* in, const float* sub, void* out, int32_t n, int32_t c, int32_t h,
                                  int32_t w, int32_t round_first, void* stream);
int stemb200_nhwc_f16_to_nchw_f32(const void* in, float* out, int32_t n, int32_t c, int32_t h, int32_t w,
                                  void* stream);
int stemb200_nhwc_f32_to_nchw_f32(const float* in, float* out, int32_t n, int32_t c, int32_t h, int32_t w,
                                  void* stream);
/* First analysis conv (3 -> N, k5 s2 p2, priors.py:422) operand staging: im2col of the NCHW fp32 frame into
 * [n*h_out*w_out][80] fp16 rows (k = (r*5+s)*3+ch for the 75 real taps*channels, then 5 zeros), with the frame
 * embedded at (pad_top, pad_left) inside a zero canvas of h_pad x w_pad (evalSTEM.py:96-109 pads to a multiple of
 * 64). The rows are the NHWC input (C = 80) of a 1x1 stemb200_conv2d_gdn_fwd whose weight is the [N][75] reshape. */
int stemb200_im2col_k5s2_c3(const float* x_nchw, void* out_rows, int32_t n, int32_t h, int32_t w,
                            int32_t h_pad, int32_t w_pad, int32_t pad_top, int32_t pad_left, void* stream);
/* Same from an 8-bit NCHW frame (pixel = float(v) / 255, see stemb200_synthesis_col2im_u8). */
int stemb200_im2col_k5s2_c3_u8(const uint8_t* x_nchw, void* out_rows, int32_t n, int32_t h, int32_t w, int32_t h_pad,
                               int32_t w_pad, int32_t pad_top, int32_t pad_left, void* stream);
/* Operand canvas of the row_taps first layer (priors.py:422 on the evalSTEM.py:96-109 padded frame): NCHW fp32
 * (n, c <= 8, h, w) -> NHWC fp16 [n][h_pad + 2*border][w_pad + 2*border][8], the frame at (pad_top + border,
 * pad_left + border), zeros elsewhere (conv padding, frame padding and channels c..7). */
int stemb200_frame_to_nhwc8(const float* x_nchw, void* canvas, int32_t n, int32_t c, int32_t h, int32_t w,
                            int32_t h_pad, int32_t w_pad, int32_t pad_top, int32_t pad_left, int32_t border,
                            void* stream);
/* Same from an 8-bit NCHW frame (pixel = float(v) / 255, see stemb200_synthesis_col2im_u8). */
int stemb200_frame_u8_to_nhwc8(const uint8_t* x_nchw, void* canvas, int32_t n, int32_t c, int32_t h, int32_t w,
                               int32_t h_pad, int32_t w_pad, int32_t pad_top, int32_t pad_left, int32_t border,
                               void* stream);

/* Canvas with 4 channels per pixel (R, G, B, 0): the operand of stemb200_conv_first_gdn_fwd. Same geometry as
 * stemb200_frame_to_nhwc8 with 8 bytes per pixel; c <= 4. */
int stemb200_frame_to_nhwc4(const float* x_nchw, void* canvas, int32_t n, int32_t c, int32_t h, int32_t w,
                            int32_t h_pad, int32_t w_pad, int32_t pad_top, int32_t pad_left, int32_t border,
                            void* stream);
int stemb200_frame_u8_to_nhwc4(const uint8_t* x_nchw, void* canvas, int32_t n, int32_t c, int32_t h, int32_t w,
                               int32_t h_pad, int32_t w_pad, int32_t pad_top, int32_t pad_left, int32_t border,
                               void* stream);
/* First analysis layer + GDN with resident weights (priors.py:422 conv(3, N) + :423 GDN(N); layers/gdn.py:52-67),
 * N = 192: conv 3 -> 192, k5, s2, p2 on the NHWC4 canvas [n][h_in + 4][w_in + 4][4] fp16 (+ 64 elements of slack;
 * h_in / w_in the even padded frame size, border = 2), then x * rsqrt(beta + gamma . x^2).
 *   packed_w0   : [192][5 kernel rows][8 taps][4 channels] fp16 = w[o][ch][r][s] at k = r*32 + s*4 + ch (taps 5..7 and
 *                 channel 3 zero): one kernel row of an output pixel is 32 contiguous fp16 of the canvas;
 *   packed_gamma: [192][192] fp16, K-major, of the re-parametrised gamma (as for stemb200_conv2d_gdn_fwd); beta fp32;
 *   sq_scale    : x is prescaled by it before squaring (power of two; fp16 range), out: NHWC fp16 [n][h_in/2][w_in/2][192].
 * W0 and gamma stay in shared memory for the whole launch; the TMA ring carries only the 8 KB operand tiles. */
int stemb200_conv_first_gdn_fwd(const void* canvas_nhwc4, int32_t n, int32_t h_in, int32_t w_in, int32_t border,
                                const void* packed_w0, const float* bias, const void* packed_gamma, const float* beta,
                                float sq_scale, void* out, void* stream);

/* stem_roi (compressai/models/stem_roi.py) staging kernels.
 * im2col_k3s1_c4: operand rows of conv(4, 192, k3, s1) on cat[x (3 ch), Qmap (1 ch)] (:379, :586):
 *   [n*h*w][40] fp16, k = (r*3+s)*4 + ch for 36 entries, then 4 zeros.
 * avgpool_nhwc_f16: F.adaptive_avg_pool2d with divisible sizes (stem_utils.py:37), NHWC fp16, mean over
 *   factor x factor blocks: in [n][h_out*f][w_out*f][c] -> out [n][h_out][w_out][c].
 * qmap_pool: the hyper-encoder's pooled quality map (:563): NCHW fp32 [n][1][h_out*f][w_out*f] -> NHWC fp16
 *   [n][h_out][w_out][8], channel 0 = block mean, channels 1..7 = 0 (so it can be a K segment of the next conv). */
int stemb200_im2col_k3s1_c4(const float* x_nchw, const float* q_nchw, void* out_rows, int32_t n, int32_t h,
                            int32_t w, void* stream);
int stemb200_im2col_k3s1_c4_u8(const uint8_t* x_nchw, const float* q_nchw, void* out_rows, int32_t n, int32_t h,
                               int32_t w, void* stream); /* 8-bit frame, fp32 quality map */
int stemb200_avgpool_nhwc_f16(const void* in, void* out, int32_t n, int32_t h_out, int32_t w_out, int32_t c,
                              int32_t factor, void* stream);
/* fp16 -> fp32 element cast (numel % 8 == 0): latents produced by an fp16 epilogue that feed the entropy kernels */
int stemb200_cast_f16_to_f32(const void* in, float* out, int64_t numel, void* stream);
int stemb200_qmap_pool(const float* q_nchw, void* out_nhwc8_f16, int32_t n, int32_t h_out, int32_t w_out,
                       int32_t factor, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Entropy-model elementwise kernels
 * ------------------------------------------------------------------------------------------------- */
/* P-frame latent staging (spatiotemporalpriors.py:570, :852-868): y NHWC fp32 ->
 *   y_f16    = fp16(y)                     (HE input, first half of the cat)
 *   yq_f16   = fp16(round(y - sub))        (context_prediction input), sub = cond (fp16 NHWC) or 0 when NULL
 *   yhat_f16 = fp16(round(y - sub) + sub)  (y_hat: next frame's y_conditioned, g_s input)
 * any output may be NULL. */
int stemb200_latent_stage(const float* y_nhwc, const void* cond_f16, void* y_f16, void* yq_f16, void* yhat_f16,
                          int64_t numel, void* stream);

/* GaussianConditional forward + build_indexes + symbols + bit count in one pass
 * (entropy_models.py:588-604, :122-150, :570-586; bound_ops.py:50-53; evalSTEM.py:133-136).
 *   inputs : y fp32 with C channels, NHWC or (y_is_nchw) NCHW; cond fp16 NHWC or NULL: when given the coded
 *            quantity is y - cond (_Res, spatiotemporalpriors.py:852); params NHWC fp32 = EPM output
 *            [pixels][2*C], scales = ch [0,C), means = ch [C,2C) (chunk(2,1), spatiotemporalpriors.py:577).
 *   yhat_mode 0: y_hat = round(y - mu) + mu (GaussianConditional output, WithoutSPM variants :188);
 *             1: y_hat = round(y - cond) + cond (SPM variants return the mean-free rounding, :570,:856-868).
 *   outputs (any may be NULL): y_hat, lik NCHW fp32; idx, sym NCHW int32;
 *            bits[frame] (double) += sum(-log2(lik)) over the frame.
 *   scale_table: fp32 [n_scales] on device (may be NULL when idx == NULL).
 * y_hat = round(y - mu) + mu; lik evaluated at |y_hat - mu| with sigma = max(sigma, scale_bound);
 * lik floored at lik_bound; idx = (n_scales-1) - #{k < n_scales-1 : sigma <= table[k]}. */
int stemb200_gaussian_conditional_fwd(const float* y, int32_t y_is_nchw, const void* cond_f16,
                                      const float* params_nhwc, int32_t n, int32_t c, int32_t h, int32_t w,
                                      const float* scale_table, int32_t n_scales, float scale_bound,
                                      float lik_bound, int32_t yhat_mode, float* y_hat_nchw, float* lik_nchw,
                                      int32_t* idx_nchw, int32_t* sym_nchw, double* bits, void* stream);
/* Same for the drop-in forward of the _Res variant (spatiotemporalpriors.py:845-868) with API tensors: y and the
 * conditioning latent both NCHW fp32; the coded quantity is y - cond in fp32 and y_hat (mode 1) = round(y - cond) + cond
 * in fp32, exactly the reference's arithmetic whatever values y_conditioned holds. */
int stemb200_gaussian_conditional_fwd_cond32(const float* y_nchw, const float* cond_f32_nchw, const float* params_nhwc,
                                             int32_t n, int32_t c, int32_t h, int32_t w, const float* scale_table,
                                             int32_t n_scales, float scale_bound, float lik_bound, int32_t yhat_mode,
                                             float* y_hat_nchw, float* lik_nchw, int32_t* idx_nchw, int32_t* sym_nchw,
                                             double* bits, void* stream);
/* entropy_parameters' last 1x1 layer and GaussianConditional in ONE kernel (BASELINE north_star item 2; reference:
 * spatiotemporalpriors.py:577-579 = EPM[4] -> chunk -> gaussian_conditional): (sigma | mu) stay in TMEM / registers and
 * never reach HBM. d: the 1x1 layer (c_in = 576, c_out = 2 C, no activation); packed_weight / bias: the layer's rows
 * INTERLEAVED per 64 channels - rows [128 t, 128 t + 64) = sigma of channels [64 t, 64 t + 64), rows [128 t + 64,
 * 128 t + 128) = mu of the same channels (pack them with stemb200_conv2d_pack_weight after reordering). y: NHWC fp32
 * [batch][h][w][C]; cond_f16 / yhat_mode / outputs / bits as in stemb200_gaussian_conditional_fwd (no indexes /
 * symbols: the forward pass). Bit-identical to stemb200_conv2d_fwd (fp32 out) + stemb200_gaussian_conditional_fwd. */
int stemb200_conv2d_gc_fwd(const stemb200_conv_desc* d, const void* const* in, const void* packed_weight,
                           const float* bias, const float* y_nhwc, const void* cond_f16, float scale_bound,
                           float lik_bound, int32_t yhat_mode, float* y_hat_nchw, float* lik_nchw, double* bits,
                           void* stream);
/* Same arithmetic on flat arrays (no layout change); the isolated parity test of a9 runs through this. */
int stemb200_gaussian_conditional_flat(const float* y, const float* scales, const float* means,
                                       int64_t numel, const float* scale_table, int32_t n_scales,
                                       float scale_bound, float lik_bound, float* y_hat, float* lik,
                                       int32_t* idx, int32_t* sym, double* bits, void* stream);

/* EntropyBottleneck forward, eval mode (entropy_models.py:424-452, :388-422): z NHWC fp32 [n][h][w][c];
 * params: fp32 [c][59] = softplus(M0)(3) b0(3) tanh(f0)(3) | softplus(M1)(9) b1(3) tanh(f1)(3) | M2.. | M3.. |
 * softplus(M4)(3) b4(1) | median(1)  (folded once per load_state_dict).
 * Outputs: z_hat NHWC fp16 (HD input), z_hat / lik NCHW fp32 (API), bits[frame] += sum(-log2 lik). */
int stemb200_entropy_bottleneck_fwd(const float* z_nhwc, const float* params, int32_t n, int32_t c,
                                    int32_t h, int32_t w, float lik_bound, void* z_hat_nhwc_f16,
                                    float* z_hat_nchw, float* lik_nchw, double* bits, void* stream);

/* Last synthesis layer, second half (priors.py:438 deconv(N, 3) + :399 clamp; evalSTEM.py:29-31,127-129 crop +
 * MSE). The transposed conv itself runs through stemb200_conv2d_fwd as a stride-2 conv over 2x2 input super pixels
 * (k5 with the r = 0 / s = 0 taps masked = a 4x4 window, 48 (+16 pad) outputs: channel (u*4+v)*3 + c holds
 * x_hat[c][4i+u][4j+v]); this entry point un-shuffles in: NHWC fp32 [n][h4][w4][64] to x_hat NCHW fp32
 * [n][3][4*h4][4*w4], clamped to [0,1] when clamp01 != 0 (stem_roi.forward returns it unclamped). When x_ref != NULL (unpadded NCHW fp32 [n][3][h_ref][w_ref], embedded at
 * pad_top/pad_left) sq_err[frame] (double) += sum (x_ref - x_hat)^2 over the un-padded area. */
int stemb200_synthesis_tail(const float* in_nhwc64, float* x_hat_nchw, int32_t n, int32_t h4, int32_t w4,
                            const float* x_ref, int32_t h_ref, int32_t w_ref, int32_t pad_top, int32_t pad_left,
                            double* sq_err, int32_t clamp01, void* stream);
int stemb200_synthesis_tail_u8(const float* in_nhwc64, float* x_hat_nchw, int32_t n, int32_t h4, int32_t w4,
                               const uint8_t* x_ref, int32_t h_ref, int32_t w_ref, int32_t pad_top, int32_t pad_left,
                               double* sq_err, int32_t clamp01, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Autoregressive coding of y for the variants with a spatial context model (SpatioTemporalPriorModel, _Res,
 * WithoutTPM: spatiotemporalpriors.py:633-678 / :729-768 / :915-961 / :1016-1055) and for the I-frame model
 * (JointAutoregressiveHierarchicalPriors, priors.py:556-600 / :651-684). The reference scans the latent grid in
 * raster order on the CPU; here one persistent 64-CTA kernel (fp32 weights resident in shared memory) runs the same
 * recurrence per latent position
 *     ctx = context_prediction(t_hat)[h, w]            (12 causal taps of the masked 5x5 conv)
 *     g   = L2(lrelu(L1(lrelu(e0[h, w] + L0_ctx . ctx))));  sigma = g[:c], mu = g[c:]
 *     idx = build_indexes(sigma);  sym = round(t - mu);  t_hat[h, w] = sym + mu
 * encode: wavefront order (w + 3h = const), all images of the batch per step; decode: raster order (the rANS stream
 * dictates it) with the rANS state machine inside the kernel. All tensors NHWC fp32 ([batch][h][w][channels]), which is
 * also the (h, w, c) symbol order of the reference's streams. e0 = the y_hat-independent part of the first EPM /
 * entropy_parameters layer (prior columns + bias, no activation), a batched 1x1 conv done by stemb200_conv2d_fwd.
 * ------------------------------------------------------------------------------------------------- */
typedef struct stemb200_ar_desc {
  int32_t batch;    /* images coded in parallel (<= 64) */
  int32_t h, w;     /* latent grid */
  int32_t c;        /* latent channels (multiple of 64, <= 256); context_prediction maps c -> 2c */
  int32_t l1, l2;   /* widths of the two hidden layers of the parameter head (multiples of 64) */
  float slope;      /* LeakyReLU slope of the head (0.01) */
  int32_t n_scales; /* scale table length */
  float scale_bound; /* GaussianConditional.lower_bound_scale (0.11): sigma is clamped to it before the table search
                        (entropy_models.py:598-604), as in stemb200_gaussian_conditional_fwd */
} stemb200_ar_desc;
/* Weights are packed on the host (fp32) as 64 consecutive per-CTA blocks; CTA j owns rows
 *   [j*rc, (j+1)*rc) of context_prediction (rc = 2c/64; each row = 12 taps x c, tap-major, taps in mask order),
 *   then its rc biases; rows [j*r1, ..) of the context columns of layer 0 (r1 = l1/64, 2c columns);
 *   rows [j*r2, ..) of layer 1 (r2 = l2/64, l1 columns) and their biases; rows {j*rg + i} (sigma) then
 *   {c + j*rg + i} (mu), i < rg = c/64, of layer 2 (l2 columns) and the 2 rg biases in the same order.
 * stemb200_ar_packed_floats returns the total float count (64 blocks). */
int64_t stemb200_ar_packed_floats(const stemb200_ar_desc* d);
int64_t stemb200_ar_workspace_bytes(const stemb200_ar_desc* d);
/* Both kernels are persistent cooperative grids whose layers are separated by a grid barrier with a watchdog: if a
 * barrier times out the grid aborts and leaves its outputs undefined. The first three 32-bit words of the workspace
 * tell the caller: [0] barrier counter, [1] abort flag (non-zero = aborted), [2] completion flag (1 = ran to the end);
 * they are zeroed by the launch, so after the stream has drained a caller must see [1] == 0 and [2] == 1. */
/* target: the coded quantity (y, or y - y_conditioned for _Res). Outputs: t_hat, symbols / indexes int32 in stream
 * order, params_out [..][2c] = sigma | mu (may be NULL, symbols may be NULL). */
int stemb200_ar_encode(const stemb200_ar_desc* d, const float* packed, const float* e0, const float* target,
                       const float* scale_table, float* t_hat, int32_t* symbols, int32_t* indexes,
                       float* params_out, void* workspace, void* stream);
/* streams: the batch's rANS byte strings on the device, image b at [stream_off[b], +stream_len[b]) with 4-byte
 * aligned offsets; cdfs / cdf_sizes / offsets: the GaussianConditional tables (int32, device; at most 64 rows);
 * cdf_total_entries = sum of cdf_sizes (the decoder keeps a compact 16-bit copy of the rows in shared memory).
 * status[b] = 0, or 1 when stream b is corrupt. */
int stemb200_ar_decode(const stemb200_ar_desc* d, const float* packed, const float* e0, const float* scale_table,
                       const uint8_t* streams, const int64_t* stream_off, const int64_t* stream_len,
                       const int32_t* cdfs, int32_t n_cdfs, int32_t cdf_stride, const int32_t* cdf_sizes,
                       const int32_t* offsets, int32_t cdf_total_entries, float* t_hat, int32_t* symbols,
                       int32_t* indexes, float* params_out, int32_t* status, void* workspace, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Host-side helper that the reference implements in C++ (compressai/cpp_exts/ops/ops.cpp:24-81)
 * ------------------------------------------------------------------------------------------------- */
/* cdf_out must hold pmf_len + 1 entries. */
int stemb200_pmf_to_quantized_cdf_host(const float* pmf, int32_t pmf_len, int32_t precision,
                                       int32_t* cdf_out);

/* ---------------------------------------------------------------------------------------------------
 * rANS entropy coder (host). Replaces compressai.ans.RansEncoder.encode_with_indexes /
 * RansDecoder.decode_with_indexes (compressai/cpp_exts/rans/rans_interface.cpp:193-275, rans64.h) with flat int32
 * arrays instead of Python lists; the byte stream is identical (little-endian u32 words written backwards, 64-bit
 * state flushed as two words, 16-bit precision, 4-bit bypass nibbles for out-of-range values).
 *   symbols, indexes: [n]; cdfs: [n_cdfs][cdf_stride] (row i valid for cdf_sizes[i] entries); offsets: [n_cdfs].
 * encode returns the number of bytes written to out (<= out_capacity) or a negative error code. */
int64_t stemb200_rans_encode_host(const int32_t* symbols, const int32_t* indexes, int64_t n, const int32_t* cdfs,
                                  int32_t n_cdfs, int32_t cdf_stride, const int32_t* cdf_sizes,
                                  const int32_t* offsets, uint8_t* out, int64_t out_capacity);
int stemb200_rans_decode_host(const uint8_t* stream, int64_t nbytes, const int32_t* indexes, int64_t n,
                              const int32_t* cdfs, int32_t n_cdfs, int32_t cdf_stride, const int32_t* cdf_sizes,
                              const int32_t* offsets, int32_t* symbols_out);

#ifdef __cplusplus
}
#endif
#endif /* STEMB200_H_ */
