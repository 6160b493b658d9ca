/*
 * mmnc_b200 — C ABI of the B200-native rate path for the ScaleHyperprior codecs of
 * narekvslife/multi-modal-neural-compression.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own: its rate path sits behind
 * CompressAI 1.2.4's Python modules and two pybind11 modules.  Every entry point
 * below names the reference call site and the CompressAI interface it replaces.  Conventions:
 *
 *   - extern "C", plain pointers and sizes; no torch / C++ types in any signature;
 *   - every pointer is a DEVICE pointer unless its name ends in `_h`;
 *   - tensors are dense, row-major ("NCHW-contiguous"): (B, C, S) means B images, C channels, S = H*W;
 *   - the caller owns and pre-allocates every buffer; the library keeps no persistent device memory;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, nothing synchronises;
 *   - return value: 0 = ok, negative = MMNC_ERR_* (message via mmnc_last_error()); never throws;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns MMNC_ERR_CUDA.
 *
 * Reference files are cited relative to /root/reference/src/models/ (mtc.py = multi_task_compressor.py).
 */
#ifndef MMNC_B200_H
#define MMNC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMNC_OK 0
#define MMNC_ERR_INVALID (-1)     /* bad argument (shape, mode, null pointer) */
#define MMNC_ERR_CUDA (-2)        /* CUDA runtime error (also: no device) */
#define MMNC_ERR_UNSUPPORTED (-3) /* valid request outside what the kernels cover */

#define MMNC_EB_PARAMS_PER_CHANNEL 58 /* filters (3,3,3,3): 33 matrix + 13 bias + 12 factor */

/* means_mode for the quantise family */
#define MMNC_MEANS_NONE 0
#define MMNC_MEANS_PER_CHANNEL 1 /* means has C entries (EntropyBottleneck medians) */
#define MMNC_MEANS_FULL 2        /* means has the shape of the input */

/* noise_mode for the forward kernels */
#define MMNC_QUANT_DEQUANTIZE 0  /* eval: round_half_even(x - m) + m */
#define MMNC_QUANT_NOISE_PHILOX 1 /* train: x + U(-1/2,1/2) from Philox4x32-10(seed, element index + offset) */
#define MMNC_QUANT_NOISE_GIVEN 2  /* train (test hook): x + noise[i] read from memory */
#define MMNC_QUANT_IDENTITY 3     /* input is already quantised: v = x */
#define MMNC_QUANT_NOISE_PHILOX_DEV 4 /* like PHILOX, but `noise` points at two device uint64 {seed, offset}: the
                                        stream can be advanced by a kernel, so a captured CUDA graph draws fresh
                                        noise on every replay */

/* likelihood_form for the EntropyBottleneck (SURVEY.md A.3 version switch) */
#define MMNC_EB_FORM_SIGN 0  /* CompressAI 1.2.x: |sigmoid(s*up) - sigmoid(s*lo)|, s = -sign(lo+up) */
#define MMNC_EB_FORM_PLAIN 1 /* later releases: sigmoid(up) - sigmoid(lo) */

/* arithmetic the caller accepts for the GDN channel contraction.  Tensor cores (tcgen05) are used when the shape
 * suits the kernel in that arithmetic; all other shapes run on the fp32 SIMT kernel (at least as accurate). */
#define MMNC_GDN_FP32 0  /* fp32 FMA only */
#define MMNC_GDN_TF32 1  /* tcgen05 kind::tf32, single pass (x^2 and gamma rounded to tf32) */
#define MMNC_GDN_3XTF32 2 /* tcgen05 kind::tf32, hi/lo split in three passes (fp32-class accuracy) */
#define MMNC_GDN_AUTO 3  /* = MMNC_GDN_TF32: what the reference's own GDN (F.conv2d under cuDNN's default TF32) does on a GPU */

/* ---------------------------------------------------------------------------------------------------------
 * Library state
 * ------------------------------------------------------------------------------------------------------- */
int mmnc_version(void);
const char *mmnc_last_error(void);
/* number of kernels this library has launched in this process (bench.py's `gpu_launches`) */
uint64_t mmnc_launch_count(void);
int mmnc_device_sm_count(void);

/* ---------------------------------------------------------------------------------------------------------
 * (a2) EntropyModel.quantize — CompressAI entropy_models.EntropyModel.quantize(inputs, mode, means),
 *      reached from EB/GC forward and compress (mtc.py:495, 509).
 * ------------------------------------------------------------------------------------------------------- */
int mmnc_quantize_noise(const float *x, int64_t n, int noise_mode, const float *noise, uint64_t seed,
                        uint64_t offset, float *out, void *stream);
int mmnc_quantize_dequantize(const float *x, int64_t B, int64_t C, int64_t S, const float *means, int means_mode,
                             float *out, void *stream);
int mmnc_quantize_symbols(const float *x, int64_t B, int64_t C, int64_t S, const float *means, int means_mode,
                          int32_t *symbols, void *stream);
/* EntropyModel.dequantize(symbols, means): float(symbols) + means */
int mmnc_dequantize_symbols(const int32_t *symbols, int64_t B, int64_t C, int64_t S, const float *means,
                            int means_mode, float *out, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * (a3) EntropyBottleneck.forward / _likelihood / _logits_cumulative (+ LowerBound floor) — mtc.py:495.
 *   x, out, lik : (B, C, S).  params : (C, 58) RAW parameters packed per channel in the order
 *     matrix0[3x1] bias0[3] factor0[3] | matrix1[3x3] bias1[3] factor1[3] | matrix2 .. | matrix3 .. |
 *     matrix4[1x3] bias4[1]          (row-major matrices; softplus / tanh are applied inside the kernel).
 *   medians : (C).  lnsum : (C) fp32, ACCUMULATED (+=) with sum over (b, s) of ln(lik) — caller zeroes; may be
 *   NULL.  noise / seed / offset per noise_mode.  likelihood_bound <= 0 disables the floor.
 * backward: given saved `out`, upstream g_out / g_lik (either may be NULL) and g_lnsum (C, may be NULL; the
 *   gradient of a loss that consumed lnsum), writes g_x (B, C, S) and ACCUMULATES g_params (C, 58) (caller zeroes).
 * ------------------------------------------------------------------------------------------------------- */
int mmnc_eb_forward(const float *x, int64_t B, int64_t C, int64_t S, const float *params, const float *medians,
                    int noise_mode, const float *noise, uint64_t seed, uint64_t offset, float likelihood_bound,
                    int likelihood_form, float *out, float *lik, float *lnsum, void *stream);
int mmnc_eb_backward(const float *out, int64_t B, int64_t C, int64_t S, const float *params, const float *g_out,
                     const float *g_lik, const float *g_lnsum, float likelihood_bound, int likelihood_form,
                     float *g_x, float *g_params, void *stream);
/* _logits_cumulative on v: (C, L) -> logits (C, L), parameters detached (EntropyBottleneck.update / loss) */
int mmnc_eb_logits(const float *v, int64_t C, int64_t L, const float *params, float *logits, void *stream);
/* (a4) EntropyBottleneck.loss() — mtc.py:386-387: loss = sum |F_c(quantiles) - target|; writes the scalar
 * loss (ACCUMULATED, caller zeroes) and d loss / d quantiles (C, 3) in one launch (parameters detached). */
int mmnc_eb_aux_loss(const float *quantiles, int64_t C, const float *params, const float *target3, float *loss,
                     float *g_quantiles, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * (a5) GaussianConditional.forward / _likelihood (+ LowerBound on scales and on the likelihood) — mtc.py:495.
 *   y : (B, C, Sy); scales, lik : (B, C, Ss) with Sy == Ss, or Sy == 1 (torch broadcasting, the shape the
 *   reference actually produces: y (B,M,1,1) against scales (B,M,4,4)).
 *   means : NULL or (B, C, Sy).  y_hat : (B, C, Sy).  lnsum : (C) accumulated sum of ln(lik), may be NULL.
 * backward writes g_y (B, C, Sy) and g_scales (B, C, Ss); g_yhat / g_lik / g_lnsum may each be NULL.
 * ------------------------------------------------------------------------------------------------------- */
int mmnc_gc_forward(const float *y, const float *scales, const float *means, int64_t B, int64_t C, int64_t Sy,
                    int64_t Ss, int noise_mode, const float *noise, uint64_t seed, uint64_t offset,
                    float scale_bound, float likelihood_bound, float *y_hat, float *lik, float *lnsum,
                    void *stream);
int mmnc_gc_backward(const float *y_hat, const float *scales, const float *means, int64_t B, int64_t C,
                     int64_t Sy, int64_t Ss, const float *g_yhat, const float *g_lik, const float *g_lnsum,
                     float scale_bound, float likelihood_bound, float *g_y, float *g_scales, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * (a6, a13) per-channel sum of ln(likelihood) for a likelihood tensor that did not come with one
 *   (MultiTaskCompressor._bits_per_pixel, mtc.py:278-293).  lik (B, C, S) -> lnsum (C), accumulated.
 *   backward: g_lik = g_lnsum[c] / lik.
 * ------------------------------------------------------------------------------------------------------- */
int mmnc_lnsum_forward(const float *lik, int64_t B, int64_t C, int64_t S, float *lnsum, void *stream);
int mmnc_lnsum_backward(const float *lik, int64_t B, int64_t C, int64_t S, const float *g_lnsum, float *g_lik,
                        void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * (a7) distortion terms of the multi-task RD loss — MultiTaskCompressor.reconstruction_loss, mtc.py:223-255.
 *   kind 0: sum (a-b)^2, kind 1: sum |a-b|.  `out` (one float) is ACCUMULATED with scale * sum.
 *   backward: g_a = (*g_scalar) * scale * d/da, with g_scalar a device scalar (no host sync).
 * ------------------------------------------------------------------------------------------------------- */
int mmnc_distortion_forward(const float *a, const float *b, int64_t n, int kind, float scale, float *out,
                            void *stream);
int mmnc_distortion_backward(const float *a, const float *b, int64_t n, int kind, float scale,
                             const float *g_scalar, float *g_a, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * (a6, a7) RD-loss epilogue — multitask_compression_loss (mtc.py:302-357, mixed_latent.py:70-118,
 *   shared_latent.py:118-147), multitask_reconstruction_loss (mtc.py:257-276), UncertaintyWeightingStrategy
 *   (loss_balancing.py:31-54), loss = lmbda * rec + comp (mtc.py:437).  One launch, one block.
 *   lnsum_y (M), lnsum_z (N): per-channel sums of ln(lik).  group_of_channel (M) int32: y channel -> rate group
 *   index in [0, n_groups) or -1 for channels that carry no rate term (Disjoint's orphaned channels).
 *   group_inv_pixels (n_groups): 1 / (B*H*W) of the task the group is normalised by; group_weight (n_groups):
 *   weight of the group's bpp in the total (1/T).  z_inv_pixels, z_weight likewise for z.
 *   task_losses (T): raw distortion per task; log_vars (T) or NULL (no weighting, SingleTask).
 *   outputs: scalars[0]=loss, [1]=rec, [2]=comp, [3]=z_bpp, [4..4+n_groups)=group bpp, then T weighted task
 *   losses.  Gradients of `loss`: g_lnsum_y (M), g_lnsum_z (N), g_task_losses (T), g_log_vars (T).
 * ------------------------------------------------------------------------------------------------------- */
int mmnc_rd_epilogue(const float *lnsum_y, int64_t M, const float *lnsum_z, int64_t N,
                     const int32_t *group_of_channel, int n_groups, const float *group_inv_pixels,
                     const float *group_weight, float z_inv_pixels, float z_weight, const float *task_losses,
                     int T, const float *log_vars, float lmbda, float *scalars, float *g_lnsum_y,
                     float *g_lnsum_z, float *g_task_losses, float *g_log_vars, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * (a8) GDN / IGDN — compressai.layers.GDN.forward, sites mtc.py:146-172, disjoint_latent.py:150-154 and the
 *   backbone's g_a / g_s.  y_i = x_i * (beta_i + sum_j gamma_ij x_j^2)^(-1/2) (GDN) or ^(+1/2) (inverse).
 *   x, y, g, dx : (B, C, HW).  beta (C), gamma (C, C): EFFECTIVE (already re-parametrised) values.
 *   backward: dx, and d beta (C) / d gamma (C, C) w.r.t. the effective values (OVERWRITTEN, not accumulated).
 *   workspace: device scratch of at least mmnc_gdn_backward_workspace_bytes(B, C, HW, precision) bytes.
 *   precision: MMNC_GDN_*.  The reparametrisation (NonNegativeParametrizer + LowerBound) is below.
 * ------------------------------------------------------------------------------------------------------- */
int mmnc_gdn_forward(const float *x, int64_t B, int64_t C, int64_t HW, const float *beta, const float *gamma,
                     int inverse, int precision, float *y, void *stream);
size_t mmnc_gdn_backward_workspace_bytes(int64_t B, int64_t C, int64_t HW, int precision);
int mmnc_gdn_backward(const float *x, const float *g, int64_t B, int64_t C, int64_t HW, const float *beta,
                      const float *gamma, int inverse, int precision, float *dx, float *dbeta, float *dgamma,
                      void *workspace, size_t workspace_bytes, void *stream);
/* Which kernel family mmnc_gdn_backward would launch for these arguments (diagnostics / tests):
 * 0 = streaming (C <= 4), 1 = fp32 SIMT, 2 = fused tcgen05, 3 = fused tcgen05 fed by TMA and software-pipelined
 * (needs HW % 128 == 0 and 16-byte aligned x / g), 4 = the same kernel with the gamma operand streamed through one
 * shared-memory buffer (112 <= C <= 128), 5 = streamed gamma + a second x buffer so that the next tile's x lands while
 * the current one computes (80 <= C <= 111, at least three tiles per CTA), 6 = the wide-layer pair (129 <= C <= 256,
 * HW % 32 == 0): a dx kernel with BOTH gamma operands streamed in K chunks through a shared-memory ring, u written to
 * the workspace, and a split-K tcgen05 GEMM over pixels for d gamma / d beta. */
int mmnc_gdn_backward_variant(const float *x, const float *g, int64_t B, int64_t C, int64_t HW, int precision);
/* Same for mmnc_gdn_forward: 0 = streaming, 1 = fp32 SIMT, 2 = tcgen05 (per-thread global loads), 3 = tcgen05 with
 * TMA in / TMA out, 6 = the wide-layer kernel (129 <= C <= 256, gamma streamed in K chunks; needs the workspace of the
 * _ws entry points below). */
int mmnc_gdn_forward_variant(const float *x, const float *y, int64_t B, int64_t C, int64_t HW, int precision);
/* Forward with a scratch buffer.  Layers of 129 .. 256 channels (IGDN(256) of config C3) run on the tensor cores only
 * through these: gamma no longer fits shared memory, so it is packed once per call into `workspace`
 * (mmnc_gdn_forward_workspace_bytes, 0 when the shape needs none) and streamed from there.  With workspace = NULL they
 * behave exactly like the calls without the suffix (fp32 SIMT kernel for such layers). */
size_t mmnc_gdn_forward_workspace_bytes(int64_t B, int64_t C, int64_t HW, int precision);
int mmnc_gdn_forward_ws(const float *x, int64_t B, int64_t C, int64_t HW, const float *beta, const float *gamma,
                        int inverse, int precision, float *y, void *workspace, size_t workspace_bytes, void *stream);
int mmnc_gdn_forward_raw_ws(const float *x, int64_t B, int64_t C, int64_t HW, const float *beta_raw,
                            const float *gamma_raw, float beta_bound, float gamma_bound, float pedestal, int inverse,
                            int precision, float *y, void *workspace, size_t workspace_bytes, void *stream);
/* Same two calls taking the RAW parameters of compressai.layers.GDN (`beta`, `gamma` as stored in the state dict):
 * the NonNegativeParametrizer re-parametrisation (effective = max(p, bound)^2 - pedestal) is applied while the
 * kernels stage the parameters, and the gradients come back w.r.t. the raw parameters with LowerBound's custom
 * gradient applied — one launch forward, two backward, instead of 3 + 5. */
int mmnc_gdn_forward_raw(const float *x, int64_t B, int64_t C, int64_t HW, const float *beta_raw,
                         const float *gamma_raw, float beta_bound, float gamma_bound, float pedestal, int inverse,
                         int precision, float *y, void *stream);
int mmnc_gdn_backward_raw(const float *x, const float *g, int64_t B, int64_t C, int64_t HW, const float *beta_raw,
                          const float *gamma_raw, float beta_bound, float gamma_bound, float pedestal, int inverse,
                          int precision, float *dx, float *dbeta_raw, float *dgamma_raw, void *workspace,
                          size_t workspace_bytes, void *stream);
/* (f1) Channels-last: the same two calls for tensors stored NHWC (torch.channels_last: element (b, c, p) at
 * ((b HW + p) C + c)), so that a model whose convolutions run channels-last on cuDNN's tensor-core kernels does not pay
 * NCHW <-> NHWC conversions around every GDN.  Native for C <= 4 and for 16 <= C <= 128 forward / 16 <= C <= 111
 * backward at single-pass TF32 (auto / tf32); mmnc_gdn_nhwc_supported says whether a call would be accepted. */
int mmnc_gdn_nhwc_supported(int64_t B, int64_t C, int64_t HW, int precision, int backward);
int mmnc_gdn_forward_raw_nhwc(const float *x, int64_t B, int64_t C, int64_t HW, const float *beta_raw,
                              const float *gamma_raw, float beta_bound, float gamma_bound, float pedestal, int inverse,
                              int precision, float *y, void *stream);
int mmnc_gdn_backward_raw_nhwc(const float *x, const float *g, int64_t B, int64_t C, int64_t HW, const float *beta_raw,
                               const float *gamma_raw, float beta_bound, float gamma_bound, float pedestal, int inverse,
                               int precision, float *dx, float *dbeta_raw, float *dgamma_raw, void *workspace,
                               size_t workspace_bytes, void *stream);
/* NonNegativeParametrizer.forward: out = max(p, bound)^2 - pedestal, and its backward with LowerBound's
 * custom gradient (pass when p >= bound or when the incoming gradient is negative). */
int mmnc_nonneg_reparam_forward(const float *p, int64_t n, float bound, float pedestal, float *out, void *stream);
int mmnc_nonneg_reparam_backward(const float *p, const float *g_out, int64_t n, float bound, float *g_p,
                                 void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * (f2) step metrics — `average_metrics`, mtc.py:359-384.  PSNR of the regression tasks comes from the distortion
 *   kernel (PSNR = -10 log10(MSE) for images scaled by 255 with data_range 255); the semantic task's metrics are
 *   computed on the argmax class-id image: logits (B, K, S) -> labels (B, 1, S) as floats (may be NULL) and
 *   *sse += sum (argmax - target)^2 (target (B, 1, S) class ids as floats; sse zero-initialised by the caller; both
 *   may be NULL when only the labels are wanted).  Ties: first maximum, like torch.argmax.
 * ------------------------------------------------------------------------------------------------------- */
int mmnc_argmax_sse(const float *logits, const float *target, int64_t B, int K, int64_t S, float *labels,
                    float *sse, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * (f2) MS-SSIM — pytorch_msssim.ms_ssim as `average_metrics` calls it on every task (mtc.py:359-384): one SCALE per call.
 *   x, y: (planes, H, W) fp32 (planes = B * C), both multiplied by `scale` on load (the reference scales by 255 first);
 *   'valid' separable 11-tap Gaussian (sigma; 1.5 in the reference) over x, y, x^2, y^2, x y;
 *   cs = (2 s12 + c2) / (s1 + s2 + c2), ssim = (2 mu1 mu2 + c1) / (mu1^2 + mu2^2 + c1) * cs;
 *   ssim_mean, cs_mean (planes): spatial means over the (H - 10) x (W - 10) valid outputs (fixed-order reduction).
 *   px, py (planes, H / 2, W / 2) or both NULL: the 2 x 2 average-pooled SCALED planes = the next scale's input
 *   (H and W even; pass scale = 1 for them).  workspace: mmnc_ssim_workspace_floats(planes, H, W) floats, 8-byte aligned.
 * ------------------------------------------------------------------------------------------------------- */
size_t mmnc_ssim_workspace_floats(int64_t planes, int H, int W);
int mmnc_ssim_scale(const float *x, const float *y, int64_t planes, int H, int W, float scale, float c1, float c2,
                    float sigma, float *workspace, float *ssim_mean, float *cs_mean, float *px, float *py, void *stream);
/* All five scales of ms_ssim in one call (H, W multiples of 16): means [5][2][planes] = per scale the SSIM means, then
 * the contrast means; the caller combines relu(cs_0..3) and relu(ssim_4) with the published exponents.
 * workspace: mmnc_ms_ssim_workspace_floats(planes, H, W) floats (tile partials + the pooled planes of scales 1-4). */
size_t mmnc_ms_ssim_workspace_floats(int64_t planes, int H, int W);
int mmnc_ms_ssim(const float *x, const float *y, int64_t planes, int H, int W, float scale, float c1, float c2,
                 float sigma, float *workspace, float *means, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * (f1) bias gradient of the convolutions: out[c] = sum over (b, s) of g[b, c, s] for an NCHW tensor (B, C, S).
 *   workspace: mmnc_channel_sum_workspace_floats(B, C, S) floats.  Fixed-order two-stage reduction (bit-reproducible).
 * ------------------------------------------------------------------------------------------------------- */
int64_t mmnc_channel_sum_workspace_floats(int64_t B, int64_t C, int64_t S);
int mmnc_channel_sum(const float *g, int64_t B, int64_t C, int64_t S, float *workspace, float *out, void *stream);
/* x[b, c, s] += bias[c] in place on an NCHW tensor (the add torch issues after a bias-free cuDNN convolution). */
int mmnc_bias_add(float *x, const float *bias, int64_t B, int64_t C, int64_t S, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * (f4) input pipeline — what `get_transform` does per sample on the host (src/datasets/transforms.py:39-131,
 *   src/datasets/clevr.py:48-83), as three kernels over a whole batch of raw decoded pixels:
 *   u8 HWC (B, HW, src_channels) -> f32 (B, dst_channels, HW), value / divisor (255 for ToTensor);
 *   u16 (n) -> f32 (n), value / divisor (2^15 - 1 for the 16-bit depth maps);
 *   labels: channel `channel` of u8 HWC pixels through a 256-entry float table (class remapping of `semantic`).
 * ------------------------------------------------------------------------------------------------------- */
int mmnc_prep_u8_hwc_to_f32_chw(const uint8_t *src, int64_t B, int64_t HW, int src_channels, int dst_channels,
                                float divisor, float *dst, void *stream);
int mmnc_prep_u16_to_f32(const uint16_t *src, int64_t n, float divisor, float *dst, void *stream);
int mmnc_prep_labels(const uint8_t *src, int64_t n_pixels, int src_channels, int channel, const float *lut256,
                     float *dst, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * (a9) compressai._CXX.pmf_to_quantized_cdf(pmf: List[float], precision) -> List[int]  — HOST function
 *   (called once per table row from EntropyBottleneck.update / GaussianConditional.update, reference call site
 *   mtc.py:486-489, off the per-step path).  pmf_h: n floats; cdf_h: n + 1 uint32, strictly increasing,
 *   cdf_h[0] = 0, cdf_h[n] = 2^precision.  Integer arithmetic identical to CompressAI's (bit-exact tables).
 * ------------------------------------------------------------------------------------------------------- */
int mmnc_pmf_to_quantized_cdf_h(const float *pmf_h, int n, int precision, uint32_t *cdf_h);

/* ---------------------------------------------------------------------------------------------------------
 * (a10) GaussianConditional.build_indexes — mtc.py:545 and inside ScaleHyperprior.compress (mtc.py:509):
 *   idx = (len-1) - #{ t in table[:-1] : max(scale, bound) <= t }.
 * ------------------------------------------------------------------------------------------------------- */
int mmnc_build_indexes(const float *scales, int64_t n, const float *scale_table, int table_len, float scale_bound,
                       int32_t *indexes, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * (a11, a12) batched rANS — replaces compressai.ans.RansEncoder.encode_with_indexes /
 *   RansDecoder.decode_with_indexes (pybind11, one stream per call) reached from mtc.py:509, 543, 546.
 *   Bit-exact with CompressAI's format: one 64-bit-state stream per image, 32-bit words, 16-bit probabilities,
 *   4-bit bypass escape, native-endian bytes.  One stream per row of `symbols`.
 *
 *   Tables: mmnc_rans_pack_tables turns `_quantized_cdf` (n_cdfs, cdf_stride) int32 + `_cdf_length` into a RAGGED
 *   uint16 table (row r = entries [row_start[r], row_start[r+1]) of `ragged`, values modulo 2^16: the final 65536 of
 *   a row is stored as 0) that fits shared memory: 53 KB instead of 802 KB for the 64 x 3133 Gaussian table.  Call it
 *   once per update(); ragged_capacity >= sum(cdf_sizes) rounded up to 8, 16-byte aligned; row_start has n_cdfs + 1
 *   entries.
 *
 *   symbols (n_streams, n_sym) int32; indexes same shape, or NULL with channel_period > 0 meaning
 *   index = (position / channel_period) % n_cdfs (EntropyBottleneck: channel id).
 *   cdf_sizes = `_cdf_length`; offsets = `_offset`.
 *   encode: `staging` is scratch of n_streams*n_sym*12 bytes (16-byte aligned); `slabs` (n_streams, slab_words)
 *   uint32 with slab_words >= mmnc_rans_slab_words(n_sym); nbytes (n_streams) int32 receives each stream's byte count
 *   (a negative value flags a malformed input for that stream).  The stream's bytes are the LAST nbytes[i]
 *   bytes of its slab.  mmnc_rans_compact packs them back to back: meta = int64 offsets[n_streams + 1] (exclusive
 *   scan) followed by int32 nbytes[n_streams] (so one device-to-host copy fetches both), packed = concatenation.
 * ------------------------------------------------------------------------------------------------------- */
int64_t mmnc_rans_slab_words(int64_t n_sym);
int mmnc_rans_pack_tables(const int32_t *cdf, const int32_t *cdf_sizes, int n_cdfs, int cdf_stride,
                          int32_t *row_start, uint16_t *ragged, int64_t ragged_capacity, void *stream);
int mmnc_rans_encode_batch(const int32_t *symbols, const int32_t *indexes, int64_t channel_period,
                           int64_t n_streams, int64_t n_sym, const uint16_t *ragged_cdf, int64_t ragged_len,
                           const int32_t *row_start, const int32_t *cdf_sizes, const int32_t *offsets, int n_cdfs,
                           void *staging, uint32_t *slabs, int64_t slab_words, int32_t *nbytes, void *stream);
int mmnc_rans_compact(const uint32_t *slabs, int64_t slab_words, const int32_t *nbytes, int64_t n_streams,
                      int64_t *meta, uint8_t *packed, int64_t packed_capacity, void *stream);
/* decode: stream i = lengths[i] bytes at packed + offsets[i]; every offset a multiple of 4 (pad between streams), a
 * length that is not a whole number of 32-bit words is reported as corrupt.  status (n_streams) int32: 0 ok,
 * negative = stream overrun / malformed. */
int mmnc_rans_decode_batch(const uint8_t *packed, const int64_t *offsets, const int32_t *lengths,
                           const int32_t *indexes, int64_t channel_period, int64_t n_streams, int64_t n_sym,
                           const uint16_t *ragged_cdf, int64_t ragged_len, const int32_t *row_start,
                           const int32_t *cdf_sizes, const int32_t *cdf_offsets, int n_cdfs, int32_t *symbols,
                           int32_t *status, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MMNC_B200_H */
