/* pcdm_b200.h — C ABI of libpcdm_b200.so: the sm_100a (B200) kernels behind the PCDMs stage-2 denoising hot path.
 *
 * The reference (tencent-ailab/PCDMs) has NO native/FFI layer of its own: on this path every operator is a PyTorch
 * library call made from diffusers 0.24.0 modules that the reference instantiates
 * (src/models/stage2_inpaint_unet_2d_condition.py:321-343,348-361,407-429) and from its pipeline loop
 * (src/pipelines/stage2_inpaint_pipeline.py:496-525).  Each entry point below therefore cites the torch/diffusers call
 * it replaces.  The binding a maintainer would add on the reference side is a ctypes stub.
 *
 * Conventions
 *   - plain C types only: device pointers as void* / float*, sizes as int / long long, the CUDA stream as void*
 *     (a cudaStream_t; NULL = legacy default stream).  No torch types cross this boundary.
 *   - activations are NHWC ("channels last"), i.e. [B, H, W, C] == row-major [B*H*W, C]; 16-bit (dtype 0 = fp16,
 *     1 = bf16); reductions/accumulators are fp32.  bias / scale vectors are fp32.
 *   - every function returns 0 on success or a negative PCDM_ERR_* code; pcdm_last_error() returns a thread-local
 *     message.  Nothing allocates, nothing synchronises; the caller owns all buffers.  There is no CPU fallback:
 *     without a CUDA device every compute entry point fails with PCDM_ERR_CUDA.
 *   - the library holds NO mutable process-wide state: scratch memory comes in per call (pcdm_ext, workspace
 *     arguments), so calls on different streams / devices / host threads are independent as long as they are given
 *     different scratch buffers.  (Tuning / experiment setters exist only in the separate experiment build,
 *     include/pcdm_b200_experiment.h.)
 */
#ifndef PCDM_B200_H_
#define PCDM_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define PCDM_ABI_VERSION 3

#define PCDM_ERR_INVALID (-1)     /* bad argument */
#define PCDM_ERR_CUDA (-2)        /* CUDA runtime / driver failure */
#define PCDM_ERR_UNSUPPORTED (-3) /* shape outside what the kernels implement */

#define PCDM_DT_F16 0
#define PCDM_DT_BF16 1

#define PCDM_FLAG_GEGLU 1   /* gemm: weight rows packed [32 value | 32 gate]; writes N/2 columns value*gelu(gate);
                             * together with PCDM_FLAG_SILU the gate activation is SiLU (SwiGLU) */
#define PCDM_FLAG_OUT_F32 2 /* gemm/conv: write fp32 instead of the 16-bit dtype */
#define PCDM_FLAG_SILU 4    /* gemm/conv epilogue and norm kernels: apply SiLU last */
#define PCDM_FLAG_GELU 8    /* gemm/conv epilogue: apply GELU (erf form) last */
#define PCDM_FLAG_PAD_BR 16 /* conv3x3 stride 2: zero padding on the bottom/right only (F.pad(x,(0,1,0,1)) + conv pad 0) */
#define PCDM_FLAG_W_STATIC 32 /* gemm: W holds model weights, i.e. it is NOT written by any kernel of the current chain of
                               * programmatically-dependent launches on this stream (with dependent launch, a kernel's
                               * prologue may run while its predecessor AND that one's predecessor are still executing) —
                               * lets the M <= 32 weight-streaming kernel request W before its dependency wait.  Leave
                               * it clear when W is an activation (QK^T / PV written as GEMMs). */

#define PCDM_FLAG_NO_SKINNY 64   /* gemm: keep M <= 32 problems on the tcgen05 tile kernel (tests: A/B of the two paths) */
#define PCDM_FLAG_GN_TWO_PASS 64 /* groupnorm: force the statistics + apply kernels (tests) */
#define PCDM_FLAG_GN_ONE_PASS 128 /* groupnorm: force the single register-resident pass whenever the shape allows (tests) */

/* Optional per-call extras of pcdm_gemm / pcdm_conv3x3 / pcdm_ln_gemm (NULL = none).  Plain pointers and sizes; set
 * `size = sizeof(pcdm_ext)` (a library built against a longer struct rejects a shorter one). */
typedef struct pcdm_ext {
  int size;
  int force_cta_group;       /* 0 = automatic; 1 = single-CTA 128-row tiles; 2 = CTA pairs (tcgen05 cta_group::2, 256-row
                              * tiles) wherever the N tile allows */
  void* workspace;           /* caller-owned scratch for split-K (tile-starved shapes: few output tiles, long K — the
                              * 4x8 / 8x16 UNet levels): fp32 partial sums.  NULL disables split-K.  One buffer per
                              * concurrently-used stream; it must outlive every launch (and every captured CUDA graph)
                              * that was given it.  pcdm_gemm_workspace_bytes() bounds what a problem can use; 64 MiB
                              * covers every BASELINE configuration.  16-byte aligned. */
  long long workspace_bytes;
  /* ---- LayerNorm folded around GEMMs (torch.nn.LayerNorm norm1/norm2/norm3 of diffusers' BasicTransformerBlock,
   *      SURVEY.md §8a a8): no separate normalisation pass over the hidden states.
   * producer — the GEMM whose 16-bit output rows the LayerNorm normalises (proj_in, attn*.to_out + residual):
   *   row_stats != NULL: the epilogue also writes, per output row, (sum, sum of squares) of the values it stores, one
   *   fp32 pair per column slice: row_stats[slot][m][2], slot < row_stats_parts.  row_stats_parts is an OUTPUT (set at
   *   call time: 2 per N tile); the buffer must hold row_stats_cap >= 2 * ceil(N / 64) slots of M pairs.
   * consumer — the GEMM that multiplies the normalised rows (to_q/k/v, ff.net.0.proj), given the RAW rows as A:
   *   ln_stats = the producer's buffer, ln_parts = its row_stats_parts.  The caller prepares the weight once:
   *   W'[n,k] = W[n,k] g[k] - mean_k(W[n,:] g) (gamma-scaled, then every ROW centred: sum_k W'[n,k] = 0, so the
   *   row mean of A cancels inside the accumulation) and bias[n] = sum_k W[n,k] beta[k] (+ the layer's own bias).  Then
   *   out[m,n] = rstd[m] * acc[m,n] + bias[n],  rstd over the K input columns (biased variance, ln_eps inside the
   *   square root: torch.nn.LayerNorm) — the epilogue costs what a plain bias add costs.  Works with PCDM_FLAG_GEGLU. ---- */
  float* row_stats;
  int row_stats_cap;
  int row_stats_parts;
  const float* ln_stats;
  int ln_parts;
  float ln_eps;
  /* ---- GroupNorm statistics out of the producing conv / GEMM epilogue (the "conv3x3 + bias + GroupNorm" fusion of the
   *      north star: diffusers ResnetBlock2D norm1/norm2, Transformer2DModel.norm, conv_norm_out).
   *   chan_stats != NULL (pcdm_gemm / pcdm_conv3x3 / pcdm_conv3x3_up2x, 16-bit output, no GEGLU): the epilogue also
   *   writes, for every 32-row slab of the output and every channel, (sum, sum of squares) of the 16-bit values AS
   *   STORED: chan_stats[slab][N][2] fp32, slab = row / 32, ceil(M / 32) slabs (pcdm_conv3x3_up2x: an image's slabs are
   *   [parity plane][32 low-resolution pixels], still contiguous per image).  The rows of a slab must belong to one
   *   image: rows_per_image (pcdm_gemm) / H*W % 32 == 0.  Deterministic (fixed reduction order); a launch that emits
   *   them never takes the split-K route.  pcdm_groupnorm_apply consumes it: GroupNorm then costs one read + one write of the activation
   *   and no pass for the statistics. ---- */
  float* chan_stats;
} pcdm_ext;

int pcdm_abi_version(void);
const char* pcdm_last_error(void);
/* Upper bound of the split-K scratch pcdm_gemm / pcdm_conv3x3 can use for an [M, N] output (conv: M = B*H*W, N = Cout). */
long long pcdm_gemm_workspace_bytes(int M, int N);

/* torch.nn.Linear (+ fused epilogues).  Replaces the nn.Linear calls inside diffusers Transformer2DModel /
 * BasicTransformerBlock / Attention / GEGLU / TimestepEmbedding and the 1x1 conv_shortcut of ResnetBlock2D
 * (SURVEY.md §8a rows a3, a5, a7, a8).
 *   out[M, N] = act( A[M, K] . W[N, K]^T + bias[N] + rowvec[m / rows_per_image, :] + residual[M, N] )
 * A may be given as two K-segments (a: columns [0, k1), a2: columns [k1, K)) — the skip-concat of the up blocks
 * is consumed in place, never materialised.  lda/lda2/ldo/ldr/ld_rowvec are row strides in elements.
 * flags: PCDM_FLAG_GEGLU | PCDM_FLAG_OUT_F32 | PCDM_FLAG_SILU (act = SiLU) | PCDM_FLAG_GELU (act = GELU; else identity).
 * K % 64 == 0, N % 32 == 0.  bn = 0 picks the N tile automatically (64/128/160/256 to force one). */
int pcdm_gemm(const void* a, long long lda, const void* a2, long long lda2, int k1, const void* w, void* out,
              long long ldo, const float* bias, const float* rowvec, long long ld_rowvec, int rows_per_image,
              const void* residual, long long ldr, int M, int N, int K, int dtype, int flags, int bn,
              pcdm_ext* ext, void* stream);

/* torch.nn.LayerNorm(K, eps) followed by torch.nn.Linear, as one call:
 *   out = act( LN(x[M, K]; gamma, beta, eps) . W[N, K]^T + bias + rowvec + residual )
 * (norm1 -> to_q/to_k/to_v and norm3 -> ff.net.0 of the stage-1 prior's BasicTransformerBlocks,
 * src/models/stage1_prior_transformer.py:112-120,283-289; norm_out -> proj_to_clip_embeddings, :286-290).  With M <= 32
 * rows (K <= 2048, M * (K + 8) * 2 <= 100 KB) it is ONE launch: the skinny kernel normalises the rows itself, rounding them to the
 * 16-bit dtype exactly as pcdm_layernorm stores them.  Otherwise pcdm_layernorm -> scratch ([M, K] 16-bit, caller-owned,
 * may be NULL only when the fused path applies) -> pcdm_gemm.  flags as pcdm_gemm without PCDM_FLAG_GEGLU. */
int pcdm_ln_gemm(const void* x, long long ldx, const float* gamma, const float* beta, float eps, void* scratch,
                 const void* w, void* out, long long ldo, const float* bias, const float* rowvec, long long ld_rowvec,
                 int rows_per_image, const void* residual, long long ldr, int M, int N, int K, int dtype, int flags,
                 pcdm_ext* ext, void* stream);

/* torch.nn.Conv2d(Cin, Cout, 3, stride, padding=1) on NHWC activations, implicit GEMM (no im2col buffer).
 * Replaces conv_in, conv1/conv2 of ResnetBlock2D, Downsample2D.conv (stride 2), Upsample2D.conv and conv_out
 * (SURVEY.md §8a a4-a6, a10).
 *   x: [B, stride*H, stride*W, Cin]; out: [B, H, W, Cout]; w_packed: [Cout][3][3][Cin] (tap-major K);
 *   rowvec: fp32, row b at rowvec + b*ld_rowvec, added per image (the resnet's time_emb_proj term);
 *   residual: [B, H, W, Cout].  Cin % 64 == 0, Cout % 32 == 0; W | 128 with (H*W) % 128 == 0 or 128 % (H*W) == 0, or
 *   W % 128 == 0 (the VAE / pose-encoder resolutions).  With stride 2, PCDM_FLAG_PAD_BR selects the asymmetric padding
 *   of the VAE encoder's Downsample2D (diffusers: F.pad(x, (0, 1, 0, 1)) then Conv2d(stride 2, padding 0)). */
int pcdm_conv3x3(const void* x, const void* w_packed, void* out, const float* bias, const float* rowvec,
                 long long ld_rowvec, const void* residual, int B, int H, int W, int Cin, int Cout, int stride,
                 int dtype, int flags, int bn, pcdm_ext* ext, void* stream);

/* diffusers Upsample2D: F.interpolate(scale_factor=2, mode="nearest") followed by Conv2d(Cin, Cout, 3, padding=1), as ONE
 * launch that never materialises the upsampled tensor (reference blocks
 * src/models/stage2_inpaint_unet_2d_condition.py:407-429).  An output pixel (2i + py, 2j + px) only sees source rows
 * {i - 1, i} (py = 0) or {i, i + 1} (py = 1), and columns alike: the layer is four 2x2 convolutions over the
 * LOW-resolution input, one per output parity, with the 3x3 taps pre-summed per parity — 16 Cin instead of 36 Cin MACs per
 * output value.  x: [B, H, W, Cin]; out: [B, 2H, 2W, Cout];
 * w_up: [4][Cout][2][2][Cin] 16-bit, plane p = 2 py + px, tap (ty, tx) = the sum of W[:, :, r, s] over the rows r / columns s
 * that collapse onto source offset (ty - 1 + py, tx - 1 + px)  (pcdms_b200.ops.pack_upsample_conv_weight).
 * Cin % 64 == 0, Cout % 64 == 0; W | 128 and H*W a multiple or divisor of 128 (as pcdm_conv3x3 at H x W), W and H*W
 * multiples or divisors of 32.  flags: PCDM_FLAG_SILU / PCDM_FLAG_GELU. */
int pcdm_conv3x3_up2x(const void* x, const void* w_up, void* out, const float* bias, int B, int H, int W, int Cin,
                      int Cout, int dtype, int flags, int bn, pcdm_ext* ext, void* stream);

/* torch.nn.GroupNorm(groups, C, eps) (+ SiLU with PCDM_FLAG_SILU) over NHWC x = [x1 | x2] (x2 may be NULL; x1 then has
 * C channels).  Replaces norm1/norm2(+nonlinearity) of ResnetBlock2D, Transformer2DModel.norm and conv_norm_out
 * (reference :817-819).  y: [B, HW, C].  workspace: pcdm_groupnorm_workspace_bytes() bytes,
 * zero-initialised once by the caller (it holds device counters the kernels return to zero); results are
 * bit-reproducible run to run (fixed reduction order, no floating-point atomics).  Small activations (<= 16 MB, >= 8
 * channels per group) run as ONE register-resident pass (one read, one write; pieces of an image share statistics
 * through a thread-block cluster); larger ones take a streaming statistics kernel + an apply kernel
 * (PCDM_FLAG_GN_TWO_PASS / PCDM_FLAG_GN_ONE_PASS force either path).  One workspace per concurrently-used stream. */
long long pcdm_groupnorm_workspace_bytes(int B, int groups);
int pcdm_groupnorm(const void* x1, const void* x2, int C1, void* y, const float* gamma, const float* beta, float eps,
                   int B, int HW, int C, int groups, int dtype, int flags, void* workspace, void* stream);

/* GroupNorm (+ SiLU) of x = [x1 | x2] whose statistics were already emitted by the producers of x1 / x2
 * (pcdm_ext.chan_stats): y = (x - mean[b, g]) * rstd[b, g] * gamma + beta, one streaming read + write; each CTA first
 * folds the per-slab partial sums of its channel groups (fp64, fixed order).  stats1 / stats2: [B * HW / 32][C1 or C - C1][2]
 * fp32 (stats2 NULL when x2 is NULL).  HW % 32 == 0, C % groups == 0, C % 8 == 0, C1 % 8 == 0.  workspace: as
 * pcdm_groupnorm (its first B * groups * 8 bytes hold the folded (mean, rstd)).  flags: PCDM_FLAG_SILU. */
int pcdm_groupnorm_apply(const void* x1, const float* stats1, const void* x2, const float* stats2, int C1, void* y,
                         const float* gamma, const float* beta, float eps, int B, int HW, int C, int groups, int dtype,
                         int flags, void* workspace, void* stream);

/* torch.nn.LayerNorm(C, eps) over rows; replaces norm1/norm2/norm3 of BasicTransformerBlock. */
int pcdm_layernorm(const void* x, long long ldx, void* y, long long ldy, const float* gamma, const float* beta,
                   float eps, int M, int C, int dtype, void* stream);

/* Per-row (sum, sum of squares) of x[M, C] (16-bit; ldx in elements) as stats[M][2] fp32: one statistics slot in the
 * layout pcdm_ext.ln_stats consumes (ln_parts = 1), for rows that did not come out of a pcdm_gemm epilogue. */
int pcdm_row_stats(const void* x, long long ldx, float* stats, int M, int C, int dtype, void* stream);

/* softmax(Q K^T * scale) V per head, head_dim 64, no mask.  Replaces xformers memory_efficient_attention /
 * F.scaled_dot_product_attention behind diffusers' attention processors (stage2_batchtest_inpaint_model.py:133;
 * SURVEY.md §8a a9).  q: element (b, s, h, d) at q[(b*Sq + s)*ldq + h*64 + d]; k, v likewise with Skv; out likewise
 * with ldo — so q/k/v may be column slices of one fused projection buffer. */
int pcdm_attention(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, void* out,
                   long long ldo, int B, int heads, int Sq, int Skv, float scale, int dtype, void* stream);
/* The same with an explicit head width: head_dim = 64 or 128; head h of a token row lives at columns
 * [h*head_dim, (h+1)*head_dim) of q / k / v / out.  Other widths are served by zero-padding the projection weights to
 * the next supported width at load time: the CLIP ViT-H/14 image encoder (transformers CLIPVisionModelWithProjection,
 * head_dim 80: stage1_batchtest_prior_model.py:61,100-101; stage2_batchtest_inpaint_model.py:97,181-183) runs with
 * head_dim = 128 and scale = 80^-1/2 — zero q/k columns add nothing to the scores, zero v columns give zero outputs. */
int pcdm_attention_hd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                      void* out, long long ldo, int B, int heads, int Sq, int Skv, int head_dim, float scale, int dtype,
                      void* stream);

/* Boundary layout conversion (the reference's tensors are NCHW).  *_dtype: 0 f16, 1 bf16, 2 f32 (destination of
 * nchw_to_nhwc_pad must be 16-bit).  Replaces nothing arithmetic: forward() entry/exit at reference :579-595,:822-825. */
int pcdm_nchw_to_nhwc_pad(const void* x, int src_dtype, void* y, int dst_dtype, int B, int C, int HW, int Cpad,
                          void* stream);
int pcdm_nhwc_to_nchw(const void* x, int src_dtype, long long ldc, void* y, int dst_dtype, int B, int C, int HW,
                      void* stream);

/* diffusers Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): out[b] = [cos(t f_i) | sin(t f_i)]
 * (reference :677-682).  t: device fp32, t_count = 1 (broadcast) or B. */
int pcdm_timestep_embedding(const float* t, int t_count, void* out, int dtype, int B, int dim, void* stream);

/* F.interpolate(scale_factor=2, mode="nearest") on NHWC (inside diffusers Upsample2D). */
int pcdm_upsample_nearest2x(const void* x, void* y, int B, int H, int W, int C, void* stream);

/* One fused launch per denoising step: CFG combine + DDIM (eta 0) update of the fp32 latents + rewrite of channels
 * 0..3 of the next UNet input for both CFG halves (reference stage2_inpaint_pipeline.py:499-501,510-512,519).
 * coef_table[step] = {1/sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev)} (float4, device);
 * step_counter: int[2] device {step, scratch=0}; the kernel advances step, so one CUDA graph serves every step.
 * t_table (steps+1 fp32 timesteps, device) / t_cur (device scalar) are optional: when given, *t_cur = t_table[step+1].
 * rescale_ratio (device, n floats from pcdm_cfg_rescale_ratio; NULL = off) / guidance_rescale: the reference's
 * rescale_noise_cfg (:52-63, applied at :514-516), i.e. eps = r * (cfg * ratio[b]) + (1 - r) * cfg. */
int pcdm_cfg_ddim_step(const void* eps, int eps_dtype, long long ld_eps, float* latents, void* x9, int x9_dtype,
                       long long ld_x9, const float* coef_table, int* step_counter, float guidance_scale, int n,
                       int HW, const float* t_table, float* t_cur, const float* rescale_ratio, float guidance_rescale,
                       void* stream);

/* Classifier-free guidance on its own (reference stage2_inpaint_pipeline.py:510-516).
 * pcdm_cfg_rescale_ratio: ratio[b] = std(eps_cond[b]) / std(cfg[b]) over (C, H, W) (unbiased, as torch.std), cfg = e_u +
 *   g (e_c - e_u); element (b, c, p) of the 2n-sample epsilon batch (samples [0, n) unconditional) is read at
 *   eps[b * stride_b + c * stride_c + p * stride_p] (strides in elements: NCHW tensors and NHWC rows both fit);
 *   one CTA per sample, fixed reduction order.
 * pcdm_cfg_combine: out[n, per_sample] = cfg of eps[2n, per_sample] (contiguous), rescaled when rescale_ratio != NULL.
 * dtypes: 0 f16, 1 bf16, 2 f32. */
int pcdm_cfg_rescale_ratio(const void* eps, int eps_dtype, long long stride_b, long long stride_c, long long stride_p,
                           int n, int C, int HW, float guidance_scale, float* ratio, void* stream);
int pcdm_cfg_combine(const void* eps, int eps_dtype, void* out, int out_dtype, int n, long long per_sample,
                     float guidance_scale, const float* rescale_ratio, float guidance_rescale, void* stream);

/* DDIMScheduler.step (eta 0) on its own, for callers that drive the scheduler protocol tensor by tensor
 * (stage2_inpaint_pipeline.py:519): prev = sqrt_a_prev * (sample - sqrt_one_minus_a_t * eps) * inv_sqrt_a_t
 *                                          + sqrt_one_minus_a_prev * eps.   dtypes: 0 f16, 1 bf16, 2 f32. */
int pcdm_ddim_step(const void* model_output, int eps_dtype, const void* sample, void* prev_sample, int dtype,
                   float inv_sqrt_a_t, float sqrt_one_minus_a_t, float sqrt_a_prev, float sqrt_one_minus_a_prev,
                   long long numel, void* stream);
/* The stochastic step (eta != 0; diffusers DDIMScheduler.step, protocol at stage2_inpaint_pipeline.py:307-322,519):
 * sigma = eta * sqrt((1 - a_prev) / (1 - a_t)) * sqrt(1 - a_t / a_prev), dir_coef = sqrt(1 - a_prev - sigma^2),
 * prev = sqrt_a_prev * x0 + dir_coef * eps + sigma * noise; noise has the sample's dtype (drawn by the caller's generator). */
int pcdm_ddim_step_eta(const void* model_output, int eps_dtype, const void* sample, const void* noise, void* prev_sample,
                       int dtype, float inv_sqrt_a_t, float sqrt_one_minus_a_t, float sqrt_a_prev, float dir_coef,
                       float sigma, long long numel, void* stream);

/* DDPMScheduler.add_noise (stage2_train_inpaint_model.py:361): out = sqrt(abar_t) x0 + sqrt(1 - abar_t) noise. */
int pcdm_add_noise(const void* x0, const void* noise, void* out, int dtype, const float* alphas_cumprod,
                   const long long* timesteps, int B, long long per_sample, void* stream);

/* UniPCMultistepScheduler.step (diffusers 0.24.0 — the scheduler stage2_batchtest_inpaint_model.py:132 installs):
 * predict_x0, solver bh1/bh2, solver_order <= 2, epsilon prediction, corrector on.  A table row holds the 16
 * schedule-only scalars of one step: {sigma_t, alpha_t (convert_model_output); use_corrector, sigma_t/sigma_s0,
 * alpha_t*h_phi_1, alpha_t*B_h, rk, rho_0, rho_last, order (multistep_uni_c_bh_update); sigma_t/sigma_s0,
 * alpha_t*h_phi_1, alpha_t*B_h, rk, rho, order (multistep_uni_p_bh_update)}.  fp32 state, IEEE round-to-nearest
 * arithmetic in the reference's operation order (bit-identical to its fp32 CPU evaluation).
 * pcdm_cfg_unipc_step: fused CFG combine + update + rewrite of the next UNet input, graph-replayable like
 * pcdm_cfg_ddim_step; state = 4 planes [n,4,HW] fp32 {sample, last_sample, model_outputs[-1], model_outputs[-2]};
 * coef_table: device, steps x 16 floats.
 * pcdm_unipc_step: the scheduler protocol's step() on same-shape contiguous tensors (dtypes 0 f16, 1 bf16, 2 f32);
 * last_sample / m0 / m1 are fp32 history buffers of numel entries owned by the caller; coef_row_host: 16 HOST floats. */
int pcdm_cfg_unipc_step(const void* eps, int eps_dtype, long long ld_eps, float* state, void* x9, int x9_dtype,
                        long long ld_x9, const float* coef_table, int* step_counter, float guidance_scale, int n,
                        int HW, const float* t_table, float* t_cur, const float* rescale_ratio, float guidance_rescale,
                        void* stream);
int pcdm_unipc_step(const void* model_output, int eps_dtype, const void* sample, void* prev_sample, int dtype,
                    float* last_sample, float* m0, float* m1, const float* coef_row_host, long long numel,
                    void* stream);

/* UnCLIPScheduler.step (diffusers 0.24.0 — the scheduler Stage1_PriorPipeline samples the stage-1 prior with,
 * src/pipelines/stage1_prior_pipeline.py:445-446,478-483): x_prev = (c_x0 * clamp(x0, -clip, clip) + c_xt * x_t) + std *
 * noise, x0 = model_output ("sample" prediction) or (x_t - sqrt_b * model_output) / sqrt_a ("epsilon").  A table row
 * holds 8 floats {c_x0, c_xt, std, clip, sqrt_a, sqrt_b, pred_is_epsilon, 0}; IEEE round-to-nearest arithmetic in the
 * reference's operation order (fp32 results bit-identical to its CPU evaluation).
 * pcdm_cfg_unclip_step: fused CFG combine (pred rows [0, n) unconditional, [n, 2n) conditional when use_cfg) + update of
 * the fp32 latents [n, E] + rewrite of the next step's 16-bit / fp32 model-input rows (both halves); coef_table
 * [steps, 8] and noise_table [steps, n, E] on the device, step_counter / t_table / t_cur as pcdm_cfg_ddim_step.
 * pcdm_unclip_step: the scheduler protocol's step() on same-shape contiguous tensors (dtypes 0 f16, 1 bf16, 2 f32);
 * noise has the sample's dtype (NULL allowed when the row's std is 0: the last step); coef_row_host: 8 HOST floats. */
int pcdm_cfg_unclip_step(const float* pred, long long ld_pred, float* latents, void* xin, int xin_dtype,
                         long long ld_xin, const float* coef_table, const float* noise_table, int* step_counter,
                         float guidance_scale, int use_cfg, int n, int E, const float* t_table, float* t_cur,
                         void* stream);
int pcdm_unclip_step(const void* model_output, int mo_dtype, const void* sample, const void* noise, void* prev_sample,
                     int dtype, const float* coef_row_host, long long numel, void* stream);

/* y[m, :] = softmax(scale * x[m, :]): fp32 scores in, 16-bit probabilities out (row strides in elements).  The VAE
 * mid-block attention (diffusers AutoencoderKL, one head of dim 512: stage2_inpaint_pipeline.py:443,528) runs as
 * pcdm_gemm (Q K^T, fp32 out) -> pcdm_softmax_rows -> pcdm_gemm (P V).  N % 4 == 0, N <= 16384. */
int pcdm_softmax_rows(const float* x, long long ldx, void* y, long long ldy, int M, int N, float scale, int dtype,
                      void* stream);

/* DiagonalGaussianDistribution.sample() / .mode() of diffusers AutoencoderKL.encode (stage2_inpaint_pipeline.py:443,
 * stage3_refined_pipeline.py:479): z = (mean + exp(0.5 * clamp(logvar, -30, 20)) * noise) * scale.
 * moments: [B, HW, ld] fp32 rows (channels [0, C) mean, [C, 2C) logvar); noise: [B, C, HW] fp32 or NULL (mode);
 * out: [B, C, HW] fp32. */
int pcdm_gaussian_sample(const float* moments, long long ld, const float* noise, float* out, int B, int C, int HW,
                         float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PCDM_B200_H_ */
