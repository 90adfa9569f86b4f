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
 */
#ifndef PCDM_B200_H_
#define PCDM_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define PCDM_ABI_VERSION 1

#define PCDM_ERR_INVALID (-1)     /* bad argument */
#define PCDM_ERR_CUDA (-2)        /* CUDA runtime / driver failure */
#define PCDM_ERR_UNSUPPORTED (-3) /* shape outside what the kernels implement */

#define PCDM_DT_F16 0
#define PCDM_DT_BF16 1

#define PCDM_FLAG_GEGLU 1   /* gemm: weight rows packed [32 value | 32 gate]; writes N/2 columns value*gelu(gate) */
#define PCDM_FLAG_OUT_F32 2 /* gemm/conv: write fp32 instead of the 16-bit dtype */
#define PCDM_FLAG_SILU 4    /* norm kernels: apply SiLU after the affine */

int pcdm_abi_version(void);
const char* pcdm_last_error(void);

/* torch.nn.Linear (+ fused epilogues).  Replaces the nn.Linear calls inside diffusers Transformer2DModel /
 * BasicTransformerBlock / Attention / GEGLU / TimestepEmbedding and the 1x1 conv_shortcut of ResnetBlock2D
 * (SURVEY.md §8a rows a3, a5, a7, a8).
 *   out[M, N] = A[M, K] . W[N, K]^T + bias[N] + rowvec[m / rows_per_image, N] + residual[M, N]
 * A may be given as two K-segments (a: columns [0, k1), a2: columns [k1, K)) — the skip-concat of the up blocks
 * is consumed in place, never materialised.  lda/lda2/ldo/ldr are row strides in elements.
 * K % 64 == 0, N % 32 == 0.  bn = 0 picks the N tile automatically (64/128/160/256 to force one). */
int pcdm_gemm(const void* a, long long lda, const void* a2, long long lda2, int k1, const void* w, void* out,
              long long ldo, const float* bias, const float* rowvec, int rows_per_image, const void* residual,
              long long ldr, int M, int N, int K, int dtype, int flags, int bn, void* stream);

/* torch.nn.Conv2d(Cin, Cout, 3, stride, padding=1) on NHWC activations, implicit GEMM (no im2col buffer).
 * Replaces conv1/conv2 of ResnetBlock2D, Downsample2D.conv (stride 2) and Upsample2D.conv.
 *   x: [B, stride*H, stride*W, Cin]; out: [B, H, W, Cout]; w_packed: [Cout][3][3][Cin] (tap-major K);
 *   rowvec: [B, Cout] fp32 added per image (the resnet's time_emb_proj term); residual: [B, H, W, Cout]. */
int pcdm_conv3x3(const void* x, const void* w_packed, void* out, const float* bias, const float* rowvec,
                 const void* residual, int B, int H, int W, int Cin, int Cout, int stride, int dtype, int flags,
                 int bn, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PCDM_B200_H_ */
