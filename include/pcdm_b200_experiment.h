/* pcdm_b200_experiment.h — tuning / experiment hooks of the EXPERIMENT build (libpcdm_b200_exp.so, compiled with
 * -DPCDM_EXPERIMENT by `python -m pcdms_b200.build --experiment`).  The release library exports none of these and has
 * no mutable process-wide state; tools/ load the experiment build through $PCDM_B200_LIB for A/B timing only.
 * All setters are process-wide and not thread-safe.
 */
#ifndef PCDM_B200_EXPERIMENT_H_
#define PCDM_B200_EXPERIMENT_H_
#include "pcdm_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
int pcdm_set_pdl(int enabled);          /* 1 (default): programmatic dependent launch on every kernel; 0: plain launches */
int pcdm_set_gemm_cta_group(int mode);  /* 0 automatic, 1 single-CTA tiles, 2 CTA pairs wherever the N tile allows */
int pcdm_set_gemm_max_stages(int n);    /* cap the GEMM/conv shared-memory ring depth (2..8) */
int pcdm_set_gemm_debug(int mask);      /* switch parts of the GEMM/conv kernel off for timing — results are WRONG while
                                         * non-zero: 1 no TMA stores, 2 no residual, 4 no bias/rowvec, 8 no epilogue body,
                                         * 16 no MMAs, 32 GEGLU epilogue without its arithmetic (1 and 8 also act on the GEGLU path) */
int pcdm_set_skinny_gemm(int enabled);  /* 0: M <= 32 GEMMs stay on the tcgen05 tiles */
int pcdm_set_attention_small(int on);   /* 0: Sq, Skv <= 32 attention stays on the tcgen05 kernel */
int pcdm_set_attention_poly(int on);    /* 1: half of the softmax exp2 on the FMA pipe (measured slower) */
int pcdm_set_attention_debug(int code); /* FMA-pipe exp2 share of the two-tile attention kernel: 0 release setting (1 of 4
                                         * column pairs), 1 none, 2 = 1 of 3, 3 = 1 of 2 */
int pcdm_set_attention_trace(void* device_buffer); /* 2 x 96 x 16 u64: cycle stamps of CTA 0 of the two-tile attention kernel (NULL: off) */
int pcdm_set_attention_v2(int on);      /* 0: head_dim 64 back on the round-1 kernel (one q-tile per CTA, two CTAs per SM) */
int pcdm_set_groupnorm_two_pass(int mode); /* 0 automatic, 1 two kernels, 2 single pass, 2 + T single pass with T threads */
#ifdef __cplusplus
}
#endif
#endif
