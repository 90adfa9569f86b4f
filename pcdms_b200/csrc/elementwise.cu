// K6 and friends — the small HBM/launch-bound kernels around the UNet: layout conversion at the boundary
// (the reference's tensors are NCHW, ours NHWC), sinusoidal timestep embedding, nearest-2x upsampling, and the fused
// per-step kernel (classifier-free-guidance combine + DDIM update + rebuild of the next UNet input).
//
// Reference call sites replaced:
//   pcdm_cfg_ddim_step  : src/pipelines/stage2_inpaint_pipeline.py:499-501 (dup + concat), :510-512 (CFG combine),
//                         :519 (scheduler.step — diffusers DDIMScheduler, eta = 0)               [SURVEY §8a a1,a11,a12]
//   pcdm_add_noise      : DDPMScheduler.add_noise, stage2_train_inpaint_model.py:361                          [a12]
//   pcdm_timestep_embedding : diffusers Timesteps(320, flip_sin_to_cos=True, freq_shift=0),
//                         src/models/stage2_inpaint_unet_2d_condition.py:677-682                              [a3]
//   pcdm_upsample_nearest2x : F.interpolate(scale_factor=2, mode="nearest") inside diffusers Upsample2D       [a6]
//   pcdm_nchw_to_nhwc_pad / pcdm_nhwc_to_nchw : the NCHW <-> NHWC boundary of UNet.forward (:579-595, :822-825)
#include "common.cuh"
#include "host_util.h"

namespace pcdm {

// src dtype codes for boundary tensors: 0 = f16, 1 = bf16, 2 = f32
__device__ __forceinline__ float load_any(const void* p, long long i, int dt) {
  if (dt == 2) return reinterpret_cast<const float*>(p)[i];
  if (dt == 0) return __half2float(reinterpret_cast<const __half*>(p)[i]);
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
__device__ __forceinline__ void store_any(void* p, long long i, int dt, float v) {
  if (dt == 2) reinterpret_cast<float*>(p)[i] = v;
  else if (dt == 0) reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
  else reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}

// [B, C, HW] (any float dtype) -> [B, HW, Cpad] 16-bit, channels >= C zero-filled.  One thread per (pixel, 8 ch).
__global__ void nchw_to_nhwc_pad_kernel(const void* __restrict__ x, int src_dt, void* __restrict__ y, int dst_dt, int B,
                                        int C, int HW, int Cpad) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = (long long)B * HW * (Cpad / 8);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % (Cpad / 8));
    const long long bp = i / (Cpad / 8);
    const int p = (int)(bp % HW);
    const int b = (int)(bp / HW);
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = cv * 8 + k;
      v[k] = (c < C) ? load_any(x, ((long long)b * C + c) * HW + p, src_dt) : 0.f;
    }
    uint4 u;
    if (dst_dt == 0) {
      u.x = pack2<DT_F16>(v[0], v[1]); u.y = pack2<DT_F16>(v[2], v[3]);
      u.z = pack2<DT_F16>(v[4], v[5]); u.w = pack2<DT_F16>(v[6], v[7]);
    } else {
      u.x = pack2<DT_BF16>(v[0], v[1]); u.y = pack2<DT_BF16>(v[2], v[3]);
      u.z = pack2<DT_BF16>(v[4], v[5]); u.w = pack2<DT_BF16>(v[6], v[7]);
    }
    reinterpret_cast<uint4*>(y)[i] = u;
  }
}

// [B, HW, ldc] (src dtype) channels [0, C) -> [B, C, HW] (dst dtype)
__global__ void nhwc_to_nchw_kernel(const void* __restrict__ x, int src_dt, long long ldc, void* __restrict__ y,
                                    int dst_dt, int B, int C, int HW) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = (long long)B * C * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const long long bc = i / HW;
    const int c = (int)(bc % C);
    const int b = (int)(bc / C);
    store_any(y, i, dst_dt, load_any(x, ((long long)b * HW + p) * ldc + c, src_dt));
  }
}

// emb[b, :] = [cos(t_b * f_i) | sin(t_b * f_i)], f_i = exp(-ln(10000) * i / half)  (flip_sin_to_cos, freq_shift 0)
// t may be a single value broadcast to all rows (t_count == 1) or one per row.  fp32 math, 16-bit store.
__global__ void timestep_embedding_kernel(const float* __restrict__ t, int t_count, void* __restrict__ out, int dt,
                                          int B, int dim) {
  pdl_launch_dependents();
  pdl_wait();
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  const int b = i / half, k = i % half;
  const float tv = t[t_count == 1 ? 0 : b];
  // double-precision libm (immune to --use_fast_math): arguments reach ~1000 rad, where __sinf/__cosf are useless.
  // The fp32 roundings mirror the reference: f = fp32(exp(.)), a = fp32(t * f), emb = fp32(cos/sin(a)).
  const float f = (float)exp(-9.210340371976184 * (double)k / (double)half);
  const float a = tv * f;
  store_any(out, (long long)b * dim + k, dt, (float)cos((double)a));
  store_any(out, (long long)b * dim + half + k, dt, (float)sin((double)a));
}

// y[b, 2h+dy, 2w+dx, :] = x[b, h, w, :]
__global__ void upsample_nearest2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int B, int H, int W,
                                          int CV) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = (long long)B * (2 * H) * (2 * W) * CV;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    long long r = i / CV;
    const int ox = (int)(r % (2 * W));
    r /= (2 * W);
    const int oy = (int)(r % (2 * H));
    const int b = (int)(r / (2 * H));
    y[i] = __ldg(&x[(((long long)b * H + (oy >> 1)) * W + (ox >> 1)) * CV + cv]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Classifier-free guidance with rescale (reference rescale_noise_cfg, src/pipelines/stage2_inpaint_pipeline.py:52-63;
// applied at :514-516 when guidance_rescale > 0 — off in the shipped drivers):
//     cfg = e_u + g (e_c - e_u);   ratio = std(e_c) / std(cfg)   per sample over (C, H, W), unbiased (torch.std);
//     out = r * (cfg * ratio) + (1 - r) * cfg
// cfg_rescale_ratio_kernel: one CTA per sample, two passes over the sample (means, then squared deviations — no
// catastrophic cancellation), fixed reduction order (bit-reproducible).  Element (b, c, p) of the 2n-row epsilon
// batch lives at eps[b * sb + c * sc + p * sp]: NCHW tensors and the UNet's NHWC output rows are both addressable.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double* red) {   // all threads get the total; fixed order
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}

__global__ void __launch_bounds__(512) cfg_rescale_ratio_kernel(const void* __restrict__ eps, int dt, long long sb,
                                                                long long sc, long long sp, int n, int C, int HW,
                                                                float guidance, float* __restrict__ ratio) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double red[16];
  const int b = blockIdx.x;
  const long long total = (long long)C * HW;
  double s_t = 0.0, s_c = 0.0;
  for (long long i = threadIdx.x; i < total; i += blockDim.x) {
    const int c = (int)(i / HW), p = (int)(i % HW);
    const float eu = load_any(eps, b * sb + c * sc + p * sp, dt);
    const float ec = load_any(eps, (long long)(b + n) * sb + c * sc + p * sp, dt);
    s_t += (double)ec;
    s_c += (double)__fadd_rn(eu, __fmul_rn(guidance, __fsub_rn(ec, eu)));
  }
  const double m_t = block_sum(s_t, red) / (double)total;
  const double m_c = block_sum(s_c, red) / (double)total;
  double q_t = 0.0, q_c = 0.0;
  for (long long i = threadIdx.x; i < total; i += blockDim.x) {
    const int c = (int)(i / HW), p = (int)(i % HW);
    const float eu = load_any(eps, b * sb + c * sc + p * sp, dt);
    const float ec = load_any(eps, (long long)(b + n) * sb + c * sc + p * sp, dt);
    const double dt_ = (double)ec - m_t;
    const double dc = (double)__fadd_rn(eu, __fmul_rn(guidance, __fsub_rn(ec, eu))) - m_c;
    q_t += dt_ * dt_;
    q_c += dc * dc;
  }
  q_t = block_sum(q_t, red);
  q_c = block_sum(q_c, red);
  if (threadIdx.x == 0) {
    const double denom = (double)(total > 1 ? total - 1 : 1);
    ratio[b] = __fdiv_rn((float)sqrt(q_t / denom), (float)sqrt(q_c / denom));
  }
}

// e = e_u + g (e_c - e_u), optionally rescaled: r * (e * ratio[b]) + (1 - r) * e  (reference operation order)
__device__ __forceinline__ float cfg_value(float eu, float ec, float guidance, const float* ratio, int b, float rescale) {
  float e = __fadd_rn(eu, __fmul_rn(guidance, __fsub_rn(ec, eu)));
  if (ratio) e = __fadd_rn(__fmul_rn(rescale, __fmul_rn(e, ratio[b])), __fmul_rn(__fsub_rn(1.0f, rescale), e));
  return e;
}

// The CFG combine on its own, for callers that drive the scheduler protocol tensor by tensor (the generic loop):
// eps: [2n, per_sample] contiguous (rows [0, n) unconditional), out: [n, per_sample].
__global__ void cfg_combine_kernel(const void* __restrict__ eps, int dt, void* __restrict__ out, int out_dt, int n,
                                   long long per_sample, float guidance, const float* __restrict__ ratio, float rescale) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = (long long)n * per_sample;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / per_sample);
    const float eu = load_any(eps, i, dt), ec = load_any(eps, i + total, dt);
    store_any(out, i, out_dt, cfg_value(eu, ec, guidance, ratio, b, rescale));
  }
}

// Fused per-step kernel.  eps: UNet output rows [2n, HW, ld_eps] (channels 0..3; [uncond ; cond] batch halves).
// latents: [n, 4, HW] fp32 scheduler state, updated in place:
//     e      = e_u + g (e_c - e_u)
//     x0     = (x - sqrt(1 - a_t) e) / sqrt(a_t)
//     x_prev = sqrt(a_prev) x0 + sqrt(1 - a_prev) e
// and the first 4 channels of the next UNet input x9 [2n, HW, ld_x9] (NHWC, 16-bit) are rewritten for both halves.
// Coefficients come from a device table coef[step] = {1/sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev)} indexed
// by *step_counter, which the kernel's last thread increments — so an identical CUDA graph replays every step; it also
// publishes the next timestep (t_table[step + 1]) into *t_cur for the next UNet evaluation's embedding.
__global__ void cfg_ddim_step_kernel(const void* __restrict__ eps, int eps_dt, long long ld_eps,
                                     float* __restrict__ latents, void* __restrict__ x9, int x9_dt, long long ld_x9,
                                     const float4* __restrict__ coef, int* __restrict__ step_counter, float guidance,
                                     int n, int HW, const float* __restrict__ t_table, float* __restrict__ t_cur,
                                     const float* __restrict__ ratio, float rescale) {
  pdl_launch_dependents();
  pdl_wait();
  const int step = *step_counter;
  const float4 cf = coef[step];
  const long long total = (long long)n * HW;
  const bool vec_eps = eps_dt == 2 && (ld_eps & 3) == 0 && (reinterpret_cast<uintptr_t>(eps) & 15) == 0;
  const bool vec_x9 = x9_dt != 2 && (ld_x9 & 3) == 0 && (reinterpret_cast<uintptr_t>(x9) & 7) == 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const int b = (int)(i / HW);
    // the UNet's fp32 output rows: the four epsilon channels of a pixel are one 16-byte load per CFG half
    float eu4[4], ec4[4];
    if (vec_eps) {
      const float4 u = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(eps) + ((long long)b * HW + p) * ld_eps);
      const float4 c = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(eps) + ((long long)(b + n) * HW + p) * ld_eps);
      eu4[0] = u.x; eu4[1] = u.y; eu4[2] = u.z; eu4[3] = u.w;
      ec4[0] = c.x; ec4[1] = c.y; ec4[2] = c.z; ec4[3] = c.w;
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        eu4[c] = load_any(eps, ((long long)b * HW + p) * ld_eps + c, eps_dt);
        ec4[c] = load_any(eps, ((long long)(b + n) * HW + p) * ld_eps + c, eps_dt);
      }
    }
    float xp4[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float eu = eu4[c], ec = ec4[c];
      const float e = ratio ? cfg_value(eu, ec, guidance, ratio, b, rescale) : eu + guidance * (ec - eu);
      const long long li = ((long long)b * 4 + c) * HW + p;
      const float x = latents[li];
      const float x0 = (x - cf.y * e) * cf.x;
      const float xp = cf.z * x0 + cf.w * e;
      latents[li] = xp;                      // NCHW planes: consecutive threads write consecutive pixels
      xp4[c] = xp;
    }
    if (vec_x9) {                            // 16-bit NHWC rows: the four latent channels are one 8-byte store per half
      uint2 o;
      if (x9_dt == 0) { o.x = pack2<DT_F16>(xp4[0], xp4[1]); o.y = pack2<DT_F16>(xp4[2], xp4[3]); }
      else { o.x = pack2<DT_BF16>(xp4[0], xp4[1]); o.y = pack2<DT_BF16>(xp4[2], xp4[3]); }
      *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(x9) + ((long long)b * HW + p) * ld_x9) = o;
      *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(x9) + ((long long)(b + n) * HW + p) * ld_x9) = o;
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        store_any(x9, ((long long)b * HW + p) * ld_x9 + c, x9_dt, xp4[c]);
        store_any(x9, ((long long)(b + n) * HW + p) * ld_x9 + c, x9_dt, xp4[c]);
      }
    }
  }
  __syncthreads();
  // grid-wide "last block" increments the step counter exactly once
  __shared__ bool last;
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned done = atomicAdd(reinterpret_cast<unsigned*>(step_counter + 1), 1u);
    last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    step_counter[1] = 0;
    step_counter[0] = step + 1;
    if (t_table && t_cur) *t_cur = t_table[step + 1];  // timestep the next UNet evaluation embeds (table has steps+1 rows)
  }
}

// out = sqrt(a_t[b]) x0 + sqrt(1 - a_t[b]) noise      (per-sample timestep)
__global__ void add_noise_kernel(const void* __restrict__ x0, const void* __restrict__ noise, void* __restrict__ out,
                                 int dt, const float* __restrict__ alphas_cumprod, const long long* __restrict__ t,
                                 int B, long long per_sample) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = (long long)B * per_sample;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / per_sample);
    const float a = alphas_cumprod[t[b]];
    const float sa = sqrtf(a), sb = sqrtf(1.f - a);
    store_any(out, i, dt, sa * load_any(x0, i, dt) + sb * load_any(noise, i, dt));
  }
}

// Vectorised forms for same-dtype 16-bit tensors (what the training / protocol callers pass): 16 bytes = 8 elements
// per thread per access, fully coalesced; identical per-element arithmetic to the scalar kernels above / below.
template <int DT>
__global__ void add_noise_vec_kernel(const uint4* __restrict__ x0, const uint4* __restrict__ noise, uint4* __restrict__ out,
                                     const float* __restrict__ alphas_cumprod, const long long* __restrict__ t,
                                     long long nvec, long long per_sample) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvec; v += (long long)gridDim.x * blockDim.x) {
    const int b = (int)((v * 8) / per_sample);          // per_sample % 8 == 0: a vector never straddles two samples
    const float a = alphas_cumprod[t[b]];
    const float sa = sqrtf(a), sb = sqrtf(1.f - a);
    const uint4 xv = x0[v], nv = noise[v];
    const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w}, ns[4] = {nv.x, nv.y, nv.z, nv.w};
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 xf = unpack2<DT>(xs[i]), nf = unpack2<DT>(ns[i]);
      o[i] = pack2<DT>(sa * xf.x + sb * nf.x, sa * xf.y + sb * nf.y);
    }
    out[v] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

template <int DT>
__global__ void ddim_step_vec_kernel(const uint4* __restrict__ eps, const uint4* __restrict__ x, uint4* __restrict__ out,
                                     float c0, float c1, float c2, float c3, long long nvec) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvec; v += (long long)gridDim.x * blockDim.x) {
    const uint4 ev = eps[v], xv = x[v];
    const uint32_t es[4] = {ev.x, ev.y, ev.z, ev.w}, xs[4] = {xv.x, xv.y, xv.z, xv.w};
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 e = unpack2<DT>(es[i]), xf = unpack2<DT>(xs[i]);
      const float a0 = (xf.x - c1 * e.x) * c0, a1 = (xf.y - c1 * e.y) * c0;
      o[i] = pack2<DT>(c2 * a0 + c3 * e.x, c2 * a1 + c3 * e.y);
    }
    out[v] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// stand-alone scheduler.step: prev = c2 * (x - c1 * eps) * c0 + c3 * eps   (same arithmetic as the fused kernel)
// (+ sigma * noise for the stochastic step, eta != 0: c3 is then sqrt(1 - a_prev - sigma^2))
__global__ void ddim_step_kernel(const void* __restrict__ eps, int eps_dt, const void* __restrict__ x,
                                 const void* __restrict__ noise, void* __restrict__ out, int dt, float c0, float c1,
                                 float c2, float c3, float sigma, long long total) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float e = load_any(eps, i, eps_dt);
    const float x0 = (load_any(x, i, dt) - c1 * e) * c0;
    float r = c2 * x0 + c3 * e;
    if (noise) r += sigma * load_any(noise, i, dt);
    store_any(out, i, dt, r);
  }
}


// ---------------------------------------------------------------------------------------------------------------
// UniPCMultistepScheduler.step (diffusers 0.24.0; predict_x0, solver bh1/bh2, solver_order <= 2, epsilon prediction) —
// the scheduler the reference's batch-test driver installs (stage2_batchtest_inpaint_model.py:132).  All step-dependent
// scalars are functions of the noise schedule only and come from a 16-float host-built table row:
//   [0] sigma_t  [1] alpha_t            convert_model_output: m_t = (x - sigma_t e) / alpha_t
//   [2] use_corrector  [3] c_a = sigma_t/sigma_s0  [4] c_b = alpha_t h_phi_1  [5] c_c = alpha_t B_h  [6] c_rk
//   [7] c_rho0  [8] c_rho_last  [9] corrector order                             multistep_uni_c_bh_update
//   [10] p_a  [11] p_b  [12] p_c  [13] p_rk  [14] p_rho  [15] predictor order   multistep_uni_p_bh_update
// Per element the arithmetic follows the reference's tensor expressions operation by operation with IEEE
// round-to-nearest intrinsics (no FMA contraction), so fp32 results are bit-identical to the torch-CPU evaluation.
// State (fp32): x = current sample, last = last_sample, ma = model_outputs[-1], mb = model_outputs[-2].
// ---------------------------------------------------------------------------------------------------------------
struct UniPCRow { float v[16]; };

__device__ __forceinline__ void unipc_update(const UniPCRow& r, float e, float& x, float& last, float& ma, float& mb) {
  const float mt = __fdiv_rn(__fsub_rn(x, __fmul_rn(r.v[0], e)), r.v[1]);
  float xc = x;
  if (r.v[2] != 0.f) {
    const float xt_ = __fsub_rn(__fmul_rn(r.v[3], last), __fmul_rn(r.v[4], ma));
    const float d1t = __fsub_rn(mt, ma);
    float corr = __fmul_rn(r.v[8], d1t);
    if (r.v[9] >= 2.f) corr = __fadd_rn(__fmul_rn(r.v[7], __fdiv_rn(__fsub_rn(mb, ma), r.v[6])), corr);
    xc = __fsub_rn(xt_, __fmul_rn(r.v[5], corr));
  }
  // shift the history: model_outputs = [old ma, mt], last_sample = corrected sample
  const float m_prev = ma;
  mb = ma;
  ma = mt;
  last = xc;
  float xn = __fsub_rn(__fmul_rn(r.v[10], xc), __fmul_rn(r.v[11], mt));
  if (r.v[15] >= 2.f) {
    const float pred = __fmul_rn(r.v[14], __fdiv_rn(__fsub_rn(m_prev, mt), r.v[13]));
    xn = __fsub_rn(xn, __fmul_rn(r.v[12], pred));
  }
  x = xn;
}

// Fused per-step kernel of the B200 pipeline: CFG combine + UniPC update + rewrite of the next UNet input, replayable
// from one CUDA graph (device step counter, as cfg_ddim_step_kernel).  state: 4 planes of [n, 4, HW] fp32.
__global__ void cfg_unipc_step_kernel(const void* __restrict__ eps, int eps_dt, long long ld_eps,
                                      float* __restrict__ state, void* __restrict__ x9, int x9_dt, long long ld_x9,
                                      const UniPCRow* __restrict__ coef, int* __restrict__ step_counter, float guidance,
                                      int n, int HW, const float* __restrict__ t_table, float* __restrict__ t_cur,
                                      const float* __restrict__ ratio, float rescale) {
  pdl_launch_dependents();
  pdl_wait();
  const int step = *step_counter;
  const UniPCRow r = coef[step];
  const long long total = (long long)n * HW;
  const long long plane = total * 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const int b = (int)(i / HW);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float eu = load_any(eps, ((long long)b * HW + p) * ld_eps + c, eps_dt);
      const float ec = load_any(eps, ((long long)(b + n) * HW + p) * ld_eps + c, eps_dt);
      const float e = cfg_value(eu, ec, guidance, ratio, b, rescale);
      const long long li = ((long long)b * 4 + c) * HW + p;
      float x = state[li], last = state[plane + li], ma = state[2 * plane + li], mb = state[3 * plane + li];
      unipc_update(r, e, x, last, ma, mb);
      state[li] = x; state[plane + li] = last; state[2 * plane + li] = ma; state[3 * plane + li] = mb;
      store_any(x9, ((long long)b * HW + p) * ld_x9 + c, x9_dt, x);
      store_any(x9, ((long long)(b + n) * HW + p) * ld_x9 + c, x9_dt, x);
    }
  }
  __syncthreads();
  __shared__ bool last_block;
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned done = atomicAdd(reinterpret_cast<unsigned*>(step_counter + 1), 1u);
    last_block = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (last_block && threadIdx.x == 0) {
    step_counter[1] = 0;
    step_counter[0] = step + 1;
    if (t_table && t_cur) *t_cur = t_table[step + 1];
  }
}

// stand-alone scheduler.step for callers that drive the scheduler protocol tensor by tensor: model_output / sample /
// prev_sample are same-shape contiguous tensors, the three history planes are fp32 buffers the scheduler object owns.
__global__ void unipc_step_kernel(const void* __restrict__ eps, int eps_dt, const void* __restrict__ sample,
                                  void* __restrict__ prev, int dt, float* __restrict__ last, float* __restrict__ ma,
                                  float* __restrict__ mb, UniPCRow r, long long total) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float x = load_any(sample, i, dt), l = last[i], a = ma[i], b = mb[i];
    unipc_update(r, load_any(eps, i, eps_dt), x, l, a, b);
    last[i] = l; ma[i] = a; mb[i] = b;
    store_any(prev, i, dt, x);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// UnCLIPScheduler.step (diffusers 0.24.0 scheduling_unclip.py; the scheduler of the stage-1 prior pipeline,
// /root/reference/src/pipelines/stage1_prior_pipeline.py:478-483):
//     x_prev = (c_x0 * clamp(x0, -clip, clip) + c_xt * x_t) + std * noise
// with x0 = model_output ("sample" prediction, the reference's configuration) or (x_t - sqrt(1-abar) eps) / sqrt(abar)
// ("epsilon").  A row holds the 8 schedule-only scalars of one step, computed by the host with the reference's fp32
// tensor arithmetic; the products / sums below are IEEE round-to-nearest in the reference's order, so fp32 results are
// bit-identical to its CPU evaluation.
struct UnCLIPRow { float c_x0, c_xt, std, clip, sqrt_a, sqrt_b, pred_eps, pad; };

__device__ __forceinline__ float unclip_update(const UnCLIPRow& r, float mo, float x, float noise) {
  float x0 = mo;
  if (r.pred_eps != 0.f) x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(r.sqrt_b, mo)), r.sqrt_a);
  x0 = fminf(fmaxf(x0, -r.clip), r.clip);
  float xp = __fadd_rn(__fmul_rn(r.c_x0, x0), __fmul_rn(r.c_xt, x));
  if (r.std != 0.f) xp = __fadd_rn(xp, __fmul_rn(r.std, noise));
  return xp;
}

// Fused per-step kernel of the stage-1 engine: CFG combine (rows [0, n) unconditional, [n, 2n) conditional when
// use_cfg) + UnCLIP update of the fp32 latents [n, E] + rewrite of the 16-bit model-input rows of the next step (both
// CFG halves); step-indexed device tables (coef [steps], noise [steps, n, E]) and the device step counter make one CUDA
// graph serve every step, as cfg_ddim_step_kernel.
__global__ void cfg_unclip_step_kernel(const float* __restrict__ pred, long long ld_pred, float* __restrict__ latents,
                                       void* __restrict__ xin, int xin_dt, long long ld_xin,
                                       const UnCLIPRow* __restrict__ coef, const float* __restrict__ noise,
                                       int* __restrict__ step_counter, float guidance, int use_cfg, int n, int E,
                                       const float* __restrict__ t_table, float* __restrict__ t_cur) {
  pdl_launch_dependents();
  pdl_wait();
  const int step = *step_counter;
  const UnCLIPRow r = coef[step];
  const long long total = (long long)n * E;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % E);
    const int b = (int)(i / E);
    float mo = pred[(long long)b * ld_pred + e];
    if (use_cfg) {
      const float c = pred[(long long)(b + n) * ld_pred + e];
      mo = __fadd_rn(mo, __fmul_rn(guidance, __fsub_rn(c, mo)));
    }
    const float xp = unclip_update(r, mo, latents[i], r.std != 0.f ? noise[(long long)step * total + i] : 0.f);
    latents[i] = xp;
    store_any(xin, (long long)b * ld_xin + e, xin_dt, xp);
    if (use_cfg) store_any(xin, (long long)(b + n) * ld_xin + e, xin_dt, xp);
  }
  __syncthreads();
  __shared__ bool last_block;
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned done = atomicAdd(reinterpret_cast<unsigned*>(step_counter + 1), 1u);
    last_block = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (last_block && threadIdx.x == 0) {
    step_counter[1] = 0;
    step_counter[0] = step + 1;
    if (t_table && t_cur) *t_cur = t_table[step + 1];
  }
}

__global__ void unclip_step_kernel(const void* __restrict__ mo, int mo_dt, const void* __restrict__ sample,
                                   const void* __restrict__ noise, void* __restrict__ prev, int dt, UnCLIPRow r,
                                   long long total) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    store_any(prev, i, dt, unclip_update(r, load_any(mo, i, mo_dt), load_any(sample, i, dt),
                                         (noise && r.std != 0.f) ? load_any(noise, i, dt) : 0.f));
}

// ---------------------------------------------------------------------------------------------------------------
// Row softmax of fp32 scores -> 16-bit probabilities: y[m, :] = softmax(scale * x[m, :]).  Used by the VAE mid-block
// attention (one head of dim 512, diffusers AutoencoderKL: stage2_inpaint_pipeline.py:443,528), whose QK^T and PV
// products run on the GEMM kernel.  One CTA per row, the row lives in registers (N <= 16384): one read, one write.
// ---------------------------------------------------------------------------------------------------------------
constexpr int SM_THREADS = 256;
constexpr int SM_MAXV = 16;   // float4 per thread

template <int DT>
__global__ void __launch_bounds__(SM_THREADS) softmax_rows_kernel(const float* __restrict__ x, long long ldx,
                                                                  void* __restrict__ y, long long ldy, int M, int N,
                                                                  float scale_log2e) {
  using T = typename TypeOf<DT>::T;
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[SM_THREADS / 32];
  __shared__ float bcast;
  const int nv = N / 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int m = blockIdx.x; m < M; m += gridDim.x) {
    const float4* xr = reinterpret_cast<const float4*>(x + (long long)m * ldx);
    float4 v[SM_MAXV];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < SM_MAXV; ++k) {
      const int i = threadIdx.x + k * SM_THREADS;
      if (i < nv) {
        v[k] = __ldg(xr + i);
        mx = fmaxf(mx, fmaxf(fmaxf(v[k].x, v[k].y), fmaxf(v[k].z, v[k].w)));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = red[0];
      for (int w = 1; w < SM_THREADS / 32; ++w) t = fmaxf(t, red[w]);
      bcast = t;
    }
    __syncthreads();
    mx = bcast;
    const float off = mx * scale_log2e;
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < SM_MAXV; ++k) {
      const int i = threadIdx.x + k * SM_THREADS;
      if (i < nv) {
        v[k].x = exp2f(fmaf(v[k].x, scale_log2e, -off));
        v[k].y = exp2f(fmaf(v[k].y, scale_log2e, -off));
        v[k].z = exp2f(fmaf(v[k].z, scale_log2e, -off));
        v[k].w = exp2f(fmaf(v[k].w, scale_log2e, -off));
        sum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
      }
    }
    sum = warp_sum(sum);
    __syncthreads();   // red / bcast reuse
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < SM_THREADS / 32; ++w) t += red[w];   // fixed order: bit-reproducible
      bcast = 1.0f / t;
    }
    __syncthreads();
    const float inv = bcast;
    uint2* yr = reinterpret_cast<uint2*>(reinterpret_cast<T*>(y) + (long long)m * ldy);
#pragma unroll
    for (int k = 0; k < SM_MAXV; ++k) {
      const int i = threadIdx.x + k * SM_THREADS;
      if (i < nv) yr[i] = make_uint2(pack2<DT>(v[k].x * inv, v[k].y * inv), pack2<DT>(v[k].z * inv, v[k].w * inv));
    }
    __syncthreads();
  }
}

// DiagonalGaussianDistribution.sample() of diffusers AutoencoderKL (stage2_inpaint_pipeline.py:443):
//   z = (mean + exp(0.5 * clamp(logvar, -30, 20)) * noise) * scale
// moments: [B, HW, ld] fp32 rows out of the quant_conv GEMM (channels [0, C) mean, [C, 2C) logvar);
// noise: [B, C, HW] fp32 or NULL (= mode()); out: [B, C, HW] fp32.
__global__ void gaussian_sample_kernel(const float* __restrict__ moments, long long ld, const float* __restrict__ noise,
                                       float* __restrict__ out, int B, int C, int HW, float scale) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = (long long)B * C * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const long long bc = i / HW;
    const int c = (int)(bc % C);
    const int b = (int)(bc / C);
    const float* row = moments + ((long long)b * HW + p) * ld;
    float v = row[c];
    if (noise) {
      const float lv = fminf(fmaxf(row[C + c], -30.f), 20.f);
      v = __fadd_rn(v, __fmul_rn(expf(0.5f * lv), noise[i]));
    }
    out[i] = v * scale;
  }
}

static inline int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace pcdm

using namespace pcdm;

extern "C" int pcdm_nchw_to_nhwc_pad(const void* x, int src_dtype, void* y, int dst_dtype, int B, int C, int HW,
                                     int Cpad, void* stream_) {
  if (!x || !y) return set_error(PCDM_ERR_INVALID, "nchw_to_nhwc_pad: null pointer");
  if (src_dtype < 0 || src_dtype > 2 || dst_dtype < 0 || dst_dtype > 1) return set_error(PCDM_ERR_INVALID, "nchw_to_nhwc_pad: bad dtype");
  if (B <= 0 || C <= 0 || HW <= 0 || Cpad < C || Cpad % 8) return set_error(PCDM_ERR_INVALID, "nchw_to_nhwc_pad: bad shape");
  const long long total = (long long)B * HW * (Cpad / 8);
  PCDM_CUDA(launch_kernel(nchw_to_nhwc_pad_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream_, 1, x, src_dtype, y, dst_dtype, B, C, HW, Cpad));
  PCDM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pcdm_nhwc_to_nchw(const void* x, int src_dtype, long long ldc, void* y, int dst_dtype, int B, int C,
                                 int HW, void* stream_) {
  if (!x || !y) return set_error(PCDM_ERR_INVALID, "nhwc_to_nchw: null pointer");
  if (src_dtype < 0 || src_dtype > 2 || dst_dtype < 0 || dst_dtype > 2) return set_error(PCDM_ERR_INVALID, "nhwc_to_nchw: bad dtype");
  if (B <= 0 || C <= 0 || HW <= 0 || ldc < C) return set_error(PCDM_ERR_INVALID, "nhwc_to_nchw: bad shape");
  const long long total = (long long)B * C * HW;
  PCDM_CUDA(launch_kernel(nhwc_to_nchw_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream_, 1, x, src_dtype, ldc, y, dst_dtype, B, C, HW));
  PCDM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pcdm_timestep_embedding(const float* t, int t_count, void* out, int dtype, int B, int dim,
                                       void* stream_) {
  if (!t || !out) return set_error(PCDM_ERR_INVALID, "timestep_embedding: null pointer");
  if (dtype < 0 || dtype > 2) return set_error(PCDM_ERR_INVALID, "timestep_embedding: bad dtype");
  if (B <= 0 || dim <= 0 || dim % 2 || (t_count != 1 && t_count != B)) return set_error(PCDM_ERR_INVALID, "timestep_embedding: bad shape");
  const int total = B * (dim / 2);
  PCDM_CUDA(launch_kernel(timestep_embedding_kernel, dim3((total + 127) / 128), dim3(128), 0, (cudaStream_t)stream_, 1, t, t_count, out, dtype, B, dim));
  PCDM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pcdm_upsample_nearest2x(const void* x, void* y, int B, int H, int W, int C, void* stream_) {
  if (!x || !y) return set_error(PCDM_ERR_INVALID, "upsample: null pointer");
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 8) return set_error(PCDM_ERR_INVALID, "upsample: bad shape (C % 8 == 0)");
  const long long total = (long long)B * 4 * H * W * (C / 8);
  PCDM_CUDA(launch_kernel(upsample_nearest2x_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream_, 1, reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), B, H, W, C / 8));
  PCDM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pcdm_cfg_ddim_step(const void* eps, int eps_dtype, long long ld_eps, float* latents, void* x9,
                                  int x9_dtype, long long ld_x9, const float* coef_table, int* step_counter,
                                  float guidance_scale, int n, int HW, const float* t_table, float* t_cur,
                                  const float* rescale_ratio, float guidance_rescale, void* stream_) {
  if (!eps || !latents || !x9 || !coef_table || !step_counter) return set_error(PCDM_ERR_INVALID, "cfg_ddim_step: null pointer");
  if (eps_dtype < 0 || eps_dtype > 2 || x9_dtype < 0 || x9_dtype > 1) return set_error(PCDM_ERR_INVALID, "cfg_ddim_step: bad dtype");
  if (n <= 0 || HW <= 0 || ld_eps < 4 || ld_x9 < 4) return set_error(PCDM_ERR_INVALID, "cfg_ddim_step: bad shape");
  if (reinterpret_cast<uintptr_t>(coef_table) & 15) return set_error(PCDM_ERR_INVALID, "cfg_ddim_step: coef table must be 16-byte aligned");
  const long long total = (long long)n * HW;
  PCDM_CUDA(launch_kernel(cfg_ddim_step_kernel, dim3(grid_for(total, 128)), dim3(128), 0, (cudaStream_t)stream_, 1, eps, eps_dtype, ld_eps, latents, x9, x9_dtype, ld_x9, reinterpret_cast<const float4*>(coef_table), step_counter, guidance_scale, n, HW, t_table, t_cur, rescale_ratio, guidance_rescale));
  PCDM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pcdm_cfg_rescale_ratio(const void* eps, int eps_dtype, long long stride_b, long long stride_c,
                                      long long stride_p, int n, int C, int HW, float guidance_scale, float* ratio,
                                      void* stream_) {
  if (!eps || !ratio) return set_error(PCDM_ERR_INVALID, "cfg_rescale_ratio: null pointer");
  if (eps_dtype < 0 || eps_dtype > 2) return set_error(PCDM_ERR_INVALID, "cfg_rescale_ratio: bad dtype");
  if (n <= 0 || C <= 0 || HW <= 0 || stride_b <= 0 || stride_c <= 0 || stride_p <= 0)
    return set_error(PCDM_ERR_INVALID, "cfg_rescale_ratio: bad shape");
  PCDM_CUDA(launch_kernel(cfg_rescale_ratio_kernel, dim3(n), dim3(512), 0, (cudaStream_t)stream_, 1, eps, eps_dtype,
                          stride_b, stride_c, stride_p, n, C, HW, guidance_scale, ratio));
  PCDM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pcdm_cfg_combine(const void* eps, int eps_dtype, void* out, int out_dtype, int n, long long per_sample,
                                float guidance_scale, const float* rescale_ratio, float guidance_rescale, void* stream_) {
  if (!eps || !out) return set_error(PCDM_ERR_INVALID, "cfg_combine: null pointer");
  if (eps_dtype < 0 || eps_dtype > 2 || out_dtype < 0 || out_dtype > 2) return set_error(PCDM_ERR_INVALID, "cfg_combine: bad dtype");
  if (n <= 0 || per_sample <= 0) return set_error(PCDM_ERR_INVALID, "cfg_combine: empty problem");
  const long long total = (long long)n * per_sample;
  PCDM_CUDA(launch_kernel(cfg_combine_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream_, 1, eps,
                          eps_dtype, out, out_dtype, n, per_sample, guidance_scale, rescale_ratio, guidance_rescale));
  PCDM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pcdm_add_noise(const void* x0, const void* noise, void* out, int dtype, const float* alphas_cumprod,
                              const long long* timesteps, int B, long long per_sample, void* stream_) {
  if (!x0 || !noise || !out || !alphas_cumprod || !timesteps) return set_error(PCDM_ERR_INVALID, "add_noise: null pointer");
  if (dtype < 0 || dtype > 2) return set_error(PCDM_ERR_INVALID, "add_noise: bad dtype");
  if (B <= 0 || per_sample <= 0) return set_error(PCDM_ERR_INVALID, "add_noise: empty problem");
  const long long total = (long long)B * per_sample;
  const bool aligned = ((reinterpret_cast<uintptr_t>(x0) | reinterpret_cast<uintptr_t>(noise) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (dtype != 2 && per_sample % 8 == 0 && aligned) {   // 16-byte vector path
    const long long nvec = total / 8;
    if (dtype == 0)
      PCDM_CUDA(launch_kernel(add_noise_vec_kernel<DT_F16>, dim3(grid_for(nvec, 256)), dim3(256), 0, (cudaStream_t)stream_, 1, reinterpret_cast<const uint4*>(x0), reinterpret_cast<const uint4*>(noise), reinterpret_cast<uint4*>(out), alphas_cumprod, timesteps, nvec, per_sample));
    else
      PCDM_CUDA(launch_kernel(add_noise_vec_kernel<DT_BF16>, dim3(grid_for(nvec, 256)), dim3(256), 0, (cudaStream_t)stream_, 1, reinterpret_cast<const uint4*>(x0), reinterpret_cast<const uint4*>(noise), reinterpret_cast<uint4*>(out), alphas_cumprod, timesteps, nvec, per_sample));
    PCDM_CUDA(cudaGetLastError());
    return 0;
  }
  PCDM_CUDA(launch_kernel(add_noise_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream_, 1, x0, noise, out, dtype, alphas_cumprod, timesteps, B, per_sample));
  PCDM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pcdm_ddim_step(const void* model_output, int eps_dtype, const void* sample, void* prev_sample, int dtype,
                              float inv_sqrt_a_t, float sqrt_one_minus_a_t, float sqrt_a_prev,
                              float sqrt_one_minus_a_prev, long long numel, void* stream_) {
  if (!model_output || !sample || !prev_sample) return set_error(PCDM_ERR_INVALID, "ddim_step: null pointer");
  if (eps_dtype < 0 || eps_dtype > 2 || dtype < 0 || dtype > 2) return set_error(PCDM_ERR_INVALID, "ddim_step: bad dtype");
  if (numel <= 0) return set_error(PCDM_ERR_INVALID, "ddim_step: empty problem");
  const bool aligned = ((reinterpret_cast<uintptr_t>(model_output) | reinterpret_cast<uintptr_t>(sample) | reinterpret_cast<uintptr_t>(prev_sample)) & 15) == 0;
  if (dtype != 2 && eps_dtype == dtype && numel % 8 == 0 && aligned) {   // 16-byte vector path
    const long long nvec = numel / 8;
    if (dtype == 0)
      PCDM_CUDA(launch_kernel(ddim_step_vec_kernel<DT_F16>, dim3(grid_for(nvec, 256)), dim3(256), 0, (cudaStream_t)stream_, 1, reinterpret_cast<const uint4*>(model_output), reinterpret_cast<const uint4*>(sample), reinterpret_cast<uint4*>(prev_sample), inv_sqrt_a_t, sqrt_one_minus_a_t, sqrt_a_prev, sqrt_one_minus_a_prev, nvec));
    else
      PCDM_CUDA(launch_kernel(ddim_step_vec_kernel<DT_BF16>, dim3(grid_for(nvec, 256)), dim3(256), 0, (cudaStream_t)stream_, 1, reinterpret_cast<const uint4*>(model_output), reinterpret_cast<const uint4*>(sample), reinterpret_cast<uint4*>(prev_sample), inv_sqrt_a_t, sqrt_one_minus_a_t, sqrt_a_prev, sqrt_one_minus_a_prev, nvec));
    PCDM_CUDA(cudaGetLastError());
    return 0;
  }
  PCDM_CUDA(launch_kernel(ddim_step_kernel, dim3(grid_for(numel, 256)), dim3(256), 0, (cudaStream_t)stream_, 1, model_output, eps_dtype, sample, (const void*)nullptr, prev_sample, dtype, inv_sqrt_a_t, sqrt_one_minus_a_t, sqrt_a_prev, sqrt_one_minus_a_prev, 0.0f, numel));
  PCDM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pcdm_ddim_step_eta(const void* model_output, int eps_dtype, const void* sample, const void* noise,
                                  void* prev_sample, int dtype, float inv_sqrt_a_t, float sqrt_one_minus_a_t,
                                  float sqrt_a_prev, float dir_coef, float sigma, long long numel, void* stream_) {
  if (!model_output || !sample || !noise || !prev_sample) return set_error(PCDM_ERR_INVALID, "ddim_step_eta: null pointer");
  if (eps_dtype < 0 || eps_dtype > 2 || dtype < 0 || dtype > 2) return set_error(PCDM_ERR_INVALID, "ddim_step_eta: bad dtype");
  if (numel <= 0) return set_error(PCDM_ERR_INVALID, "ddim_step_eta: empty problem");
  PCDM_CUDA(launch_kernel(ddim_step_kernel, dim3(grid_for(numel, 256)), dim3(256), 0, (cudaStream_t)stream_, 1, model_output, eps_dtype, sample, noise, prev_sample, dtype, inv_sqrt_a_t, sqrt_one_minus_a_t, sqrt_a_prev, dir_coef, sigma, numel));
  PCDM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pcdm_cfg_unipc_step(const void* eps, int eps_dtype, long long ld_eps, float* state, void* x9,
                                   int x9_dtype, long long ld_x9, const float* coef_table, int* step_counter,
                                   float guidance_scale, int n, int HW, const float* t_table, float* t_cur,
                                   const float* rescale_ratio, float guidance_rescale, void* stream_) {
  if (!eps || !state || !x9 || !coef_table || !step_counter) return set_error(PCDM_ERR_INVALID, "cfg_unipc_step: null pointer");
  if (eps_dtype < 0 || eps_dtype > 2 || x9_dtype < 0 || x9_dtype > 1) return set_error(PCDM_ERR_INVALID, "cfg_unipc_step: bad dtype");
  if (n <= 0 || HW <= 0 || ld_eps < 4 || ld_x9 < 4) return set_error(PCDM_ERR_INVALID, "cfg_unipc_step: bad shape");
  const long long total = (long long)n * HW;
  PCDM_CUDA(launch_kernel(cfg_unipc_step_kernel, dim3(grid_for(total, 128)), dim3(128), 0, (cudaStream_t)stream_, 1, eps, eps_dtype, ld_eps, state, x9, x9_dtype, ld_x9, reinterpret_cast<const UniPCRow*>(coef_table), step_counter, guidance_scale, n, HW, t_table, t_cur, rescale_ratio, guidance_rescale));
  PCDM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pcdm_unipc_step(const void* model_output, int eps_dtype, const void* sample, void* prev_sample, int dtype,
                               float* last_sample, float* m0, float* m1, const float* coef_row_host, long long numel,
                               void* stream_) {
  if (!model_output || !sample || !prev_sample || !last_sample || !m0 || !m1 || !coef_row_host)
    return set_error(PCDM_ERR_INVALID, "unipc_step: null pointer");
  if (eps_dtype < 0 || eps_dtype > 2 || dtype < 0 || dtype > 2) return set_error(PCDM_ERR_INVALID, "unipc_step: bad dtype");
  if (numel <= 0) return set_error(PCDM_ERR_INVALID, "unipc_step: empty problem");
  UniPCRow r;
  memcpy(r.v, coef_row_host, sizeof(r.v));
  PCDM_CUDA(launch_kernel(unipc_step_kernel, dim3(grid_for(numel, 256)), dim3(256), 0, (cudaStream_t)stream_, 1, model_output, eps_dtype, sample, prev_sample, dtype, last_sample, m0, m1, r, numel));
  PCDM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pcdm_cfg_unclip_step(const float* pred, long long ld_pred, float* latents, void* xin, int xin_dtype,
                                    long long ld_xin, const float* coef_table, const float* noise_table,
                                    int* step_counter, float guidance_scale, int use_cfg, int n, int E,
                                    const float* t_table, float* t_cur, void* stream_) {
  if (!pred || !latents || !xin || !coef_table || !noise_table || !step_counter)
    return set_error(PCDM_ERR_INVALID, "cfg_unclip_step: null pointer");
  if (xin_dtype < 0 || xin_dtype > 2) return set_error(PCDM_ERR_INVALID, "cfg_unclip_step: bad dtype");
  if (n <= 0 || E <= 0 || ld_pred < E || ld_xin < E) return set_error(PCDM_ERR_INVALID, "cfg_unclip_step: bad shape");
  if (reinterpret_cast<uintptr_t>(coef_table) & 15) return set_error(PCDM_ERR_INVALID, "cfg_unclip_step: coef table must be 16-byte aligned");
  const long long total = (long long)n * E;
  PCDM_CUDA(launch_kernel(cfg_unclip_step_kernel, dim3(grid_for(total, 128)), dim3(128), 0, (cudaStream_t)stream_, 1, pred, ld_pred, latents, xin, xin_dtype, ld_xin, reinterpret_cast<const UnCLIPRow*>(coef_table), noise_table, step_counter, guidance_scale, use_cfg, n, E, t_table, t_cur));
  PCDM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pcdm_unclip_step(const void* model_output, int mo_dtype, const void* sample, const void* noise,
                                void* prev_sample, int dtype, const float* coef_row_host, long long numel,
                                void* stream_) {
  if (!model_output || !sample || !prev_sample || !coef_row_host) return set_error(PCDM_ERR_INVALID, "unclip_step: null pointer");
  if (mo_dtype < 0 || mo_dtype > 2 || dtype < 0 || dtype > 2) return set_error(PCDM_ERR_INVALID, "unclip_step: bad dtype");
  if (numel <= 0) return set_error(PCDM_ERR_INVALID, "unclip_step: empty problem");
  UnCLIPRow r;
  memcpy(&r, coef_row_host, sizeof(r));
  if (r.std != 0.f && !noise) return set_error(PCDM_ERR_INVALID, "unclip_step: this step adds noise (std != 0) but noise is NULL");
  PCDM_CUDA(launch_kernel(unclip_step_kernel, dim3(grid_for(numel, 256)), dim3(256), 0, (cudaStream_t)stream_, 1, model_output, mo_dtype, sample, noise, prev_sample, dtype, r, numel));
  PCDM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pcdm_softmax_rows(const float* x, long long ldx, void* y, long long ldy, int M, int N, float scale,
                                 int dtype, void* stream_) {
  if (!x || !y) return set_error(PCDM_ERR_INVALID, "softmax_rows: null pointer");
  if (dtype != DT_F16 && dtype != DT_BF16) return set_error(PCDM_ERR_INVALID, "softmax_rows: bad dtype");
  if (M <= 0 || N <= 0) return set_error(PCDM_ERR_INVALID, "softmax_rows: empty problem");
  if (N % 4 || N > 4 * SM_THREADS * SM_MAXV || (ldx % 4) || (ldy % 4))
    return set_error(PCDM_ERR_UNSUPPORTED, "softmax_rows: N % 4 == 0, N <= 16384, strides % 4 == 0");
  int grid = M < 8 * num_sms() ? M : 8 * num_sms();
  const float sl2 = scale * 1.4426950408889634f;
  if (dtype == DT_F16)
    PCDM_CUDA(launch_kernel(softmax_rows_kernel<DT_F16>, dim3(grid), dim3(SM_THREADS), 0, (cudaStream_t)stream_, 1, x, ldx, y, ldy, M, N, sl2));
  else
    PCDM_CUDA(launch_kernel(softmax_rows_kernel<DT_BF16>, dim3(grid), dim3(SM_THREADS), 0, (cudaStream_t)stream_, 1, x, ldx, y, ldy, M, N, sl2));
  PCDM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pcdm_gaussian_sample(const float* moments, long long ld, const float* noise, float* out, int B, int C,
                                    int HW, float scale, void* stream_) {
  if (!moments || !out) return set_error(PCDM_ERR_INVALID, "gaussian_sample: null pointer");
  if (B <= 0 || C <= 0 || HW <= 0 || ld < 2 * C) return set_error(PCDM_ERR_INVALID, "gaussian_sample: bad shape");
  const long long total = (long long)B * C * HW;
  PCDM_CUDA(launch_kernel(gaussian_sample_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream_, 1, moments, ld, noise, out, B, C, HW, scale));
  PCDM_CUDA(cudaGetLastError());
  return 0;
}
