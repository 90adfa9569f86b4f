// K4 GroupNorm(32)(+SiLU) and K5 LayerNorm on NHWC activations.  Both are HBM/L2-bandwidth bound streaming kernels:
// 16-byte vector loads, fp32 statistics (fp64 for the cross-CTA GroupNorm accumulation), one read for the statistics
// and one read + one write for the apply pass.
//
// GroupNorm replaces torch.nn.GroupNorm + SiLU in diffusers' ResnetBlock2D (norm1/norm2 + nonlinearity), the
// Transformer2DModel input norm (eps 1e-6, no activation) and conv_norm_out + conv_act
// (reference call site src/models/stage2_inpaint_unet_2d_condition.py:817-819).
// The input may be given as two channel segments [x1 | x2] — the skip concat of the up blocks is normalised
// straight from its two source tensors (groups may straddle the boundary; statistics are per channel first).
// LayerNorm replaces torch.nn.LayerNorm(norm1/2/3) of BasicTransformerBlock (row a8).
#include "common.cuh"
#include "host_util.h"

namespace pcdm {

// ---------------------------------------------------------------------------------------------------------------
// GroupNorm pass 1: per-(image, group) sum / sum-of-squares into stats[B][G][2] (double, pre-zeroed)
// block = (C/8, PY); thread (cv, py) owns 8 consecutive channels of pixels py, py+PY, ... within the CTA's chunk
// ---------------------------------------------------------------------------------------------------------------
template <int DT>
__global__ void gn_stats_kernel(const void* __restrict__ x1, const void* __restrict__ x2, int C1, int C, int HW,
                                int groups, int pix_per_cta, double* __restrict__ stats) {
  using T = typename TypeOf<DT>::T;
  extern __shared__ float sm[];  // [2][C] per-channel partials, then [2][groups]
  const int b = blockIdx.y;
  const int cv = threadIdx.x, py = threadIdx.y, PY = blockDim.y;
  const int c0 = cv * 8;
  const int p_begin = blockIdx.x * pix_per_cta;
  const int p_end = min(HW, p_begin + pix_per_cta);
  const T* src;
  int cs, coff;
  if (c0 < C1) { src = reinterpret_cast<const T*>(x1); cs = C1; coff = c0; }
  else { src = reinterpret_cast<const T*>(x2); cs = C - C1; coff = c0 - C1; }
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; }
  for (int p = p_begin + py; p < p_end; p += PY) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + ((size_t)b * HW + p) * cs + coff));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = unpack2<DT>(w[i]);
      s[2 * i] += f.x; q[2 * i] += f.x * f.x;
      s[2 * i + 1] += f.y; q[2 * i + 1] += f.y * f.y;
    }
  }
  float* ssum = sm;
  float* ssq = sm + C;
  for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < 2 * C; i += blockDim.x * blockDim.y) sm[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    atomicAdd(&ssum[c0 + i], s[i]);
    atomicAdd(&ssq[c0 + i], q[i]);
  }
  __syncthreads();
  const int cpg = C / groups;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if (tid < groups) {
    double a = 0.0, bq = 0.0;
    for (int c = tid * cpg; c < (tid + 1) * cpg; ++c) { a += (double)ssum[c]; bq += (double)ssq[c]; }
    atomicAdd(&stats[((size_t)b * groups + tid) * 2 + 0], a);
    atomicAdd(&stats[((size_t)b * groups + tid) * 2 + 1], bq);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// GroupNorm pass 2: y = (x - mean) * rstd * gamma + beta (+ SiLU); per-image per-channel scale/shift staged in smem
// ---------------------------------------------------------------------------------------------------------------
template <int DT>
__global__ void gn_apply_kernel(const void* __restrict__ x1, const void* __restrict__ x2, int C1, int C, int HW,
                                int groups, int pix_per_cta, const double* __restrict__ stats,
                                const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int silu,
                                void* __restrict__ y) {
  using T = typename TypeOf<DT>::T;
  extern __shared__ float sm[];  // scale[C], shift[C]
  float* scale = sm;
  float* shift = sm + C;
  const int b = blockIdx.y;
  const int cpg = C / groups;
  const double inv_n = 1.0 / ((double)cpg * (double)HW);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const double mean = stats[((size_t)b * groups + g) * 2 + 0] * inv_n;
    double var = stats[((size_t)b * groups + g) * 2 + 1] * inv_n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = rstd * gamma[c];
    scale[c] = sc;
    shift[c] = beta[c] - (float)mean * sc;
  }
  __syncthreads();
  const int cvs = C / 8;
  const int p_begin = blockIdx.x * pix_per_cta;
  const int p_end = min(HW, p_begin + pix_per_cta);
  const int total = (p_end - p_begin) * cvs;
  const int C2 = C - C1;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int p = p_begin + i / cvs;
    const int c0 = (i % cvs) * 8;
    const T* src = (c0 < C1) ? reinterpret_cast<const T*>(x1) + ((size_t)b * HW + p) * C1 + c0
                             : reinterpret_cast<const T*>(x2) + ((size_t)b * HW + p) * C2 + (c0 - C1);
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(src));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack2<DT>(w[k]);
      float a = f.x * scale[c0 + 2 * k] + shift[c0 + 2 * k];
      float bb = f.y * scale[c0 + 2 * k + 1] + shift[c0 + 2 * k + 1];
      if (silu) { a = silu_f(a); bb = silu_f(bb); }
      o[k] = pack2<DT>(a, bb);
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<T*>(y) + ((size_t)b * HW + p) * C + c0) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row kept in registers (two-pass variance), C <= 2048, C % 8 == 0
// ---------------------------------------------------------------------------------------------------------------
template <int DT, int VPL>  // VPL = 16-byte vectors per lane
__global__ void layernorm_kernel(const void* __restrict__ x, long long ldx, void* __restrict__ y, long long ldy,
                                 const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int M,
                                 int C) {
  using T = typename TypeOf<DT>::T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= M) return;
  const int nvec = C / 8;
  const T* xr = reinterpret_cast<const T*>(x) + row * ldx;
  float v[VPL][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(xr + vi * 8));
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = unpack2<DT>(w[k]);
        v[i][2 * k] = f.x; v[i][2 * k + 1] = f.y;
        sum += f.x + f.y;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[i][k] = 0.f;
    }
  }
  const float mean = warp_sum(sum) / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
#pragma unroll
      for (int k = 0; k < 8; ++k) { const float d = v[i][k] - mean; sq += d * d; }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / (float)C + eps);
  T* yr = reinterpret_cast<T*>(y) + row * ldy;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8 + 4));
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      uint32_t o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        o[k] = pack2<DT>((v[i][2 * k] - mean) * rstd * g[2 * k] + bb[2 * k],
                         (v[i][2 * k + 1] - mean) * rstd * g[2 * k + 1] + bb[2 * k + 1]);
      *reinterpret_cast<uint4*>(yr + vi * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

}  // namespace pcdm

using namespace pcdm;

extern "C" long long pcdm_groupnorm_workspace_bytes(int B, int groups) { return (long long)B * groups * 2 * 8; }

extern "C" int pcdm_groupnorm(const void* x1, const void* x2, int C1, void* y, const float* gamma, const float* beta,
                              float eps, int B, int HW, int C, int groups, int dtype, int flags, void* workspace,
                              void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!x1 || !y || !gamma || !beta || !workspace) return set_error(PCDM_ERR_INVALID, "groupnorm: null pointer");
  if (dtype != DT_F16 && dtype != DT_BF16) return set_error(PCDM_ERR_INVALID, "groupnorm: bad dtype");
  if (B <= 0 || HW <= 0 || C <= 0 || groups <= 0) return set_error(PCDM_ERR_INVALID, "groupnorm: empty problem");
  if (C % groups || C % 8 || C > 4096) return set_error(PCDM_ERR_UNSUPPORTED, "groupnorm: C must be a multiple of groups and 8, <= 4096");
  if (!x2) C1 = C;
  if (C1 % 8 || C1 <= 0 || C1 > C) return set_error(PCDM_ERR_INVALID, "groupnorm: bad channel split");
  if (groups > 256) return set_error(PCDM_ERR_UNSUPPORTED, "groupnorm: groups > 256");
  double* stats = reinterpret_cast<double*>(workspace);
  PCDM_CUDA(cudaMemsetAsync(stats, 0, (size_t)B * groups * 2 * sizeof(double), stream));
  const int cvs = C / 8;
  int PY = 512 / cvs;
  if (PY < 1) PY = 1;
  // ~4096 elements per thread-row pass; keep >= 2 waves of CTAs when the tensor is big
  int pix_per_cta = (64 * 1024) / C;
  if (pix_per_cta < PY) pix_per_cta = PY;
  if (pix_per_cta > HW) pix_per_cta = HW;
  const int chunks = (HW + pix_per_cta - 1) / pix_per_cta;
  const dim3 grid(chunks, B);
  const size_t smem = (size_t)2 * C * sizeof(float);
  if (dtype == DT_F16) {
    gn_stats_kernel<DT_F16><<<grid, dim3(cvs, PY), smem, stream>>>(x1, x2, C1, C, HW, groups, pix_per_cta, stats);
    gn_apply_kernel<DT_F16><<<grid, 256, smem, stream>>>(x1, x2, C1, C, HW, groups, pix_per_cta, stats, gamma, beta,
                                                          eps, (flags & PCDM_FLAG_SILU) ? 1 : 0, y);
  } else {
    gn_stats_kernel<DT_BF16><<<grid, dim3(cvs, PY), smem, stream>>>(x1, x2, C1, C, HW, groups, pix_per_cta, stats);
    gn_apply_kernel<DT_BF16><<<grid, 256, smem, stream>>>(x1, x2, C1, C, HW, groups, pix_per_cta, stats, gamma, beta,
                                                           eps, (flags & PCDM_FLAG_SILU) ? 1 : 0, y);
  }
  PCDM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pcdm_layernorm(const void* x, long long ldx, void* y, long long ldy, const float* gamma,
                              const float* beta, float eps, int M, int C, int dtype, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!x || !y || !gamma || !beta) return set_error(PCDM_ERR_INVALID, "layernorm: null pointer");
  if (dtype != DT_F16 && dtype != DT_BF16) return set_error(PCDM_ERR_INVALID, "layernorm: bad dtype");
  if (M <= 0 || C <= 0) return set_error(PCDM_ERR_INVALID, "layernorm: empty problem");
  if (C % 8 || C > 2048 || (ldx % 8) || (ldy % 8)) return set_error(PCDM_ERR_UNSUPPORTED, "layernorm: C % 8 == 0, C <= 2048, strides % 8 == 0");
  const int vpl = (C / 8 + 31) / 32;
  const int rows_per_cta = 8;
  const int grid = (M + rows_per_cta - 1) / rows_per_cta;
#define LN_LAUNCH(V)                                                                                              \
  do {                                                                                                            \
    if (dtype == DT_F16) layernorm_kernel<DT_F16, V><<<grid, 256, 0, stream>>>(x, ldx, y, ldy, gamma, beta, eps, M, C); \
    else layernorm_kernel<DT_BF16, V><<<grid, 256, 0, stream>>>(x, ldx, y, ldy, gamma, beta, eps, M, C);          \
  } while (0)
  switch (vpl) {
    case 1: LN_LAUNCH(1); break;
    case 2: LN_LAUNCH(2); break;
    case 3: LN_LAUNCH(3); break;
    case 4: LN_LAUNCH(4); break;
    case 5: LN_LAUNCH(5); break;
    case 6: LN_LAUNCH(6); break;
    case 7: LN_LAUNCH(7); break;
    default: LN_LAUNCH(8); break;
  }
#undef LN_LAUNCH
  PCDM_CUDA(cudaGetLastError());
  return 0;
}
