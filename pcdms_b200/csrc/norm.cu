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
// GroupNorm pass 1: statistics.  block = (C/8, PY); thread (cv, py) owns 8 consecutive channels of pixels py, py+PY, ...
// of the CTA's pixel chunk.  Reduction order is fixed at every level (registers -> smem over py -> channels of a group
// -> chunks of an image), so results are bit-reproducible run to run: no floating-point atomics anywhere.  The last
// CTA of an image (device counter) folds the per-chunk partials into (mean, rstd) per group.
// ---------------------------------------------------------------------------------------------------------------
struct GnWorkspace {
  float2* final_stats;   // [B][groups] (mean, rstd)
  unsigned* counters;    // [B], zero between launches
  double2* partial;      // [B][chunks][groups] (sum, sumsq)
};

__host__ __device__ inline size_t gn_align(size_t x) { return (x + 255) / 256 * 256; }

template <int DT>
__global__ void gn_stats_kernel(const void* __restrict__ x1, const void* __restrict__ x2, int C1, int C, int HW,
                                int groups, int pix_per_cta, float eps, GnWorkspace ws) {
  pdl_launch_dependents();
  pdl_wait();
  using T = typename TypeOf<DT>::T;
  extern __shared__ float sm[];  // [2][PY][C] per-(py, channel) partials; then [2][C] per-channel sums
  const int b = blockIdx.y;
  const int cv = threadIdx.x, py = threadIdx.y, PY = blockDim.y;
  const int nthreads = blockDim.x * blockDim.y;
  const int tid = py * blockDim.x + cv;
  const int c0 = cv * 8;
  const int p_begin = blockIdx.x * pix_per_cta;
  const int p_end = min(HW, p_begin + pix_per_cta);
  const T* src;
  int cs, coff;
  if (c0 < C1) { src = reinterpret_cast<const T*>(x1); cs = C1; coff = c0; }
  else { src = reinterpret_cast<const T*>(x2); cs = C - C1; coff = c0 - C1; }
  // four pixels in flight per thread; fp32 sums on the packed FFMA2 path (one add + one fma per 32-bit word)
  float2 s2[4], q2[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { s2[k] = make_float2(0.f, 0.f); q2[k] = make_float2(0.f, 0.f); }
  const T* sp = src + (size_t)b * HW * cs + coff;
  for (int p = p_begin + py; p < p_end; p += 4 * PY) {
    uint4 u[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      u[j] = (p + j * PY < p_end) ? __ldg(reinterpret_cast<const uint4*>(sp + (size_t)(p + j * PY) * cs))
                                  : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t w[4] = {u[j].x, u[j].y, u[j].z, u[j].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = unpack2<DT>(w[k]);
        s2[k] = __fadd2_rn(s2[k], f);
        q2[k] = __ffma2_rn(f, f, q2[k]);
      }
    }
  }
  const float s[8] = {s2[0].x, s2[0].y, s2[1].x, s2[1].y, s2[2].x, s2[2].y, s2[3].x, s2[3].y};
  const float q[8] = {q2[0].x, q2[0].y, q2[1].x, q2[1].y, q2[2].x, q2[2].y, q2[3].x, q2[3].y};
  float* ps = sm;                   // [PY][C]
  float* pq = sm + (size_t)PY * C;  // [PY][C]
  *reinterpret_cast<float4*>(ps + (size_t)py * C + c0) = make_float4(s[0], s[1], s[2], s[3]);
  *reinterpret_cast<float4*>(ps + (size_t)py * C + c0 + 4) = make_float4(s[4], s[5], s[6], s[7]);
  *reinterpret_cast<float4*>(pq + (size_t)py * C + c0) = make_float4(q[0], q[1], q[2], q[3]);
  *reinterpret_cast<float4*>(pq + (size_t)py * C + c0 + 4) = make_float4(q[4], q[5], q[6], q[7]);
  __syncthreads();
  // one warp per group: lanes stride over the group's (py, channel) partials in a fixed order, then a shuffle tree —
  // no thread ever walks a long dependent chain (the tail of this kernel is pure latency)
  const int cpg = C / groups;
  const int chunks = gridDim.x;
  const int warp = tid >> 5, lane = tid & 31, full_warps = nthreads >> 5;   // host guarantees nthreads >= 32
  if (warp < full_warps) {
    for (int g = warp; g < groups; g += full_warps) {
      double a = 0.0, bq = 0.0;
      for (int idx = lane; idx < PY * cpg; idx += 32) {
        const int yy = idx / cpg, c = g * cpg + idx - yy * cpg;
        a += (double)ps[(size_t)yy * C + c];
        bq += (double)pq[(size_t)yy * C + c];
      }
      a = warp_sum(a);
      bq = warp_sum(bq);
      if (lane == 0) ws.partial[((size_t)b * chunks + blockIdx.x) * groups + g] = make_double2(a, bq);
    }
  }
  __threadfence();
  __syncthreads();
  __shared__ bool is_last;
  if (tid == 0) is_last = (atomicAdd(&ws.counters[b], 1u) == (unsigned)(chunks - 1));
  __syncthreads();
  if (is_last) {
    __threadfence();
    if (warp < full_warps) {
      for (int g0 = warp; g0 < groups; g0 += 4 * full_warps) {   // four groups' loads in flight per warp
        double a[4], bq[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int g = g0 + j * full_warps;
          a[j] = 0.0; bq[j] = 0.0;
          if (g < groups) {
            for (int k = lane; k < chunks; k += 32) {
              const double2 v = ws.partial[((size_t)b * chunks + k) * groups + g];
              a[j] += v.x; bq[j] += v.y;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int g = g0 + j * full_warps;
          const double sa = warp_sum(a[j]), sq = warp_sum(bq[j]);
          if (g < groups && lane == 0) {
            const double inv_n = 1.0 / ((double)cpg * (double)HW);
            const double mean = sa * inv_n;
            double var = sq * inv_n - mean * mean;
            if (var < 0.0) var = 0.0;
            ws.final_stats[(size_t)b * groups + g] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
          }
        }
      }
    }
    if (tid == 0) ws.counters[b] = 0;  // ready for the next launch (and for CUDA-graph replays)
  }
}

// ---------------------------------------------------------------------------------------------------------------
// GroupNorm pass 2: y = (x - mean) * rstd * gamma + beta (+ SiLU).  Same (C/8, PY) thread layout as pass 1: a thread
// owns 8 fixed channels, so its 8 scale/shift pairs live in registers and the pixel loop is pure streaming.
// ---------------------------------------------------------------------------------------------------------------
template <int DT>
__global__ void gn_apply_kernel(const void* __restrict__ x1, const void* __restrict__ x2, int C1, int C, int HW,
                                int groups, int pix_per_cta, const float2* __restrict__ final_stats,
                                const float* __restrict__ gamma, const float* __restrict__ beta, int silu,
                                void* __restrict__ y) {
  pdl_launch_dependents();
  using T = typename TypeOf<DT>::T;
  const int b = blockIdx.y;
  const int cv = threadIdx.x, py = threadIdx.y, PY = blockDim.y;
  const int c0 = cv * 8;
  const int cpg = C / groups;
  // Per-channel scale / shift ONCE PER CTA, through shared memory (round 2: every thread used to derive its eight pairs
  // itself — eight integer divisions and sixteen dependent global loads in front of a loop of only ~10 pixels per
  // thread).  gamma / beta are parameters: fetched before the wait on the statistics kernel.
  extern __shared__ __align__(16) float gn_apply_smem[];   // [C] scale | [C] shift
  float* s_sc = gn_apply_smem;
  float* s_sh = gn_apply_smem + C;
  const int tid = py * blockDim.x + cv, nthr = blockDim.x * PY;
  float gm_r[2], bt_r[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = tid + i * nthr;
    gm_r[i] = c < C ? __ldg(gamma + c) : 0.f;
    bt_r[i] = c < C ? __ldg(beta + c) : 0.f;
  }
  pdl_wait();
  const float2* stats_b = final_stats + (size_t)b * groups;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = tid + i * nthr;
    if (c < C) {
      const float2 st = stats_b[c / cpg];
      const float sc = st.y * gm_r[i];
      s_sc[c] = sc;
      s_sh[c] = bt_r[i] - st.x * sc;
    }
  }
  for (int c = tid + 2 * nthr; c < C; c += nthr) {   // C > 2 x threads: not a shape of this network
    const float2 st = stats_b[c / cpg];
    const float sc = st.y * __ldg(gamma + c);
    s_sc[c] = sc;
    s_sh[c] = __ldg(beta + c) - st.x * sc;
  }
  __syncthreads();
  float2 sc2[4], sh2[4];
  {
    const float4 a0 = *reinterpret_cast<const float4*>(s_sc + c0), a1 = *reinterpret_cast<const float4*>(s_sc + c0 + 4);
    const float4 h0 = *reinterpret_cast<const float4*>(s_sh + c0), h1 = *reinterpret_cast<const float4*>(s_sh + c0 + 4);
    sc2[0] = make_float2(a0.x, a0.y); sc2[1] = make_float2(a0.z, a0.w);
    sc2[2] = make_float2(a1.x, a1.y); sc2[3] = make_float2(a1.z, a1.w);
    sh2[0] = make_float2(h0.x, h0.y); sh2[1] = make_float2(h0.z, h0.w);
    sh2[2] = make_float2(h1.x, h1.y); sh2[3] = make_float2(h1.z, h1.w);
  }
  const T* src;
  int cs, coff;
  if (c0 < C1) { src = reinterpret_cast<const T*>(x1); cs = C1; coff = c0; }
  else { src = reinterpret_cast<const T*>(x2); cs = C - C1; coff = c0 - C1; }
  const int p_begin = blockIdx.x * pix_per_cta;
  const int p_end = min(HW, p_begin + pix_per_cta);
  const T* sp = src + (size_t)b * HW * cs + coff;
  T* dp = reinterpret_cast<T*>(y) + (size_t)b * HW * C + c0;
  for (int p = p_begin + py; p < p_end; p += 4 * PY) {   // four pixels in flight per thread
    uint4 u[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (p + j * PY < p_end) u[j] = __ldg(reinterpret_cast<const uint4*>(sp + (size_t)(p + j * PY) * cs));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (p + j * PY < p_end) {
        const uint32_t w[4] = {u[j].x, u[j].y, u[j].z, u[j].w};
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float2 v = __ffma2_rn(unpack2<DT>(w[k]), sc2[k], sh2[k]);
          if (silu) v = silu2_f<DT>(v);
          o[k] = pack2<DT>(v.x, v.y);
        }
        *reinterpret_cast<uint4*>(dp + (size_t)(p + j * PY) * C) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// GroupNorm, single pass (the path every SD-2.1 shape takes): one read + one write of the activation.
//
// A group's statistics only involve that group's channels, so the problem splits into independent (image, group-set)
// pieces.  A piece — HW pixels x (gset groups = nv 16-byte vectors per pixel) — is small enough to live in the
// REGISTERS of one CTA, or of a cluster of S <= 8 CTAs that split its pixels: every thread keeps its K <= 6 vectors,
// the block reduces (sum, sum of squares) per group in a fixed order, the cluster folds its S partials through
// distributed shared memory (rank order, fp64), and the same registers are then normalised (+SiLU) and stored.
// No second read, no cross-CTA atomics, bit-reproducible.  block = (nv, PY); thread (cv, py) owns channels
// [cset + 8 cv, +8) of pixels p_begin + py + i PY.  A 16-byte vector touches at most two groups (cpg >= 8).
// ---------------------------------------------------------------------------------------------------------------
template <int DT, int K>
__global__ void __launch_bounds__(1024, 1)
gn_fused_kernel(const void* __restrict__ x1, const void* __restrict__ x2, int C1, int C, int HW, int cpg, int gset,
                int pix_per_cta, int S, float eps, const float* __restrict__ gamma, const float* __restrict__ beta,
                int silu, void* __restrict__ y) {
  pdl_launch_dependents();
  using T = typename TypeOf<DT>::T;
  extern __shared__ __align__(16) unsigned char gsm[];
  const int nv = blockDim.x, PY = blockDim.y;
  const int cv = threadIdx.x, py = threadIdx.y;
  const int nthreads = nv * PY, tid = py * nv + cv;
  float* red = reinterpret_cast<float*>(gsm);            // [4][nthreads]: s_lo, q_lo, s_hi, q_hi
  float* gam = red + 4 * (size_t)nthreads;               // [nv * 8]
  float* bet = gam + nv * 8;                             // [nv * 8]
  double2* cl = reinterpret_cast<double2*>(bet + nv * 8);  // [8] this CTA's (sum, sumsq) per group of the set
  float2* stat = reinterpret_cast<float2*>(cl + 8);        // [8] (mean, rstd)
  const int rank = blockIdx.x, set = blockIdx.y, b = blockIdx.z;
  const int cset = set * gset * cpg;
  const int crel = cv * 8;
  const int c0 = cset + crel;
  for (int i = tid; i < nv * 8; i += nthreads) {   // parameters do not depend on the producer kernel: before the wait
    gam[i] = __ldg(gamma + cset + i);
    bet[i] = __ldg(beta + cset + i);
  }
  pdl_wait();
  const T* src;
  int cs, coff;
  if (c0 < C1) { src = reinterpret_cast<const T*>(x1); cs = C1; coff = c0; }
  else { src = reinterpret_cast<const T*>(x2); cs = C - C1; coff = c0 - C1; }
  const int p_begin = rank * pix_per_cta;
  const int p_end = min(HW, p_begin + pix_per_cta);
  const T* sp = src + ((size_t)b * HW + p_begin + py) * cs + coff;
  uint4 u[K];
#pragma unroll
  for (int i = 0; i < K; ++i) {
    if (p_begin + py + i * PY < p_end) u[i] = __ldg(reinterpret_cast<const uint4*>(sp + (size_t)i * PY * cs));
    else u[i] = make_uint4(0u, 0u, 0u, 0u);   // zeros add nothing to either sum
  }
  const int glo = crel / cpg;                         // group (within the set) of this thread's first channel
  const int split = min(8, (glo + 1) * cpg - crel);   // elements [0, split) belong to glo, [split, 8) to glo + 1
  // These kernels are issue-bound before they are bandwidth-bound (an SM's share of HBM is ~12 elements per clock),
  // so the fp32 arithmetic runs on the packed FFMA2 path: per 32-bit word, one add and one fma for both channels.
  float2 s2[4], q2[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { s2[k] = make_float2(0.f, 0.f); q2[k] = make_float2(0.f, 0.f); }
#pragma unroll
  for (int i = 0; i < K; ++i) {
    const uint32_t w[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack2<DT>(w[k]);
      s2[k] = __fadd2_rn(s2[k], f);
      q2[k] = __ffma2_rn(f, f, q2[k]);
    }
  }
  float s_lo = 0.f, q_lo = 0.f, s_hi = 0.f, q_hi = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {   // channel position e = 2k, 2k+1 of the vector -> its group (fixed order)
    if (2 * k < split) { s_lo += s2[k].x; q_lo += q2[k].x; } else { s_hi += s2[k].x; q_hi += q2[k].x; }
    if (2 * k + 1 < split) { s_lo += s2[k].y; q_lo += q2[k].y; } else { s_hi += s2[k].y; q_hi += q2[k].y; }
  }
  red[tid] = s_lo;
  red[nthreads + tid] = q_lo;
  red[2 * nthreads + tid] = s_hi;
  red[3 * nthreads + tid] = q_hi;
  __syncthreads();
  {
    const int warp = tid >> 5, lane = tid & 31, full_warps = nthreads >> 5;   // host guarantees nthreads >= 32
    if (warp < full_warps) {
      for (int j = warp; j < gset; j += full_warps) {
        const int cfirst = (j * cpg) >> 3, clast = ((j + 1) * cpg - 1) >> 3;
        double a = 0.0, q = 0.0;
        for (int c = cfirst; c <= clast; ++c) {
          const float* ps = red + (((c * 8) / cpg == j) ? 0 : 2 * nthreads);   // j is that column's lo or hi group
          const float* pq = ps + nthreads;
          for (int yy = lane; yy < PY; yy += 32) { a += (double)ps[yy * nv + c]; q += (double)pq[yy * nv + c]; }
        }
        a = warp_sum(a);
        q = warp_sum(q);
        if (lane == 0) cl[j] = make_double2(a, q);
      }
    }
  }
  if (S > 1) cluster_sync_all(); else __syncthreads();
  if (tid < 32) {   // warp 0: lane (g, r) fetches rank r's partial of group g, S lanes fold in a fixed xor tree
    for (int base = 0; base < gset * S; base += 32) {
      const int item = base + tid, g = item / S, r = item - g * S;
      double a = 0.0, q = 0.0;
      if (item < gset * S) {
        if (S > 1) {
          const double2 v = ld_dsmem_f64x2(mapa_u32(smem_u32(&cl[g]), (uint32_t)r));
          a = v.x; q = v.y;
        } else {
          a = cl[g].x; q = cl[g].y;
        }
      }
      for (int o = 1; o < S; o <<= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
      }
      if (item < gset * S && r == 0) {
        const double inv_n = 1.0 / ((double)cpg * (double)HW);
        const double mean = a * inv_n;
        double var = q * inv_n - mean * mean;
        if (var < 0.0) var = 0.0;
        stat[g] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
      }
    }
  }
  __syncthreads();
  float2 sc2[4], sh2[4];
  {
    const float2 st_lo = stat[glo], st_hi = stat[min(glo + 1, gset - 1)];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 sa = (2 * k < split) ? st_lo : st_hi, sb = (2 * k + 1 < split) ? st_lo : st_hi;
      const float2 g = *reinterpret_cast<const float2*>(gam + crel + 2 * k);
      const float2 bt = *reinterpret_cast<const float2*>(bet + crel + 2 * k);
      sc2[k] = make_float2(sa.y * g.x, sb.y * g.y);
      sh2[k] = make_float2(bt.x - sa.x * sc2[k].x, bt.y - sb.x * sc2[k].y);
    }
  }
  T* dp = reinterpret_cast<T*>(y) + ((size_t)b * HW + p_begin + py) * C + c0;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    if (p_begin + py + i * PY < p_end) {
      const uint32_t w[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
      uint32_t o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 v = __ffma2_rn(unpack2<DT>(w[k]), sc2[k], sh2[k]);
        if (silu) v = silu2_f<DT>(v);
        o[k] = pack2<DT>(v.x, v.y);
      }
      *reinterpret_cast<uint4*>(dp + (size_t)i * PY * C) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
  if (S > 1) cluster_sync_all();   // cl[] must outlive every peer's read of it
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm.  LPR lanes share a row (32 / LPR rows per warp) and each lane keeps VPL 16-byte vectors of it, chosen so
// that LPR * VPL covers C / 8 with the least idle lanes (C = 320, 640, 1280 -> 8, 16, 32 lanes x 5 vectors exactly).
// Row groups are strided over a grid of ~4 CTAs per SM; gamma/beta are staged in shared memory before the
// programmatic-dependency wait (they are parameters); the NEXT row group of the warp is already in flight while the
// current one is reduced (two-pass variance, xor-shuffles within the LPR lanes), normalised and stored.  The fp32
// arithmetic is on the packed FFMA2 path: the kernel is issue-bound before it is bandwidth-bound.
// ---------------------------------------------------------------------------------------------------------------
template <int LPR>
__device__ __forceinline__ float row_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int DT, int VPL, int LPR>
__global__ void __launch_bounds__(256)
layernorm_kernel(const void* __restrict__ x, long long ldx, void* __restrict__ y, long long ldy,
                 const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int M, int C) {
  pdl_launch_dependents();
  using T = typename TypeOf<DT>::T;
  constexpr int RPW = 32 / LPR;                  // rows per warp
  extern __shared__ __align__(16) float lsm[];   // gamma[C] | beta[C]
  for (int i = threadIdx.x; i < C / 4; i += blockDim.x) {
    reinterpret_cast<float4*>(lsm)[i] = __ldg(reinterpret_cast<const float4*>(gamma) + i);
    reinterpret_cast<float4*>(lsm + C)[i] = __ldg(reinterpret_cast<const float4*>(beta) + i);
  }
  pdl_wait();
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane % LPR;
  const int wpc = blockDim.x >> 5;
  const long long stride = (long long)gridDim.x * wpc * RPW;
  const int nvec = C / 8;
  const T* xb = reinterpret_cast<const T*>(x);
  T* yb = reinterpret_cast<T*>(y);
  long long row = ((long long)blockIdx.x * wpc + warp) * RPW + lane / LPR;
  uint4 cur[VPL], nxt[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i)
    cur[i] = (row < M && sub + i * LPR < nvec) ? __ldg(reinterpret_cast<const uint4*>(xb + row * ldx + (sub + i * LPR) * 8))
                                               : make_uint4(0u, 0u, 0u, 0u);
  const float inv_c = 1.0f / (float)C;
  // every lane of the warp runs the same number of iterations (the shuffles need all 32): bound by the warp's first row
  for (long long base = row - lane / LPR; base < M; base += stride, row += stride) {
    const long long rn = row + stride;
#pragma unroll
    for (int i = 0; i < VPL; ++i)
      nxt[i] = (rn < M && sub + i * LPR < nvec) ? __ldg(reinterpret_cast<const uint4*>(xb + rn * ldx + (sub + i * LPR) * 8))
                                                : make_uint4(0u, 0u, 0u, 0u);
    float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {   // out-of-range vectors are zeros
      const uint32_t w[4] = {cur[i].x, cur[i].y, cur[i].z, cur[i].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) sum2 = __fadd2_rn(sum2, unpack2<DT>(w[k]));
    }
    const float mean = row_sum<LPR>(sum2.x + sum2.y) * inv_c;
    const float2 nmean2 = make_float2(-mean, -mean);
    float2 sq2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (sub + i * LPR < nvec) {
        const uint32_t w[4] = {cur[i].x, cur[i].y, cur[i].z, cur[i].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 d = __fadd2_rn(unpack2<DT>(w[k]), nmean2);
          sq2 = __ffma2_rn(d, d, sq2);
        }
      }
    }
    const float rstd = rsqrtf(row_sum<LPR>(sq2.x + sq2.y) * inv_c + eps);
    const float2 rstd2 = make_float2(rstd, rstd), shift2 = make_float2(-mean * rstd, -mean * rstd);
    if (row < M) {
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int vi = sub + i * LPR;
        if (vi < nvec) {
          const float4 g0 = *reinterpret_cast<const float4*>(lsm + vi * 8);
          const float4 g1 = *reinterpret_cast<const float4*>(lsm + vi * 8 + 4);
          const float4 b0 = *reinterpret_cast<const float4*>(lsm + C + vi * 8);
          const float4 b1 = *reinterpret_cast<const float4*>(lsm + C + vi * 8 + 4);
          const float2 g[4] = {make_float2(g0.x, g0.y), make_float2(g0.z, g0.w), make_float2(g1.x, g1.y), make_float2(g1.z, g1.w)};
          const float2 bb[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
          const uint32_t w[4] = {cur[i].x, cur[i].y, cur[i].z, cur[i].w};
          uint32_t o[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 t = __ffma2_rn(unpack2<DT>(w[k]), rstd2, shift2);   // (x - mean) * rstd
            const float2 v = __ffma2_rn(t, g[k], bb[k]);
            o[k] = pack2<DT>(v.x, v.y);
          }
          *reinterpret_cast<uint4*>(yb + row * ldy + vi * 8) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) cur[i] = nxt[i];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// GroupNorm statistics from the PRODUCER's epilogue (pcdm_ext.chan_stats: per 32-row slab and channel, (sum, sum of
// squares) of the stored values) -> (mean, rstd) per (image, group).  One 128-thread CTA per (image, group): threads
// stride over the group's (slab, channel) partials in a fixed order, fp64 accumulation, fixed tree.  The activation itself is
// not read: the statistics pass of GroupNorm has disappeared into the conv / GEMM that wrote the tensor.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) gn_fold_kernel(const float2* __restrict__ st1, const float2* __restrict__ st2,
                                                      int C1, int C, int slabs_per_image, int groups, int B, int HW,
                                                      float eps, float2* __restrict__ final_stats) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double red[2][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x;                 // one CTA per (image, group): the fold is latency-, not bandwidth-bound
  const int b = item / groups, g = item - b * groups;
  const int cpg = C / groups, C2 = C - C1;
  const int total = slabs_per_image * cpg;
  double a = 0.0, q = 0.0;
  for (int base = 0; base < total; base += 4 * 128) {   // four independent loads in flight per thread
    float2 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = base + j * 128 + (int)threadIdx.x;
      v[j] = make_float2(0.f, 0.f);
      if (idx < total) {
        const int sl = idx / cpg, c = g * cpg + (idx - sl * cpg);
        const long long slab = (long long)b * slabs_per_image + sl;
        v[j] = c < C1 ? __ldg(st1 + slab * C1 + c) : __ldg(st2 + slab * C2 + (c - C1));
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { a += (double)v[j].x; q += (double)v[j].y; }
  }
  a = warp_sum(a);
  q = warp_sum(q);
  if (lane == 0) { red[0][warp] = a; red[1][warp] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double sa = ((red[0][0] + red[0][1]) + red[0][2]) + red[0][3];
    const double sq = ((red[1][0] + red[1][1]) + red[1][2]) + red[1][3];
    const double inv_n = 1.0 / ((double)cpg * (double)HW);
    const double mean = sa * inv_n;
    double var = sq * inv_n - mean * mean;
    if (var < 0.0) var = 0.0;
    final_stats[item] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
  }
}

// Per-row (sum, sum of squares) of a [M, C] 16-bit matrix: the statistics slot a GEMM with a folded LayerNorm consumes
// (pcdm_ext.ln_stats with ln_parts = 1) when the rows were NOT produced by one of this library's GEMM epilogues (the
// custom-attention-processor path).  One warp per row, fixed reduction order.
template <int DT>
__global__ void __launch_bounds__(256) row_stats_kernel(const void* __restrict__ x, long long ldx, float2* __restrict__ out,
                                                        int M, int C) {
  pdl_launch_dependents();
  pdl_wait();
  using T = typename TypeOf<DT>::T;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const T* xr = reinterpret_cast<const T*>(x) + row * ldx;
  float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
  for (int v = lane; v < C / 8; v += 32) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(xr + v * 8));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack2<DT>(w[k]);
      s2 = __fadd2_rn(s2, f);
      q2 = __ffma2_rn(f, f, q2);
    }
  }
  const float s = warp_sum(s2.x + s2.y), q = warp_sum(q2.x + q2.y);
  if (lane == 0) out[row] = make_float2(s, q);
}

}  // namespace pcdm

using namespace pcdm;

extern "C" int pcdm_row_stats(const void* x, long long ldx, float* stats, int M, int C, int dtype, void* stream_) {
  if (!x || !stats) return set_error(PCDM_ERR_INVALID, "row_stats: null pointer");
  if (dtype != DT_F16 && dtype != DT_BF16) return set_error(PCDM_ERR_INVALID, "row_stats: bad dtype");
  if (M <= 0 || C <= 0) return set_error(PCDM_ERR_INVALID, "row_stats: empty problem");
  if (C % 8 || (ldx % 8) || (reinterpret_cast<uintptr_t>(stats) & 7))
    return set_error(PCDM_ERR_UNSUPPORTED, "row_stats: C % 8 == 0, ldx % 8 == 0, 8-byte aligned output");
  const dim3 grid((M + 7) / 8), block(256);
  if (dtype == DT_F16)
    PCDM_CUDA(launch_kernel(row_stats_kernel<DT_F16>, grid, block, 0, (cudaStream_t)stream_, 1, x, ldx, reinterpret_cast<float2*>(stats), M, C));
  else
    PCDM_CUDA(launch_kernel(row_stats_kernel<DT_BF16>, grid, block, 0, (cudaStream_t)stream_, 1, x, ldx, reinterpret_cast<float2*>(stats), M, C));
  PCDM_CUDA(cudaGetLastError());
  return 0;
}

static void gn_grid(int B, int HW, int C, int* PY_, int* pix_per_cta_, int* chunks_) {
  const int cvs = C / 8;
  int PY = 512 / cvs;
  if (PY < 1) PY = 1;
  // ~2 CTAs of ~480 threads per SM across the batch, four 16-byte loads in flight per thread (~60 KB per SM); keep
  // at least one pixel per thread row
  int chunks_wanted = (2 * num_sms()) / B;   // rounded DOWN: one CTA over the resident capacity is a whole second wave
  if (chunks_wanted < 1) chunks_wanted = 1;
  int pix_per_cta = (HW + chunks_wanted - 1) / chunks_wanted;
  if (pix_per_cta < PY) pix_per_cta = PY;
  if (pix_per_cta > HW) pix_per_cta = HW;
  if (PY > pix_per_cta) PY = pix_per_cta;
  while (cvs * PY < 32) ++PY;   // the group reductions are warp-wide: at least one full warp
  *PY_ = PY;
  *pix_per_cta_ = pix_per_cta;
  *chunks_ = (HW + pix_per_cta - 1) / pix_per_cta;
}

struct GnFusedCfg { int gset, nv, PY, S, pix_per_cta, K; };

// Shape of the single-pass launch, or false when the problem needs the two-kernel path (cpg < 8, pieces too large).
// Measured (tools/dev_norm_perf.py): the single pass wins while the whole activation is a wave or two of CTAs
// (<= 16 MB: one launch, one read); above that its load / reduce+cluster-sync / store phases run in lockstep across
// the chip and the two streaming kernels (statistics, then apply) are faster.
static bool gn_fused_config(int B, int HW, int C, int C1, int groups, GnFusedCfg* c) {
  const int cpg = C / groups;
  if (cpg < 8 || (C1 % 8)) return false;
  int gmin = 1;
  while (gmin <= 8 && (gmin * cpg) % 8) gmin *= 2;
  if (gmin > 8 || groups % gmin) return false;
  static const int kThreads[3] = {1024, 512, 256};
  for (int ti = 0; ti < 3; ++ti) {
    const int T = kThreads[ti];
    if (g_tune.gn_mode > 2 && T != g_tune.gn_mode - 2) continue;   // experiment hook: force T threads per CTA
    for (int gset = 8; gset >= gmin; gset /= 2) {   // wider sets = longer contiguous runs per pixel
      if (gset % gmin || groups % gset) continue;
      const int nv = gset * cpg / 8;
      if (nv > 64 || (gset > gmin && (long long)B * (groups / gset) < 128)) continue;
      const int PYmax = T / nv;
      if (PYmax < 1) continue;
      for (int S = 1; S <= 8; S *= 2) {
        const int ppc = (HW + S - 1) / S;
        int PY = ppc < PYmax ? ppc : PYmax;
        const int PYmin = (32 + nv - 1) / nv;
        if (PY < PYmin) PY = PYmin;
        if (PY * nv > 1024) break;
        const int K = (ppc + PY - 1) / PY;
        if (K > 6) continue;
        // spread a small batch over more CTAs while every thread keeps >= 2 vectors
        if ((long long)B * (groups / gset) * S < 128 && S < 8 && K >= 4) continue;
        c->gset = gset; c->nv = nv; c->PY = PY; c->S = S; c->pix_per_cta = ppc;
        c->K = K <= 4 ? K : 6;
        return true;
      }
    }
  }
  return false;
}

template <int DT>
static cudaError_t launch_gn_fused(const GnFusedCfg& c, cudaStream_t stream, const void* x1, const void* x2, int C1,
                                   int C, int HW, int groups, int B, float eps, const float* gamma, const float* beta,
                                   int silu, void* y) {
  const dim3 grid(c.S, groups / c.gset, B), block(c.nv, c.PY);
  const size_t smem = (size_t)16 * c.nv * c.PY + (size_t)64 * c.nv + 8 * sizeof(double2) + 8 * sizeof(float2);
  const int cpg = C / groups;
#define GN_FUSED(KK)                                                                                              \
  return launch_kernel(gn_fused_kernel<DT, KK>, grid, block, smem, stream, c.S, x1, x2, C1, C, HW, cpg, c.gset,   \
                       c.pix_per_cta, c.S, eps, gamma, beta, silu, y)
  switch (c.K) {
    case 1: GN_FUSED(1);
    case 2: GN_FUSED(2);
    case 3: GN_FUSED(3);
    case 4: GN_FUSED(4);
    default: GN_FUSED(6);
  }
#undef GN_FUSED
}

constexpr long long kGnFusedMaxBytes = 16LL << 20;

// workspace = [final (mean, rstd) float2 x B x groups][counters x B][partials double2 x B x max_chunks x groups];
// it must be zero-initialised ONCE by the caller (the counters), afterwards the kernels keep it consistent.
extern "C" long long pcdm_groupnorm_workspace_bytes(int B, int groups) {
  const long long max_cta = 4LL * num_sms() + 2LL * B;   // B * chunks never exceeds this (see gn_grid)
  return (long long)(gn_align((size_t)B * groups * sizeof(float2)) + gn_align((size_t)B * sizeof(unsigned)) +
                     gn_align((size_t)max_cta * groups * sizeof(double2)));
}

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
  const int silu = (flags & PCDM_FLAG_SILU) ? 1 : 0;
  GnFusedCfg fc;
  const long long in_bytes = 2LL * B * HW * C;
  // path choice: per call through PCDM_FLAG_GN_TWO_PASS / PCDM_FLAG_GN_ONE_PASS (tests), else by size
  const int mode = (flags & PCDM_FLAG_GN_TWO_PASS) ? 1 : ((flags & PCDM_FLAG_GN_ONE_PASS) ? 2 : g_tune.gn_mode);
  const bool want_fused = mode >= 2 || (mode == 0 && in_bytes <= kGnFusedMaxBytes);
  if (want_fused && gn_fused_config(B, HW, C, C1, groups, &fc)) {
    PCDM_CUDA(dtype == DT_F16
                  ? launch_gn_fused<DT_F16>(fc, stream, x1, x2, C1, C, HW, groups, B, eps, gamma, beta, silu, y)
                  : launch_gn_fused<DT_BF16>(fc, stream, x1, x2, C1, C, HW, groups, B, eps, gamma, beta, silu, y));
    PCDM_CUDA(cudaGetLastError());
    return 0;
  }
  int PY, pix_per_cta, chunks;
  gn_grid(B, HW, C, &PY, &pix_per_cta, &chunks);
  GnWorkspace ws;
  char* base = reinterpret_cast<char*>(workspace);
  ws.final_stats = reinterpret_cast<float2*>(base);
  base += gn_align((size_t)B * groups * sizeof(float2));
  ws.counters = reinterpret_cast<unsigned*>(base);
  base += gn_align((size_t)B * sizeof(unsigned));
  ws.partial = reinterpret_cast<double2*>(base);
  const dim3 grid(chunks, B);
  const dim3 block(C / 8, PY);
  const size_t smem = (size_t)2 * PY * C * sizeof(float);
  if (smem > 48 * 1024) return set_error(PCDM_ERR_UNSUPPORTED, "groupnorm: statistics staging exceeds 48 KB");
  if (dtype == DT_F16) {
    PCDM_CUDA(launch_kernel(gn_stats_kernel<DT_F16>, grid, block, smem, stream, 1, x1, x2, C1, C, HW, groups,
                            pix_per_cta, eps, ws));
    PCDM_CUDA(launch_kernel(gn_apply_kernel<DT_F16>, grid, block, (size_t)2 * C * sizeof(float), stream, 1, x1, x2, C1, C, HW, groups, pix_per_cta,
                            (const float2*)ws.final_stats, gamma, beta, silu, y));
  } else {
    PCDM_CUDA(launch_kernel(gn_stats_kernel<DT_BF16>, grid, block, smem, stream, 1, x1, x2, C1, C, HW, groups,
                            pix_per_cta, eps, ws));
    PCDM_CUDA(launch_kernel(gn_apply_kernel<DT_BF16>, grid, block, (size_t)2 * C * sizeof(float), stream, 1, x1, x2, C1, C, HW, groups, pix_per_cta,
                            (const float2*)ws.final_stats, gamma, beta, silu, y));
  }
  PCDM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pcdm_groupnorm_apply(const void* x1, const float* stats1, const void* x2, const float* stats2, int C1,
                                    void* y, const float* gamma, const float* beta, float eps, int B, int HW, int C,
                                    int groups, int dtype, int flags, void* workspace, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!x1 || !stats1 || !y || !gamma || !beta || !workspace) return set_error(PCDM_ERR_INVALID, "groupnorm_apply: null pointer");
  if (dtype != DT_F16 && dtype != DT_BF16) return set_error(PCDM_ERR_INVALID, "groupnorm_apply: bad dtype");
  if (B <= 0 || HW <= 0 || C <= 0 || groups <= 0) return set_error(PCDM_ERR_INVALID, "groupnorm_apply: empty problem");
  if (C % groups || C % 8 || C > 4096) return set_error(PCDM_ERR_UNSUPPORTED, "groupnorm_apply: C must be a multiple of groups and 8, <= 4096");
  if (HW % 32) return set_error(PCDM_ERR_UNSUPPORTED, "groupnorm_apply: H*W must be a multiple of 32 (statistics come per 32-row slab)");
  if (!x2) C1 = C;
  if (C1 % 8 || C1 <= 0 || C1 > C || (x2 && !stats2)) return set_error(PCDM_ERR_INVALID, "groupnorm_apply: bad channel split");
  if ((reinterpret_cast<uintptr_t>(stats1) | reinterpret_cast<uintptr_t>(stats2)) & 7)
    return set_error(PCDM_ERR_INVALID, "groupnorm_apply: statistics must be 8-byte aligned");
  float2* final_stats = reinterpret_cast<float2*>(workspace);   // the first region of a pcdm_groupnorm workspace
  PCDM_CUDA(launch_kernel(gn_fold_kernel, dim3(B * groups), dim3(128), 0, stream, 1,
                          reinterpret_cast<const float2*>(stats1), reinterpret_cast<const float2*>(stats2), C1, C, HW / 32,
                          groups, B, HW, eps, final_stats));
  int PY, pix_per_cta, chunks;
  gn_grid(B, HW, C, &PY, &pix_per_cta, &chunks);
  const dim3 grid(chunks, B), block(C / 8, PY);
  const int silu = (flags & PCDM_FLAG_SILU) ? 1 : 0;
  if (dtype == DT_F16)
    PCDM_CUDA(launch_kernel(gn_apply_kernel<DT_F16>, grid, block, (size_t)2 * C * sizeof(float), stream, 1, x1, x2, C1, C, HW, groups, pix_per_cta,
                            (const float2*)final_stats, gamma, beta, silu, y));
  else
    PCDM_CUDA(launch_kernel(gn_apply_kernel<DT_BF16>, grid, block, (size_t)2 * C * sizeof(float), stream, 1, x1, x2, C1, C, HW, groups, pix_per_cta,
                            (const float2*)final_stats, gamma, beta, silu, y));
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
  // lanes per row / vectors per lane with the fewest idle lanes (ties -> more lanes per row)
  const int nvec = C / 8;
  int lpr = 32, vpl = (nvec + 31) / 32;
  for (int l = 16; l >= 8; l /= 2) {
    const int v = (nvec + l - 1) / l;
    if (v <= 6 && v * l < vpl * lpr) { lpr = l; vpl = v; }
  }
  const int warps_per_cta = 8, rows_per_cta = warps_per_cta * (32 / lpr);
  int grid = (M + rows_per_cta - 1) / rows_per_cta;
  if (grid > 4 * num_sms()) grid = 4 * num_sms();
  const size_t ln_smem = (size_t)2 * C * sizeof(float);
  cudaError_t le = cudaErrorInvalidValue;
#define LN_LAUNCH(V, LP)                                                                                          \
  le = (dtype == DT_F16)                                                                                          \
      ? launch_kernel(layernorm_kernel<DT_F16, V, LP>, dim3(grid), dim3(256), ln_smem, stream, 1, x, ldx, y, ldy, gamma, beta, eps, M, C)  \
      : launch_kernel(layernorm_kernel<DT_BF16, V, LP>, dim3(grid), dim3(256), ln_smem, stream, 1, x, ldx, y, ldy, gamma, beta, eps, M, C)
#define LN_CASES(LP)                                                                                              \
  switch (vpl) {                                                                                                  \
    case 1: LN_LAUNCH(1, LP); break;                                                                              \
    case 2: LN_LAUNCH(2, LP); break;                                                                              \
    case 3: LN_LAUNCH(3, LP); break;                                                                              \
    case 4: LN_LAUNCH(4, LP); break;                                                                              \
    case 5: LN_LAUNCH(5, LP); break;                                                                              \
    default: LN_LAUNCH(6, LP); break;                                                                             \
  }
  if (lpr == 8) { LN_CASES(8) }
  else if (lpr == 16) { LN_CASES(16) }
  else if (vpl <= 6) { LN_CASES(32) }
  else if (vpl == 7) { LN_LAUNCH(7, 32); }
  else { LN_LAUNCH(8, 32); }
  if (le != cudaSuccess) return set_error(PCDM_ERR_CUDA, "layernorm launch failed: %s", cudaGetErrorString(le));
#undef LN_CASES
#undef LN_LAUNCH
  PCDM_CUDA(cudaGetLastError());
  return 0;
}
