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
  float* ps = sm;                   // [PY][C]
  float* pq = sm + (size_t)PY * C;  // [PY][C]
  *reinterpret_cast<float4*>(ps + (size_t)py * C + c0) = make_float4(s[0], s[1], s[2], s[3]);
  *reinterpret_cast<float4*>(ps + (size_t)py * C + c0 + 4) = make_float4(s[4], s[5], s[6], s[7]);
  *reinterpret_cast<float4*>(pq + (size_t)py * C + c0) = make_float4(q[0], q[1], q[2], q[3]);
  *reinterpret_cast<float4*>(pq + (size_t)py * C + c0 + 4) = make_float4(q[4], q[5], q[6], q[7]);
  __syncthreads();
  // per-channel totals over py (fixed order), written back into row 0
  for (int c = tid; c < C; c += nthreads) {
    float a = ps[c], bq = pq[c];
    for (int y = 1; y < PY; ++y) { a += ps[(size_t)y * C + c]; bq += pq[(size_t)y * C + c]; }
    ps[c] = a;
    pq[c] = bq;
  }
  __syncthreads();
  const int cpg = C / groups;
  const int chunks = gridDim.x;
  if (tid < groups) {
    double a = 0.0, bq = 0.0;
    for (int c = tid * cpg; c < (tid + 1) * cpg; ++c) { a += (double)ps[c]; bq += (double)pq[c]; }
    ws.partial[((size_t)b * chunks + blockIdx.x) * groups + tid] = make_double2(a, bq);
  }
  __threadfence();
  __syncthreads();
  __shared__ bool is_last;
  if (tid == 0) is_last = (atomicAdd(&ws.counters[b], 1u) == (unsigned)(chunks - 1));
  __syncthreads();
  if (is_last) {
    __threadfence();
    if (tid < groups) {
      double a = 0.0, bq = 0.0;
      for (int k = 0; k < chunks; ++k) {
        const double2 v = ws.partial[((size_t)b * chunks + k) * groups + tid];
        a += v.x; bq += v.y;
      }
      const double inv_n = 1.0 / ((double)cpg * (double)HW);
      const double mean = a * inv_n;
      double var = bq * inv_n - mean * mean;
      if (var < 0.0) var = 0.0;
      ws.final_stats[(size_t)b * groups + tid] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
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
  pdl_wait();
  using T = typename TypeOf<DT>::T;
  const int b = blockIdx.y;
  const int cv = threadIdx.x, py = threadIdx.y, PY = blockDim.y;
  const int c0 = cv * 8;
  const int cpg = C / groups;
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + i;
    const float2 st = __ldg(final_stats + (size_t)b * groups + c / cpg);
    const float s_ = st.y * __ldg(gamma + c);
    sc[i] = s_;
    sh[i] = __ldg(beta + c) - st.x * s_;
  }
  const T* src;
  int cs, coff;
  if (c0 < C1) { src = reinterpret_cast<const T*>(x1); cs = C1; coff = c0; }
  else { src = reinterpret_cast<const T*>(x2); cs = C - C1; coff = c0 - C1; }
  const int p_begin = blockIdx.x * pix_per_cta;
  const int p_end = min(HW, p_begin + pix_per_cta);
  T* dst = reinterpret_cast<T*>(y);
  for (int p = p_begin + py; p < p_end; p += PY) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + ((size_t)b * HW + p) * cs + coff));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack2<DT>(w[k]);
      float a = f.x * sc[2 * k] + sh[2 * k];
      float bb = f.y * sc[2 * k + 1] + sh[2 * k + 1];
      if (silu) { a = silu_f(a); bb = silu_f(bb); }
      o[k] = pack2<DT>(a, bb);
    }
    *reinterpret_cast<uint4*>(dst + ((size_t)b * HW + p) * C + c0) = make_uint4(o[0], o[1], o[2], o[3]);
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
  float s_lo = 0.f, q_lo = 0.f, s_hi = 0.f, q_hi = 0.f;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    const uint32_t w[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack2<DT>(w[k]);
      const float a0 = (2 * k < split) ? f.x : 0.f, b0 = (2 * k < split) ? 0.f : f.x;
      const float a1 = (2 * k + 1 < split) ? f.y : 0.f, b1 = (2 * k + 1 < split) ? 0.f : f.y;
      s_lo += a0; q_lo += a0 * a0; s_hi += b0; q_hi += b0 * b0;
      s_lo += a1; q_lo += a1 * a1; s_hi += b1; q_hi += b1 * b1;
    }
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
  if (tid < gset) {
    double a = 0.0, q = 0.0;
    if (S > 1) {
      const uint32_t local = smem_u32(&cl[tid]);
      for (int r = 0; r < S; ++r) {
        const double2 v = ld_dsmem_f64x2(mapa_u32(local, (uint32_t)r));
        a += v.x; q += v.y;
      }
    } else {
      a = cl[tid].x; q = cl[tid].y;
    }
    const double inv_n = 1.0 / ((double)cpg * (double)HW);
    const double mean = a * inv_n;
    double var = q * inv_n - mean * mean;
    if (var < 0.0) var = 0.0;
    stat[tid] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
  }
  __syncthreads();
  float sc[8], sh[8];
  {
    const float2 st_lo = stat[glo], st_hi = stat[min(glo + 1, gset - 1)];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float2 st = (e < split) ? st_lo : st_hi;
      const float s_ = st.y * gam[crel + e];
      sc[e] = s_;
      sh[e] = bet[crel + e] - st.x * s_;
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
        const float2 f = unpack2<DT>(w[k]);
        float a = f.x * sc[2 * k] + sh[2 * k];
        float bb = f.y * sc[2 * k + 1] + sh[2 * k + 1];
        if (silu) { a = silu_f(a); bb = silu_f(bb); }
        o[k] = pack2<DT>(a, bb);
      }
      *reinterpret_cast<uint4*>(dp + (size_t)i * PY * C) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
  if (S > 1) cluster_sync_all();   // cl[] must outlive every peer's read of it
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, rows strided over a grid of ~4 CTAs per SM.  gamma/beta are staged in shared memory
// before the programmatic-dependency wait (they are parameters), the row lives in registers (two-pass variance), and
// the NEXT row of the warp is already in flight while the current one is reduced and stored.  C <= 2048, C % 8 == 0.
// ---------------------------------------------------------------------------------------------------------------
template <int DT, int VPL>  // VPL = 16-byte vectors per lane
__global__ void __launch_bounds__(256)
layernorm_kernel(const void* __restrict__ x, long long ldx, void* __restrict__ y, long long ldy,
                 const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int M, int C) {
  pdl_launch_dependents();
  using T = typename TypeOf<DT>::T;
  extern __shared__ __align__(16) float lsm[];   // gamma[C] | beta[C]
  for (int i = threadIdx.x; i < C / 4; i += blockDim.x) {
    reinterpret_cast<float4*>(lsm)[i] = __ldg(reinterpret_cast<const float4*>(gamma) + i);
    reinterpret_cast<float4*>(lsm + C)[i] = __ldg(reinterpret_cast<const float4*>(beta) + i);
  }
  pdl_wait();
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpc = blockDim.x >> 5;
  const long long nwarps = (long long)gridDim.x * wpc;
  const int nvec = C / 8;
  const T* xb = reinterpret_cast<const T*>(x);
  T* yb = reinterpret_cast<T*>(y);
  long long row = (long long)blockIdx.x * wpc + warp;
  uint4 cur[VPL], nxt[VPL];
  if (row < M) {
#pragma unroll
    for (int i = 0; i < VPL; ++i)
      cur[i] = (lane + i * 32 < nvec) ? __ldg(reinterpret_cast<const uint4*>(xb + row * ldx + (lane + i * 32) * 8))
                                      : make_uint4(0u, 0u, 0u, 0u);
  }
  const float inv_c = 1.0f / (float)C;
  for (; row < M; row += nwarps) {
    const long long rn = row + nwarps;
    if (rn < M) {
#pragma unroll
      for (int i = 0; i < VPL; ++i)
        nxt[i] = (lane + i * 32 < nvec) ? __ldg(reinterpret_cast<const uint4*>(xb + rn * ldx + (lane + i * 32) * 8))
                                        : make_uint4(0u, 0u, 0u, 0u);
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const uint32_t w[4] = {cur[i].x, cur[i].y, cur[i].z, cur[i].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = unpack2<DT>(w[k]);
        sum += f.x + f.y;
      }
    }
    const float mean = warp_sum(sum) * inv_c;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (lane + i * 32 < nvec) {
        const uint32_t w[4] = {cur[i].x, cur[i].y, cur[i].z, cur[i].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = unpack2<DT>(w[k]);
          const float d0 = f.x - mean, d1 = f.y - mean;
          sq += d0 * d0 + d1 * d1;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) * inv_c + eps);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        const float4 g0 = *reinterpret_cast<const float4*>(lsm + vi * 8);
        const float4 g1 = *reinterpret_cast<const float4*>(lsm + vi * 8 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(lsm + C + vi * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(lsm + C + vi * 8 + 4);
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        const uint32_t w[4] = {cur[i].x, cur[i].y, cur[i].z, cur[i].w};
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = unpack2<DT>(w[k]);
          o[k] = pack2<DT>((f.x - mean) * rstd * g[2 * k] + bb[2 * k], (f.y - mean) * rstd * g[2 * k + 1] + bb[2 * k + 1]);
        }
        *reinterpret_cast<uint4*>(yb + row * ldy + vi * 8) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) cur[i] = nxt[i];
  }
}

}  // namespace pcdm

using namespace pcdm;

static void gn_grid(int B, int HW, int C, int* PY_, int* pix_per_cta_, int* chunks_) {
  const int cvs = C / 8;
  int PY = 512 / cvs;
  if (PY < 1) PY = 1;
  // >= ~4 CTAs per SM across the batch (these kernels are latency/bandwidth bound: parallelism first), but keep at
  // least one pixel per thread row
  const int chunks_wanted = (4 * num_sms() + B - 1) / B;
  int pix_per_cta = (HW + chunks_wanted - 1) / chunks_wanted;
  if (pix_per_cta < PY) pix_per_cta = PY;
  if (pix_per_cta > HW) pix_per_cta = HW;
  if (PY > pix_per_cta) PY = pix_per_cta;
  *PY_ = PY;
  *pix_per_cta_ = pix_per_cta;
  *chunks_ = (HW + pix_per_cta - 1) / pix_per_cta;
}

struct GnFusedCfg { int gset, nv, PY, S, pix_per_cta, K; };

// Shape of the single-pass launch, or false when the problem needs the two-kernel path (cpg < 8, pieces too large).
static bool gn_fused_config(int B, int HW, int C, int C1, int groups, GnFusedCfg* c) {
  const int cpg = C / groups;
  if (cpg < 8 || (C1 % 8)) return false;
  int gmin = 1;
  while (gmin <= 8 && (gmin * cpg) % 8) gmin *= 2;
  if (gmin > 8 || groups % gmin) return false;
  int gset = gmin;
  for (int g = 8; g > gmin; g /= 2) {   // wider sets = longer contiguous runs per pixel, while >= 128 pieces remain
    if (g % gmin || groups % g || g * cpg / 8 > 64) continue;
    if ((long long)B * (groups / g) >= 128) { gset = g; break; }
  }
  const int nv = gset * cpg / 8;
  if (nv > 128) return false;
  int PYmax = 1024 / nv;
  int S = 1;
  for (;;) {
    const int ppc = (HW + S - 1) / S;
    int PY = ppc < PYmax ? ppc : PYmax;
    const int PYmin = (32 + nv - 1) / nv;
    if (PY < PYmin) PY = PYmin;
    const int K = (ppc + PY - 1) / PY;
    const bool more_parallel = (long long)B * (groups / gset) * S < 128 && ppc >= 2 * PY;
    if (K > 6 || more_parallel) {
      if (S >= 8) { if (K > 6) return false; }
      else { S *= 2; continue; }
    }
    c->gset = gset; c->nv = nv; c->PY = PY; c->S = S; c->pix_per_cta = ppc;
    c->K = K <= 4 ? K : 6;
    return true;
  }
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

static int g_gn_two_pass = 0;   // pcdm_set_groupnorm_two_pass(): force the two-kernel path (tests / A-B timing)
extern "C" int pcdm_set_groupnorm_two_pass(int enabled) {
  g_gn_two_pass = enabled ? 1 : 0;
  return 0;
}

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
  if (!g_gn_two_pass && gn_fused_config(B, HW, C, C1, groups, &fc)) {
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
    PCDM_CUDA(launch_kernel(gn_apply_kernel<DT_F16>, grid, block, 0, stream, 1, x1, x2, C1, C, HW, groups, pix_per_cta,
                            (const float2*)ws.final_stats, gamma, beta, silu, y));
  } else {
    PCDM_CUDA(launch_kernel(gn_stats_kernel<DT_BF16>, grid, block, smem, stream, 1, x1, x2, C1, C, HW, groups,
                            pix_per_cta, eps, ws));
    PCDM_CUDA(launch_kernel(gn_apply_kernel<DT_BF16>, grid, block, 0, stream, 1, x1, x2, C1, C, HW, groups, pix_per_cta,
                            (const float2*)ws.final_stats, gamma, beta, silu, y));
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
  const int rows_per_cta = 8;   // warps per CTA
  int grid = (M + rows_per_cta - 1) / rows_per_cta;
  if (grid > 4 * num_sms()) grid = 4 * num_sms();
  const size_t ln_smem = (size_t)2 * C * sizeof(float);
#define LN_LAUNCH(V)                                                                                              \
  do {                                                                                                            \
    cudaError_t _le = (dtype == DT_F16)                                                                           \
        ? launch_kernel(layernorm_kernel<DT_F16, V>, dim3(grid), dim3(256), ln_smem, stream, 1, x, ldx, y, ldy, gamma, beta, eps, M, C)  \
        : launch_kernel(layernorm_kernel<DT_BF16, V>, dim3(grid), dim3(256), ln_smem, stream, 1, x, ldx, y, ldy, gamma, beta, eps, M, C); \
    if (_le != cudaSuccess) return set_error(PCDM_ERR_CUDA, "layernorm launch failed: %s", cudaGetErrorString(_le)); \
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
