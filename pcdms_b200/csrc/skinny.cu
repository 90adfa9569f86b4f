// Skinny GEMM: out[M, N] = act( A[M, K] . W[N, K]^T + bias + rowvec + residual ) for M <= 32 activation rows — the
// stage-1 prior (6 or 12 token rows against ~1.0 G weight parameters per step: /root/reference/src/models/
// stage1_prior_transformer.py:112-131,283-293) and the per-step embedding GEMMs of the UNet (time / class embedding,
// the stacked time_emb_proj: M = UNet batch).  With so few rows the problem is a WEIGHT STREAM: every weight byte is
// read once from HBM and used M times, so the kernel is organised around memory-level parallelism, not around the
// 128-row tcgen05 tile (which would run > 90 % empty and serialise the stream behind one TMA ring per CTA):
//
//   * roles swapped — the weight rows are the M dimension of `mma.sync.m16n8k16` (16 output features per CTA), the
//     activation rows are its 8-wide N dimension (one n-tile per 8 activation rows);
//   * a CTA works on 16 weight rows and ALL of K at a time: its 16 warps interleave over 64-element K chunks, so the
//     CTA reads 2 KB contiguous per row per round and every lane issues plain 16-byte loads that are consumed in
//     registers in exactly the layout the MMA fragments want (the k-slots of a fragment are a permutation of physical
//     k, the same one for both operands — a dot product does not care); two or four chunks are in flight per warp
//     (128 / 256 B of weights per lane), the first of them requested before the programmatic-dependency wait (weights
//     do not depend on the previous kernel), and the kernel is persistent over row blocks so that the stream keeps
//     flowing while a row block is reduced and written;
//   * the 16 per-warp partial tiles meet in shared memory and are summed in fixed warp order (bit-reproducible);
//     bias / per-image row vector / residual / SiLU / GELU are applied there, 16-bit or fp32 output.
//
// HBM roofline: N * K * 2 bytes per launch (activations and outputs are noise).  pcdm_gemm routes here for M <= 32
// (no GEGLU, one K segment); pcdm_set_skinny_gemm(0) restores the tcgen05 path for A/B timing.
#include "common.cuh"
#include "host_util.h"

namespace pcdm {

constexpr int SK_WARPS = 16;
constexpr int SK_THREADS = SK_WARPS * 32;
constexpr int SK_ROWS = 16;   // weight rows (output features) per CTA

struct SkinnyParams {
  const void* a; long long lda;
  const void* w;                 // [N][K] row-major 16-bit
  void* out; long long ldo;
  const float* bias;
  const float* rowvec; long long ld_rowvec; int hw;
  const void* residual; long long ldr;
  int M, N, K;
  int act;                       // 0 none, 1 SiLU, 2 GELU(erf)
  int out_f32;
  const float* ln_gamma;         // LN variant: A is layer-normalised over K on the way in (fused torch.nn.LayerNorm)
  const float* ln_beta;
  float ln_eps;
  int w_static;                  // W is not written by the stream predecessor: it may be read before the dependency wait
};

template <int DT>
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  if (DT == DT_F16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
                 "{%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
                 "{%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint4 ldg_stream(const void* p) {   // weights: read once, do not pollute L1
  uint4 u;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p));
  return u;
}

// NT = number of 8-row activation tiles (M <= 8 NT).  PF = weight chunks in flight per warp (2: 64 registers, two CTAs
// per SM; 3 / 4: one CTA per SM).  LN: the activation rows are layer-normalised first (each CTA normalises the <= 32 rows
// itself — 4 KB per row out of L2 — into shared memory, rounded to 16 bits exactly as the stand-alone LayerNorm kernel
// would have stored them).
//
// The kernel is PERSISTENT over 16-row blocks of W (grid = min(N / 16, resident CTAs)): a warp walks the flattened
// sequence (row block, its K chunks) with PF chunks always in flight, so the loads of the next row block are already
// under way while the current one is reduced and written — the stream does not drain between row blocks.
template <int DT, int NT, bool LN, int PF>
__global__ void __launch_bounds__(SK_THREADS, (NT == 1 && !LN && PF == 2) ? 2 : 1)
skinny_gemm_kernel(const SkinnyParams p) {
  __shared__ float red[SK_WARPS][NT][SK_ROWS][8 + 1];
  extern __shared__ __align__(16) uint8_t sk_dyn[];   // LN: normalised activations [M][K + 8] 16-bit
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int nchunks = p.K >> 6;
  const int n_rb = p.N / SK_ROWS;
  using T = typename TypeOf<DT>::T;
  const T* W = reinterpret_cast<const T*>(p.w);
  const T* A = reinterpret_cast<const T*>(p.a);
  const T* w_lane = W + (long long)g * p.K + 16 * t;   // + row block * 16 * K + chunk * 64; rows g and g + 8
  const long long hi = 8LL * p.K;
  const T* a_row[NT];
  bool a_ok[NT];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    a_ok[nt] = nt * 8 + g < p.M;
    a_row[nt] = A + (long long)(a_ok[nt] ? nt * 8 + g : 0) * p.lda + 16 * t;
  }

  // Model weights do not depend on the stream predecessor (PCDM_FLAG_W_STATIC): the first PF chunks of this warp are
  // then requested BEFORE the programmatic-dependency wait, so the weight stream of this kernel overlaps the tail (and,
  // in a chain of small kernels, most of the body) of the previous one; only activations wait.
  auto load_w = [&](uint4 (&dst)[4], int rb, int chunk) {
    const T* q = w_lane + (long long)rb * SK_ROWS * p.K + (chunk << 6);
    dst[0] = ldg_stream(q); dst[1] = ldg_stream(q + 8);
    dst[2] = ldg_stream(q + hi); dst[3] = ldg_stream(q + hi + 8);
  };
  // Register ring with STATIC slots: a row block is walked in groups of PF chunks, chunk j of a group always lives in
  // buf[j]; when slot j has been consumed it is refilled with the chunk PF positions ahead — in the same row block, or
  // slot j of the first group of this CTA's NEXT row block — so PF chunks (or the whole next row block) are in flight.
  uint4 buf[PF][4];
  if (!p.w_static) pdl_wait();   // W produced on this stream (e.g. K of an attention written as a GEMM): wait first
#pragma unroll
  for (int j = 0; j < PF; ++j) {
    const int c = warp + j * SK_WARPS;
    if (c < nchunks && (int)blockIdx.x < n_rb) load_w(buf[j], blockIdx.x, c);
  }
  if (p.w_static) pdl_wait();
  const int lds = p.K + 8;   // padded row stride of the normalised activations
  if (LN) {
    T* act_s = reinterpret_cast<T*>(sk_dyn);
    for (int r = warp; r < p.M; r += SK_WARPS) {   // one warp per row, the row in registers (K <= 2048: 8 x 16 B per lane)
      const T* x = A + (long long)r * p.lda;
      uint4 xr[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = lane * 8 + i * 256;
        xr[i] = k < p.K ? __ldg(reinterpret_cast<const uint4*>(x + k)) : make_uint4(0u, 0u, 0u, 0u);
      }
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 f0 = unpack2<DT>(xr[i].x), f1 = unpack2<DT>(xr[i].y), f2 = unpack2<DT>(xr[i].z), f3 = unpack2<DT>(xr[i].w);
        sum += ((f0.x + f0.y) + (f1.x + f1.y)) + ((f2.x + f2.y) + (f3.x + f3.y));   // zero padding adds nothing
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mean = sum / (float)p.K;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (lane * 8 + i * 256 < p.K) {
          const uint32_t uu[4] = {xr[i].x, xr[i].y, xr[i].z, xr[i].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = unpack2<DT>(uu[e]);
            sq += (f.x - mean) * (f.x - mean) + (f.y - mean) * (f.y - mean);
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      const float rstd = rsqrtf(sq / (float)p.K + p.ln_eps);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = lane * 8 + i * 256;
        if (k < p.K) {
          const uint32_t uu[4] = {xr[i].x, xr[i].y, xr[i].z, xr[i].w};
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.ln_gamma + k));
          const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.ln_gamma + k + 4));
          const float4 e0 = __ldg(reinterpret_cast<const float4*>(p.ln_beta + k));
          const float4 e1 = __ldg(reinterpret_cast<const float4*>(p.ln_beta + k + 4));
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          const float ee[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
          uint32_t o4[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = unpack2<DT>(uu[e]);
            o4[e] = pack2<DT>((f.x - mean) * rstd * gg[2 * e] + ee[2 * e], (f.y - mean) * rstd * gg[2 * e + 1] + ee[2 * e + 1]);
          }
          *reinterpret_cast<uint4*>(act_s + (long long)r * lds + k) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
        }
      }
    }
    __syncthreads();
  }
  for (int rb = blockIdx.x; rb < n_rb; rb += gridDim.x) {
    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
    for (int c0 = warp; c0 < nchunks; c0 += PF * SK_WARPS) {
#pragma unroll
      for (int j = 0; j < PF; ++j) {
        const int c = c0 + j * SK_WARPS;
        if (c < nchunks) {
          uint4 b0[NT], b1[NT];
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            b0[nt] = make_uint4(0u, 0u, 0u, 0u);
            b1[nt] = b0[nt];
            if (a_ok[nt]) {
              if (LN) {
                const T* ar = reinterpret_cast<const T*>(sk_dyn) + (long long)(nt * 8 + g) * lds + 16 * t + (c << 6);
                b0[nt] = *reinterpret_cast<const uint4*>(ar);
                b1[nt] = *reinterpret_cast<const uint4*>(ar + 8);
              } else {
                b0[nt] = __ldg(reinterpret_cast<const uint4*>(a_row[nt] + (c << 6)));
                b1[nt] = __ldg(reinterpret_cast<const uint4*>(a_row[nt] + (c << 6) + 8));
              }
            }
          }
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            mma16816<DT>(acc[nt], buf[j][0].x, buf[j][2].x, buf[j][0].y, buf[j][2].y, b0[nt].x, b0[nt].y);
            mma16816<DT>(acc[nt], buf[j][0].z, buf[j][2].z, buf[j][0].w, buf[j][2].w, b0[nt].z, b0[nt].w);
            mma16816<DT>(acc[nt], buf[j][1].x, buf[j][3].x, buf[j][1].y, buf[j][3].y, b1[nt].x, b1[nt].y);
            mma16816<DT>(acc[nt], buf[j][1].z, buf[j][3].z, buf[j][1].w, buf[j][3].w, b1[nt].z, b1[nt].w);
          }
          // refill this slot with the chunk PF positions ahead
          int nc = c + PF * SK_WARPS, nrb = rb;
          if (nc >= nchunks) { nc = warp + j * SK_WARPS; nrb = rb + gridDim.x; }
          if (nrb < n_rb && nc < nchunks) load_w(buf[j], nrb, nc);
        }
      }
    }
    // accumulator fragment: c0/c1 = (weight row g, activation rows 2t, 2t+1), c2/c3 = (weight row g + 8, same)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      red[warp][nt][g][2 * t] = acc[nt][0];
      red[warp][nt][g][2 * t + 1] = acc[nt][1];
      red[warp][nt][g + 8][2 * t] = acc[nt][2];
      red[warp][nt][g + 8][2 * t + 1] = acc[nt][3];
    }
    __syncthreads();
    // one thread per output element: fixed summation order over the warps, then the epilogue
    const int n0 = rb * SK_ROWS;
    for (int idx = threadIdx.x; idx < NT * 8 * SK_ROWS; idx += SK_THREADS) {
      const int r = idx & (SK_ROWS - 1);      // weight row inside the block (fastest: consecutive output columns)
      const int m = idx >> 4;                 // activation row
      if (m >= p.M) continue;
      const int nt = m >> 3, mc = m & 7;
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < SK_WARPS; ++w) v += red[w][nt][r][mc];
      const int n = n0 + r;
      if (p.bias) v += p.bias[n];
      if (p.rowvec) v += p.rowvec[(long long)(m / p.hw) * p.ld_rowvec + n];
      if (p.residual) v += (float)reinterpret_cast<const T*>(p.residual)[(long long)m * p.ldr + n];
      if (p.act == 1) v = silu_f(v);
      else if (p.act == 2) v = gelu_erf_f(v);
      if (p.out_f32) reinterpret_cast<float*>(p.out)[(long long)m * p.ldo + n] = v;
      else reinterpret_cast<T*>(p.out)[(long long)m * p.ldo + n] = (T)v;
    }
    __syncthreads();   // the partial-sum buffer is reused by the next row block
  }
}


constexpr int SK_LN_SMEM_LIMIT = 100 * 1024;   // fused LayerNorm: M * (K + 8) * 2 bytes of normalised rows must fit

template <int DT, bool LN>
static cudaError_t launch_skinny(const SkinnyParams& p, cudaStream_t stream) {
  const int n_rb = p.N / SK_ROWS;
  const int nt = (p.M + 7) / 8;
  const size_t dyn = LN ? (size_t)p.M * (p.K + 8) * 2 : 0;
  // Two CTAs per SM (64 registers, two chunks in flight per warp) when the problem has more row blocks than SMs;
  // otherwise — or with more than 8 rows, where the accumulators take the registers anyway — one CTA per SM with four
  // chunks in flight per warp.
  const bool deep = LN || nt > 1 || n_rb <= num_sms();
  const int resident = (deep ? 1 : 2) * num_sms();
  const dim3 grid(n_rb < resident ? n_rb : resident), block(SK_THREADS);
#define SK_CASE(NT_, PF_)                                                                                            \
  do {                                                                                                               \
    if (LN) {                                                                                                        \
      static bool _cfg[64] = {};                                                                                     \
      int _dev = 0;                                                                                                  \
      cudaGetDevice(&_dev);                                                                                          \
      if (_dev < 0 || _dev >= 64 || !_cfg[_dev]) {                                                                   \
        cudaFuncSetAttribute(skinny_gemm_kernel<DT, NT_, LN, PF_>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                             SK_LN_SMEM_LIMIT);                                                                      \
        if (_dev >= 0 && _dev < 64) _cfg[_dev] = true;                                                               \
      }                                                                                                              \
    }                                                                                                                \
    return launch_kernel(skinny_gemm_kernel<DT, NT_, LN, PF_>, grid, block, dyn, stream, 1, p);                      \
  } while (0)
  switch (nt) {   // the LayerNorm variants keep two chunks in flight (the row registers of the prologue crowd the ring)
    case 1:
      if constexpr (LN) SK_CASE(1, 2);
      else if (deep) SK_CASE(1, 4);
      else SK_CASE(1, 2);
    case 2: SK_CASE(2, (LN ? 2 : 3));
    case 3: SK_CASE(3, (LN ? 2 : 3));
    default: SK_CASE(4, (LN ? 2 : 3));
  }
#undef SK_CASE
}

// returns 1 when the problem was taken (launched), 0 when pcdm_gemm should use the tcgen05 path, < 0 on error
int skinny_gemm_try(const void* a, long long lda, const void* w, void* out, long long ldo, const float* bias,
                    const float* rowvec, long long ld_rowvec, int rows_per_image, const void* residual, long long ldr,
                    int M, int N, int K, int dtype, int flags, cudaStream_t stream, const float* ln_gamma,
                    const float* ln_beta, float ln_eps) {
  if (!g_tune.skinny || M > 32 || (flags & PCDM_FLAG_GEGLU) || (N % SK_ROWS) || (K % 64)) return 0;
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(w)) & 15) return 0;   // 16-byte vector loads
  if (ln_gamma && ((size_t)M * (K + 8) * 2 > (size_t)SK_LN_SMEM_LIMIT || K > 2048)) return 0;   // row held in registers
  SkinnyParams p;
  p.ln_gamma = ln_gamma; p.ln_beta = ln_beta; p.ln_eps = ln_eps;
  p.w_static = (flags & PCDM_FLAG_W_STATIC) ? 1 : 0;
  p.a = a; p.lda = lda; p.w = w; p.out = out; p.ldo = ldo; p.bias = bias;
  p.rowvec = rowvec; p.ld_rowvec = ld_rowvec; p.hw = rows_per_image > 0 ? rows_per_image : 1;
  p.residual = residual; p.ldr = ldr;
  p.M = M; p.N = N; p.K = K;
  p.act = (flags & PCDM_FLAG_SILU) ? 1 : ((flags & PCDM_FLAG_GELU) ? 2 : 0);
  p.out_f32 = (flags & PCDM_FLAG_OUT_F32) ? 1 : 0;
  cudaError_t e;
  if (ln_gamma) e = dtype == DT_F16 ? launch_skinny<DT_F16, true>(p, stream) : launch_skinny<DT_BF16, true>(p, stream);
  else e = dtype == DT_F16 ? launch_skinny<DT_F16, false>(p, stream) : launch_skinny<DT_BF16, false>(p, stream);
  if (e != cudaSuccess) return set_error(PCDM_ERR_CUDA, "skinny gemm launch failed: %s", cudaGetErrorString(e));
  if (cudaGetLastError() != cudaSuccess) return set_error(PCDM_ERR_CUDA, "skinny gemm launch failed");
  return 1;
}

}  // namespace pcdm

