// K1/K2 — implicit-GEMM convolution (3x3, stride 1/2, im2col-free) and linear GEMM on tcgen05 tensor cores.
//
//   D[M, N] = A[M, K] . W[N, K]^T  (+bias[N]) (+rowvec[image(m), N]) (+residual[M, N])   or GEGLU pairing
//
// * A is never materialised as an im2col matrix: activations are NHWC and every (tap, 64-channel) K-block of an
//   output tile is ONE 4-D TMA box load (64 ch x W x tile_h x tile_b) whose start coordinate carries the tap offset;
//   TMA's out-of-bounds zero fill IS the conv padding.  Stride-2 convs read through four "parity" tensor maps laid
//   over the same NHWC buffer (element strides doubled), so they use the identical box loads.
// * A 128 x BN fp32 accumulator lives in TMEM (double-buffered: the epilogue of tile i overlaps the MMAs of tile
//   i+1); one elected thread issues tcgen05.mma (M=128, N=BN, K=16) from 128B-swizzled shared-memory operands.
// * Warp roles: warp 0 = TMA producer (activations), warp 3 = TMA producer (weights), warp 1 = MMA issuer, warp 2 = TMEM
//   allocator, warps 4-11 = epilogue.  Producer and MMA warps stay converged and elect one lane per issue.
//   (tcgen05.ld -> fp32 math -> 16-bit stores).  Persistent: each CTA walks tiles blockIdx.x, +gridDim.x, ...
//
// Replaces, on the reference path, torch.nn.Conv2d / torch.nn.Linear as called from diffusers' ResnetBlock2D,
// Downsample2D/Upsample2D, Transformer2DModel and BasicTransformerBlock (reference call
// sites src/models/stage2_inpaint_unet_2d_condition.py:321-343,348-361,407-429).
#include "common.cuh"
#include "host_util.h"

namespace pcdm {

#ifdef PCDM_EXPERIMENT
#define IG_DBG(p, bit) ((p).dbg & (bit))   // experiment build: parts of the kernel can be switched off for timing
#else
#define IG_DBG(p, bit) 0
#endif

constexpr int IG_THREADS = 384;        // warps 0-3: TMA(A) / MMA / TMEM-alloc / TMA(B); warps 4-11: epilogue
constexpr int IG_EPI_WARPS = 8;        // EW of the kernel template: 8, or 16 for the GEGLU-only instances (640 threads)
static_assert(IG_THREADS == (4 + IG_EPI_WARPS) * 32, "default instance: 4 role warps + 8 epilogue warps");
constexpr int IG_SLOT_BYTES = 32 * 64; // one epilogue staging slot: 32 rows x 32 columns x 16 bit (64-byte swizzle)
constexpr int IG_RES_SLOTS = 5;        // residual slots per epilogue warp: the chunks one warp owns in a <= 320-wide tile
constexpr int IG_MAX_STAGES = 8;

struct IGemmParams {
  CUtensorMap tmA[4];
  CUtensorMap tmB;
  CUtensorMap tmOut;   // [M, N_out] 16-bit output, box 32 x 32, 64B swizzle (unused for fp32 output)
  CUtensorMap tmRes;   // residual, same geometry
  int M, N;
  int num_kb;      // K / 64
  int m_tiles, n_tiles;
  int mode;        // 0 = plain GEMM (A row-major [M, K], up to two K-segments), 1 = conv3x3 s1, 2 = conv3x3 s2 (pad 1),
                   // 3 = conv3x3 s2 with bottom/right padding only, 4 = nearest-2x upsample + conv3x3 as four 2x2
                   // convolutions over the LOW-resolution input, one per output-pixel parity ("plane": the tile index
                   // enumerates it through `splits` = 4; H / W / M are low-resolution; see pcdm_conv3x3_up2x)
  CUtensorMap tmOutP[4];   // mode 4: the output pixels of parity (py, px) as a 4-D strided view [Cout, W, H, B]
  int obox_w, obox_h;      // mode 4: the 32 pixels a warp stores at once as a (obox_w x obox_h x 32/(w*h)) box
  int H, W;        // conv: output height / width
  int cblocks;     // conv: Cin / 64
  int kb_split;    // plain: k-blocks taken from tmA[0]; the rest come from tmA[1]
  uint32_t a_bytes, b_bytes;  // bytes one A / B box load delivers (boxes are clamped to the tensor extent)
  int splits;      // split-K factor (1 = off): tile index also enumerates the K slice; partials go to fp32 scratch
  int kb_per_split;
  int stages;      // smem ring depth (runtime: depends on BN and on whether residual staging is needed)
  int nbuf;        // staging slots per epilogue warp (2; 4 with a residual: one per 32-column chunk a warp owns in a
                   // tile, so the whole residual tile is in flight while the tile's MMAs run)
  const float* bias;
  const float* rowvec;
  long long ld_rowvec;
  int hw;          // rows per image for rowvec indexing
  int has_res;
  int dbg;         // experiment mask (pcdm_set_gemm_debug; results are WRONG when non-zero): 1 no TMA stores, 2 no
                   // residual, 4 no bias / rowvec, 8 epilogue body skipped, 16 no MMAs, 32 GEGLU without its arithmetic
  void* out;       // only dereferenced for fp32 output
  long long ldo;
  int geglu;
  int out_f32;
  int silu;        // activation: 0 none, 1 SiLU, 2 GELU (erf)
  // ---- LayerNorm around the GEMM (16-bit output paths; see pcdm_ext in include/pcdm_b200.h) ----
  float2* stats_out;        // producer: per-row (sum, sum of squares) of the outputs, one slot per (n-tile, column-half):
                            // [2 * n_tiles][M]; the LayerNorm statistics of the NEXT op without another pass over the rows
  const float2* ln_stats;   // consumer: LayerNorm folded into this GEMM — the rows arrive raw, W was pre-scaled by gamma
  int ln_parts;             //   AND row-centred (sum_k W'[n,k] = 0, so the mean cancels inside the accumulation):
  float ln_eps, ln_inv_k;   //   out = rstd[m] * acc + bias[n]   (bias[n] carries W . beta), rstd from the producer's sums
  // ---- GroupNorm statistics of the output (pcdm_ext.chan_stats): per 32-row slab and per channel, (sum, sum of
  //      squares) of the 16-bit values as stored: [slabs][N][2] fp32.  Rows of a slab belong to one image (H*W % 32 == 0).
  float* chan_stats;
};

template <int BN, int CG>
struct IGemmCfg {
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = (BN / CG) * 128;   // a CTA pair splits the weight tile: each CTA stages BN/2 rows
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // accumulators: double-buffered in TMEM (the epilogue of tile i overlaps the MMAs of tile i + 1) up to 256 columns;
  // the 320-wide pair tile (two 160-column MMAs fed from ONE activation stage) takes 320 of the 512 columns once
  static constexpr int NACC = BN <= 256 ? 2 : 1;
  static constexpr int ACC_STRIDE = BN <= 128 ? 128 : (BN <= 256 ? 256 : 512);
  static constexpr int TMEM_COLS = NACC * ACC_STRIDE;
  static constexpr int NSUB = BN <= 256 ? 1 : 2;          // MMAs per k16 step (N <= 256 per tcgen05.mma)
  static constexpr int SUB_N = BN / NSUB;
};

// byte offset of 16-byte chunk j of row r inside a [rows x 64 B] tile written with the TMA 64-byte swizzle
__device__ __forceinline__ uint32_t sw64(int r, int j) { return (uint32_t)(r * 64 + ((j ^ ((r >> 1) & 3)) << 4)); }
// staging-slot accesses in the shared state space proper (a generic ST/LD to shared memory made the compiler put a
// MEMBAR.ALL.CTA in front of every fence.proxy.async, and 64-bit address arithmetic around every access)
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 u;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(addr));
  return u;
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 f;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f.x), "=f"(f.y), "=f"(f.z), "=f"(f.w) : "r"(addr));
  return f;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 u) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}

// rstd of row m of the LayerNorm input from the PRODUCING GEMM's per-(n-tile, half) partial sums.  Fixed summation
// order (bit-reproducible).
__device__ __forceinline__ float2 ln_row_rstd(const IGemmParams& p, long long m, bool valid) {
  float sum = 0.f, sq = 0.f;
  if (valid) {
    for (int i0 = 0; i0 < p.ln_parts; i0 += 4) {   // four independent loads in flight (a dependent chain of L2
      float2 t[4];                                 // latencies per tile sat on the epilogue's critical path otherwise)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        t[j] = (i0 + j < p.ln_parts) ? __ldg(p.ln_stats + (long long)(i0 + j) * p.M + m) : make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 4; ++j) { sum += t[j].x; sq += t[j].y; }
    }
  }
  const float mean = sum * p.ln_inv_k;
  const float var = fmaxf(sq * p.ln_inv_k - mean * mean, 0.f);
  const float rstd = rsqrtf(var + p.ln_eps);
  return make_float2(rstd, rstd);
}
// ... and the NEXT tile's statistics are pulled into L1 while this tile's epilogue runs (no registers held)
__device__ __forceinline__ void ln_prefetch(const IGemmParams& p, long long m) {
  if (m < p.M)
    for (int i = 0; i < p.ln_parts; ++i)
      asm volatile("prefetch.global.L1 [%0];" ::"l"(p.ln_stats + (long long)i * p.M + m));
}

// One GEGLU chunk of a row: 32 value + 32 gate accumulators -> 32 outputs  value * act(gate)  (+ bias, + folded LayerNorm:
// rstd * acc + bias', rstd == 1 otherwise).  ACT: 1 = SiLU (SwiGLU, DINOv2's FFN), 2 = GELU erf (GEGLU, diffusers).
// Straight-line code: the eight iterations are independent and the compiler interleaves their MUFU / FFMA2 chains.
template <int DT, int ACT, int NC>
__device__ __forceinline__ void geglu_math(const uint32_t (&rh)[NC], const uint32_t (&rg)[NC], const float* __restrict__ bias,
                                           int n0, int off, float2 rstd2, uint32_t (&o)[NC / 2], uint32_t bias_s = 0) {
  // rh / rg: NC value / gate accumulators of chunk columns [off, off + NC); bias = [32 value | 32 gate] per chunk,
  // read from the warp's staged copy in shared memory (bias_s: its 256 bytes for this chunk) when there is one
  const bool has_bias = bias != nullptr;
  const float* bp = has_bias ? bias + n0 + off : nullptr;
#pragma unroll
  for (int j = 0; j < NC; j += 4) {
    float4 bh = make_float4(0.f, 0.f, 0.f, 0.f), bg = bh;
    if (has_bias && bias_s) {
      bh = lds128f(bias_s + (uint32_t)(off + j) * 4u);
      bg = lds128f(bias_s + (uint32_t)(32 + off + j) * 4u);
    } else if (has_bias) {
      bh = __ldg(reinterpret_cast<const float4*>(bp + j));
      bg = __ldg(reinterpret_cast<const float4*>(bp + 32 + j));
    }
    const float2 h0 = __ffma2_rn(rstd2, make_float2(__uint_as_float(rh[j]), __uint_as_float(rh[j + 1])), make_float2(bh.x, bh.y));
    const float2 h1 = __ffma2_rn(rstd2, make_float2(__uint_as_float(rh[j + 2]), __uint_as_float(rh[j + 3])), make_float2(bh.z, bh.w));
    const float2 g0 = __ffma2_rn(rstd2, make_float2(__uint_as_float(rg[j]), __uint_as_float(rg[j + 1])), make_float2(bg.x, bg.y));
    const float2 g1 = __ffma2_rn(rstd2, make_float2(__uint_as_float(rg[j + 2]), __uint_as_float(rg[j + 3])), make_float2(bg.z, bg.w));
    const float2 v0 = __fmul2_rn(h0, ACT == 1 ? silu2_exact(g0) : gelu_erf2_f<DT>(g0));
    const float2 v1 = __fmul2_rn(h1, ACT == 1 ? silu2_exact(g1) : gelu_erf2_f<DT>(g1));
    o[j / 2] = pack2<DT>(v0.x, v0.y);
    o[j / 2 + 1] = pack2<DT>(v1.x, v1.y);
  }
}

// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) per 256 x BN tile —
// each CTA loads its own 128 activation rows and HALF of the weight tile, the leader CTA issues the M = 256 MMAs, each
// CTA's TMEM receives (and each CTA's epilogue stores) its own 128 rows.
// EW = epilogue warps.  8 everywhere except the GEGLU GEMMs with a short K, whose epilogue (two TMEM reads, a GELU and a
// product per output) out-lasts the tile's MMAs 2.4 : 1 with two warps per scheduler issuing on 17 % of their cycles
// (profiles/r2_s3_epilogue.md): EW = 16 puts four warps on every scheduler, one 64-column chunk of the tile each.
// An EW = 16 instance compiles the GEGLU epilogue only.
template <int BN, int DT, int CG, int EW = IG_EPI_WARPS>
__global__ void __launch_bounds__((4 + EW) * 32, 1) igemm_kernel(const __grid_constant__ IGemmParams p) {
  static_assert(EW == 8 || EW == 16, "epilogue warps");
  using Cfg = IGemmCfg<BN, CG>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_smem = smem + p.stages * Cfg::STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(epi_smem + EW * p.nbuf * IG_SLOT_BYTES);
  uint64_t* empty = full + IG_MAX_STAGES;
  uint64_t* tfull = empty + IG_MAX_STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* res_bar = tempty + 2;                 // [EW][IG_RES_SLOTS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + EW * IG_RES_SLOTS);
  // per-warp copy of the bias values of the chunks the warp owns in the current tile (see the epilogue): [EW][512 B]
  uint8_t* bias_smem = epi_smem + EW * p.nbuf * IG_SLOT_BYTES + 1024;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int mn_tiles = p.m_tiles * p.n_tiles;       // m_tiles counts 128*CG-row tiles
  const int total_tiles = mn_tiles * p.splits;      // split-K: tile = split * mn_tiles + m_blk * n_tiles + n_blk
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  pdl_launch_dependents();
  const int first_tile = (CG == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = (CG == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmA[0]);
    if (!p.out_f32) tma_prefetch_desc(p.mode == 4 ? &p.tmOutP[0] : &p.tmOut);
    if (p.has_res) tma_prefetch_desc(&p.tmRes);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full[i], 2);   // the activation producer's and the weight producer's expect_tx arrivals
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], EW * CG);
    }
    for (int i = 0; i < EW * IG_RES_SLOTS; ++i) mbar_init(&res_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    if (CG == 2) tmem_alloc_cg2(tmem_slot, Cfg::TMEM_COLS);
    else tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();   // peer barriers are initialised before anyone signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel's tail

  if (warp == 0) {
    // ===================== TMA producer, activations (whole warp, converged; one elected lane issues) ==========
    // (the weight tiles come from warp 3: two short issue chains in parallel instead of one long one per k-block)
    int stage = 0;
    uint32_t phase = 0;
    const int hw = p.H * p.W;
    const int nstages = p.stages;
    const uint32_t a_bytes = p.a_bytes;
    for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
      const int split = tile / mn_tiles, mn = tile - split * mn_tiles;
      const int m_blk = mn / p.n_tiles, n_blk = mn % p.n_tiles;
      const int m0 = (m_blk * CG + (int)rank) * 128;
      int b0 = 0, y0 = 0, x0 = 0;   // x0 != 0 only for rows wider than a tile (W > 128: 128-pixel row segments)
      if (p.mode != 0) {
        b0 = m0 / hw;
        const int rem = m0 - b0 * hw;
        y0 = rem / p.W;
        x0 = rem - y0 * p.W;
      }
      const int plane = p.mode == 4 ? split : 0;
      const int kb_begin = (p.mode == 4 ? 0 : split) * p.kb_per_split;
      const int kb_end = min(p.num_kb, kb_begin + p.kb_per_split);
      // The K range is walked segment by segment — one filter tap of a conv, or one of the (up to two) concatenated
      // A matrices of a GEMM — so that inside a segment only the channel coordinate moves and the per-k-block loop is
      // a barrier wait, one elected lane's expect_tx + two TMA issues, and a few adds: this thread's latency per
      // k-block has to stay below the MMA time of the k-block.
      int kb = kb_begin;
      while (kb < kb_end) {
        const CUtensorMap* tma = &p.tmA[0];
        int seg_end, c0, c1, c2 = 0, c3 = 0;
        if (p.mode == 0) {
          if (kb < p.kb_split) { c0 = kb * 64; seg_end = min(kb_end, p.kb_split); }
          else { tma = &p.tmA[1]; c0 = (kb - p.kb_split) * 64; seg_end = kb_end; }
          c1 = m0;
        } else {
          const int tap = kb / p.cblocks;
          const int r = tap / 3, s = tap - r * 3;
          c0 = (kb - tap * p.cblocks) * 64;
          seg_end = min(kb_end, (tap + 1) * p.cblocks);
          c3 = b0;
          if (p.mode == 4) {
            // output row 2i + py reads upsampled rows 2i + py - 1 .. + 1 = source rows {i - 1, i} (py = 0) or {i, i + 1}
            // (py = 1): a 2-tap filter whose taps were pre-summed on the host; columns alike
            c1 = x0 + (tap & 1) - 1 + (plane & 1);
            c2 = y0 + (tap >> 1) - 1 + (plane >> 1);
          } else if (p.mode == 1) {
            c1 = x0 + s - 1;
            c2 = y0 + r - 1;
          } else if (p.mode == 2) {
            // input row 2y + r - 1: r=0 -> odd plane, row y-1; r=1 -> even plane, row y; r=2 -> odd plane, row y
            const int py = (r != 1), px = (s != 1);
            tma = &p.tmA[py * 2 + px];
            c1 = x0 + ((s == 0) ? -1 : 0);
            c2 = y0 + ((r == 0) ? -1 : 0);
          } else {
            // mode 3, padding on the bottom/right only: input row 2y + r: r=0 -> even plane, row y; r=1 -> odd plane,
            // row y; r=2 -> even plane, row y+1 (row H of the plane is out of bounds = the zero padding)
            const int py = (r == 1), px = (s == 1);
            tma = &p.tmA[py * 2 + px];
            c1 = x0 + ((s == 2) ? 1 : 0);
            c2 = y0 + ((r == 2) ? 1 : 0);
          }
        }
        const bool is_gemm = p.mode == 0;
        for (; kb < seg_end; ++kb, c0 += 64) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (elect_one()) {
            uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
            if (CG == 1) {
              mbar_expect_tx(&full[stage], a_bytes);
              if (is_gemm) tma_load_2d(sa, tma, &full[stage], c0, c1);
              else tma_load_4d(sa, tma, &full[stage], c0, c1, c2, c3);
            } else {
              // pair: both CTAs' loads complete on the LEADER's full barrier, which expects the bytes of both
              if (rank == 0) mbar_expect_tx(&full[stage], 2u * a_bytes);
              const uint32_t fbar = mapa_u32(smem_u32(&full[stage]), 0);
              if (is_gemm) tma_load_2d_cg2(sa, tma, fbar, c0, c1);
              else tma_load_4d_cg2(sa, tma, fbar, c0, c1, c2, c3);
            }
          }
          __syncwarp();
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== TMA producer, weights =====================
    int stage = 0;
    uint32_t phase = 0;
    const int nstages = p.stages;
    const uint32_t b_bytes = p.b_bytes;
    for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
      const int split = tile / mn_tiles, mn = tile - split * mn_tiles;
      const int n_blk = mn % p.n_tiles;
      // mode 4: the four parity planes' weights are stacked along N ([4 * Cout, 4 * Cin])
      const int n0 = n_blk * BN + ((CG == 2) ? (int)rank * (Cfg::SUB_N / 2) : 0) + (p.mode == 4 ? split * p.N : 0);
      const int kb_begin = (p.mode == 4 ? 0 : split) * p.kb_per_split;
      const int kb_end = min(p.num_kb, kb_begin + p.kb_per_split);
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sb = smem + stage * Cfg::STAGE_BYTES + Cfg::A_BYTES;
          if (CG == 1) {
            mbar_expect_tx(&full[stage], b_bytes);
            tma_load_2d(sb, &p.tmB, &full[stage], kb * 64, n0);
          } else {
            if (rank == 0) mbar_expect_tx(&full[stage], 2u * b_bytes);
            const uint32_t fbar = mapa_u32(smem_u32(&full[stage]), 0);
            if (Cfg::NSUB == 1) {
              tma_load_2d_cg2(sb, &p.tmB, fbar, kb * 64, n0);
            } else {
              // 320-wide tile = two 160-column MMAs; of each, this CTA supplies 80 weight rows (box = 80 rows)
              tma_load_2d_cg2(sb, &p.tmB, fbar, kb * 64, n0);
              tma_load_2d_cg2(sb + (Cfg::SUB_N / 2) * 128, &p.tmB, fbar, kb * 64, n0 + Cfg::SUB_N);
            }
          }
        }
        __syncwarp();
        if (++stage == nstages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===================== MMA issuer (leader CTA of a pair; whole warp, one elected lane issues) =====================
    constexpr uint32_t idesc = make_idesc(DT, 128 * CG, Cfg::SUB_N, 0, 0);
    // SW128 K-major descriptor: high word constant (SBO 1024 B, version 1, 128-byte swizzle); low word = addr >> 4 |
    // LBO(16 B) << 16, advanced by plain adds (shared addresses < 256 KB never carry out of the 14-bit field)
    constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t desc_lo0 = ((smem_u32(smem) >> 4) & 0x3FFFu) | (1u << 16);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_STRIDE;
      const int split = p.mode == 4 ? 0 : tile / mn_tiles;
      const int nkb = min(p.num_kb, (split + 1) * p.kb_per_split) - split * p.kb_per_split;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = desc_lo0 + (uint32_t)stage * (Cfg::STAGE_BYTES >> 4);
          const uint32_t b_lo = a_lo + (Cfg::A_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < 4 && !IG_DBG(p, 16); ++k) {
            const uint64_t da = ((uint64_t)kDescHi << 32) | (a_lo + 2u * k);   // +32 B per k16 inside the swizzle atom
#pragma unroll
            for (int h = 0; h < Cfg::NSUB; ++h) {   // 320-wide pair tile: two N = 160 MMAs share the activation operand
              const uint64_t db = ((uint64_t)kDescHi << 32) | (b_lo + 2u * k + (uint32_t)h * ((Cfg::SUB_N / CG) * 128 >> 4));
              if (CG == 2) umma_ss_cg2(d_tmem + h * Cfg::SUB_N, da, db, idesc, (kb | k) != 0);
              else umma_ss(d_tmem + h * Cfg::SUB_N, da, db, idesc, (kb | k) != 0);
            }
          }
          // commits come from the SAME lane as the MMAs they track
          if (CG == 2) tc_commit_cg2(&empty[stage]);   // frees the stage in both CTAs
          else tc_commit(&empty[stage]);
          if (kb == nkb - 1) {
            if (CG == 2) tc_commit_cg2(&tfull[acc]);    // both CTAs' epilogues
            else tc_commit(&tfull[acc]);
          }
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (++acc == Cfg::NACC) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (EW warps: TMEM quadrant q, column-chunk phase `half` of CSTEP) ==============
    using T = typename TypeOf<DT>::T;
    const int e = warp - 4;
    const int q = e & 3;      // == warp % 4: the TMEM lane quadrant this warp may touch
    const int half = e >> 2;
    constexpr int CSTEP = EW / 4;   // warps per TMEM quadrant: each takes every CSTEP-th column chunk of the tile
    const int row = q * 32 + lane;
    uint8_t* slots = epi_smem + e * p.nbuf * IG_SLOT_BYTES;
    const uint32_t slots_s = smem_u32(slots);
    uint64_t* rbar = res_bar + e * IG_RES_SLOTS;
    uint32_t cnt = 0;         // chunks staged by this warp so far (slot rotation when there is no residual)
    uint32_t rphase = 0;      // bit i: parity of the next completion of residual barrier i
    int acc = 0;
    uint32_t acc_phase = 0;
    const int nbuf = p.nbuf;
    const uint32_t tempty_addr[2] = {mapa_u32(smem_u32(&tempty[0]), 0), mapa_u32(smem_u32(&tempty[1]), 0)};
    auto next_tile_row = [&](int t) -> long long {   // this thread's output row in tile t (>= M when there is none)
      if (t >= total_tiles) return (long long)p.M;
      const int mn_ = t % mn_tiles;
      return (long long)((mn_ / p.n_tiles) * CG + (int)rank) * 128 + row;
    };
    for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
      const int split = tile / mn_tiles, mn = tile - split * mn_tiles;
      const int m_blk = mn / p.n_tiles, n_blk = mn % p.n_tiles;
      const int m_cta0 = (m_blk * CG + (int)rank) * 128;
      const int m_warp0 = m_cta0 + q * 32;
      const long long m = (long long)m_cta0 + row;
      const bool valid = m < p.M;
      const int n_tile0 = n_blk * BN;
      const int cols_here = min(BN, p.N - n_tile0);
      const uint32_t t_row = tmem_base + acc * Cfg::ACC_STRIDE + ((uint32_t)(q * 32) << 16);
      const float* rv = (p.rowvec && valid) ? p.rowvec + (long long)(m / p.hw) * p.ld_rowvec : nullptr;
      if (EW == 8 && p.out_f32) {
        // ---- fp32 output (conv_out, stacked time-embedding projection): direct per-row stores ----
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = half; c * 32 < cols_here; c += CSTEP) {
          const int n0 = n_tile0 + c * 32;
          uint32_t r[32];
          tmem_ld32(t_row + c * 32, r);
          tc_wait_ld();
          if (valid) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            if (p.bias) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
              }
            }
            if (rv) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(rv + n0 + j));
                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
              }
            }
            if (p.silu == 1) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = silu_f(v[j]);
            } else if (p.silu == 2) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = gelu_erf_f(v[j]);
            }
            float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) +
                                                   ((long long)split * p.M + m) * p.ldo + n0);
#pragma unroll
            for (int j = 0; j < 8; ++j) op[j] = make_float4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
          }
        }
      } else if (EW == 8 && !p.geglu) {
        // ---- 16-bit output: TMEM -> registers -> (+bias, +rowvec, +residual, act) -> swizzled smem slot -> TMA store.
        //      ALL residual chunks this warp owns in the tile are TMA-prefetched (one slot each) before the wait for the
        //      tile's MMAs: their HBM latency overlaps the mainloop instead of being paid chunk by chunk (a short-K
        //      GEMM — K = 320..1280, ~1 us of MMAs per tile — was bound by exactly that latency chain). ----
        const bool use_res = p.has_res && !IG_DBG(p, 2);
        float2 ln_rstd2 = make_float2(1.f, 1.f);
        if (p.ln_stats) {
          ln_rstd2 = ln_row_rstd(p, m, valid);
          ln_prefetch(p, next_tile_row(tile + tile_step));
        }
        float2 st_sum = make_float2(0.f, 0.f), st_sq = make_float2(0.f, 0.f);
        if (use_res) {
          if (lane == 0) {
            bulk_wait_read<0>();   // this warp's earlier stores have finished reading their slots
            int i = 0;
            for (int c = half; c * 32 < cols_here; c += CSTEP, ++i) {
              mbar_expect_tx(&rbar[i], IG_SLOT_BYTES);
              tma_load_2d(slots + i * IG_SLOT_BYTES, &p.tmRes, &rbar[i], n_tile0 + c * 32, m_warp0);
            }
          }
        }
#ifndef PCDM_NO_BIAS_SMEM
        // The bias values of this warp's (up to four) chunks go through a warp-private 512 B of shared memory: one
        // coalesced 16-byte load per lane, issued BEFORE the wait for the tile's MMAs, instead of eight dependent
        // broadcast loads per chunk between the TMEM read-out and the arithmetic.
        const uint32_t bias_s = smem_u32(bias_smem) + (uint32_t)e * 512u;
        const bool bias_staged = p.bias != nullptr;
        if (bias_staged) {
          const int c = half + (lane >> 3) * CSTEP;
          if (c * 32 < cols_here)
            sts128(bias_s + (uint32_t)lane * 16u,
                   __ldg(reinterpret_cast<const uint4*>(p.bias + n_tile0 + c * 32 + (lane & 7) * 4)));
          __syncwarp();
        }
#define IG_BIAS4(ci_, n0_, j_) ((bias_staged && (ci_) < 4) ? lds128f(bias_s + (uint32_t)(ci_) * 128u + (uint32_t)(j_) * 4u) \
                                                        : __ldg(reinterpret_cast<const float4*>(p.bias + (n0_) + (j_))))
#else
#define IG_BIAS4(ci_, n0_, j_) __ldg(reinterpret_cast<const float4*>(p.bias + (n0_) + (j_)))
#endif
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        int ci = 0;
#pragma unroll 1
        for (int c = half; c * 32 < cols_here && !IG_DBG(p, 8); c += CSTEP, ++ci) {
          const int n0 = n_tile0 + c * 32;
          const uint32_t slot = use_res ? (uint32_t)ci : (cnt & (uint32_t)(nbuf - 1));   // nbuf is 2 or 4
          uint8_t* sl = slots + slot * IG_SLOT_BYTES;
          const uint32_t sl_s = slots_s + slot * IG_SLOT_BYTES;
          if (!use_res) {
            if (lane == 0) bulk_wait_read<1>();   // every store but the most recent has finished reading its slot
            __syncwarp();
          }
          // (requesting chunk c + 2 here, under the arithmetic of chunk c, was measured 10-15 % SLOWER on the short-K GEMMs:
          //  profiles/r2_s3_epilogue.md)
          uint32_t r[32];
          tmem_ld32(t_row + c * 32, r);
          tc_wait_ld();
          float2 v[16];   // column pairs, packed fp32 (FFMA2 path)
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = make_float2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
          if (p.ln_stats) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {   // rstd * acc + bias'[n]: the same issue slots as a plain bias add
              const float4 b = IG_BIAS4(ci, n0, j);
              v[j / 2] = __ffma2_rn(ln_rstd2, v[j / 2], make_float2(b.x, b.y));
              v[j / 2 + 1] = __ffma2_rn(ln_rstd2, v[j / 2 + 1], make_float2(b.z, b.w));
            }
          } else if (p.bias && !IG_DBG(p, 4)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = IG_BIAS4(ci, n0, j);
              v[j / 2] = __fadd2_rn(v[j / 2], make_float2(b.x, b.y));
              v[j / 2 + 1] = __fadd2_rn(v[j / 2 + 1], make_float2(b.z, b.w));
            }
          }
          if (rv && !IG_DBG(p, 4)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(rv + n0 + j));
              v[j / 2] = __fadd2_rn(v[j / 2], make_float2(b.x, b.y));
              v[j / 2 + 1] = __fadd2_rn(v[j / 2 + 1], make_float2(b.z, b.w));
            }
          }
          if (use_res) {
            mbar_wait(&rbar[slot], (rphase >> slot) & 1u);
            rphase ^= 1u << slot;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 u = lds128(sl_s + sw64(lane, j));
              v[j * 4 + 0] = __fadd2_rn(v[j * 4 + 0], unpack2<DT>(u.x));
              v[j * 4 + 1] = __fadd2_rn(v[j * 4 + 1], unpack2<DT>(u.y));
              v[j * 4 + 2] = __fadd2_rn(v[j * 4 + 2], unpack2<DT>(u.z));
              v[j * 4 + 3] = __fadd2_rn(v[j * 4 + 3], unpack2<DT>(u.w));
            }
          }
          if (p.silu == 1) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = silu2_exact(v[j]);
          } else if (p.silu == 2) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = gelu_erf2_f<DT>(v[j]);
          }
          if (p.stats_out) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              st_sum = __fadd2_rn(st_sum, v[j]);
              st_sq = __ffma2_rn(v[j], v[j], st_sq);
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 u;
            u.x = pack2<DT>(v[j * 4 + 0].x, v[j * 4 + 0].y);
            u.y = pack2<DT>(v[j * 4 + 1].x, v[j * 4 + 1].y);
            u.z = pack2<DT>(v[j * 4 + 2].x, v[j * 4 + 2].y);
            u.w = pack2<DT>(v[j * 4 + 3].x, v[j * 4 + 3].y);
            sts128(sl_s + sw64(lane, j), u);
          }
          fence_proxy_async();
          __syncwarp();
          if (p.chan_stats) {
            // column sums over the warp's 32 rows, read back from the staged (rounded) tile: lane (w, par) takes the two
            // columns of 32-bit word w from rows par, par + 2, ... — even rows sit in banks 0-15, odd rows in 16-31, so
            // the 16 loads are conflict-free; one xor-16 shuffle folds the two row parities (fixed order)
            const int w = lane & 15, par = lane >> 4;
            const int rows_ok = p.M - m_warp0;   // rows of this slab that exist (>= 32: all)
            float2 cs = make_float2(0.f, 0.f), cq = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int r = 2 * i + par;
              uint32_t word;
              asm volatile("ld.shared.u32 %0, [%1];" : "=r"(word) : "r"(sl_s + sw64(r, w >> 2) + (uint32_t)(w & 3) * 4u));
              if (r < rows_ok) {
                const float2 f = unpack2<DT>(word);
                cs = __fadd2_rn(cs, f);
                cq = __ffma2_rn(f, f, cq);
              }
            }
            cs.x += __shfl_xor_sync(0xffffffffu, cs.x, 16);
            cs.y += __shfl_xor_sync(0xffffffffu, cs.y, 16);
            cq.x += __shfl_xor_sync(0xffffffffu, cq.x, 16);
            cq.y += __shfl_xor_sync(0xffffffffu, cq.y, 16);
            if (par == 0 && rows_ok > 0) {
              long long slab = m_warp0 >> 5;
              if (p.mode == 4) {   // an image's slabs stay contiguous: [image][parity plane][32-pixel slab of the plane]
                const int hw = p.H * p.W;
                const int bw = m_warp0 / hw;
                slab = (long long)bw * (hw >> 3) + (long long)split * (hw >> 5) + ((m_warp0 - bw * hw) >> 5);
              }
              *reinterpret_cast<float4*>(p.chan_stats + (slab * p.N + n0 + 2 * w) * 2) = make_float4(cs.x, cq.x, cs.y, cq.y);
            }
          }
          if (lane == 0 && !IG_DBG(p, 1)) {
            if (p.mode == 4) {   // the warp's 32 low-resolution pixels land on every second pixel of every second row
              const int hw = p.H * p.W;
              const int bw = m_warp0 / hw, rem = m_warp0 - bw * hw;
              tma_store_4d(&p.tmOutP[split], sl, n0, rem % p.W, rem / p.W, bw);
            } else {
              tma_store_2d(&p.tmOut, sl, n0, m_warp0);   // rows >= M / columns >= N are clipped by the tensor map
            }
            bulk_commit();
          }
          ++cnt;
        }
        if (p.stats_out && valid)   // one slot per (n-tile, column half); a warp without chunks still writes its zeros
          p.stats_out[(long long)(n_blk * 2 + half) * p.M + m] = make_float2(st_sum.x + st_sum.y, st_sq.x + st_sq.y);
      } else {
        // ---- GEGLU: weight rows were packed as groups of [32 value | 32 gate]; out[:, g*32 + j] = val * gelu(gate) ----
        float2 ln_rstd2 = make_float2(1.f, 1.f);   // stays 1 without a folded LayerNorm: rstd * acc + bias == acc + bias
        if (p.ln_stats) {
          ln_rstd2 = ln_row_rstd(p, m, valid);
          ln_prefetch(p, next_tile_row(tile + tile_step));
        }
        uint32_t gb_s = 0;   // this warp's staged bias (two 256-byte chunks at EW = 8, one at EW = 16)
#ifndef PCDM_NO_BIAS_SMEM
        if (p.bias) {
          gb_s = smem_u32(bias_smem) + (uint32_t)e * 512u;
          const int c = half + (lane >> 4) * CSTEP;
          if (c * 64 < cols_here && (EW == 8 || lane < 16))
            sts128(gb_s + (uint32_t)lane * 16u,
                   __ldg(reinterpret_cast<const uint4*>(p.bias + n_tile0 + c * 64 + (lane & 15) * 4)));
          __syncwarp();
        }
#endif
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        int gci = 0;
#pragma unroll 1
        for (int c = half; c * 64 < cols_here && !IG_DBG(p, 8); c += CSTEP, ++gci) {
          const int n0 = n_tile0 + c * 64;
          const uint32_t slot = cnt & (uint32_t)(nbuf - 1);
          uint8_t* sl = slots + slot * IG_SLOT_BYTES;
          const uint32_t sl_s = slots_s + slot * IG_SLOT_BYTES;
          if (lane == 0) bulk_wait_read<1>();
          __syncwarp();
          uint32_t rh[32], rg[32];
          tmem_ld32(t_row + c * 64, rh);
          tmem_ld32(t_row + c * 64 + 32, rg);
          tc_wait_ld();
          // packed fp32 (FFMA2) arithmetic.  The gate activation is chosen OUTSIDE the unrolled loop: a warp-uniform
          // `silu ? a : b` inside it compiled to a branch per element pair, which fenced the eight independent GELU
          // chains off from each other (no instruction-level parallelism: the K = 320 GEGLU GEMM ran at 0.4 of its bound).
          // (Streaming the chunk in two 16 + 16 column halves with the TMEM reads one half ahead of the GELUs was
          //  measured 8 % SLOWER: profiles/r2_s3_epilogue.md.)
          uint32_t o[16];
          if (IG_DBG(p, 32)) {   // experiment build: no arithmetic, the accumulators go out as they are
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = pack2<DT>(__uint_as_float(rh[2 * j]), __uint_as_float(rg[2 * j + 1]));
          } else if (p.silu == 1) geglu_math<DT, 1, 32>(rh, rg, p.bias, n0, 0, ln_rstd2, o, (gb_s && gci < 2) ? gb_s + gci * 256u : 0u);
          else geglu_math<DT, 2, 32>(rh, rg, p.bias, n0, 0, ln_rstd2, o, (gb_s && gci < 2) ? gb_s + gci * 256u : 0u);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128(sl_s + sw64(lane, j), make_uint4(o[j * 4], o[j * 4 + 1], o[j * 4 + 2], o[j * 4 + 3]));
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && !IG_DBG(p, 1)) {
            tma_store_2d(&p.tmOut, sl, n0 >> 1, m_warp0);
            bulk_commit();
          }
          ++cnt;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(tempty_addr[acc]);   // the MMA issuer lives in the leader CTA
        else mbar_arrive(&tempty[acc]);
      }
      if (++acc == Cfg::NACC) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) bulk_wait<0>();   // all TMA stores have landed before the CTA (and its shared memory) goes away
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all();   // nobody leaves (or frees TMEM) while the peer may still signal / read this CTA
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_cg2(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// Split-K finish: out[m, n] = act( sum_s part[s][m][n] + bias[n] + rowvec[m / hw][n] + residual[m][n] ), fixed
// summation order (bit-reproducible); 8 columns per thread.
template <int DT>
__global__ void splitk_finish_kernel(const float* __restrict__ part, int splits, int M, int N,
                                     const float* __restrict__ bias, const float* __restrict__ rowvec,
                                     long long ld_rowvec, int hw, const void* __restrict__ residual, long long ldr,
                                     void* __restrict__ out, long long ldo, int silu) {
  using T = typename TypeOf<DT>::T;
  pdl_launch_dependents();
  pdl_wait();
  const int nv = N / 8;
  const long long total = (long long)M * nv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n0 = (int)(i % nv) * 8;
    const long long m = i / nv;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    for (int sidx = 0; sidx < splits; ++sidx) {
      const float4* pp = reinterpret_cast<const float4*>(part + ((long long)sidx * M + m) * N + n0);
      const float4 a = __ldg(pp), b = __ldg(pp + 1);
      v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    if (bias) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(bias + n0)), b = __ldg(reinterpret_cast<const float4*>(bias + n0 + 4));
      v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    if (rowvec) {
      const float* rv = rowvec + (m / hw) * ld_rowvec + n0;
      const float4 a = __ldg(reinterpret_cast<const float4*>(rv)), b = __ldg(reinterpret_cast<const float4*>(rv + 4));
      v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    if (residual) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(residual) + m * ldr + n0));
      float2 f;
      f = unpack2<DT>(u.x); v[0] += f.x; v[1] += f.y;
      f = unpack2<DT>(u.y); v[2] += f.x; v[3] += f.y;
      f = unpack2<DT>(u.z); v[4] += f.x; v[5] += f.y;
      f = unpack2<DT>(u.w); v[6] += f.x; v[7] += f.y;
    }
    if (silu == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = silu_f(v[j]);
    } else if (silu == 2) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = gelu_erf_f(v[j]);
    }
    uint4 o;
    o.x = pack2<DT>(v[0], v[1]); o.y = pack2<DT>(v[2], v[3]); o.z = pack2<DT>(v[4], v[5]); o.w = pack2<DT>(v[6], v[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<T*>(out) + m * ldo + n0) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
constexpr int IG_SMEM_LIMIT = 227 * 1024;

template <int BN, int DT, int CG, int EW = IG_EPI_WARPS>
static int launch_igemm(IGemmParams& p, cudaStream_t stream) {
  using Cfg = IGemmCfg<BN, CG>;
  PCDM_ENSURE_SMEM(IG_SMEM_LIMIT, igemm_kernel<BN, DT, CG, EW>);
  p.nbuf = p.has_res ? (BN > 256 ? 5 : 4) : 2;   // residual: one slot per 32-column chunk a warp owns in the tile
  p.dbg = g_tune.gemm_dbg;
  const int fixed = 1024 /*align slack*/ + 1024 /*barriers*/ + EW * 512 /*bias staging*/ + EW * p.nbuf * IG_SLOT_BYTES;
  int stages = (IG_SMEM_LIMIT - fixed) / Cfg::STAGE_BYTES;
  if (stages > IG_MAX_STAGES) stages = IG_MAX_STAGES;
  if (stages > g_tune.max_stages) stages = g_tune.max_stages;
  if (stages < 2) return set_error(PCDM_ERR_UNSUPPORTED, "igemm: not enough shared memory for a 2-stage ring");
  p.stages = stages;
  const int smem_bytes = fixed + stages * Cfg::STAGE_BYTES;
  const int total = p.m_tiles * p.n_tiles * p.splits;
  constexpr int threads = (4 + EW) * 32;
  if (CG == 1) {
    const int grid = total < num_sms() ? total : num_sms();
    PCDM_CUDA(launch_kernel(igemm_kernel<BN, DT, CG, EW>, dim3(grid), dim3(threads), smem_bytes, stream, 1, p));
  } else {
    const int pairs = num_sms() / 2;
    const int grid = 2 * (total < pairs ? total : pairs);
    PCDM_CUDA(launch_kernel(igemm_kernel<BN, DT, CG, EW>, dim3(grid), dim3(threads), smem_bytes, stream, 2, p));
  }
  PCDM_CUDA(cudaGetLastError());
  return 0;
}

// GEGLU GEMMs on 256-wide CTA-pair tiles run the 16-epilogue-warp instance up to this many 64-deep k-blocks per tile
// (measured, profiles/r2_s3_epilogue.md: K = 320 54.9 -> 51.6 us; K = 640 38.5 -> 38.9, K = 1280 unchanged — there the
// tile's MMAs out-last the epilogue anyway and the 32 KB of extra staging slots cost a pipeline stage)
#ifndef PCDM_GEGLU_EW16_MAX_KB
#define PCDM_GEGLU_EW16_MAX_KB 5
#endif

// Tile-shape selection from a cost model fitted to B200 measurements (tools/autotune.py, profiles/r1_autotune2.json):
// cycles per 64-deep k-block of one CTA in steady state.  Once the issue chains were shortened the loop is bound by
// L2 -> shared-memory delivery (~70 B/clk/SM: (128 + BN / cta_group) x 128 B per k-block) or by the MMAs (2 BN
// cycles), which makes 160-wide tiles on CTA pairs the best shape for this UNet (every N is a multiple of 160).
// The 320-wide pair tile is chosen automatically only where it was measured faster (tools/dev_bn320.py,
// profiles/r2_bn320.md): N = 320 with K >= 5760 (the 640 -> 320 and 960 -> 320 convs of the 32x64 level: the k-loop of a
// 160-wide tile is paced by the activation fill, 26 KB per 320 MMA cycles; the 320-wide one stages 36 KB per 640 and
// its unhidden epilogue is amortised over >= 90 k-blocks).  Extending it to N = 640 at the 16x32 level was mixed
// (640 -> 640 47.1 vs 50.6 us, 1280 -> 640 93.4 vs 85.4): not taken.  Everywhere else 160 / 256 win; explicit bn = 320 stays.
#ifdef PCDM_NO_AUTO_BN320
static bool g_tune_bn320(int, int, int) { return false; }
#else
static bool g_tune_bn320(int N, int num_kb, int geglu) { return N == 320 && num_kb >= 90 && !geglu; }
#endif

static int kb_cycles(int bn, int cg) {
  if (cg == 2) return bn == 128 ? 361 : (bn == 160 ? 398 : (bn == 320 ? 680 : 629));
  return bn == 64 ? 446 : (bn == 128 ? 479 : (bn == 160 ? 549 : 619));
}

static void pick_tile(int M, int N, int num_kb, int geglu, int has_res, int force_cg, int planes, int* bn_out, int* cg_out) {
  const int cand[5] = {160, 256, 128, 64, 320};
  long long best = 1LL << 62;
  *bn_out = 128;
  *cg_out = 1;
  for (int cg = 1; cg <= 2; ++cg) {
    if (force_cg && cg != force_cg) continue;
    if (cg == 2 && M <= 128) continue;
    for (int i = 0; i < 5; ++i) {
      const int bn = cand[i];
      if (geglu && (bn % 64)) continue;
      if (bn == 160 && (N % 160)) continue;
      if (bn == 320 && (cg != 2 || (N % 320) || !g_tune_bn320(N, num_kb, geglu))) continue;
      if (cg == 2 && bn < 128) continue;
      if (planes > 1 && (N % bn)) continue;   // stacked per-plane weights: an N tile must not straddle two planes
      const long long tiles = (long long)((M + 128 * cg - 1) / (128 * cg)) * ((N + bn - 1) / bn) * planes;
      const long long slots = num_sms() / cg;
      const long long waves = (tiles + slots - 1) / slots;
      // per tile: the k loop, plus the part of the epilogue / pipeline turn-around that is not hidden
      // (the 320-wide tile's accumulator is single-buffered: its epilogue is not hidden behind the next tile's MMAs)
      const long long tile_cycles = (long long)num_kb * kb_cycles(bn, cg) + 1500 + (geglu ? 1500 : 0) + (has_res ? 300 : 0) +
                                    (cg == 2 ? 400 : 0) + (bn == 320 ? 2500 : 0);
      long long cost = waves * tile_cycles;
      if (bn == 320) cost += cost / 12;   // ties go to the double-buffered tiles (32x64 640 -> 640: measured 162 vs 193 us)
      if (cost < best) { best = cost; *bn_out = bn; *cg_out = cg; }
    }
  }
  // GEGLU (epilogue-paced): 256-wide tiles on CTA pairs measured 3-8 % faster than on single CTAs at M = 2048 .. 32768
  // (tools/dev_cg.py) — half the weight traffic per SM leaves the epilogue's TMA stores more of the fabric
  if (geglu && !force_cg && planes == 1 && M >= 2048 && (N % 256) == 0) { *bn_out = 256; *cg_out = 2; }
}

struct ExtArgs {   // what the caller's pcdm_ext carries (all optional)
  int force_cg = 0;          // 0 = auto, 1 / 2 = force
  void* ws = nullptr;        // fp32 scratch for split-K partials; without it the launch never splits K
  long long ws_bytes = 0;
  float* row_stats = nullptr;   // producer side of a folded LayerNorm
  int row_stats_cap = 0;
  const float* ln_stats = nullptr;   // consumer side
  int ln_parts = 0;
  float ln_eps = 0.f;
  float* chan_stats = nullptr;  // GroupNorm statistics of the output, per 32-row slab and channel
  pcdm_ext* raw = nullptr;      // for the output field row_stats_parts
};

static int read_ext(pcdm_ext* ext, ExtArgs* e) {
  e->force_cg = g_tune.force_cg;
  if (!ext) return 0;
  if (ext->size < (int)sizeof(pcdm_ext)) return set_error(PCDM_ERR_INVALID, "pcdm_ext: size field does not match this library's struct");
  if (ext->force_cta_group < 0 || ext->force_cta_group > 2) return set_error(PCDM_ERR_INVALID, "pcdm_ext: force_cta_group must be 0, 1 or 2");
  if (ext->workspace_bytes < 0 || (!ext->workspace && ext->workspace_bytes != 0))
    return set_error(PCDM_ERR_INVALID, "pcdm_ext: bad workspace arguments");
  if (reinterpret_cast<uintptr_t>(ext->workspace) & 15) return set_error(PCDM_ERR_INVALID, "pcdm_ext: workspace must be 16-byte aligned");
  if (ext->force_cta_group) e->force_cg = ext->force_cta_group;
  e->ws = ext->workspace;
  e->ws_bytes = ext->workspace_bytes;
  ext->row_stats_parts = 0;
  if (ext->row_stats) {
    if (ext->row_stats_cap <= 0 || (reinterpret_cast<uintptr_t>(ext->row_stats) & 7))
      return set_error(PCDM_ERR_INVALID, "pcdm_ext: row_stats needs row_stats_cap > 0 and an 8-byte aligned buffer");
    e->row_stats = ext->row_stats;
    e->row_stats_cap = ext->row_stats_cap;
  }
  if (ext->ln_stats) {
    if (ext->ln_parts <= 0 || !(ext->ln_eps >= 0.f) || (reinterpret_cast<uintptr_t>(ext->ln_stats) & 7))
      return set_error(PCDM_ERR_INVALID, "pcdm_ext: ln_stats needs ln_parts > 0, ln_eps >= 0 and an 8-byte aligned buffer");
    e->ln_stats = ext->ln_stats;
    e->ln_parts = ext->ln_parts;
    e->ln_eps = ext->ln_eps;
  }
  if (ext->chan_stats) {
    if (reinterpret_cast<uintptr_t>(ext->chan_stats) & 15) return set_error(PCDM_ERR_INVALID, "pcdm_ext: chan_stats must be 16-byte aligned");
    e->chan_stats = ext->chan_stats;
  }
  e->raw = ext;
  return 0;
}

struct EpiArgs {   // the fused-epilogue operands as the caller gave them
  const float* bias;
  const float* rowvec;
  long long ld_rowvec;
  int hw;
  const void* residual;
  long long ldr;
  void* out;
  long long ldo;
  int silu;
};

static int dispatch_igemm(IGemmParams& p, int dt, int bn, const void* w, int K, const void* residual, long long ldr,
                          const ExtArgs& ext, cudaStream_t stream, int planes = 1) {
  void* const g_ws = ext.ws;
  const long long g_ws_bytes = ext.ws_bytes;
  const int g_force_cg = ext.force_cg;
  p.splits = 1;
  p.kb_per_split = p.num_kb;
  // ---- split-K for tile-starved problems (the 4x8 / 8x16 levels: M = 64..2048 rows but K up to 23 040): slice K over
  //      otherwise idle SMs into fp32 partials, then one small finishing kernel applies the epilogue ----
  EpiArgs epi = {p.bias, p.rowvec, p.ld_rowvec, p.hw, residual, ldr, p.out, p.ldo, p.silu};
  bool split = false;
  if (ext.ln_stats) {
    if (p.out_f32 || !p.bias || p.mode != 0)
      return set_error(PCDM_ERR_UNSUPPORTED, "gemm: a folded LayerNorm needs a 16-bit-output GEMM with the bias vector");
    p.ln_stats = reinterpret_cast<const float2*>(ext.ln_stats);
    p.ln_parts = ext.ln_parts;
    p.ln_eps = ext.ln_eps;
    p.ln_inv_k = 1.0f / (float)K;
  }
  if (ext.row_stats && (p.out_f32 || p.geglu))
    return set_error(PCDM_ERR_UNSUPPORTED, "gemm: row statistics come with the plain 16-bit-output epilogue only");
  if (ext.chan_stats) {
    if (p.out_f32 || p.geglu) return set_error(PCDM_ERR_UNSUPPORTED, "channel statistics come with the plain 16-bit-output epilogue only");
    if (p.hw % 32) return set_error(PCDM_ERR_UNSUPPORTED, "channel statistics need rows-per-image (H*W) to be a multiple of 32");
    p.chan_stats = ext.chan_stats;   // (a launch that emits statistics never takes the split-K route)
  }
  if (bn == 0 && g_ws && !p.out_f32 && !p.geglu && p.num_kb >= 64 && (p.N % 8) == 0 && !ext.ln_stats && !ext.row_stats && !ext.chan_stats && planes == 1) {   // K >= 4096: below, one launch wins
    const int sbn = (p.N % 160 == 0) ? 160 : 128;
    const int tiles = p.m_tiles * ((p.N + sbn - 1) / sbn);
    if (tiles * 2 <= num_sms()) {
      int splits = num_sms() / tiles;
      if (splits > 8) splits = 8;
      while (splits > 1 && p.num_kb / splits < 8) --splits;
      const int kbps = (p.num_kb + splits - 1) / splits;
      splits = (p.num_kb + kbps - 1) / kbps;
      if (splits > 1 && (long long)splits * p.M * p.N * 4 <= g_ws_bytes) {
        split = true;
        bn = sbn;
        p.splits = splits;
        p.kb_per_split = kbps;
        p.bias = nullptr; p.rowvec = nullptr; p.silu = 0;
        p.out = g_ws; p.ldo = p.N; p.out_f32 = 1;
        residual = nullptr;
      }
    }
  }
  int cg = 1;
  if (bn == 0) {
    pick_tile(p.M, p.N, p.kb_per_split, p.geglu, residual != nullptr, split ? 1 : g_force_cg, planes, &bn, &cg);
  } else {
    // explicit tile width (tests / tuning): CTA pairs (256-row tiles) when forced, or by the same model
    int bn_model;
    pick_tile(p.M, p.N, p.kb_per_split, p.geglu, residual != nullptr, split ? 1 : g_force_cg, planes, &bn_model, &cg);
    if (planes > 1 && (p.N % bn)) return set_error(PCDM_ERR_INVALID, "conv3x3_up2x: the forced N tile must divide Cout");
    if (g_force_cg == 0) cg = (p.M > 128 && bn >= 128 && !split && (bn == 320 || kb_cycles(bn, 2) < kb_cycles(bn, 1))) ? 2 : 1;
    if (bn < 128 || p.M <= 128 || split) cg = 1;
  }
  if (cg == 2) p.m_tiles = (p.M + 255) / 256;
  if (planes > 1) { p.splits = planes; p.kb_per_split = p.num_kb; }
  p.has_res = residual ? 1 : 0;
  if (p.has_res && (p.geglu || p.out_f32))
    return set_error(PCDM_ERR_UNSUPPORTED, "igemm: residual cannot be combined with GEGLU or fp32 output");
  if (!p.out_f32 && planes == 1) {
    const uint64_t n_out = p.geglu ? (uint64_t)p.N / 2 : (uint64_t)p.N;
    const uint64_t dims[2] = {n_out, (uint64_t)p.M};
    const uint64_t strides[1] = {(uint64_t)p.ldo * 2};
    const uint32_t box[2] = {32, 32};
    PCDM_CHECK(make_tmap(&p.tmOut, p.out, 2, dims, strides, box, 64), "output tensor map");
    if (p.has_res) {
      const uint64_t rdims[2] = {(uint64_t)p.N, (uint64_t)p.M};
      const uint64_t rstrides[1] = {(uint64_t)ldr * 2};
      PCDM_CHECK(make_tmap(&p.tmRes, residual, 2, rdims, rstrides, box, 64), "residual tensor map");
    }
  }
  p.n_tiles = (p.N + bn - 1) / bn;
  if (ext.row_stats) {
    if (2 * p.n_tiles > ext.row_stats_cap)
      return set_error(PCDM_ERR_INVALID, "gemm: row_stats_cap %d is below the %d slots this problem writes (2 * N / 64 always fits)",
                       ext.row_stats_cap, 2 * p.n_tiles);
    p.stats_out = reinterpret_cast<float2*>(ext.row_stats);
    ext.raw->row_stats_parts = 2 * p.n_tiles;
  }
  const int bbox = bn == 320 ? 80 : bn / cg;    // 320-wide pair tile: two boxes of 80 rows per CTA and k-block
  const int brows = p.N < bbox ? p.N : bbox;
  p.b_bytes = (uint32_t)brows * 128u * (bn == 320 ? 2u : 1u);
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)p.N * planes};
    const uint64_t strides[1] = {(uint64_t)K * 2};
    const uint32_t box[2] = {64, (uint32_t)brows};
    PCDM_CHECK(make_tmap(&p.tmB, w, 2, dims, strides, box), "weight tensor map");
  }
  int rc;
#define PCDM_LAUNCH(BN_, CG_)                                                      \
  (dt == DT_F16 ? launch_igemm<BN_, DT_F16, CG_>(p, stream) : launch_igemm<BN_, DT_BF16, CG_>(p, stream))
  switch (bn) {
    case 64: rc = PCDM_LAUNCH(64, 1); break;
    case 128: rc = cg == 2 ? PCDM_LAUNCH(128, 2) : PCDM_LAUNCH(128, 1); break;
    case 160: rc = cg == 2 ? PCDM_LAUNCH(160, 2) : PCDM_LAUNCH(160, 1); break;
    case 256:
      if (cg == 2 && p.geglu && p.kb_per_split <= PCDM_GEGLU_EW16_MAX_KB)
        rc = dt == DT_F16 ? launch_igemm<256, DT_F16, 2, 16>(p, stream) : launch_igemm<256, DT_BF16, 2, 16>(p, stream);
      else
        rc = cg == 2 ? PCDM_LAUNCH(256, 2) : PCDM_LAUNCH(256, 1);
      break;
    case 320:
      if (cg != 2 || (p.N % 320)) return set_error(PCDM_ERR_UNSUPPORTED, "igemm: the 320-wide tile needs a CTA pair (M > 128) and N % 320 == 0");
      rc = PCDM_LAUNCH(320, 2);
      break;
    default: return set_error(PCDM_ERR_INVALID, "igemm: BN must be 0, 64, 128, 160, 256 or 320");
  }
#undef PCDM_LAUNCH
  if (rc != 0 || !split) return rc;
  const long long total = (long long)p.M * (p.N / 8);
  long long grid = (total + 255) / 256;
  if (grid > (long long)num_sms() * 8) grid = (long long)num_sms() * 8;
  if (dt == DT_F16)
    PCDM_CUDA(launch_kernel(splitk_finish_kernel<DT_F16>, dim3((int)grid), dim3(256), 0, stream, 1,
                            reinterpret_cast<const float*>(g_ws), p.splits, p.M, p.N, epi.bias, epi.rowvec,
                            epi.ld_rowvec, epi.hw, epi.residual, epi.ldr, epi.out, epi.ldo, epi.silu));
  else
    PCDM_CUDA(launch_kernel(splitk_finish_kernel<DT_BF16>, dim3((int)grid), dim3(256), 0, stream, 1,
                            reinterpret_cast<const float*>(g_ws), p.splits, p.M, p.N, epi.bias, epi.rowvec,
                            epi.ld_rowvec, epi.hw, epi.residual, epi.ldr, epi.out, epi.ldo, epi.silu));
  PCDM_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace pcdm

namespace pcdm {
int skinny_gemm_try(const void* a, long long lda, const void* w, void* out, long long ldo, const float* bias,
                    const float* rowvec, long long ld_rowvec, int rows_per_image, const void* residual, long long ldr,
                    int M, int N, int K, int dtype, int flags, cudaStream_t stream, const float* ln_gamma,
                    const float* ln_beta, float ln_eps);   // skinny.cu
}

using namespace pcdm;

extern "C" int pcdm_gemm(const void* a, long long lda, const void* a2, long long lda2, int k1, const void* w,
                         void* out, long long ldo, const float* bias, const float* rowvec, long long ld_rowvec,
                         int rows_per_image, const void* residual, long long ldr, int M, int N, int K, int dtype, int flags, int bn,
                         pcdm_ext* ext_, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ExtArgs ext;
  PCDM_CHECK(read_ext(ext_, &ext), "ext");
  if (!a || !w || !out) return set_error(PCDM_ERR_INVALID, "gemm: null pointer");
  if (dtype != DT_F16 && dtype != DT_BF16) return set_error(PCDM_ERR_INVALID, "gemm: dtype must be 0 (f16) or 1 (bf16)");
  if (M <= 0 || N <= 0 || K <= 0) return set_error(PCDM_ERR_INVALID, "gemm: empty problem");
  if (K % 64 || N % 32) return set_error(PCDM_ERR_UNSUPPORTED, "gemm: K must be a multiple of 64 and N of 32");
  if (a2 && (k1 % 64 || k1 <= 0 || k1 >= K)) return set_error(PCDM_ERR_INVALID, "gemm: bad K split");
  const bool geglu = flags & PCDM_FLAG_GEGLU;
  if (geglu && (N % 64)) return set_error(PCDM_ERR_UNSUPPORTED, "gemm: GEGLU needs N % 64 == 0");
  if ((lda % 8) || (ldo % 8) || (residual && (ldr % 8))) return set_error(PCDM_ERR_UNSUPPORTED, "gemm: strides must be multiples of 8");
  if (!a2 && bn == 0 && !(flags & PCDM_FLAG_NO_SKINNY) && !ext.ln_stats && !ext.row_stats) {   // M <= 32 activation rows: a weight stream, not a 128-row tile problem (skinny.cu)
    const int taken = skinny_gemm_try(a, lda, w, out, ldo, bias, rowvec, ld_rowvec, rows_per_image, residual, ldr, M, N,
                                      K, dtype, flags, stream, nullptr, nullptr, 0.f);
    if (taken != 0) return taken < 0 ? taken : 0;
  }
  IGemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.num_kb = K / 64;
  p.m_tiles = (M + 127) / 128;
  p.mode = 0;
  p.H = 1; p.W = 1;
  p.kb_split = a2 ? k1 / 64 : p.num_kb;
  const int arows = M < 128 ? M : 128;
  p.a_bytes = (uint32_t)arows * 128u;
  {
    const int ka = a2 ? k1 : K;
    const uint64_t dims[2] = {(uint64_t)ka, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)lda * 2};
    const uint32_t box[2] = {64, (uint32_t)arows};
    PCDM_CHECK(make_tmap(&p.tmA[0], a, 2, dims, strides, box), "A tensor map");
    if (a2) {
      const uint64_t dims2[2] = {(uint64_t)(K - k1), (uint64_t)M};
      const uint64_t strides2[1] = {(uint64_t)lda2 * 2};
      PCDM_CHECK(make_tmap(&p.tmA[1], a2, 2, dims2, strides2, box), "A2 tensor map");
    }
  }
  p.bias = bias; p.rowvec = rowvec; p.ld_rowvec = ld_rowvec; p.hw = rows_per_image > 0 ? rows_per_image : 1;
  p.out = out; p.ldo = ldo;
  p.geglu = geglu; p.out_f32 = (flags & PCDM_FLAG_OUT_F32) ? 1 : 0; p.silu = (flags & PCDM_FLAG_SILU) ? 1 : ((flags & PCDM_FLAG_GELU) ? 2 : 0);
  if (rowvec && (ld_rowvec % 4)) return set_error(PCDM_ERR_UNSUPPORTED, "gemm: rowvec stride must be a multiple of 4");
  if (p.geglu && p.out_f32) return set_error(PCDM_ERR_UNSUPPORTED, "gemm: GEGLU with fp32 output");
  return dispatch_igemm(p, dtype, bn, w, K, residual, ldr, ext, stream);
}

extern "C" int pcdm_conv3x3(const void* x, const void* w_packed, void* out, const float* bias, const float* rowvec,
                            long long ld_rowvec, const void* residual, int B, int H, int W, int Cin, int Cout, int stride, int dtype,
                            int flags, int bn, pcdm_ext* ext_, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ExtArgs ext;
  PCDM_CHECK(read_ext(ext_, &ext), "ext");
  if (!x || !w_packed || !out) return set_error(PCDM_ERR_INVALID, "conv3x3: null pointer");
  if (dtype != DT_F16 && dtype != DT_BF16) return set_error(PCDM_ERR_INVALID, "conv3x3: dtype must be 0 (f16) or 1 (bf16)");
  if (stride != 1 && stride != 2) return set_error(PCDM_ERR_UNSUPPORTED, "conv3x3: stride must be 1 or 2");
  if (B <= 0 || H <= 0 || W <= 0) return set_error(PCDM_ERR_INVALID, "conv3x3: empty problem");
  if (Cin % 64 || Cout % 32) return set_error(PCDM_ERR_UNSUPPORTED, "conv3x3: Cin must be a multiple of 64 and Cout of 32");
  // output tile = 128 consecutive NHWC pixels = tile_b images x tile_h rows x W columns, or (rows wider than a tile,
  // W % 128 == 0: the VAE / pose-encoder resolutions) one 128-pixel segment of a row
  const bool wide = W > 128;
  if (wide ? (W % 128) : (128 % W)) return set_error(PCDM_ERR_UNSUPPORTED, "conv3x3: output width must divide 128 or be a multiple of it");
  const bool pad_br = (flags & PCDM_FLAG_PAD_BR) != 0;
  if (pad_br && stride != 2) return set_error(PCDM_ERR_INVALID, "conv3x3: PCDM_FLAG_PAD_BR goes with stride 2");
  const int hw = H * W;
  int tile_h, tile_b, tile_w = wide ? 128 : W;
  if (wide) {
    tile_h = 1; tile_b = 1;
  } else if (hw >= 128) {
    if (hw % 128) return set_error(PCDM_ERR_UNSUPPORTED, "conv3x3: H*W must be a multiple of 128 (or divide it)");
    tile_h = 128 / W; tile_b = 1;
  } else {
    if (128 % hw) return set_error(PCDM_ERR_UNSUPPORTED, "conv3x3: H*W must divide 128");
    tile_h = H; tile_b = 128 / hw;
  }
  if (tile_b > B) tile_b = B;
  if ((long long)B * hw > 0x7fffffffLL - 256) return set_error(PCDM_ERR_UNSUPPORTED, "conv3x3: B*H*W exceeds 2^31");
  IGemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = B * hw; p.N = Cout; p.num_kb = 9 * (Cin / 64);
  p.m_tiles = (p.M + 127) / 128;
  p.mode = stride == 1 ? 1 : (pad_br ? 3 : 2); p.H = H; p.W = W; p.cblocks = Cin / 64; p.kb_split = p.num_kb;
  p.a_bytes = 128u * (uint32_t)(tile_w * tile_h * tile_b);
  const uint32_t box[4] = {64, (uint32_t)tile_w, (uint32_t)tile_h, (uint32_t)tile_b};
  if (stride == 1) {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)hw * Cin * 2};
    PCDM_CHECK(make_tmap(&p.tmA[0], x, 4, dims, strides, box), "conv input tensor map");
  } else {
    const int Hin = 2 * H, Win = 2 * W;
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)2 * Cin * 2, (uint64_t)2 * Win * Cin * 2, (uint64_t)Hin * Win * Cin * 2};
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        const char* base = reinterpret_cast<const char*>(x) + ((size_t)py * Win + px) * Cin * 2;
        PCDM_CHECK(make_tmap(&p.tmA[py * 2 + px], base, 4, dims, strides, box), "conv parity tensor map");
      }
  }
  p.bias = bias; p.rowvec = rowvec; p.ld_rowvec = ld_rowvec; p.hw = hw;
  p.out = out; p.ldo = Cout;
  p.geglu = 0; p.out_f32 = (flags & PCDM_FLAG_OUT_F32) ? 1 : 0; p.silu = (flags & PCDM_FLAG_SILU) ? 1 : ((flags & PCDM_FLAG_GELU) ? 2 : 0);
  if (rowvec && (ld_rowvec % 4)) return set_error(PCDM_ERR_UNSUPPORTED, "conv3x3: rowvec stride must be a multiple of 4");
  return dispatch_igemm(p, dtype, bn, w_packed, 9 * Cin, residual, Cout, ext, stream);
}

extern "C" int pcdm_conv3x3_up2x(const void* x, const void* w_up, void* out, const float* bias, int B, int H, int W, int Cin,
                                 int Cout, int dtype, int flags, int bn, pcdm_ext* ext_, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ExtArgs ext;
  PCDM_CHECK(read_ext(ext_, &ext), "ext");
  if (!x || !w_up || !out) return set_error(PCDM_ERR_INVALID, "conv3x3_up2x: null pointer");
  if (dtype != DT_F16 && dtype != DT_BF16) return set_error(PCDM_ERR_INVALID, "conv3x3_up2x: dtype must be 0 (f16) or 1 (bf16)");
  if (B <= 0 || H <= 0 || W <= 0) return set_error(PCDM_ERR_INVALID, "conv3x3_up2x: empty problem");
  if (Cin % 64 || Cout % 64) return set_error(PCDM_ERR_UNSUPPORTED, "conv3x3_up2x: Cin and Cout must be multiples of 64");
  if (flags & ~(PCDM_FLAG_SILU | PCDM_FLAG_GELU)) return set_error(PCDM_ERR_UNSUPPORTED, "conv3x3_up2x: only activation flags");
  if (ext.row_stats || ext.ln_stats) return set_error(PCDM_ERR_UNSUPPORTED, "conv3x3_up2x: no LayerNorm extras");
  // tiles are 128 consecutive LOW-resolution pixels (the same geometry as pcdm_conv3x3 at H x W)
  if (W > 128 || (128 % W)) return set_error(PCDM_ERR_UNSUPPORTED, "conv3x3_up2x: input width must divide 128");
  const int hw = H * W;
  int tile_h, tile_b;
  if (hw >= 128) {
    if (hw % 128) return set_error(PCDM_ERR_UNSUPPORTED, "conv3x3_up2x: H*W must be a multiple of 128 (or divide it)");
    tile_h = 128 / W; tile_b = 1;
  } else {
    if (128 % hw) return set_error(PCDM_ERR_UNSUPPORTED, "conv3x3_up2x: H*W must divide 128");
    tile_h = H; tile_b = 128 / hw;
  }
  if (tile_b > B) tile_b = B;
  if ((long long)B * hw > 0x7fffffffLL / 4 - 256) return set_error(PCDM_ERR_UNSUPPORTED, "conv3x3_up2x: B*H*W too large");
  // a warp stores 32 consecutive low-resolution pixels per TMA box: (obox_w x obox_h x images)
  int obw, obh, obb;
  if (W >= 32) { if (W % 32) return set_error(PCDM_ERR_UNSUPPORTED, "conv3x3_up2x: W must be a multiple or divisor of 32"); obw = 32; obh = 1; obb = 1; }
  else if (hw >= 32) { if ((32 % W) || (hw % 32)) return set_error(PCDM_ERR_UNSUPPORTED, "conv3x3_up2x: W | 32 and 32 | H*W"); obw = W; obh = 32 / W; obb = 1; }
  else { if (32 % hw) return set_error(PCDM_ERR_UNSUPPORTED, "conv3x3_up2x: H*W must divide 32"); obw = W; obh = H; obb = 32 / hw; }
  IGemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = B * hw; p.N = Cout; p.num_kb = 4 * (Cin / 64);
  p.m_tiles = (p.M + 127) / 128;
  p.mode = 4; p.H = H; p.W = W; p.cblocks = Cin / 64; p.kb_split = p.num_kb;
  p.obox_w = obw; p.obox_h = obh;
  p.a_bytes = 128u * (uint32_t)(W * tile_h * tile_b);
  {
    const uint32_t box[4] = {64, (uint32_t)W, (uint32_t)tile_h, (uint32_t)tile_b};
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)hw * Cin * 2};
    PCDM_CHECK(make_tmap(&p.tmA[0], x, 4, dims, strides, box), "conv input tensor map");
  }
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      // output pixel (2i + py, 2j + px) of the [B, 2H, 2W, Cout] result, viewed per parity as [Cout, W, H, B]
      char* base = reinterpret_cast<char*>(out) + ((size_t)py * 2 * W + px) * Cout * 2;
      const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)B};
      const uint64_t strides[3] = {(uint64_t)2 * Cout * 2, (uint64_t)2 * 2 * W * Cout * 2, (uint64_t)4 * hw * Cout * 2};
      const uint32_t box[4] = {32, (uint32_t)obw, (uint32_t)obh, (uint32_t)obb};
      PCDM_CHECK(make_tmap(&p.tmOutP[py * 2 + px], base, 4, dims, strides, box, 64), "parity output tensor map");
    }
  p.bias = bias; p.hw = hw;
  p.out = out; p.ldo = Cout;
  p.silu = (flags & PCDM_FLAG_SILU) ? 1 : ((flags & PCDM_FLAG_GELU) ? 2 : 0);
  return dispatch_igemm(p, dtype, bn, w_up, 4 * Cin, nullptr, Cout, ext, stream, 4);
}

extern "C" long long pcdm_gemm_workspace_bytes(int M, int N) {
  if (M <= 0 || N <= 0) return 0;
  return 8LL * M * N * 4;   // at most 8 K-slices of fp32 partials
}

/* LayerNorm over K fused in front of a GEMM: out = act( LN(x; gamma, beta, eps) . W^T + bias + rowvec + residual ).
 * For M <= 32 rows that fit (K <= 2048, M * (K + 8) * 2 <= 100 KB) this is ONE launch of the skinny kernel, which normalises the
 * rows itself; otherwise pcdm_layernorm writes the normalised rows to `scratch` ([M, K] 16-bit, row stride K) and
 * pcdm_gemm follows. */
extern "C" int pcdm_ln_gemm(const void* x, long long ldx, const float* gamma, const float* beta, float eps, void* scratch,
                            const void* w, void* out, long long ldo, const float* bias, const float* rowvec,
                            long long ld_rowvec, int rows_per_image, const void* residual, long long ldr, int M, int N,
                            int K, int dtype, int flags, pcdm_ext* ext_, void* stream_) {
  if (!x || !gamma || !beta || !w || !out) return set_error(PCDM_ERR_INVALID, "ln_gemm: null pointer");
  if (dtype != DT_F16 && dtype != DT_BF16) return set_error(PCDM_ERR_INVALID, "ln_gemm: dtype must be 0 (f16) or 1 (bf16)");
  if (M <= 0 || N <= 0 || K <= 0) return set_error(PCDM_ERR_INVALID, "ln_gemm: empty problem");
  if (K % 64 || N % 32) return set_error(PCDM_ERR_UNSUPPORTED, "ln_gemm: K must be a multiple of 64 and N of 32");
  if ((ldx % 8) || (ldo % 8) || (residual && (ldr % 8))) return set_error(PCDM_ERR_UNSUPPORTED, "ln_gemm: strides must be multiples of 8");
  const int taken = skinny_gemm_try(x, ldx, w, out, ldo, bias, rowvec, ld_rowvec, rows_per_image, residual, ldr, M, N, K,
                                    dtype, flags, (cudaStream_t)stream_, gamma, beta, eps);
  if (taken != 0) return taken < 0 ? taken : 0;
  if (!scratch) return set_error(PCDM_ERR_INVALID, "ln_gemm: this shape needs the scratch buffer");
  const int rc = pcdm_layernorm(x, ldx, scratch, K, gamma, beta, eps, M, K, dtype, stream_);
  if (rc != 0) return rc;
  return pcdm_gemm(scratch, K, nullptr, 0, 0, w, out, ldo, bias, rowvec, ld_rowvec, rows_per_image, residual, ldr, M, N,
                   K, dtype, flags, 0, ext_, stream_);
}
