// K3 — fused multi-head attention, no mask:  O = softmax(Q K^T * scale) V   (flash-style, online softmax), tcgen05.
//
// Two head_dim-64 kernels (plus the head_dim-128 instantiation of the first, and a CUDA-core kernel for <= 32 tokens):
//
//  * attention2_kernel (round 2; sequences of two or more 128-query tiles — every UNet level above 8x16, DINOv2):
//    one PERSISTENT CTA per SM works through (batch, head, q-tile pair) items.  TWO 128-row query tiles share every
//    K/V block TMA brings in and ping-pong through the tensor pipe; a softmax thread owns one full score row and
//    streams it out of TMEM 32 columns at a time (next chunk's tcgen05.ld in flight under the current chunk's
//    exponentials, ~100 live registers), with SPECULATIVE exponentials: probabilities are taken against the row's
//    current reference maximum while the block's true maximum is tracked in the shadow of the MUFUs, and only a growth
//    beyond 2^8 redoes the block from the scores still in TMEM.  One of every four column pairs takes its exp2 on the
//    FMA pipe (Cody-Waite + degree-4 polynomial).  Description at the kernel.
//
//  * attention_kernel (round 1; single-tile sequences, head_dim 128): one CTA per (batch, head, 128-query tile), two
//    CTAs per SM at d = 64.  320 threads: warp 0 = TMA producer (Q once, then a 2-deep K/V ring) + TMEM allocator,
//    warp 1 = tcgen05.mma issuer, warps 2-9 = softmax: two warps per TMEM lane quadrant, each owning one query row per
//    thread and HALF of the 128 score columns (the two half-row maxima / sums meet in shared memory).
//    TMEM (256 columns per CTA):  S : 128 fp32 score columns (QK^T, SS MMA) | P : 64 columns = 128 x 128 16-bit
//    probabilities, consumed by the PV MMA straight from tensor memory (A operand in TMEM) | O : 64 fp32 columns.
//    The running maximum is applied lazily: O and l are rescaled only when a row's maximum grows by more than 2^8.
//
// In both, P never touches shared memory, V is consumed in its natural [kv, d] layout as an MN-major B operand and K as
// a K-major B operand, and a ragged last K/V block costs MMAs / exponentials for the rows that exist only.
//
// Replaces xformers.memory_efficient_attention / F.scaled_dot_product_attention as enabled by the reference at
// stage2_batchtest_inpaint_model.py:133 and used via diffusers' attention processors (the
// processor protocol is mirrored at /root/reference/src/pipelines/PCDMs_pipeline.py:78-153).
#include <type_traits>

#include "common.cuh"
#include "host_util.h"

namespace pcdm {

struct AttnParams {
  CUtensorMap tmQ, tmK, tmV;
  int B, heads, Sq, Skv;
  int q_tiles;
  void* out;
  long long ldo;       // row stride of O in elements
  float scale_log2;    // softmax scale * log2(e)
  unsigned long long* trace;   // experiment build: per-block cycle stamps of CTA 0 (tools/dev_attn_trace.py), else NULL
};

constexpr int ATT_THREADS = 320;   // warp 0: TMA + TMEM alloc, warp 1: MMA, warps 2..9: softmax (2 per TMEM quadrant)
constexpr int ATT_KV_STAGES = 2;
constexpr int ATT_TILE_BYTES = 128 * 128;  // 128 rows x 64 x 2 B
constexpr int ATT_XCH_BYTES = 2 * 2 * 128 * 4 + 2 * 128 * 4;   // row-max exchange [parity][half][row] + row-sum [half][row]
// head_dim HD = 64 (the UNet / DINOv2 / prior transformer) or 128 (CLIP ViT-H: 80 real columns per head, zero-padded by
// the caller's weight packing).  HD = 128 keeps every operand as 64-column 128B-swizzled tiles (two per Q / K / V),
// needs 128 + 64 + 128 TMEM columns (allocation 512) and 160 KB of shared memory: one CTA per SM.
__host__ __device__ constexpr int att_smem_bytes(int HD) { return ATT_TILE_BYTES * (HD / 64) * (1 + 2 * ATT_KV_STAGES) + ATT_XCH_BYTES + 1024 + 256; }
__host__ __device__ constexpr int att_tmem_cols(int HD) { return HD == 64 ? 256 : 512; }
constexpr int ATT_TMEM_S = 0, ATT_TMEM_P = 128, ATT_TMEM_O = 192;

__device__ __forceinline__ void pair_barrier(int q) {   // the two softmax warps that share TMEM quadrant q
  asm volatile("bar.sync %0, 64;" ::"r"(q + 1) : "memory");
}

template <int DT, int ATT_POLY_OF_4, int HD>
__global__ void __launch_bounds__(ATT_THREADS, HD == 64 ? 2 : 1) attention_kernel(const __grid_constant__ AttnParams p) {
  constexpr int NCH = HD / 64;                       // 64-column operand tiles per Q / K / V block
  constexpr int ATT_STAGE_BYTES = 2 * NCH * ATT_TILE_BYTES;
  constexpr int ATT_TMEM_COLS = att_tmem_cols(HD);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + NCH * ATT_TILE_BYTES;  // stage s: K tiles at s*STAGE, V tiles at s*STAGE + NCH*TILE
  float* xmax = reinterpret_cast<float*>(smem + NCH * ATT_TILE_BYTES * (1 + 2 * ATT_KV_STAGES));   // [2][2][128]
  float* xsum = xmax + 2 * 2 * 128;                                                          // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xmax) + ATT_XCH_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                  // [2]
  uint64_t* kv_empty = kv_full + ATT_KV_STAGES;  // [2]
  uint64_t* s_full = kv_empty + ATT_KV_STAGES;   // MMA -> softmax: S_j landed
  uint64_t* s_free = s_full + 1;                 // softmax -> MMA: S_j is in registers (8 warp arrivals)
  uint64_t* p_full = s_free + 1;                 // softmax -> MMA: P_j written, O corrected (8 warp arrivals)
  uint64_t* pv_done = p_full + 1;                // MMA -> softmax: PV_j retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x % p.q_tiles;
  const int bh = blockIdx.x / p.q_tiles;
  const int h = bh % p.heads, b = bh / p.heads;
  const int n_kv = (p.Skv + 127) / 128;

  if (warp == 1 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < ATT_KV_STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(s_free, 8);
    mbar_init(p_full, 8);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, ATT_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();

  // The TMA and MMA warps stay converged and elect one lane at each issue: operands then sit in uniform registers
  // (a `lane == 0` branch costs ~80 cycles of vector->uniform moves per tcgen05.mma, measured in igemm.cu).
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_expect_tx(q_full, NCH * ATT_TILE_BYTES);
#pragma unroll
      for (int c = 0; c < NCH; ++c) tma_load_4d(sQ + c * ATT_TILE_BYTES, &p.tmQ, q_full, c * 64, qt * 128, h, b);
    }
    __syncwarp();
    for (int j = 0; j < n_kv; ++j) {
      const int stage = j & 1;
      mbar_wait(&kv_empty[stage], (uint32_t)(((j >> 1) & 1) ^ 1));
      uint8_t* sk = sKV + stage * ATT_STAGE_BYTES;
      if (elect_one()) {
        mbar_expect_tx(&kv_full[stage], ATT_STAGE_BYTES);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          tma_load_4d(sk + c * ATT_TILE_BYTES, &p.tmK, &kv_full[stage], c * 64, j * 128, h, b);
          tma_load_4d(sk + (NCH + c) * ATT_TILE_BYTES, &p.tmV, &kv_full[stage], c * 64, j * 128, h, b);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // S = Q K^T : both operands K-major, N = the block's kv rows rounded up to 16 (a ragged last block — 2 rows of the
    // 258 cross-attention tokens — costs a 16-wide MMA, not a 128-wide one); O += P V : A from TMEM, B (V) MN-major,
    // one K=16 MMA per 16 kv rows that exist
    constexpr uint32_t idesc_qk0 = make_idesc(DT, 128, 0, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc(DT, 128, 64, 0, 1);
    auto n16_of = [&](int j) { return (min(128, p.Skv - j * 128) + 15) >> 4; };
    // shared-memory descriptors: constant high word (SBO 1024 B, version 1, 128-byte swizzle), low word = addr >> 4
    // | LBO >> 4 << 16, advanced by plain adds
    constexpr uint64_t kDescHi = (uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;
    const uint32_t q_lo = ((smem_u32(sQ) >> 4) & 0x3FFFu) | ((16u >> 4) << 16);
    const uint32_t kv_lo0 = (smem_u32(sKV) >> 4) & 0x3FFFu;
    mbar_wait(q_full, 0);
    auto issue_qk = [&](int j) {
      const int stage = j & 1;
      mbar_wait(&kv_full[stage], (uint32_t)((j >> 1) & 1));
      if (j > 0) mbar_wait(s_free, (uint32_t)((j - 1) & 1));   // softmax holds S_{j-1} in registers
      tc_fence_after();
      if (elect_one()) {
        const uint32_t k_lo = (kv_lo0 + (uint32_t)stage * (ATT_STAGE_BYTES >> 4)) | ((16u >> 4) << 16);
        const uint32_t idesc_qk = idesc_qk0 | ((uint32_t)(n16_of(j) * 2) << 17);   // N >> 3 at bits [17, 23)
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ss(tmem + ATT_TMEM_S, kDescHi | (q_lo + (uint32_t)c * (ATT_TILE_BYTES >> 4) + 2u * k),
                    kDescHi | (k_lo + (uint32_t)c * (ATT_TILE_BYTES >> 4) + 2u * k), idesc_qk, (c | k) != 0);
        tc_commit(s_full);
      }
      __syncwarp();
    };
    issue_qk(0);
    for (int j = 0; j < n_kv; ++j) {
      if (j + 1 < n_kv) issue_qk(j + 1);
      mbar_wait(p_full, (uint32_t)(j & 1));
      tc_fence_after();
      const int stage = j & 1;
      if (elect_one()) {
        const uint32_t v_lo = (kv_lo0 + (uint32_t)(stage * (ATT_STAGE_BYTES >> 4) + NCH * (ATT_TILE_BYTES >> 4))) |
                              ((1024u >> 4) << 16);
        const int n16 = n16_of(j);
#pragma unroll
        for (int c = 0; c < NCH; ++c)   // one N = 64 MMA chain per 64-column tile of V -> output columns [64c, 64c+64)
#pragma unroll
          for (int k = 0; k < 8; ++k)  // K = 16 kv rows per MMA: 8 packed P columns, 16 V rows (2048 B)
            if (k < n16)
              umma_ts(tmem + ATT_TMEM_O + c * 64, tmem + ATT_TMEM_P + k * 8,
                      kDescHi | (v_lo + (uint32_t)c * (ATT_TILE_BYTES >> 4) + (2048u >> 4) * k), idesc_pv, (j | k) != 0);
        tc_commit(&kv_empty[stage]);
        tc_commit(pv_done);
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax / correction / epilogue =====================
    // warps 2..9: TMEM quadrant q = warp & 3 (rows 32q..32q+31); `half` selects which 64 of the 128 score columns
    // (and which HD/2 of the HD output columns) this warp owns.  The two warps of a quadrant exchange row maxima and,
    // at the end, row sums through shared memory.
    using T = typename TypeOf<DT>::T;
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    float m_ref = -INFINITY;  // reference maximum (log2 domain) the accumulators are relative to
    float l = 0.f;            // this warp's share of the row sum
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(s_full, (uint32_t)(j & 1));
      tc_fence_after();
      const int kv_left = p.Skv - j * 128 - half * 64;  // valid columns among this warp's 64
      uint32_t s[64];
      {
        uint32_t (&s0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[0]);
        uint32_t (&s1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[32]);
        tmem_ld32(tmem + lane_off + ATT_TMEM_S + half * 64, s0);
        tmem_ld32(tmem + lane_off + ATT_TMEM_S + half * 64 + 32, s1);
        tc_wait_ld();
      }
      // S_j now lives in registers: the MMA warp may overwrite the score columns with Q K_{j+1}^T
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      float mx = -INFINITY;
      if (kv_left >= 64) {
#pragma unroll
        for (int i = 0; i < 64; i += 2) mx = fmax3(mx, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
      } else {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i < kv_left) mx = fmaxf(mx, __uint_as_float(s[i]));
      }
      float* xm = xmax + (j & 1) * 256;
      xm[half * 128 + row] = mx;
      pair_barrier(q);
      mx = fmaxf(mx, xm[(half ^ 1) * 128 + row]) * p.scale_log2;
      // Reference maximum: decided now; the rescale of O and the P stores wait for PV_{j-1} AFTER this block's
      // exponentials (round 2: the exponentials overlap that MMA instead of queueing behind it).  The packed
      // probabilities overwrite the scores in place, so nothing extra stays live across the wait.
      float alpha = 1.0f;
      bool rescale = false;
      if (j == 0) {
        m_ref = mx;
      } else {
        const bool grow = mx > m_ref + 8.0f;
        rescale = __any_sync(0xffffffffu, grow);   // both warps of the pair see the same rows => the same decision
        if (rescale) {
          const float m_new = grow ? mx : m_ref;
          alpha = exp2f(m_ref - m_new);
          l *= alpha;
          m_ref = m_new;
        }
      }
      float2 sum2 = make_float2(0.f, 0.f);
      const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(-m_ref, -m_ref);
      auto exp_pair = [&](int c, int i) {
        const float2 x = __ffma2_rn(make_float2(__uint_as_float(s[c * 32 + 2 * i]), __uint_as_float(s[c * 32 + 2 * i + 1])),
                                    sc2, nm2);
        float2 e;
        if ((i & 3) < ATT_POLY_OF_4) {
          e = exp2_poly2(x);
        } else {
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(x.x));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(x.y));
        }
        return e;
      };
      int nst = 2;   // P chunks (16 packed columns each) this warp stores
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (kv_left >= 64) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 e = exp_pair(c, i);
            sum2 = __fadd2_rn(sum2, e);
            s[c * 32 + i] = pack2<DT>(e.x, e.y);
          }
        } else {
          // ragged last block: columns past the sequence contribute nothing and cost nothing (warp-uniform skips);
          // the PV MMAs read only the P columns of 16-row groups that exist, so chunks beyond them are not written
          if (c * 32 >= ((kv_left + 15) & ~15)) { nst = min(nst, c); continue; }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (c * 32 + 2 * i < kv_left) {
              float2 e = exp_pair(c, i);
              if (c * 32 + 2 * i + 1 >= kv_left) e.y = 0.f;
              sum2 = __fadd2_rn(sum2, e);
              s[c * 32 + i] = pack2<DT>(e.x, e.y);
            } else {
              s[c * 32 + i] = 0u;
            }
          }
        }
      }
      // P (and O, if a rescale is needed) may only be touched once PV_{j-1} has retired
      if (j > 0) {
        mbar_wait(pv_done, (uint32_t)((j - 1) & 1));
        tc_fence_after();
      }
      if (rescale) {
#pragma unroll 1
        for (int c = 0; c < HD / 32; ++c) {
          uint32_t o[16];
          tmem_ld16(tmem + lane_off + ATT_TMEM_O + half * (HD / 2) + c * 16, o);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st16(tmem + lane_off + ATT_TMEM_O + half * (HD / 2) + c * 16, o);
        }
      }
#pragma unroll
      for (int c = 0; c < 2; ++c)
        if (c < nst) tmem_st16(tmem + lane_off + ATT_TMEM_P + half * 32 + c * 16, *reinterpret_cast<uint32_t (*)[16]>(&s[c * 32]));
      const float sum = sum2.x + sum2.y;
      l += sum;
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // epilogue: O / l  (row sum = both halves)
    xsum[half * 128 + row] = l;
    pair_barrier(q);
    const float inv_l = 1.0f / (l + xsum[(half ^ 1) * 128 + row]);
    mbar_wait(pv_done, (uint32_t)((n_kv - 1) & 1));
    tc_fence_after();
    const long long qrow = (long long)qt * 128 + row;
    T* op = reinterpret_cast<T*>(p.out) + ((long long)b * p.Sq + qrow) * p.ldo + h * HD + half * (HD / 2);
#pragma unroll 1
    for (int c = 0; c < HD / 32; ++c) {
      uint32_t o[16];
      tmem_ld16(tmem + lane_off + ATT_TMEM_O + half * (HD / 2) + c * 16, o);
      tc_wait_ld();
      if (qrow < p.Sq) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          uint4 u;
          u.x = pack2<DT>(__uint_as_float(o[i * 8 + 0]) * inv_l, __uint_as_float(o[i * 8 + 1]) * inv_l);
          u.y = pack2<DT>(__uint_as_float(o[i * 8 + 2]) * inv_l, __uint_as_float(o[i * 8 + 3]) * inv_l);
          u.z = pack2<DT>(__uint_as_float(o[i * 8 + 4]) * inv_l, __uint_as_float(o[i * 8 + 5]) * inv_l);
          u.w = pack2<DT>(__uint_as_float(o[i * 8 + 6]) * inv_l, __uint_as_float(o[i * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(op + c * 16 + i * 8) = u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, ATT_TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Round 2: head_dim 64 attention restructured (VERDICT r1 "next" item 3).  One PERSISTENT CTA per SM works through
// (batch, head, q-tile pair) items; inside an item TWO 128-row query tiles ping-pong through the tensor pipe and share
// every K/V block that TMA brings in (half the K/V traffic per query row).  384 threads: warp 0 = TMA producer (Q
// double-buffered across items, 3-deep K/V ring running ahead across item boundaries), warp 1 = tcgen05.mma issuer,
// warps 2-3 idle (a warp may only touch TMEM lane quadrant warp % 4, so the softmax groups start at a multiple of 4),
// warps 4-7 = softmax of tile 0, warps 8-11 = softmax of tile 1.  A softmax thread owns ONE FULL score row: no half-row
// exchange through shared memory, no pair barrier, the rescale of O is a per-warp decision.
//
// The scores stay in TMEM while the row streams through them in 32-column chunks (registers: two chunk buffers + 16
// packed probabilities — the first version of this kernel held all 128 scores of a row in registers, could not
// overlap anything inside a warp and was SLOWER than the round-1 kernel, profiles/r2_attention.md).  Per tile the chain
// is  S(j) -> softmax(j) -> { Q K(j+1)^T, P(j) V(j) }; the bubble of one tile is filled by the other tile's
// exponentials (the static issue order QK0, PV0, QK1, PV1 makes the tiles settle half a period apart).
//
// TMEM (512 columns): S0 [0,128) S1 [128,256) fp32 scores | O0 [256,320) O1 [320,384) fp32 accumulators |
//                     P0 [384,448) P1 [448,512) 16-bit probabilities (A operand of the PV MMA, never in shared memory)
// Barriers (phases by use counters; "^1" waits pass on a fresh barrier):
//   q_full[2]/q_empty[2] per Q buffer, kv_full[3]/kv_empty[3] per ring stage, and per tile t:
//   s_full[t] (MMA->softmax: S landed), p_full[t] (softmax->MMA: block done — P written, O rescaled, S free),
//   pv_done[t] (MMA->softmax: PV retired, P / O may be touched), o_free[t] (softmax->MMA: O read out).
// Tail balance: the items of the last partial round are issued as SINGLE tiles (a lone tile gets the SM's whole SFU).
// Measured (B200, bf16, B 16, inside a CUDA graph; round-1 kernel -> this one): 2048 x 2048 142 -> 124 us, 2048 x 258
// 40.2 -> 30.6, 512 x 512 28.6 -> 23.9, 512 x 258 23.0 -> 18.5, 8192 x 8192 (B 8) 971 -> 910.
// ---------------------------------------------------------------------------------------------------------------
constexpr int A2_THREADS = 384;
constexpr int A2_THREADS_HALF = 576;   // half-row variant: warps 2-17 = softmax, two warps per (tile, TMEM lane quadrant)
constexpr int A2_KV_STAGES = 3;
constexpr int A2_SMEM = 2 * 2 * ATT_TILE_BYTES + A2_KV_STAGES * 2 * ATT_TILE_BYTES + 1024 + 512 + 2048;   // + half-row exchange [tile][half][128]
constexpr int A2_S = 0, A2_O = 256, A2_P = 384;   // TMEM column bases (tile t: + t * 128 / 64 / 64)

struct A2Item { int b, h, qt, nt; };   // nt = 1 or 2 query tiles (qt, qt + 1) of head (b, h); nt = 0: nothing

// item k of this CTA -> its tiles.  `pairs` = ceil(q_tiles / 2) pair slots per (b, h); items [0, full) are pairs, the
// remaining pair slots are issued as two single-tile items each.
// (item boundaries are on every role's critical path: the two divisions go through the float reciprocal + one exact
// correction step instead of the ~40-instruction integer division sequence)
__device__ __forceinline__ int a2_div(int x, int d, int* rem) {
  int qv = __float2int_rz(__fdividef((float)x + 0.5f, (float)d));
  int r = x - qv * d;
  if (r < 0) { --qv; r += d; }
  else if (r >= d) { ++qv; r -= d; }
  *rem = r;
  return qv;
}
__device__ __forceinline__ A2Item a2_item(const AttnParams& p, int item, int pairs, int full, int total_items) {
  A2Item it;
  it.nt = 0; it.b = it.h = it.qt = 0;
  if (item >= total_items) return it;
  int slot, want;   // slot = tile slot index (2 per pair)
  if (item < full) { slot = 2 * item; want = 2; }
  else { slot = 2 * full + (item - full); want = 1; }
  int qt, h;
  const int bh = a2_div(slot, 2 * pairs, &qt);
  it.b = a2_div(bh, p.heads, &h);
  it.h = h; it.qt = qt;
  it.nt = qt >= p.q_tiles ? 0 : ((want == 2 && qt + 1 < p.q_tiles) ? 2 : 1);
  return it;
}

// One 32-column chunk of a score row -> 32 probabilities against the reference maximum `mref` (log2 domain), 16 packed
// 16-bit pairs in pk.  PN of every PM column pairs take their exp2 on the FMA pipe (exp2_poly2), the rest on the SFU.
// TRACK: the chunk's raw maximum is folded into mx on the side (FMNMX3 in the shadow of the MUFUs).
template <int DT, int PN, int PM, bool TRACK>
__device__ __forceinline__ void a2_chunk(const uint32_t (&s)[32], float2 sc2, float mref, float& mx, float2& sum2,
                                         uint32_t (&pk)[16]) {
  const float2 nm2 = make_float2(-mref, -mref);
  float mxa = mx, mxb = -INFINITY;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float s0 = __uint_as_float(s[2 * i]), s1 = __uint_as_float(s[2 * i + 1]);
    const float2 x = __ffma2_rn(make_float2(s0, s1), sc2, nm2);
    float2 e;
    if ((i % PM) < PN) {
      e = exp2_poly2(x);
    } else {
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(x.x));
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(x.y));
    }
    if (TRACK) {
      if (i & 1) mxb = fmax3(mxb, s0, s1);
      else mxa = fmax3(mxa, s0, s1);
    }
    sum2 = __fadd2_rn(sum2, e);
    pk[i] = pack2<DT>(e.x, e.y);
  }
  if (TRACK) mx = fmaxf(mxa, mxb);
}

// Half-row variant: one 16-column piece of a score row -> 16 probabilities (8 packed pairs); PN of every PM pairs on the
// FMA pipe.  TRACK: the piece's raw maximum is folded into mx on the side.
template <int DT, int PN, int PM, bool TRACK>
__device__ __forceinline__ void a2_piece(const uint32_t (&s)[16], float2 sc2, float mref, float& mx, float2& sum2,
                                         uint32_t (&pk)[8]) {
  const float2 nm2 = make_float2(-mref, -mref);
  float mxa = mx, mxb = -INFINITY;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float s0 = __uint_as_float(s[2 * i]), s1 = __uint_as_float(s[2 * i + 1]);
    const float2 x = __ffma2_rn(make_float2(s0, s1), sc2, nm2);
    float2 e;
    if ((i % PM) < PN) {
      e = exp2_poly2(x);
    } else {
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(x.x));
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(x.y));
    }
    if (TRACK) {
      if (i & 1) mxb = fmax3(mxb, s0, s1);
      else mxa = fmax3(mxa, s0, s1);
    }
    sum2 = __fadd2_rn(sum2, e);
    pk[i] = pack2<DT>(e.x, e.y);
  }
  if (TRACK) mx = fmaxf(mxa, mxb);
}

// barrier over the 64 threads of a half-row warp pair that also ORs a predicate across them
__device__ __forceinline__ bool pair_barrier_or(int id, bool pred) {
  uint32_t r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %2, 0;\n\t"
      "barrier.red.or.pred q, %1, 64, p;\n\t"
      "selp.u32 %0, 1, 0, q;\n\t}"
      : "=r"(r) : "r"(id), "r"((uint32_t)pred) : "memory");
  return r != 0;
}
__device__ __forceinline__ void pair_barrier_id(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// TRACE (experiment build): CTA 0 stamps clock64() at the hand-shakes of its first A2_TRACE_N blocks per tile —
//   softmax warp of quadrant 0: [0] S landed, [1] chunk 0 in registers, [2] chunk 0 exponentiated, [3] PV(n - 1) seen
//   retired, [4] / [5] / [6] chunks 1 / 2 / 3 done, [7] P stores complete, p_full arrive;
//   MMA warp: [8] about to wait for block n - 1, [9] wait passed, [10] QK(n) issued, [11] PV(n - 1) issued.
constexpr int A2_TRACE_N = 96;
#define A2_STAMP(t_, n_, k_)                                                                                   \
  do {                                                                                                         \
    if (TRACE && p.trace && blockIdx.x == 0 && (n_) < A2_TRACE_N && lane == 0)                                 \
      p.trace[((t_) * A2_TRACE_N + (n_)) * 16 + (k_)] = (unsigned long long)clock64();                          \
  } while (0)

template <int DT, int PN, int PM, bool EARLY, bool HALF, bool TRACE = false>
__global__ void __launch_bounds__(HALF ? A2_THREADS_HALF : A2_THREADS, 1) attention2_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // [buf][tile][16 KB]
  uint8_t* sKV = smem + 4 * ATT_TILE_BYTES;             // [stage][K 16 KB | V 16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + A2_KV_STAGES * 2 * ATT_TILE_BYTES);
  uint64_t* q_full = bars;              // [2]
  uint64_t* q_empty = bars + 2;         // [2]
  uint64_t* kv_full = bars + 4;         // [3]
  uint64_t* kv_empty = bars + 7;        // [3]
  uint64_t* s_full = bars + 10;         // [2]
  uint64_t* p_full = bars + 12;         // [2]  softmax -> MMA: block done (P written, O rescaled, S no longer needed)
  uint64_t* pv_done = bars + 14;        // [2]
  uint64_t* o_free = bars + 16;         // [2]
  uint64_t* s_free = bars + 18;         // [2]  EARLY: softmax -> MMA: the block's scores are in registers / checked
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  float* xch = reinterpret_cast<float*>(bars + 24);   // HALF: [tile][half][128 rows] maxima / sums exchanged by a warp pair

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kv = (p.Skv + 127) / 128;
  const int pairs = (p.q_tiles + 1) / 2;
  const int pair_slots = p.B * p.heads * pairs;
  const int G = gridDim.x;
  const int full = (pair_slots / G) * G;                         // pair slots issued as pairs
  const int total_items = full + 2 * (pair_slots - full);        // the rest as single tiles

  if (warp == 1 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
    for (int i = 0; i < A2_KV_STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&p_full[i], HALF ? 8 : 4); mbar_init(&s_free[i], 4);
      mbar_init(&pv_done[i], 1); mbar_init(&o_free[i], HALF ? 8 : 4);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    int item_it = 0, kv_it = 0;   // item_it counts the items actually processed (every role skips the same empty ones)
    for (int item = blockIdx.x; item < total_items; item += G) {
      const A2Item it = a2_item(p, item, pairs, full, total_items);
      if (it.nt == 0) continue;
      const int qb = item_it & 1;
      mbar_wait(&q_empty[qb], (uint32_t)(((item_it >> 1) & 1) ^ 1));
      if (elect_one()) {
        mbar_expect_tx(&q_full[qb], (uint32_t)it.nt * ATT_TILE_BYTES);
        for (int t = 0; t < it.nt; ++t)
          tma_load_4d(sQ + (qb * 2 + t) * ATT_TILE_BYTES, &p.tmQ, &q_full[qb], 0, (it.qt + t) * 128, it.h, it.b);
      }
      __syncwarp();
      for (int j = 0; j < n_kv; ++j, ++kv_it) {
        const int stage = kv_it % A2_KV_STAGES;
        mbar_wait(&kv_empty[stage], (uint32_t)(((kv_it / A2_KV_STAGES) & 1) ^ 1));
        if (elect_one()) {
          uint8_t* sk = sKV + stage * 2 * ATT_TILE_BYTES;
          mbar_expect_tx(&kv_full[stage], 2 * ATT_TILE_BYTES);
          tma_load_4d(sk, &p.tmK, &kv_full[stage], 0, j * 128, it.h, it.b);
          tma_load_4d(sk + ATT_TILE_BYTES, &p.tmV, &kv_full[stage], 0, j * 128, it.h, it.b);
        }
        __syncwarp();
      }
      ++item_it;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Per tile the chain is  S(j) -> softmax(j) -> { Q K(j+1)^T, P(j) V(j) }: the scores stay in TMEM while the softmax
    // warps stream through them, so S_t is rewritten only once block j is done — that bubble of one tile (one QK^T
    // plus hand-shakes) is filled by the exponentials of the OTHER tile, whose blocks end half a period later.
    constexpr uint32_t idesc_qk0 = make_idesc(DT, 128, 0, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc(DT, 128, 64, 0, 1);
    constexpr uint64_t kDescHi = (uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;
    const uint32_t q_lo0 = ((smem_u32(sQ) >> 4) & 0x3FFFu) | ((16u >> 4) << 16);
    const uint32_t kv_lo0 = (smem_u32(sKV) >> 4) & 0x3FFFu;
    int item_it = 0, kv_it = 0;
    int blk[2] = {0, 0}, items_done[2] = {0, 0};
    auto n16_of = [&](int j) { return (min(128, p.Skv - j * 128) + 15) >> 4; };
    auto issue_qk = [&](int t, int j, int stage, int qb, bool last) {   // S_t = Q_t K(j)^T  (S_t is free: block n - 1 done)
      if (elect_one()) {
        const uint32_t q_lo = q_lo0 + (uint32_t)(qb * 2 + t) * (ATT_TILE_BYTES >> 4);
        const uint32_t k_lo = (kv_lo0 + (uint32_t)stage * (2 * ATT_TILE_BYTES >> 4)) | ((16u >> 4) << 16);
        const uint32_t idesc_qk = idesc_qk0 | ((uint32_t)(n16_of(j) * 2) << 17);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss(tmem + A2_S + t * 128, kDescHi | (q_lo + 2u * k), kDescHi | (k_lo + 2u * k), idesc_qk, k != 0);
        tc_commit(&s_full[t]);
        if (last) tc_commit(&q_empty[qb]);   // the item's last read of its Q tiles
      }
      __syncwarp();
    };
    auto issue_pv = [&](int t, int j, int stage) {   // O_t (+)= P_t(j) V(j)
      if (j == 0) {
        mbar_wait(&o_free[t], (uint32_t)((items_done[t] & 1) ^ 1));   // O of the previous item was read out
        tc_fence_after();
      }
      if (elect_one()) {
        const uint32_t v_lo = (kv_lo0 + (uint32_t)(stage * (2 * ATT_TILE_BYTES >> 4) + (ATT_TILE_BYTES >> 4))) |
                              ((1024u >> 4) << 16);
        const int n16 = n16_of(j);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < n16)
            umma_ts(tmem + A2_O + t * 64, tmem + A2_P + t * 64 + k * 8, kDescHi | (v_lo + (2048u >> 4) * k), idesc_pv,
                    (j | k) != 0);
        tc_commit(&pv_done[t]);
      }
      __syncwarp();
    };
    for (int item = blockIdx.x; item < total_items; item += G) {
      const A2Item it = a2_item(p, item, pairs, full, total_items);
      if (it.nt == 0) continue;
      const int qb = item_it & 1;
      mbar_wait(&q_full[qb], (uint32_t)((item_it >> 1) & 1));
      for (int j = 0; j <= n_kv; ++j) {
        const int stage = (kv_it + j) % A2_KV_STAGES, pstage = (kv_it + j - 1) % A2_KV_STAGES;
        if (j < n_kv) mbar_wait(&kv_full[stage], (uint32_t)(((kv_it + j) / A2_KV_STAGES) & 1));
        for (int t = 0; t < it.nt; ++t) {
          const int n = blk[t] + j;                   // global block counter of tile t
          if (EARLY) {
            if (j < n_kv) {
              if (n > 0) mbar_wait(&s_free[t], (uint32_t)((n - 1) & 1));   // block n - 1's scores are out of TMEM (also across items)
              tc_fence_after();
              issue_qk(t, j, stage, qb, j == n_kv - 1 && t == it.nt - 1);
            }
            if (j > 0) {
              mbar_wait(&p_full[t], (uint32_t)((n - 1) & 1));            // softmax finished block n - 1
              tc_fence_after();
              issue_pv(t, j - 1, pstage);
            }
          } else {
            A2_STAMP(t, n, 8);
            if (n > 0) mbar_wait(&p_full[t], (uint32_t)((n - 1) & 1));   // softmax finished block n - 1 (also across items)
            tc_fence_after();
            A2_STAMP(t, n, 9);
            if (j < n_kv) issue_qk(t, j, stage, qb, j == n_kv - 1 && t == it.nt - 1);
            A2_STAMP(t, n, 10);
            if (j > 0) issue_pv(t, j - 1, pstage);
            A2_STAMP(t, n, 11);
          }
        }
        if (j > 0) {
          if (elect_one()) tc_commit(&kv_empty[pstage]);
          __syncwarp();
        }
      }
      kv_it += n_kv;
      ++item_it;
      for (int t = 0; t < it.nt; ++t) { blk[t] += n_kv; ++items_done[t]; }
    }
  } else if (HALF) {
    // ===================== HALF-ROW softmax: warps 2-17; (tile, column half, TMEM quadrant) = (idx >> 3, (idx >> 2) & 1,
    // warp & 3) with idx = warp - 2.  Four softmax warps per scheduler instead of two: the full-row variant leaves the
    // issue slots half empty (each warp stalls on its own MUFU / TMEM latencies and there is only one other warp to
    // switch to).  A thread owns 64 columns of a row, streamed in 16-column pieces; the two halves of a row agree on
    // the reference maximum through shared memory — at the first block of an item, and when one of them sees its
    // maximum grow past 2^8 (a barrier.red.or per block tells both).
    using T = typename TypeOf<DT>::T;
    const int idx = warp - 2;
    const int t = idx >> 3, half = (idx >> 2) & 1, q = warp & 3;
    const int row = q * 32 + lane;
    const int bar_id = 1 + t * 4 + q;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t tS = tmem + lane_off + A2_S + t * 128 + half * 64, tO = tmem + lane_off + A2_O + t * 64 + half * 32,
                   tP = tmem + lane_off + A2_P + t * 64 + half * 32;
    float* xm = xch + (t * 2 + half) * 128, *xo = xch + (t * 2 + (half ^ 1)) * 128;
    const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
    int n = 0;
    auto rescale_o = [&](float alpha) {
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t o[16];
        tmem_ld16(tO + c * 16, o);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st16(tO + c * 16, o);
      }
    };
    for (int item = blockIdx.x; item < total_items; item += G) {
      const A2Item it = a2_item(p, item, pairs, full, total_items);
      if (t >= it.nt) continue;
      float m_ref = -INFINITY, l = 0.f;
      for (int j = 0; j < n_kv; ++j, ++n) {
        mbar_wait(&s_full[t], (uint32_t)(n & 1));
        tc_fence_after();
        const int kv_left = p.Skv - j * 128 - half * 64;   // valid columns among this thread's 64 (may be <= 0)
        uint32_t sa[16], sb[16], pk[8];
        if (p.Skv - j * 128 >= 128) {
          tmem_ld16(tS, sa);
          tc_wait_ld();
          tmem_ld16(tS + 16, sb);
          if (j == 0) {   // first block of the item: reference maximum = max over both halves' first 16 scores
            float a = -INFINITY, b = -INFINITY;
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              a = fmax3(a, __uint_as_float(sa[i]), __uint_as_float(sa[i + 1]));
              b = fmax3(b, __uint_as_float(sa[i + 2]), __uint_as_float(sa[i + 3]));
            }
            const float mine = fmaxf(a, b);
            xm[row] = mine;
            pair_barrier_id(bar_id);
            m_ref = fmaxf(mine, xo[row]) * p.scale_log2;
            pair_barrier_id(bar_id);   // the slot is reused by the growth exchange below
          }
          float mx = -INFINITY;
          float2 sum2 = make_float2(0.f, 0.f);
          a2_piece<DT, PN, PM, true>(sa, sc2, m_ref, mx, sum2, pk);
          mbar_wait(&pv_done[t], (uint32_t)((n & 1) ^ 1));   // P_t (and O_t) may only be touched once PV_t(n - 1) has retired
          tc_fence_after();
          tmem_st8(tP, pk);
          tc_wait_ld();
          tmem_ld16(tS + 32, sa);
          a2_piece<DT, PN, PM, true>(sb, sc2, m_ref, mx, sum2, pk);
          tmem_st8(tP + 8, pk);
          tc_wait_ld();
          tmem_ld16(tS + 48, sb);
          a2_piece<DT, PN, PM, true>(sa, sc2, m_ref, mx, sum2, pk);
          tmem_st8(tP + 16, pk);
          tc_wait_ld();
          a2_piece<DT, PN, PM, true>(sb, sc2, m_ref, mx, sum2, pk);
          tmem_st8(tP + 24, pk);
          mx *= p.scale_log2;
          if (pair_barrier_or(bar_id, mx > m_ref + 8.0f)) {   // rare: some row of this quadrant grew (either half)
            xm[row] = mx;
            pair_barrier_id(bar_id);
            const float mrow = fmaxf(mx, xo[row]);
            pair_barrier_id(bar_id);
            const float m_new = mrow > m_ref + 8.0f ? mrow : m_ref;
            const float alpha = exp2f(m_ref - m_new);
            sum2 = make_float2(0.f, 0.f);
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
              tmem_ld16(tS + c * 16, sa);
              tc_wait_ld();
              a2_piece<DT, 0, 1, false>(sa, sc2, m_new, mx, sum2, pk);
              tmem_st8(tP + c * 8, pk);
            }
            if (j > 0) rescale_o(alpha);
            l *= alpha;
            m_ref = m_new;
          }
          l += sum2.x + sum2.y;
        } else {
          // ---- ragged last block: classic order; this half may own no valid column at all
          const int nv = kv_left < 0 ? 0 : (kv_left > 64 ? 64 : kv_left);
          const int ncols = ((p.Skv - j * 128 + 15) & ~15) - half * 64;          // P columns the PV MMAs read, of this half
          const int npc = ncols <= 0 ? 0 : (ncols > 64 ? 4 : (ncols + 15) >> 4);  // 16-column pieces to write
          float mx = -INFINITY;
#pragma unroll 1
          for (int c = 0; c * 16 < nv; ++c) {
            tmem_ld16(tS + c * 16, sa);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (c * 16 + i < nv) mx = fmaxf(mx, __uint_as_float(sa[i]));
          }
          xm[row] = mx;
          pair_barrier_id(bar_id);
          mx = fmaxf(mx, xo[row]) * p.scale_log2;
          pair_barrier_id(bar_id);
          float alpha = 1.0f;
          bool rescale = false;
          if (j == 0) {
            m_ref = mx;
          } else {
            const bool grow = mx > m_ref + 8.0f;          // identical in both halves (same mx, same m_ref)
            rescale = __any_sync(0xffffffffu, grow);
            if (rescale) {
              const float m_new = grow ? mx : m_ref;
              alpha = exp2f(m_ref - m_new);
              l *= alpha;
              m_ref = m_new;
            }
          }
          mbar_wait(&pv_done[t], (uint32_t)((n & 1) ^ 1));
          tc_fence_after();
          if (rescale) rescale_o(alpha);
          float2 sum2 = make_float2(0.f, 0.f);
          const float2 nm2 = make_float2(-m_ref, -m_ref);
#pragma unroll 1
          for (int c = 0; c < npc; ++c) {
            if (c * 16 < nv) {
              tmem_ld16(tS + c * 16, sa);
              tc_wait_ld();
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int col = c * 16 + 2 * i;
              float2 e = make_float2(0.f, 0.f);
              if (col < nv) {
                const float2 x = __ffma2_rn(make_float2(__uint_as_float(sa[2 * i]), __uint_as_float(sa[2 * i + 1])), sc2, nm2);
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(x.x));
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(x.y));
                if (col + 1 >= nv) e.y = 0.f;
              }
              sum2 = __fadd2_rn(sum2, e);
              pk[i] = pack2<DT>(e.x, e.y);
            }
            tmem_st8(tP + c * 8, pk);
          }
          l += sum2.x + sum2.y;
        }
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
      }
      // epilogue of the item: O / l, this half's 32 output columns; the row sum is the two halves'
      xm[row] = l;
      pair_barrier_id(bar_id);
      const float inv_l = 1.0f / (l + xo[row]);
      mbar_wait(&pv_done[t], (uint32_t)((n - 1) & 1));
      tc_fence_after();
      const long long qrow = (long long)(it.qt + t) * 128 + row;
      T* op = reinterpret_cast<T*>(p.out) + ((long long)it.b * p.Sq + qrow) * p.ldo + it.h * 64 + half * 32;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t o[16];
        tmem_ld16(tO + c * 16, o);
        tc_wait_ld();
        if (qrow < p.Sq) {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            uint4 u;
            u.x = pack2<DT>(__uint_as_float(o[i * 8 + 0]) * inv_l, __uint_as_float(o[i * 8 + 1]) * inv_l);
            u.y = pack2<DT>(__uint_as_float(o[i * 8 + 2]) * inv_l, __uint_as_float(o[i * 8 + 3]) * inv_l);
            u.z = pack2<DT>(__uint_as_float(o[i * 8 + 4]) * inv_l, __uint_as_float(o[i * 8 + 5]) * inv_l);
            u.w = pack2<DT>(__uint_as_float(o[i * 8 + 6]) * inv_l, __uint_as_float(o[i * 8 + 7]) * inv_l);
            *reinterpret_cast<uint4*>(op + c * 16 + i * 8) = u;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[t]);         // the next item's first PV may overwrite O_t
      pair_barrier_id(bar_id);                         // the exchange slot is rewritten by the next item's first block
    }
  } else if (warp >= 4 && warp < 12) {
    // ===================== softmax / correction / epilogue: warps 4-7 tile 0, warps 8-11 tile 1 =====================
    using T = typename TypeOf<DT>::T;
    const int t = (warp - 4) >> 2;
    const int q = warp & 3;                       // TMEM lane quadrant this warp may touch
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t tS = tmem + lane_off + A2_S + t * 128, tO = tmem + lane_off + A2_O + t * 64,
                   tP = tmem + lane_off + A2_P + t * 64;
    const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
    int n = 0;   // blocks this tile has processed so far (all items)
    auto rescale_o = [&](float alpha) {
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t o[16];
        tmem_ld16(tO + c * 16, o);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st16(tO + c * 16, o);
      }
    };
    for (int item = blockIdx.x; item < total_items; item += G) {
      const A2Item it = a2_item(p, item, pairs, full, total_items);
      if (t >= it.nt) continue;
      float m_ref = -INFINITY, l = 0.f;
      for (int j = 0; j < n_kv; ++j, ++n) {
        mbar_wait(&s_full[t], (uint32_t)(n & 1));
        tc_fence_after();
        if (q == 0) A2_STAMP(t, n, 0);
        const int kv_left = p.Skv - j * 128;          // valid columns of this block (>= 128: all)
        uint32_t sa[32], sb[32], pk[16];
        if (kv_left >= 128) {
          // ---- full block, SPECULATIVE exponentials, 32 columns at a time straight from TMEM (the next chunk's
          // tcgen05.ld is in flight while this one is exponentiated; ~100 live registers, no spills).  Probabilities are
          // computed against the reference maximum the row already has (first block: the maximum of its first 32
          // scores) while the true maximum is tracked on the side; only if some row's maximum exceeds its reference by
          // more than 2^8 is the block redone from the scores still in TMEM, and O rescaled.  P <= 2^8 as before.
          tmem_ld32(tS, sa);
          tc_wait_ld();
          if (q == 0) A2_STAMP(t, n, 1);
          tmem_ld32(tS + 32, sb);
          if (j == 0) {
            float a = -INFINITY, b = -INFINITY;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              a = fmax3(a, __uint_as_float(sa[i]), __uint_as_float(sa[i + 1]));
              b = fmax3(b, __uint_as_float(sa[i + 2]), __uint_as_float(sa[i + 3]));
            }
            m_ref = fmaxf(a, b) * p.scale_log2;
          }
          float mx = -INFINITY;
          float2 sum2 = make_float2(0.f, 0.f);
          a2_chunk<DT, PN, PM, true>(sa, sc2, m_ref, mx, sum2, pk);
          if (q == 0) A2_STAMP(t, n, 2);
          mbar_wait(&pv_done[t], (uint32_t)((n & 1) ^ 1));   // P_t (and O_t) may only be touched once PV_t(n - 1) has retired
          tc_fence_after();
          if (q == 0) A2_STAMP(t, n, 3);
          tmem_st16(tP, pk);
          tc_wait_ld();
          tmem_ld32(tS + 64, sa);
          a2_chunk<DT, PN, PM, true>(sb, sc2, m_ref, mx, sum2, pk);
          tmem_st16(tP + 16, pk);
          if (q == 0) A2_STAMP(t, n, 4);
          tc_wait_ld();
          tmem_ld32(tS + 96, sb);
          a2_chunk<DT, PN, PM, true>(sa, sc2, m_ref, mx, sum2, pk);
          tmem_st16(tP + 32, pk);
          if (q == 0) A2_STAMP(t, n, 5);
          tc_wait_ld();
          if (EARLY) {
            // the last chunk's maximum directly, and the block's growth check BEFORE the score columns are handed back
            // to the MMA warp (a redo re-reads them): Q K(j+1)^T then runs under this block's last 32 exponentials
            float a = -INFINITY, b = -INFINITY;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              a = fmax3(a, __uint_as_float(sb[i]), __uint_as_float(sb[i + 1]));
              b = fmax3(b, __uint_as_float(sb[i + 2]), __uint_as_float(sb[i + 3]));
            }
            mx = fmaxf(mx, fmaxf(a, b)) * p.scale_log2;
            const bool grow = mx > m_ref + 8.0f;
            if (__any_sync(0xffffffffu, grow)) {        // rare; this warp's 32 rows only
              const float m_new = grow ? mx : m_ref;
              const float alpha = exp2f(m_ref - m_new);
              sum2 = make_float2(0.f, 0.f);
#pragma unroll 1
              for (int c = 0; c < 3; ++c) {
                tmem_ld32(tS + c * 32, sa);
                tc_wait_ld();
                a2_chunk<DT, 0, 1, false>(sa, sc2, m_new, mx, sum2, pk);
                tmem_st16(tP + c * 16, pk);
              }
              if (j > 0) rescale_o(alpha);
              l *= alpha;
              m_ref = m_new;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[t]);
            a2_chunk<DT, PN, PM, false>(sb, sc2, m_ref, mx, sum2, pk);
            tmem_st16(tP + 48, pk);
          } else {
            a2_chunk<DT, PN, PM, true>(sb, sc2, m_ref, mx, sum2, pk);
            tmem_st16(tP + 48, pk);
            mx *= p.scale_log2;
            const bool grow = mx > m_ref + 8.0f;
            if (__any_sync(0xffffffffu, grow)) {        // rare; this warp's 32 rows only
              const float m_new = grow ? mx : m_ref;
              const float alpha = exp2f(m_ref - m_new);
              sum2 = make_float2(0.f, 0.f);
#pragma unroll 1
              for (int c = 0; c < 4; ++c) {
                tmem_ld32(tS + c * 32, sa);
                tc_wait_ld();
                a2_chunk<DT, 0, 1, false>(sa, sc2, m_new, mx, sum2, pk);
                tmem_st16(tP + c * 16, pk);
              }
              if (j > 0) rescale_o(alpha);
              l *= alpha;
              m_ref = m_new;
            }
          }
          l += sum2.x + sum2.y;
          if (q == 0) A2_STAMP(t, n, 6);
        } else {
          // ---- ragged last block: the classic order, two passes over the scores in TMEM — true maximum of the valid
          // columns first, then their exponentials (columns past the sequence contribute zeros; chunks the MMAs never
          // read are not touched)
          const int nch = (((kv_left + 15) & ~15) + 31) >> 5;   // 32-column chunks the MMAs touch
          float mx = -INFINITY;
#pragma unroll 1
          for (int c = 0; c < nch; ++c) {
            tmem_ld32(tS + c * 32, sa);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i < kv_left) mx = fmaxf(mx, __uint_as_float(sa[i]));
          }
          mx *= p.scale_log2;
          float alpha = 1.0f;
          bool rescale = false;
          if (j == 0) {
            m_ref = mx;
          } else {
            const bool grow = mx > m_ref + 8.0f;
            rescale = __any_sync(0xffffffffu, grow);
            if (rescale) {
              const float m_new = grow ? mx : m_ref;
              alpha = exp2f(m_ref - m_new);
              l *= alpha;
              m_ref = m_new;
            }
          }
          mbar_wait(&pv_done[t], (uint32_t)((n & 1) ^ 1));
          tc_fence_after();
          if (rescale) rescale_o(alpha);
          float2 sum2 = make_float2(0.f, 0.f);
          const float2 nm2 = make_float2(-m_ref, -m_ref);
#pragma unroll 1
          for (int c = 0; c < nch; ++c) {
            tmem_ld32(tS + c * 32, sa);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int col = c * 32 + 2 * i;
              float2 e = make_float2(0.f, 0.f);
              if (col < kv_left) {
                const float2 x = __ffma2_rn(make_float2(__uint_as_float(sa[2 * i]), __uint_as_float(sa[2 * i + 1])), sc2, nm2);
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(x.x));
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(x.y));
                if (col + 1 >= kv_left) e.y = 0.f;
              }
              sum2 = __fadd2_rn(sum2, e);
              pk[i] = pack2<DT>(e.x, e.y);
            }
            tmem_st16(tP + c * 16, pk);
          }
          l += sum2.x + sum2.y;
          if (EARLY) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[t]);
          }
        }
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
        if (q == 0) A2_STAMP(t, n, 7);
      }
      // epilogue of the item: O / l
      mbar_wait(&pv_done[t], (uint32_t)((n - 1) & 1));
      tc_fence_after();
      const float inv_l = 1.0f / l;
      const long long qrow = (long long)(it.qt + t) * 128 + row;
      T* op = reinterpret_cast<T*>(p.out) + ((long long)it.b * p.Sq + qrow) * p.ldo + it.h * 64;
      {
        // all 64 accumulator columns in one TMEM round trip (four dependent load -> wait -> store rounds cost an item
        // ~2800 cycles here, r2_attn_trace_258_items.log), then eight 16-byte stores per row
        uint32_t o0[32], o1[32];
        tmem_ld32(tO, o0);
        tmem_ld32(tO + 32, o1);
        tc_wait_ld();
        if (qrow < p.Sq) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 u;
            u.x = pack2<DT>(__uint_as_float(o0[i * 8 + 0]) * inv_l, __uint_as_float(o0[i * 8 + 1]) * inv_l);
            u.y = pack2<DT>(__uint_as_float(o0[i * 8 + 2]) * inv_l, __uint_as_float(o0[i * 8 + 3]) * inv_l);
            u.z = pack2<DT>(__uint_as_float(o0[i * 8 + 4]) * inv_l, __uint_as_float(o0[i * 8 + 5]) * inv_l);
            u.w = pack2<DT>(__uint_as_float(o0[i * 8 + 6]) * inv_l, __uint_as_float(o0[i * 8 + 7]) * inv_l);
            *reinterpret_cast<uint4*>(op + i * 8) = u;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 u;
            u.x = pack2<DT>(__uint_as_float(o1[i * 8 + 0]) * inv_l, __uint_as_float(o1[i * 8 + 1]) * inv_l);
            u.y = pack2<DT>(__uint_as_float(o1[i * 8 + 2]) * inv_l, __uint_as_float(o1[i * 8 + 3]) * inv_l);
            u.z = pack2<DT>(__uint_as_float(o1[i * 8 + 4]) * inv_l, __uint_as_float(o1[i * 8 + 5]) * inv_l);
            u.w = pack2<DT>(__uint_as_float(o1[i * 8 + 6]) * inv_l, __uint_as_float(o1[i * 8 + 7]) * inv_l);
            *reinterpret_cast<uint4*>(op + 32 + i * 8) = u;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[t]);         // the next item's first PV may overwrite O_t
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Short sequences (Sq, Skv <= 32; head_dim 64): the six-token sequence of the stage-1 prior
// (/root/reference/src/models/stage1_prior_transformer.py:262-285).  A 128-row tcgen05 tile would be > 95 % padding and
// its set-up (TMEM allocation, tensor maps, barrier ring) is pure latency in a chain of launch-bound kernels, so this is
// plain CUDA-core code: one warp per (batch, head).  Lane s owns key row s (its 64 channels in registers) and computes
// the score of every query against it; softmax statistics are warp reductions (fixed xor order); then lane j owns
// output channels 2j, 2j+1 and accumulates sum_s p_s V[s] with p_s broadcast by shuffle.  fp32 throughout; P is NOT
// rounded to 16 bits here (closer to the fp32 reference than the tensor-core kernel).
template <int DT>
__global__ void __launch_bounds__(128) attention_small_kernel(const void* __restrict__ q, long long ldq,
                                                             const void* __restrict__ k, long long ldk,
                                                             const void* __restrict__ v, long long ldv,
                                                             void* __restrict__ out, long long ldo, int B, int heads,
                                                             int Sq, int Skv, float scale_log2) {
  using T = typename TypeOf<DT>::T;
  __shared__ __align__(16) uint32_t sq[4][32][32];   // per warp: the query rows, 64 channels as 32 packed pairs
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.x * 4 + warp;
  pdl_wait();
  if (bh >= B * heads) return;
  const int h = bh % heads, b = bh / heads;
  const T* qp = reinterpret_cast<const T*>(q) + (long long)b * Sq * ldq + h * 64;
  const T* kp = reinterpret_cast<const T*>(k) + (long long)b * Skv * ldk + h * 64;
  const T* vp = reinterpret_cast<const T*>(v) + (long long)b * Skv * ldv + h * 64;
  T* op = reinterpret_cast<T*>(out) + (long long)b * Sq * ldo + h * 64;
  for (int i = 0; i < Sq; ++i) sq[warp][i][lane] = *reinterpret_cast<const uint32_t*>(qp + (long long)i * ldq + 2 * lane);
  uint4 kr[8];                                       // key row `lane`: 64 channels
#pragma unroll
  for (int j = 0; j < 8; ++j)
    kr[j] = lane < Skv ? *reinterpret_cast<const uint4*>(kp + (long long)lane * ldk + 8 * j) : make_uint4(0u, 0u, 0u, 0u);
  uint32_t vr[32];                                   // value channels 2*lane, 2*lane+1 of every kv row
#pragma unroll
  for (int s2 = 0; s2 < 32; ++s2)
    vr[s2] = s2 < Skv ? *reinterpret_cast<const uint32_t*>(vp + (long long)s2 * ldv + 2 * lane) : 0u;
  __syncwarp();
  for (int i = 0; i < Sq; ++i) {
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint4 qq = *reinterpret_cast<const uint4*>(&sq[warp][i][4 * j]);   // broadcast read
      const uint32_t qa[4] = {qq.x, qq.y, qq.z, qq.w}, ka[4] = {kr[j].x, kr[j].y, kr[j].z, kr[j].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 a = unpack2<DT>(qa[e]), c = unpack2<DT>(ka[e]);
        dot = fmaf(a.x, c.x, dot);
        dot = fmaf(a.y, c.y, dot);
      }
    }
    const float sc = lane < Skv ? dot * scale_log2 : -INFINITY;
    float mx = sc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float pr = lane < Skv ? exp2f(sc - mx) : 0.f;
    float sum = pr;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int s2 = 0; s2 < 32; ++s2) {
      const float ps = __shfl_sync(0xffffffffu, pr, s2);
      const float2 vv = unpack2<DT>(vr[s2]);
      acc.x = fmaf(ps, vv.x, acc.x);
      acc.y = fmaf(ps, vv.y, acc.y);
    }
    const float inv = 1.0f / sum;
    *reinterpret_cast<uint32_t*>(op + (long long)i * ldo + 2 * lane) = pack2<DT>(acc.x * inv, acc.y * inv);
  }
}

}  // namespace pcdm

using namespace pcdm;

static int make_qkv_map(CUtensorMap* m, const void* base, long long ld, int S, int heads, int B, int box_rows, int hd) {
  // element (b, s, h, d) at ((b*S + s) * ld + h*hd + d); boxes are 64 columns wide (one 128B-swizzled operand tile)
  const uint64_t dims[4] = {(uint64_t)hd, (uint64_t)S, (uint64_t)heads, (uint64_t)B};
  const uint64_t strides[3] = {(uint64_t)ld * 2, (uint64_t)hd * 2, (uint64_t)S * ld * 2};
  const uint32_t box[4] = {64, (uint32_t)box_rows, 1, 1};
  return make_tmap(m, base, 4, dims, strides, box);
}


extern "C" int pcdm_attention(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                              void* out, long long ldo, int B, int heads, int Sq, int Skv, float scale, int dtype,
                              void* stream_) {
  return pcdm_attention_hd(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, Sq, Skv, 64, scale, dtype, stream_);
}

extern "C" int pcdm_attention_hd(const void* q, long long ldq, const void* k, long long ldk, const void* v,
                                 long long ldv, void* out, long long ldo, int B, int heads, int Sq, int Skv,
                                 int head_dim, float scale, int dtype, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (head_dim != 64 && head_dim != 128)
    return set_error(PCDM_ERR_UNSUPPORTED, "attention: head_dim must be 64 or 128 (pad other widths with zero columns)");
  if (!q || !k || !v || !out) return set_error(PCDM_ERR_INVALID, "attention: null pointer");
  if (dtype != DT_F16 && dtype != DT_BF16) return set_error(PCDM_ERR_INVALID, "attention: bad dtype");
  if (B <= 0 || heads <= 0 || Sq <= 0 || Skv <= 0) return set_error(PCDM_ERR_INVALID, "attention: empty problem");
  if ((ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 8)) return set_error(PCDM_ERR_UNSUPPORTED, "attention: strides must be multiples of 8");
  if (g_tune.att_small && head_dim == 64 && Sq <= 32 && Skv <= 32) {   // short sequences: CUDA-core kernel, one warp per (batch, head)
    const int grid_s = (B * heads + 3) / 4;
    const float sl2 = scale * 1.4426950408889634f;
    if (dtype == DT_F16)
      PCDM_CUDA(launch_kernel(attention_small_kernel<DT_F16>, dim3(grid_s), dim3(128), 0, stream, 1, q, ldq, k, ldk, v,
                              ldv, out, ldo, B, heads, Sq, Skv, sl2));
    else
      PCDM_CUDA(launch_kernel(attention_small_kernel<DT_BF16>, dim3(grid_s), dim3(128), 0, stream, 1, q, ldq, k, ldk, v,
                              ldv, out, ldo, B, heads, Sq, Skv, sl2));
    PCDM_CUDA(cudaGetLastError());
    return 0;
  }
  AttnParams p;
  memset(&p, 0, sizeof(p));
  // TMA boxes are always 128 rows: rows past the end of a (b, h) sequence are out of bounds for the 4-D map and
  // are zero-filled, so short sequences need no special casing (zero K rows are masked, zero V rows add nothing).
  PCDM_CHECK(make_qkv_map(&p.tmQ, q, ldq, Sq, heads, B, 128, head_dim), "Q map");
  PCDM_CHECK(make_qkv_map(&p.tmK, k, ldk, Skv, heads, B, 128, head_dim), "K map");
  PCDM_CHECK(make_qkv_map(&p.tmV, v, ldv, Skv, heads, B, 128, head_dim), "V map");
  p.B = B; p.heads = heads; p.Sq = Sq; p.Skv = Skv;
  p.q_tiles = (Sq + 127) / 128;
  p.out = out; p.ldo = ldo;
  p.scale_log2 = scale * 1.4426950408889634f;
  // head_dim 64, two or more query tiles: the persistent two-tile kernel; single-tile sequences stay on the one-CTA-
  // per-tile kernel (measured equal or a few % faster there)
  if (head_dim == 64 && g_tune.att_v2 && p.q_tiles >= 2) {
    const long long pair_slots = (long long)B * heads * ((p.q_tiles + 1) / 2);
    const int grid2 = (int)(pair_slots < num_sms() ? pair_slots : num_sms());
#define A2_LAUNCH(DT_, PN_, PM_, EARLY_, HALF_) A2_LAUNCH_T(DT_, PN_, PM_, EARLY_, HALF_, false)
#define A2_LAUNCH_T(DT_, PN_, PM_, EARLY_, HALF_, TRACE_)                                                                                    \
  do {                                                                                                                \
    PCDM_ENSURE_SMEM(A2_SMEM, (attention2_kernel<DT_, PN_, PM_, EARLY_, HALF_, TRACE_>));                             \
    PCDM_CUDA(launch_kernel(attention2_kernel<DT_, PN_, PM_, EARLY_, HALF_, TRACE_>, dim3(grid2),                     \
                            dim3(HALF_ ? A2_THREADS_HALF : A2_THREADS), A2_SMEM, stream, 1, p));                      \
  } while (0)
#ifdef PCDM_EXPERIMENT
    // g_tune.att_dbg: 0 the release setting (1 of 4 column pairs' exp2 on the FMA pipe), 1 none, 2 = 1 of 3 (tools/
    // dev_attn3.py: 124 / 133 / 131 us at 2048 x 2048; 1 of 2: 142), 3 = release setting + early hand-back of the score
    // columns, 4 = no FMA-pipe exp2 + early hand-back, 5 = half-row threads (16 softmax warps) with the release exp2 share,
    // 6 = half-row threads without FMA-pipe exp2
    if (g_tune.att_trace && dtype == DT_BF16) {   // cycle stamps of CTA 0 (release configuration of the kernel)
      p.trace = reinterpret_cast<unsigned long long*>(g_tune.att_trace);
      A2_LAUNCH_T(DT_BF16, 1, 4, false, false, true);
      PCDM_CUDA(cudaGetLastError());
      return 0;
    }
    if (g_tune.att_dbg) {
      if (dtype == DT_F16) {
        switch (g_tune.att_dbg) {
          case 1: A2_LAUNCH(DT_F16, 0, 1, false, false); break;
          case 2: A2_LAUNCH(DT_F16, 1, 3, false, false); break;
          case 3: A2_LAUNCH(DT_F16, 1, 4, true, false); break;
          case 5: A2_LAUNCH(DT_F16, 1, 4, false, true); break;
          case 6: A2_LAUNCH(DT_F16, 0, 1, false, true); break;
          default: A2_LAUNCH(DT_F16, 0, 1, true, false); break;
        }
      } else {
        switch (g_tune.att_dbg) {
          case 1: A2_LAUNCH(DT_BF16, 0, 1, false, false); break;
          case 2: A2_LAUNCH(DT_BF16, 1, 3, false, false); break;
          case 3: A2_LAUNCH(DT_BF16, 1, 4, true, false); break;
          case 5: A2_LAUNCH(DT_BF16, 1, 4, false, true); break;
          case 6: A2_LAUNCH(DT_BF16, 0, 1, false, true); break;
          default: A2_LAUNCH(DT_BF16, 0, 1, true, false); break;
        }
      }
      PCDM_CUDA(cudaGetLastError());
      return 0;
    }
#endif
    if (dtype == DT_F16) A2_LAUNCH(DT_F16, 1, 4, false, false);
    else A2_LAUNCH(DT_BF16, 1, 4, false, false);
#undef A2_LAUNCH
#undef A2_LAUNCH_T
    PCDM_CUDA(cudaGetLastError());
    return 0;
  }
  const int grid = B * heads * p.q_tiles;
#define ATT_LAUNCH(DT_, P_, HD_)                                                                                    \
  do {                                                                                                              \
    PCDM_ENSURE_SMEM(att_smem_bytes(HD_), attention_kernel<DT_, P_, HD_>);                                          \
    PCDM_CUDA(launch_kernel(attention_kernel<DT_, P_, HD_>, dim3(grid), dim3(ATT_THREADS), att_smem_bytes(HD_),     \
                            stream, 1, p));                                                                         \
  } while (0)
  if (head_dim == 128) {
    if (dtype == DT_F16) ATT_LAUNCH(DT_F16, 0, 128);
    else ATT_LAUNCH(DT_BF16, 0, 128);
  } else if (dtype == DT_F16) {
    if (g_tune.att_poly) ATT_LAUNCH(DT_F16, 2, 64);
    else ATT_LAUNCH(DT_F16, 0, 64);
  } else {
    if (g_tune.att_poly) ATT_LAUNCH(DT_BF16, 2, 64);
    else ATT_LAUNCH(DT_BF16, 0, 64);
  }
#undef ATT_LAUNCH
  PCDM_CUDA(cudaGetLastError());
  return 0;
}
