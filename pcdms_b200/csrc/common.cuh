// pcdms_b200 — sm_100a device-side primitives (inline PTX wrappers).
// mbarrier / TMA / tcgen05 / TMEM helpers shared by every kernel in csrc/.
// Everything here is written for sm_100a only (B200); there is no fallback path.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace pcdm {

// ----------------------------------------------------------------------------------------------
// dtype tags (the C-ABI passes these as ints)
// ----------------------------------------------------------------------------------------------
enum : int { DT_F16 = 0, DT_BF16 = 1 };

template <int DT> struct TypeOf;
template <> struct TypeOf<DT_F16>  { using T = __half;         using T2 = __half2; };
template <> struct TypeOf<DT_BF16> { using T = __nv_bfloat16;  using T2 = __nv_bfloat162; };

template <int DT> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<DT_F16>(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<DT_BF16>(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <int DT> __device__ __forceinline__ float2 unpack2(uint32_t v);
template <> __device__ __forceinline__ float2 unpack2<DT_F16>(uint32_t v) {
  return __half22float2(*reinterpret_cast<__half2*>(&v));
}
template <> __device__ __forceinline__ float2 unpack2<DT_BF16>(uint32_t v) {
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v));
}
template <int DT> __device__ __forceinline__ float to_f32(typename TypeOf<DT>::T v);
template <> __device__ __forceinline__ float to_f32<DT_F16>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<DT_BF16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <int DT> __device__ __forceinline__ typename TypeOf<DT>::T from_f32(float v);
template <> __device__ __forceinline__ __half from_f32<DT_F16>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<DT_BF16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a broken pipeline traps (CUDA error on the host) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xFFFu) == 0) {
      uint64_t now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) {  // 4 s
        printf("pcdm: mbarrier wait timeout block %d thread %d\n", (int)blockIdx.x, (int)threadIdx.x);
        __trap();
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), tile mode, completion on an mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ---- programmatic dependent launch (PDL): every kernel lets its successor start early and itself waits for its
//      predecessor right before the first global-memory access, so launch latency + prologues overlap kernel tails.
//      Both instructions are no-ops when the launch carries no programmatic dependency.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// One lane of a CONVERGED warp.  The TMA / MMA issuing warps stay warp-uniform and elect at the instruction: the
// operands of UTMALDG / UTCHMMA then live in uniform registers; a `lane == 0` branch instead makes the compiler move
// every operand vector->uniform (R2UR) inside a per-thread serialisation loop — ~80 cycles per MMA issue, measured.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---- cluster / CTA-pair (cta_group::2) helpers -------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {  // every thread of every CTA in the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ double2 ld_dsmem_f64x2(uint32_t cluster_addr) {
  double a, b;
  asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(cluster_addr) : "memory");
  return make_double2(a, b);
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default semantics (.release at CTA scope), as CUTLASS's ClusterBarrier::arrive(cta_id): what this arrive hands over
  // lives in TMEM and is ordered by the tcgen05 fences either side; `.release.cluster` cost a MEMBAR.ALL.GPU + ERRBAR
  // per epilogue warp per tile (11 % of the samples of a short-K GEMM, profiles/r1_s3_shortk_gemm.md)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads issued by either CTA of a pair whose completion bytes are credited to a barrier in the LEADER CTA
__device__ __forceinline__ void tma_load_2d_cg2(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_cg2(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}

// TMA store (shared -> global), bulk-group completion
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {  // <= N most recent groups may still be reading shared memory
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // whole warp, ncols pow2 >= 32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* smem_dst, uint32_t ncols) {  // one warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// tcgen05.commit: arrive on an mbarrier once every MMA issued so far by this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// pair variant: arrives on the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit_cg2(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}

// Shared-memory matrix descriptor (sm_100 "version 1").  K-major, 128-byte swizzle: rows are 128 B
// (64 x 16-bit), groups of 8 rows are 1024 B apart (SBO); LBO is unused for swizzled K-major.
// Bit layout: [0,14) addr>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout (2 = SW128).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}

// Instruction descriptor, kind::f16, fp32 accumulate.
// [4,6) D fmt (1 = f32) | [7,10) A fmt | [10,13) B fmt (0 = f16, 1 = bf16) | 15 A major | 16 B major (1 = MN)
// [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc(int dt, int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | ((uint32_t)dt << 7) | ((uint32_t)dt << 10) | ((uint32_t)a_mn_major << 15) |
         ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// CTA-pair MMA: D[256 x N] split over the two CTAs' TMEM (128 rows each); each CTA supplies its 128 rows of A and
// half of the N rows of B from its own shared memory (same offsets in both CTAs).  Issued by the leader CTA only.
__device__ __forceinline__ void umma_ss_cg2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (thread t gets lane t's row).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// registers -> TMEM, 32 lanes x 16 columns (used to stage the fp16/bf16 P operand of attention).
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// registers -> TMEM, 32 lanes x 8 columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
      : "memory");
}

// ----------------------------------------------------------------------------------------------
// misc
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
// Two SiLUs with the fp32 arithmetic on the packed FFMA2 path (one issue slot per pair); ex2 / rcp stay on the SFU.
__device__ __forceinline__ float2 silu2_exact(float2 x) {
  const float2 t = __fmul2_rn(x, make_float2(-1.4426950408889634f, -1.4426950408889634f));
  float e0, e1, r0, r1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t.y));
  const float2 d = __fadd2_rn(make_float2(e0, e1), make_float2(1.0f, 1.0f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d.y));
  return __fmul2_rn(x, make_float2(r0, r1));
}
// x sigmoid(x) = h + h tanh(h), h = x / 2: ONE SFU op per element instead of two.  tanh.approx is good to 2^-11
// relative, a quarter of a bf16 ulp of the result — used for bf16 outputs only (fp16 keeps the ex2 / rcp form).
__device__ __forceinline__ float2 silu2_tanh(float2 x) {
  const float2 h = __fmul2_rn(x, make_float2(0.5f, 0.5f));
  float t0, t1;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h.y));
  return __ffma2_rn(h, make_float2(t0, t1), h);
}
template <int DT> __device__ __forceinline__ float2 silu2_f(float2 x) {
  return DT == DT_BF16 ? silu2_tanh(x) : silu2_exact(x);
}
// exact-form GELU 0.5 x (1 + erf(x / sqrt 2)) with erf from Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, far below
// the 16-bit output rounding): one rcp + one ex2 on the SFU and a degree-5 Horner chain, instead of libm erff's
// two-branch polynomial — the GEGLU epilogue evaluates this 42 M times per UNet step.
__device__ __forceinline__ float gelu_erf_f(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t;   // rcp.approx: a bare MUFU.RCP (1 ulp) — __frcp_rn would add a range check + slow-path call per element
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float erf_abs = 1.0f - poly * exp2f(-1.4426950408889634f * z * z);
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}

// Two exp2 on the FMA pipe instead of the SFU (16 ex2/clk/SM is what bounds attention at head_dim 64).  Cody-Waite:
// n = round(x) through the 1.5 * 2^23 magic add, f = x - n in [-0.5, 0.5], 2^f by a degree-4 polynomial (relative error
// 4e-5: a sixth of an fp16 ulp of the 16-bit P it feeds), exponent spliced in with one integer shift-add.
// x <= ~8 by construction (lazy rescale threshold); very negative x is clamped so that n stays in the 9-bit field.
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -120.0f);
  x.y = fmaxf(x.y, -120.0f);
  const float2 magic = make_float2(12582912.0f, 12582912.0f);
  const float2 r = __fadd2_rn(x, magic);
  const float2 n = __fadd2_rn(r, make_float2(-12582912.0f, -12582912.0f));
  const float2 f = __fadd2_rn(x, make_float2(-n.x, -n.y));
  float2 p = __ffma2_rn(make_float2(0.0096181291f, 0.0096181291f), f, make_float2(0.0555041087f, 0.0555041087f));
  p = __ffma2_rn(p, f, make_float2(0.2402265070f, 0.2402265070f));
  p = __ffma2_rn(p, f, make_float2(0.6931471806f, 0.6931471806f));
  p = __ffma2_rn(p, f, make_float2(1.0f, 1.0f));
  return make_float2(__uint_as_float(__float_as_uint(p.x) + (__float_as_uint(r.x) << 23)),
                     __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(r.y) << 23)));
}

// Two GELUs at once with ONE SFU op per element (round 2; the GEGLU epilogue evaluates this 42 M times per UNet step and
// its three per-tile costs — SFU ops, TMEM reads, issue slots — add up instead of overlapping).  With t = |x|:
//   gelu(x) = x Phi(x) = max(x, 0) - t Phi(-t),   Phi(-t) = 2^P(t),   P = polynomial fit of log2 Phi(-t)
// (no reciprocal, no cancellation in the negative tail: the small quantity is produced directly by the ex2).  The fit is
// weighted by t Phi(-t), i.e. minimises the absolute error of the result: DEG 8 -> 2.5e-7 (the A&S 7.1.26 form it
// replaces: 1.5e-7 |x| / 2), DEG 5 -> 6.4e-7 — both at the level of one fp32 rounding of the result and 2-3 orders of
// magnitude below the 16-bit output rounding (tools/fit_gelu.py regenerates and checks the coefficients).  The leading
// coefficients are negative, so P -> -inf and the correction term vanishes for |x| beyond the fitted range.
// ~10 (DEG 5) / ~13 (DEG 8) issue slots + 2 SFU ops per pair, against ~19 + 4 for the form below.
template <int DEG>
__device__ __forceinline__ float2 gelu2_phi(float2 x) {
  static_assert(DEG == 5 || DEG == 8, "fitted degrees");
  const float2 t = make_float2(fabsf(x.x), fabsf(x.y));
  float2 p;
#define PCDM_C2(v) make_float2(v, v)
  if (DEG == 5) {
    p = __ffma2_rn(PCDM_C2(-4.732939706e-04f), t, PCDM_C2(7.084460929e-03f));
    p = __ffma2_rn(p, t, PCDM_C2(-5.182716995e-02f));
    p = __ffma2_rn(p, t, PCDM_C2(-4.599926472e-01f));
    p = __ffma2_rn(p, t, PCDM_C2(-1.150787711e+00f));
    p = __ffma2_rn(p, t, PCDM_C2(-1.000037670e+00f));
  } else {
    p = __ffma2_rn(PCDM_C2(-1.690381168e-06f), t, PCDM_C2(2.508305806e-05f));
    p = __ffma2_rn(p, t, PCDM_C2(-1.144614507e-04f));
    p = __ffma2_rn(p, t, PCDM_C2(-3.233452735e-04f));
    p = __ffma2_rn(p, t, PCDM_C2(7.333388552e-03f));
    p = __ffma2_rn(p, t, PCDM_C2(-5.271420255e-02f));
    p = __ffma2_rn(p, t, PCDM_C2(-4.591154456e-01f));
    p = __ffma2_rn(p, t, PCDM_C2(-1.151123285e+00f));
    p = __ffma2_rn(p, t, PCDM_C2(-9.999988675e-01f));
  }
#undef PCDM_C2
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(p.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(p.y));
  return __ffma2_rn(make_float2(-t.x, -t.y), make_float2(e0, e1), make_float2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f)));
}

// The former two-SFU-op form (A&S 7.1.26: rcp + 5-term Horner + ex2), kept for the A/B build (-DPCDM_GELU_AS).
__device__ __forceinline__ float2 gelu_erf2_as(float2 x) {
  const float2 z = __fmul2_rn(make_float2(fabsf(x.x), fabsf(x.y)), make_float2(0.70710678118654752f, 0.70710678118654752f));
  const float2 d = __ffma2_rn(make_float2(0.3275911f, 0.3275911f), z, make_float2(1.0f, 1.0f));
  float t0, t1, e0, e1;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d.y));
  const float2 t = make_float2(t0, t1);
  float2 poly = __ffma2_rn(make_float2(1.061405429f, 1.061405429f), t, make_float2(-1.453152027f, -1.453152027f));
  poly = __ffma2_rn(poly, t, make_float2(1.421413741f, 1.421413741f));
  poly = __ffma2_rn(poly, t, make_float2(-0.284496736f, -0.284496736f));
  poly = __ffma2_rn(poly, t, make_float2(0.254829592f, 0.254829592f));
  poly = __fmul2_rn(poly, t);
  const float2 a = __fmul2_rn(__fmul2_rn(z, make_float2(-1.4426950408889634f, -1.4426950408889634f)), z);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a.y));
  const float2 erf_abs = __ffma2_rn(make_float2(-poly.x, -poly.y), make_float2(e0, e1), make_float2(1.0f, 1.0f));
  const float2 erf = make_float2(copysignf(erf_abs.x, x.x), copysignf(erf_abs.y, x.y));
  const float2 h = __fmul2_rn(x, make_float2(0.5f, 0.5f));
  return __ffma2_rn(h, erf, h);
}

// exact-form GELU for a 16-bit output of type DT: bf16 takes the degree-5 fit, fp16 the degree-8 one
template <int DT> __device__ __forceinline__ float2 gelu_erf2_f(float2 x) {
#ifdef PCDM_GELU_AS
  return gelu_erf2_as(x);
#else
  return gelu2_phi<DT == DT_BF16 ? 5 : 8>(x);
#endif
}

// max of three in ONE instruction (FMNMX3, sm_100): halves the row-maximum pass of the attention softmax
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace pcdm
