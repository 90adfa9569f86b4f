// Host-side helpers shared by the C-ABI entry points: error reporting, device info, TMA tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/pcdm_b200.h"

namespace pcdm {

int set_error(int code, const char* fmt, ...);
int num_sms();
// rank-D tiled tensor map over a 16-bit tensor, 128-byte swizzle, zero OOB fill.
// dims/box are innermost-first; strides_bytes has rank-1 entries (stride of dims 1..rank-1).
int make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, int swizzle_bytes = 128);

// Every tuning knob of the library in one place.  The RELEASE library (libpcdm_b200.so) holds them as compile-time
// constants: it has no mutable process-wide state, and every entry point is safe to call concurrently from several
// host threads / on several streams / on several devices (per-call scratch comes in through pcdm_ext).  Built with
// -DPCDM_EXPERIMENT (libpcdm_b200_exp.so, used by tools/ only) the knobs become a mutable struct behind the
// pcdm_set_* hooks declared in include/pcdm_b200_experiment.h.
struct Tuning {
  int pdl = 1;             // launch with programmatic stream serialization
  int force_cg = 0;        // 0 auto, 1 single-CTA tiles, 2 CTA pairs wherever the N tile allows
  int max_stages = 8;      // cap on the GEMM/conv shared-memory ring depth
  int gemm_dbg = 0;        // experiment mask (results WRONG when non-zero)
  int skinny = 1;          // M <= 32 GEMMs on the weight-streaming kernel
  int att_small = 1;       // Sq, Skv <= 32 attention on the one-warp-per-head kernel
  int att_poly = 0;        // half of the softmax exp2 on the FMA pipe
  int att_v2 = 1;          // head_dim 64, >= 2 query tiles: the persistent two-tile kernel (0: the round-1 kernel, for A/B)
  void* att_trace = nullptr;   // experiment build: device buffer for the two-tile attention kernel's cycle stamps
  int att_dbg = 0;         // experiment build: FMA-pipe exp2 share of that kernel (0 release setting 1/4, 1 none, 2 = 1/3, 3 = 1/2)
  int gn_mode = 0;         // 0 auto, 1 two kernels, 2 single pass, 2 + T single pass with T threads per CTA
};
#ifdef PCDM_EXPERIMENT
extern Tuning g_tune;
#else
constexpr Tuning g_tune{};
#endif

// Launch `kernel` with the PDL attribute (and an optional 1-D cluster size).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 int cluster, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  int n = 0;
  if (g_tune.pdl) {
    attrs[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster > 1) {
    attrs[n].id = cudaLaunchAttributeClusterDimension;
    attrs[n].val.clusterDim.x = cluster;
    attrs[n].val.clusterDim.y = 1;
    attrs[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attrs;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#define PCDM_CUDA(expr)                                                                      \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return ::pcdm::set_error(PCDM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                               __LINE__);                                                    \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is per device: remember which devices a kernel was configured on
#define PCDM_ENSURE_SMEM(bytes, ...)                                                                             \
  do {                                                                                                           \
    static bool _cfg[64] = {};                                                                                   \
    int _dev = 0;                                                                                                \
    PCDM_CUDA(cudaGetDevice(&_dev));                                                                             \
    if (_dev < 0 || _dev >= 64 || !_cfg[_dev]) {                                                                 \
      PCDM_CUDA(cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributeMaxDynamicSharedMemorySize, (bytes)));        \
      if (_dev >= 0 && _dev < 64) _cfg[_dev] = true;                                                             \
    }                                                                                                            \
  } while (0)

#define PCDM_CHECK(expr, what)                  \
  do {                                          \
    int _r = (expr);                            \
    if (_r != 0) return _r;                     \
  } while (0)

}  // namespace pcdm
