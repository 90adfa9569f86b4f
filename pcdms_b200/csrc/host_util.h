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

extern int g_pdl_enabled;   // pcdm_set_pdl(): 1 = launch with programmatic stream serialization (default), 0 = plain

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
  if (g_pdl_enabled) {
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

#define PCDM_CHECK(expr, what)                  \
  do {                                          \
    int _r = (expr);                            \
    if (_r != 0) return _r;                     \
  } while (0)

}  // namespace pcdm
