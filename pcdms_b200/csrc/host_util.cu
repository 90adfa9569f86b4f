#include "host_util.h"
#include "../../include/pcdm_b200_experiment.h"
#include <stdarg.h>

namespace pcdm {

static thread_local char g_err[512] = "";
#ifdef PCDM_EXPERIMENT
Tuning g_tune;
#endif

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int num_sms() {   // of the CURRENT device (cached per device)
  static int n[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (n[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
    n[dev] = v;
  }
  return n[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(PCDM_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  if (reinterpret_cast<uintptr_t>(base) & 15)
    return set_error(PCDM_ERR_INVALID, "tensor map: base pointer must be 16-byte aligned");
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                       : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return set_error(PCDM_ERR_CUDA,
                     "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu] box [%u %u %u %u]", (int)r,
                     rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                     (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
                     rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
  }
  return 0;
}

}  // namespace pcdm

extern "C" const char* pcdm_last_error(void) { return pcdm::g_err; }

extern "C" int pcdm_abi_version(void) { return PCDM_ABI_VERSION; }

#ifdef PCDM_EXPERIMENT
// ---- experiment hooks (include/pcdm_b200_experiment.h): present in libpcdm_b200_exp.so only ----
extern "C" int pcdm_set_pdl(int enabled) { pcdm::g_tune.pdl = enabled ? 1 : 0; return 0; }
extern "C" int pcdm_set_gemm_cta_group(int mode) {
  if (mode < 0 || mode > 2) return pcdm::set_error(PCDM_ERR_INVALID, "gemm cta group mode must be 0, 1 or 2");
  pcdm::g_tune.force_cg = mode;
  return 0;
}
extern "C" int pcdm_set_gemm_max_stages(int n) {
  if (n < 2 || n > 8) return pcdm::set_error(PCDM_ERR_INVALID, "gemm max stages must be in [2, 8]");
  pcdm::g_tune.max_stages = n;
  return 0;
}
extern "C" int pcdm_set_gemm_debug(int mask) { pcdm::g_tune.gemm_dbg = mask; return 0; }
extern "C" int pcdm_set_skinny_gemm(int enabled) { pcdm::g_tune.skinny = enabled ? 1 : 0; return 0; }
extern "C" int pcdm_set_attention_small(int on) { pcdm::g_tune.att_small = on ? 1 : 0; return 0; }
extern "C" int pcdm_set_attention_poly(int on) { pcdm::g_tune.att_poly = on ? 1 : 0; return 0; }
extern "C" int pcdm_set_attention_v2(int on) { pcdm::g_tune.att_v2 = on ? 1 : 0; return 0; }
extern "C" int pcdm_set_attention_debug(int mask) { pcdm::g_tune.att_dbg = mask; return 0; }
extern "C" int pcdm_set_attention_trace(void* buf) { pcdm::g_tune.att_trace = buf; return 0; }
extern "C" int pcdm_set_groupnorm_two_pass(int mode) { pcdm::g_tune.gn_mode = mode; return 0; }
#endif
