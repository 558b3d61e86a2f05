#include "host_util.h"

#include <atomic>
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <string.h>

namespace ldm_host {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

const char* last_error() { return g_err; }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(LDM_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return LDM_OK;
}

#ifdef LDM_DIAG
int diag_env(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
bool diag_env_has(const char* name, const char* word) {
  const char* e = getenv(name);
  return e && strstr(e, word);
}
#endif

// SM count of the CURRENT device (a process may drive several): immutable per-device cache
int num_sms() {
  static std::atomic<int> cache[64];
  int dev = 0;
  cudaGetDevice(&dev);
  std::atomic<int>& slot = cache[dev & 63];
  int sms = slot.load(std::memory_order_relaxed);
  if (sms == 0) {
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    slot.store(sms, std::memory_order_relaxed);
  }
  return sms;
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode() {
  static encode_tiled_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<encode_tiled_fn>(p);
  }
  return fn;
}

int make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, int elem_bytes, bool swizzle128, const uint32_t* elem_strides) {
  encode_tiled_fn enc = get_encode();
  if (!enc) return set_error(LDM_ERR_DRIVER, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0)
    return set_error(LDM_ERR_ALIGNMENT, "tensor map base %p not 16-byte aligned", base);
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides ? elem_strides[i] : 1;  // > 1: every es-th element; box[i] is then in traversal space
    if (box[i] == 0 || box[i] > 256) return set_error(LDM_ERR_BAD_SHAPE, "tensor map box[%d]=%u out of range", i, box[i]);
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gs[i] = strides_bytes[i];
    if (gs[i] % 16 != 0) return set_error(LDM_ERR_ALIGNMENT, "tensor map stride[%d]=%llu not a multiple of 16", i,
                                          (unsigned long long)gs[i]);
  }
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(LDM_ERR_DRIVER, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return LDM_OK;
}

const char* last_error();
long long launches();

}  // namespace ldm_host

extern "C" {

int ldm_abi_version(void) { return 4; }  // 4: ldm_gemm_desc.row_stats_out / ln_stats (LayerNorm fold); 3: ldm_gemm_desc.splitk_ws; 2: ldm_gemm_desc.qkv_part0, ldm_attn_desc.kv_seq
const char* ldm_last_error(void) { return ldm_host::last_error(); }
long long ldm_launch_count(void) { return ldm_host::launches(); }

int ldm_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return ldm_host::set_error(LDM_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10 || minor != 0)
    return ldm_host::set_error(LDM_ERR_ARCH, "libldmseg_b200 is built for sm_100a only; device is sm_%d%d", major,
                               minor);
  return LDM_OK;
}

}  // extern "C"
