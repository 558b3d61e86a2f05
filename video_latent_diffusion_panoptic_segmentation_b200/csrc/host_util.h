// Host-side helpers shared by the C-ABI translation units: error reporting, launch accounting,
// TMA tensor-map encoding through the driver entry point (no link-time libcuda dependency).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ldmseg_b200.h"

namespace ldm_host {

int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);  // cudaGetLastError -> status

// rank-`rank` bf16 (or other 2-byte / 4-byte) tiled tensor map. dims/box innermost first; strides in bytes for
// dims 1..rank-1. swizzle128: inner box must span exactly 128 bytes. elem_strides (optional, per dim): s > 1 makes the
// TMA unit touch every s-th element of that dimension, for loads and for stores; box[i] is then the extent in traversal
// space (s * elements delivered) -- measured, tools/microbench/tma_stride_test.cu.
int make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, int elem_bytes, bool swizzle128, const uint32_t* elem_strides = nullptr);

inline cudaStream_t as_stream(ldm_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Diagnostic switches (A/B timing by the scripts under tools/): compiled in only with -DLDM_DIAG
// (`LDM_BUILD_DIAG=1 python -m video_latent_diffusion_panoptic_segmentation_b200.build --force`). The product library reads
// NO environment variable; its only process-wide state are immutable per-device caches (SM count).
#ifdef LDM_DIAG
int diag_env(const char* name, int dflt);              // atoi(getenv(name)) or dflt
bool diag_env_has(const char* name, const char* word);  // strstr(getenv(name), word)
#else
inline int diag_env(const char*, int dflt) { return dflt; }
inline bool diag_env_has(const char*, const char*) { return false; }
#endif

// LDM_PDL=1 (diagnostic build) turns programmatic dependent launch on. Measured on B200 inside the CUDA graph of the UNet
// plan: 8.58 fps with it, 8.63 without at 8 frames per batch, +0.3 % at one frame (launch gaps are already hidden by
// the graph; the fat GEMM / attention CTAs cannot co-reside with their predecessor's anyway), so it is off.
inline bool pdl_enabled() { return diag_env("LDM_PDL", 0) != 0; }

// Launch with programmatic stream serialisation (the kernel must call pdl_wait() before its first global access) and,
// optionally, a thread-block cluster of `cluster` CTAs.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster,
                       Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
int num_sms();

}  // namespace ldm_host

#define LDM_REQUIRE(cond, code, ...)                          \
  do {                                                        \
    if (!(cond)) return ldm_host::set_error(code, __VA_ARGS__); \
  } while (0)
