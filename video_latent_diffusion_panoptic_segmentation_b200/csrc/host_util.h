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
// dims 1..rank-1. swizzle128: inner box must span exactly 128 bytes.
int make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, int elem_bytes, bool swizzle128);

inline cudaStream_t as_stream(ldm_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
int num_sms();

}  // namespace ldm_host

#define LDM_REQUIRE(cond, code, ...)                          \
  do {                                                        \
    if (!(cond)) return ldm_host::set_error(code, __VA_ARGS__); \
  } while (0)
