// Shared by the attention kernels (attn_tc.cu: generic head dims; attn40_tc.cu: the 40-wide heads of the 48x156 level):
// launch parameters, the softmax math helpers and the trace hooks of tools/microbench/attn_trace.cu.
#pragma once
#include "common.cuh"
#include "host_util.h"

// Trace harness only (tools/microbench/attn_trace.cu includes the kernels with LDM_ATTN_TRACE defined): lane 0 of every
// warp of CTA (0, 0) records clock64() at the marked points. The product build compiles TRACE() to nothing.
#ifdef LDM_ATTN_TRACE
#define LDM_TRACE_SLOTS 8
#define LDM_TRACE_BLOCKS 96
#define LDM_TRACE_WARPS 20
extern __device__ long long g_attn_trace[LDM_TRACE_WARPS * LDM_TRACE_BLOCKS * LDM_TRACE_SLOTS];
#define TRACE(slot, blk)                                                                                        \
  do {                                                                                                          \
    if (blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x & 31) == 0 && (blk) < LDM_TRACE_BLOCKS)              \
      g_attn_trace[((threadIdx.x >> 5) * LDM_TRACE_BLOCKS + (blk)) * LDM_TRACE_SLOTS + (slot)] = clock64();     \
  } while (0)
#else
#define TRACE(slot, blk) \
  do {                   \
  } while (0)
#endif

namespace ldm_attn {
using namespace ldm;

struct AttnParams {
  int seq, heads, head_dim;
  int kv_seq;  // keys / values per (image, head): seq for self-attention, the context length for cross-attention
  float scale_log2;  // scale * log2(e)
  __nv_bfloat16* out;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 on the FMA pipe (Cody-Waite split + degree-3 minimax of 2^f on [-0.5, 0.5], max rel. error 1.0e-4 -- 40x below
// the bf16 rounding of P): the softmax of the 40-wide heads is bound by the 16/clk/SM MUFU unit, so kPoly of every 8
// exponentials are computed here instead. x <= 8 (lazy rescale) and x may be -inf (masked tail keys).
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;         // 1.5 * 2^23: round(x) lands in the low mantissa bits
  const float f = x - (t - 12582912.0f);  // [-0.5, 0.5]
  float p = fmaf(f, 0.05500891f, 0.24221097f);
  p = fmaf(p, f, 0.69328293f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// Packed fp32 pairs (FFMA2 / FADD2, sm_100): one issue slot for two lanes' worth of work. The softmax loop is bound by
// issue slots shared between the MUFU and the FMA / ALU pipes, so everything that has a packed form uses it.
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t fsub2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// ex2_poly on a packed pair
__device__ __forceinline__ void ex2_poly2(uint64_t x, float& e0, float& e1) {
  float x0, x1;
  unpack2(x, x0, x1);
  const uint64_t xc = pack2(fmaxf(x0, -125.0f), fmaxf(x1, -125.0f));
  const uint64_t magic = pack2(12582912.0f, 12582912.0f);
  const uint64_t t = fadd2(xc, magic);
  const uint64_t f = fsub2(xc, fsub2(t, magic));
  uint64_t q = ffma2(f, pack2(0.05500891f, 0.05500891f), pack2(0.24221097f, 0.24221097f));
  q = ffma2(q, f, pack2(0.69328293f, 0.69328293f));
  q = ffma2(q, f, pack2(1.0f, 1.0f));
  float q0, q1, t0, t1;
  unpack2(q, q0, q1);
  unpack2(t, t0, t1);
  e0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  e1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

}  // namespace ldm_attn

// the 40-wide heads (attn40_tc.cu)
int ldm_launch_attn40(const ldm_attn_desc* d, cudaStream_t s);
