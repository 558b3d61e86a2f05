// Trace harness for the attention kernels: links csrc/attn_tc.cu + attn40_tc.cu built with LDM_ATTN_TRACE, runs the L0 shape of configs[1]
// (B = 8 images x 8 heads, 7 488 tokens, d = 40) once and prints, for CTA (0, 0), the clock64() stamps lane 0 of every
// warp took per KV block: softmax warps -- 0 before / 1 after the s_full wait, 2 scores in registers, 3 exp pass done,
// 4 arrived on p_full; MMA warp (warp 1) -- 2+g S issue, 4+g PV issue.
// Build (from the repo root):
//   C=video_latent_diffusion_panoptic_segmentation_b200/csrc
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -rdc=true -DLDM_ATTN_TRACE \
//     -Iinclude -I$C -o tools/microbench/attn_trace_bin tools/microbench/attn_trace.cu $C/attn_tc.cu $C/attn40_tc.cu \
//     $C/host_util.cu -lcuda
// Run: LDM_ATTN40=0|1 tools/microbench/attn_trace_bin [first_block]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "attn_common.cuh"

__device__ long long g_attn_trace[LDM_TRACE_WARPS * LDM_TRACE_BLOCKS * LDM_TRACE_SLOTS];
static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return (uint16_t)((u + 0x7fff + ((u >> 16) & 1)) >> 16);
}

int main(int argc, char** argv) {
  const int B = 8, heads = 8, seq = 7488, d = 40, dpad = 64, seq_pad = 7488, vt_rows = 48;
  const int BH = B * heads;
  const size_t nq = (size_t)BH * seq * dpad, nv = (size_t)BH * vt_rows * seq_pad, no = (size_t)B * seq * heads * d;
  std::vector<uint16_t> hq(nq, 0), hk(nq, 0), hv(nv, 0);
  srand(1);
  const bool fold = getenv("LDM_TRACE_FOLD") && atoi(getenv("LDM_TRACE_FOLD"));  // scores in log2 units (scale = ln 2)
  auto rnd = [] { return ((rand() & 0xffff) / 65536.0f - 0.5f) * 4.0f; };
  for (size_t r = 0; r < (size_t)BH * seq; ++r)
    for (int c = 0; c < d; ++c) {
      hq[r * dpad + c] = f2bf(rnd() * (fold ? 0.15811388f * 1.44269504f : 1.0f));
      hk[r * dpad + c] = f2bf(rnd());
    }
  for (int bh = 0; bh < BH; ++bh)
    for (int c = 0; c <= d; ++c)
      for (int t = 0; t < seq; ++t) hv[((size_t)bh * vt_rows + c) * seq_pad + t] = c == d ? f2bf(1.0f) : f2bf(rnd());
  void *q, *k, *vt, *out;
  cudaMalloc(&q, nq * 2);
  cudaMalloc(&k, nq * 2);
  cudaMalloc(&vt, nv * 2);
  cudaMalloc(&out, no * 2);
  cudaMemcpy(q, hq.data(), nq * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(k, hk.data(), nq * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(vt, hv.data(), nv * 2, cudaMemcpyHostToDevice);
  ldm_attn_desc desc = {q, k, vt, out, B, heads, seq, d, dpad, seq_pad, vt_rows, fold ? 0.69314718f : 0.15811388f};
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int it = 0; it < 3; ++it) {
    cudaEventRecord(e0);
    int rc = ldm_flash_attn_fwd(&desc, nullptr);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (rc || e != cudaSuccess) {
      printf("error rc=%d %s %s\n", rc, ldm_last_error(), cudaGetErrorString(e));
      return 1;
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("launch %d: %.3f ms\n", it, ms);
  }
  static long long h[LDM_TRACE_WARPS * LDM_TRACE_BLOCKS * LDM_TRACE_SLOTS];
  cudaMemcpyFromSymbol(h, g_attn_trace, sizeof(h));
  auto at = [&](int w, int j, int s) { return h[(w * LDM_TRACE_BLOCKS + j) * LDM_TRACE_SLOTS + s]; };
  const long long t0 = at(4, 0, 0);
  const int j0 = argc > 1 ? atoi(argv[1]) : 20, j1 = j0 + 6;
  printf("times in cycles relative to warp 4's first stamp; blocks %d..%d\n", j0, j1 - 1);
  for (int j = j0; j < j1; ++j) {
    printf("blk %2d  MMA: S(g0,%d)@%lld S(g1,%d)@%lld PV(g0)@%lld PV(g1)@%lld\n", j, j, at(1, j, 2) - t0, j, at(1, j, 3) - t0,
           at(1, j, 4) - t0, at(1, j, 5) - t0);
    for (int w : {4, 8, 12}) {
      printf("   warp %2d (g%d): wait_s %6lld  s_full %6lld (+%4lld)  ld_done +%4lld  exp_done +%4lld  p_full +%4lld   block period %lld\n",
             w, (w - 4) / 4, at(w, j, 0) - t0, at(w, j, 1) - t0, at(w, j, 1) - at(w, j, 0), at(w, j, 2) - at(w, j, 1),
             at(w, j, 3) - at(w, j, 2), at(w, j, 4) - at(w, j, 3), at(w, j + 1, 0) - at(w, j, 0));
    }
  }
  // averages over blocks 8..50 for all softmax warps
  const char* e40 = getenv("LDM_ATTN40");
  const int nwarps = (e40 && atoi(e40) == 0) ? 12 : 20;
  for (int w = 4; w < nwarps; ++w) {
    double a[5] = {0, 0, 0, 0, 0};
    int n = 0;
    for (int j = 8; j < 50; ++j, ++n) {
      a[0] += at(w, j, 1) - at(w, j, 0);
      a[1] += at(w, j, 2) - at(w, j, 1);
      a[2] += at(w, j, 3) - at(w, j, 2);
      a[3] += at(w, j, 4) - at(w, j, 3);
      a[4] += at(w, j + 1, 0) - at(w, j, 0);
    }
    double b5 = 0, b6 = 0, b7 = 0;
    for (int j = 8; j < 50; ++j) {
      b5 += at(w, j, 5) - at(w, j, 2);
      b6 += at(w, j, 6) - at(w, j, 5);
      b7 += at(w, j, 7) - at(w, j, 6);
    }
    printf("warp %2d avg: s_full wait %5.0f  tmem ld %5.0f  exp pass %5.0f  st wait %5.0f  period %5.0f | fold: to prefill %5.0f  o_full wait %5.0f  fill %5.0f\n", w, a[0] / n,
           a[1] / n, a[2] / n, a[3] / n, a[4] / n, b5 / n, b6 / n, b7 / n);
  }
  return 0;
}
