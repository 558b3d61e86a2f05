// HBM-bound normalisation kernels on NHWC bf16 activations: GroupNorm(+SiLU) over a (virtual) channel concat and
// LayerNorm over the channel dim. 16-byte vector loads/stores, warp-shuffle + shared-memory reductions, fp32 math,
// fp64 cross-CTA accumulation of the GroupNorm moments.
#include <stdlib.h>

#include "common.cuh"
#include "host_util.h"

namespace {
using namespace ldm;

// ---------------------------------------------------------------------------------------------------------
// GroupNorm pass 1: per-(image, chunk, group) partial sum and sum of squares, fully deterministic (no atomics).
// grid = (chunks, B); block = (C/8 vectors, ppb pixels). Each thread owns one 8-channel vector position and walks
// pixels with stride ppb*chunks (4 independent 16-byte loads in flight); per-thread sums go to shared memory and one
// thread per group folds its channels x pixel-lanes in a fixed order. partial: f32 [B, chunks, groups, 2].
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void gn_stats_body(float* sh, const __nv_bfloat16* __restrict__ x1,
                                              const __nv_bfloat16* __restrict__ x2, int c1, int c2, int HW, int groups,
                                              float* __restrict__ partial) {
  const int C = c1 + c2;
  const int cpg = C / groups;
  const int b = blockIdx.y;
  const int v = threadIdx.x;
  const int ppb = blockDim.y;
  const int c0 = v * 8;
  const __nv_bfloat16* src;
  int cs, coff;
  if (c0 < c1) {
    src = x1; cs = c1; coff = c0;
  } else {
    src = x2; cs = c2; coff = c0 - c1;
  }
  src += (long long)b * HW * cs + coff;
  float s[8], ss[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.f;
  const int stride = gridDim.x * ppb;
  int pix = blockIdx.x * ppb + threadIdx.y;
  auto acc8 = [&](const uint4& u) {
    const float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y), f2 = unpack_bf16(u.z), f3 = unpack_bf16(u.w);
    const float f[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j] += f[j];
      ss[j] = fmaf(f[j], f[j], ss[j]);
    }
  };
  for (; pix + 3 * stride < HW; pix += 4 * stride) {
    const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(src + (long long)pix * cs));
    const uint4 u1 = __ldg(reinterpret_cast<const uint4*>(src + (long long)(pix + stride) * cs));
    const uint4 u2 = __ldg(reinterpret_cast<const uint4*>(src + (long long)(pix + 2 * stride) * cs));
    const uint4 u3 = __ldg(reinterpret_cast<const uint4*>(src + (long long)(pix + 3 * stride) * cs));
    acc8(u0); acc8(u1); acc8(u2); acc8(u3);
  }
  for (; pix < HW; pix += stride) acc8(__ldg(reinterpret_cast<const uint4*>(src + (long long)pix * cs)));
  float* shs = sh + threadIdx.y * C + c0;
  float* shq = sh + ppb * C + threadIdx.y * C + c0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    shs[j] = s[j];
    shq[j] = ss[j];
  }
  __syncthreads();
  // fold in two steps, both in a fixed order: every channel over the ppb pixel lanes, then every group over its channels
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int nthr = blockDim.x * blockDim.y;
  for (int c = tid; c < 2 * C; c += nthr) {  // c < C: sums, c >= C: sums of squares
    const float* col = sh + (c < C ? c : ppb * C + (c - C));
    float a = 0.f;
    for (int y = 0; y < ppb; ++y) a += col[y * C];
    sh[(c < C ? c : ppb * C + (c - C))] = a;  // row 0 of each half now holds the per-channel totals
  }
  __syncthreads();
  if (tid < groups) {
    float a = 0.f, q = 0.f;
    const float* rs = sh + tid * cpg;
    const float* rq = sh + ppb * C + tid * cpg;
    for (int c = 0; c < cpg; ++c) {
      a += rs[c];
      q += rq[c];
    }
    float* dst = partial + (((long long)b * gridDim.x + blockIdx.x) * groups + tid) * 2;
    dst[0] = a;
    dst[1] = q;
  }
}

__global__ void gn_stats_kernel(const __nv_bfloat16* __restrict__ x1, const __nv_bfloat16* __restrict__ x2, int c1,
                                int c2, int HW, int groups, float* __restrict__ partial) {
  extern __shared__ float sh[];  // [ppb][C] sums, then [ppb][C] sums of squares
  pdl_launch_dependents();
  pdl_wait();
  gn_stats_body(sh, x1, x2, c1, c2, HW, groups, partial);
}

// GroupNorm pass 2: fold the partials (fixed order, fp64), then normalise + affine (+ SiLU). Same thread layout as
// pass 1: a thread owns one 8-channel vector position, so its scale/shift (rstd*gamma, beta - mean*rstd*gamma) are
// computed once and the pixel loop is load -> 8 FMA (+ SiLU) -> store with 4 independent 16-byte loads in flight.
__device__ __forceinline__ void gn_apply_body(float* sh, const __nv_bfloat16* __restrict__ x1,
                                              const __nv_bfloat16* __restrict__ x2, int c1, int c2, int HW, int groups,
                                              const float* partial, int chunks, const float* __restrict__ gamma,
                                              const float* __restrict__ beta, float eps, int silu,
                                              __nv_bfloat16* __restrict__ out) {
  const int C = c1 + c2;
  const int cpg = C / groups;
  const int b = blockIdx.y;
  const int ppb = blockDim.y;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const double inv_n = 1.0 / ((double)cpg * (double)HW);  // (one division, off the dependent chain below)
  // fold the per-chunk partials in fp64: kFoldSlices threads per group take every kFoldSlices-th chunk, then one
  // thread per group adds the slices in order (fixed order -> bit-reproducible). (A warp per group with a shuffle tree
  // was measured slower: three rounds of L2 latency for the 32 groups instead of one.)
  constexpr int kFoldSlices = 8;
  double* shd = reinterpret_cast<double*>(sh + 2 * groups);  // [kFoldSlices][groups][2]
  const int nthr = blockDim.x * blockDim.y;
  for (int t = tid; t < kFoldSlices * groups; t += nthr) {
    const int g = t % groups, sl = t / groups;
    double a = 0.0, q = 0.0;
    const float* pp = partial + ((long long)b * chunks * groups + g) * 2;
    for (int k = sl; k < chunks; k += kFoldSlices) {
      // (L2 loads: in the fused kernel other CTAs wrote these after this kernel started)
      const float2 v = __ldcg(reinterpret_cast<const float2*>(pp + (long long)k * groups * 2));
      a += (double)v.x;
      q += (double)v.y;
    }
    shd[(sl * groups + g) * 2] = a;
    shd[(sl * groups + g) * 2 + 1] = q;
  }
  __syncthreads();
  for (int g = tid; g < groups; g += nthr) {
    double a = 0.0, q = 0.0;
    for (int sl = 0; sl < kFoldSlices; ++sl) {
      a += shd[(sl * groups + g) * 2];
      q += shd[(sl * groups + g) * 2 + 1];
    }
    // (fp64 only for the E[x^2] - mean^2 cancellation: no fp64 division / square root on the critical path)
    const double m = a * inv_n;
    const double var = fma(-m, m, q * inv_n);
    sh[g] = (float)m;
    sh[groups + g] = rsqrtf(fmaxf((float)var, 0.f) + eps);
  }
  __syncthreads();
  const int c0 = threadIdx.x * 8;
  float sc[8], sf[8];
  {
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c0));
    const float4 gb = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
    const float4 ba = __ldg(reinterpret_cast<const float4*>(beta + c0));
    const float4 bb = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
    const float gm[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
    const float bt[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (c0 + j) / cpg;
      sc[j] = sh[groups + g] * gm[j];
      sf[j] = fmaf(-sh[g], sc[j], bt[j]);
    }
  }
  const __nv_bfloat16* src;
  int cs;
  if (c0 < c1) {
    src = x1 + (long long)b * HW * c1 + c0; cs = c1;
  } else {
    src = x2 + (long long)b * HW * c2 + (c0 - c1); cs = c2;
  }
  __nv_bfloat16* dst = out + (long long)b * HW * C + c0;
  auto apply8 = [&](const uint4& u, int pix) {
    const float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y), f2 = unpack_bf16(u.z), f3 = unpack_bf16(u.w);
    float f[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float y = fmaf(f[j], sc[j], sf[j]);
      f[j] = silu ? silu_f(y) : y;
    }
    uint4 o;
    o.x = pack_bf16(f[0], f[1]);
    o.y = pack_bf16(f[2], f[3]);
    o.z = pack_bf16(f[4], f[5]);
    o.w = pack_bf16(f[6], f[7]);
    *reinterpret_cast<uint4*>(dst + (long long)pix * C) = o;
  };
  const int stride = gridDim.x * ppb;
  int pix = blockIdx.x * ppb + threadIdx.y;
  for (; pix + 3 * stride < HW; pix += 4 * stride) {
    const uint4 u0 = ld_nc_v4(src + (long long)pix * cs);
    const uint4 u1 = ld_nc_v4(src + (long long)(pix + stride) * cs);
    const uint4 u2 = ld_nc_v4(src + (long long)(pix + 2 * stride) * cs);
    const uint4 u3 = ld_nc_v4(src + (long long)(pix + 3 * stride) * cs);
    apply8(u0, pix); apply8(u1, pix + stride); apply8(u2, pix + 2 * stride); apply8(u3, pix + 3 * stride);
  }
  for (; pix < HW; pix += stride) apply8(ld_nc_v4(src + (long long)pix * cs), pix);
}

__global__ void gn_apply_kernel(const __nv_bfloat16* __restrict__ x1, const __nv_bfloat16* __restrict__ x2, int c1,
                                int c2, int HW, int groups, const float* __restrict__ partial, int chunks,
                                const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int silu,
                                __nv_bfloat16* __restrict__ out) {
  extern __shared__ float sh[];  // mean[groups], rstd[groups], fp64 fold scratch
  pdl_launch_dependents();
  pdl_wait();
  gn_apply_body(sh, x1, x2, c1, c2, HW, groups, partial, chunks, gamma, beta, eps, silu, out);
}

// Both passes in ONE cooperative launch (all CTAs co-resident: grid <= 2 x SM count): the CTAs of an image meet at a
// counter barrier between the passes, so a GroupNorm costs one launch and the second read of x comes out of L2 while
// it is still warm. Half of the 61 GroupNorms of a UNet forward are small (7-30 MB) and were bound by the two launches,
// not by bandwidth. counters: u32 [2][B], zero before the first use; the kernel leaves them zero.
__global__ void __launch_bounds__(512, 2)
gn_fused_kernel(const __nv_bfloat16* __restrict__ x1, const __nv_bfloat16* __restrict__ x2, int c1, int c2, int HW,
                int groups, float* partial, unsigned int* counters, const float* __restrict__ gamma,
                const float* __restrict__ beta, float eps, int silu, __nv_bfloat16* __restrict__ out) {
  extern __shared__ float sh[];
  const int b = blockIdx.y, B = gridDim.y, chunks = gridDim.x;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  gn_stats_body(sh, x1, x2, c1, c2, HW, groups, partial);
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    atomicAdd(&counters[b], 1u);
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counters + b) : "memory");
      if (seen < (unsigned int)chunks) __nanosleep(64);
    } while (seen < (unsigned int)chunks);
  }
  __syncthreads();
  gn_apply_body(sh, x1, x2, c1, c2, HW, groups, partial, chunks, gamma, beta, eps, silu, out);
  if (tid == 0) {
    // the last CTA of the image to get here has seen everybody pass the barrier: reset for the next launch
    if (atomicAdd(&counters[B + b], 1u) == (unsigned int)chunks - 1u) {
      counters[b] = 0u;
      counters[B + b] = 0u;
    }
  }
}

// GroupNorm(+SiLU) for the shapes whose image fits the register files of a few SMs (everything below the 48x156 level):
// ONE pass over HBM, no cooperative launch, no global barrier. A cluster of kGnClu CTAs owns (image, channel set): CTA r
// takes the r-th pixel slice, keeps its vectors in REGISTERS (kGnRegVec x 16 bytes per thread; a longer slice re-reads
// its tail from L2), the per-group moments of the slices meet through distributed shared memory (one cluster barrier,
// folded in rank order in fp64: deterministic), and the apply pass runs from the registers. The channel sets
// (whole groups, blockIdx.y) only add parallelism. The cooperative kernel above sat on a 15 us floor of launch +
// global counter barrier + second sweep for these shapes (7 - 57 MB at B = 8, all of them at B = 1).
constexpr int kGnCluMax = 8;

__device__ __forceinline__ float2 ld_dsmem_f32x2(const float2* p, uint32_t rank) {
  uint32_t a;
  float2 v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(smem_u32(p)), "r"(rank));
  asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
  return v;
}

template <int kGnRegVec>
__global__ void __launch_bounds__(512, 1)
gn_cluster_kernel(const __nv_bfloat16* __restrict__ x1, const __nv_bfloat16* __restrict__ x2, int c1, int c2, int HW,
                  int groups, int slab_c, double inv_n, const float* __restrict__ gamma, const float* __restrict__ beta,
                  float eps, int silu, __nv_bfloat16* __restrict__ out) {
  extern __shared__ float sh[];  // [ppb][slab_c] sums | [ppb][slab_c] sums of squares | slice moments | mean, rstd
  pdl_launch_dependents();
  pdl_wait();
  const int C = c1 + c2;
  const int cpg = C / groups;
  const int gps = slab_c / cpg;  // groups of this channel set
  const uint32_t rank = cluster_ctarank();
  const int nclu = (int)gridDim.x;  // the cluster spans the x dimension: pixel slices of one (image, channel set)
  const int set = blockIdx.y, b = blockIdx.z;
  const int v = threadIdx.x, y = threadIdx.y, vps = blockDim.x, ppb = blockDim.y;
  const int tid = y * vps + v, nthr = vps * ppb;
  const int cl = v * 8;                 // channel inside the set
  const int c0 = set * slab_c + cl;     // channel of the virtual concat
  const __nv_bfloat16* src;
  int cs;
  if (c0 < c1) {
    src = x1 + (long long)b * HW * c1 + c0; cs = c1;
  } else {
    src = x2 + (long long)b * HW * c2 + (c0 - c1); cs = c2;
  }
  // (scale / shift operands: in flight while the statistics are taken)
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c0)), gb = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
  const float4 ba = __ldg(reinterpret_cast<const float4*>(beta + c0)), bb = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
  const int per = (HW + nclu - 1) / nclu;
  const int p0 = (int)rank * per + y;
  const int p1 = min(HW, ((int)rank + 1) * per);

  uint4 u[kGnRegVec];
#pragma unroll
  for (int k = 0; k < kGnRegVec; ++k) {
    const int pix = p0 + k * ppb;
    u[k] = pix < p1 ? ld_nc_v4(src + pix * cs) : make_uint4(0u, 0u, 0u, 0u);
  }
  float s[8], ss[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.f;
  auto acc8 = [&](const uint4& w) {
    const float2 f0 = unpack_bf16(w.x), f1 = unpack_bf16(w.y), f2 = unpack_bf16(w.z), f3 = unpack_bf16(w.w);
    const float f[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j] += f[j];
      ss[j] = fmaf(f[j], f[j], ss[j]);
    }
  };
#pragma unroll
  for (int k = 0; k < kGnRegVec; ++k) acc8(u[k]);  // (zeros past the slice add nothing)
  {
    int pix = p0 + kGnRegVec * ppb;  // a slice longer than the register cache: its tail is read again by the apply pass
    for (; pix + 3 * ppb < p1; pix += 4 * ppb) {
      const uint4 w0 = ld_nc_v4(src + pix * cs), w1 = ld_nc_v4(src + (pix + ppb) * cs);
      const uint4 w2 = ld_nc_v4(src + (pix + 2 * ppb) * cs), w3 = ld_nc_v4(src + (pix + 3 * ppb) * cs);
      acc8(w0); acc8(w1); acc8(w2); acc8(w3);
    }
    for (; pix < p1; pix += ppb) acc8(ld_nc_v4(src + pix * cs));
  }

  float* shs = sh + y * slab_c + cl;
  float* shq = sh + (ppb + y) * slab_c + cl;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    shs[j] = s[j];
    shq[j] = ss[j];
  }
  __syncthreads();
  // Fixed-order folds, short dependency chains (a serial fp64 fold per group cost more than the loads):
  // every channel over the pixel lanes (fp32, <= 64 terms) ...
  for (int c = tid; c < 2 * slab_c; c += nthr) {
    float* col = sh + (c < slab_c ? c : ppb * slab_c + (c - slab_c));
    float a = 0.f;
    for (int yy = 0; yy < ppb; ++yy) a += col[yy * slab_c];
    col[0] = a;
  }
  __syncthreads();
  // ... every group over its channels: one warp per group, lanes stride the channels, xor tree ...
  float2* part = reinterpret_cast<float2*>(sh + 2 * ppb * slab_c);  // [gps] (sum, sum of squares) of this slice
  float* stat = reinterpret_cast<float*>(part + gps);                // mean[gps], rstd[gps]
  {
    const int warp = tid >> 5, lane = tid & 31, nfull = nthr >> 5;  // (a ragged last warp stays out of the shuffles)
    if (warp < nfull) {
      for (int g = warp; g < gps; g += nfull) {
        float a = 0.f, q = 0.f;
        for (int c = lane; c < cpg; c += 32) {
          a += sh[g * cpg + c];
          q += sh[ppb * slab_c + g * cpg + c];
        }
        a = warp_sum(a);
        q = warp_sum(q);
        if (lane == 0) part[g] = make_float2(a, q);
      }
    }
  }
  cluster_sync_all();  // every slice's moments are in its CTA's shared memory
  // ... and the slices in rank order (fp64 for the E[x^2] - mean^2 cancellation only)
  if (tid < gps) {
    float2 pr[kGnCluMax];
#pragma unroll
    for (int r = 0; r < kGnCluMax; ++r) pr[r] = r < nclu ? ld_dsmem_f32x2(part + tid, (uint32_t)r) : make_float2(0.f, 0.f);
    double a = 0.0, q = 0.0;
#pragma unroll
    for (int r = 0; r < kGnCluMax; ++r) {
      a += (double)pr[r].x;
      q += (double)pr[r].y;
    }
    const double m = a * inv_n;
    double var = fma(-m, m, q * inv_n);
    stat[tid] = (float)m;
    stat[gps + tid] = rsqrtf(fmaxf((float)var, 0.f) + eps);
  }
  __syncthreads();
  // no CTA may exit (its shared memory would go away) before every peer has read its moments: arrive now, wait last
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");

  float sc[8], sf[8];
  {
    const float gm[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
    const float bt[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (cl + j) / cpg;
      sc[j] = stat[gps + g] * gm[j];
      sf[j] = fmaf(-stat[g], sc[j], bt[j]);
    }
  }
  __nv_bfloat16* dst = out + (long long)b * HW * C + c0;
  auto apply8 = [&](const uint4& w, int pix) {
    const float2 f0 = unpack_bf16(w.x), f1 = unpack_bf16(w.y), f2 = unpack_bf16(w.z), f3 = unpack_bf16(w.w);
    float f[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float t = fmaf(f[j], sc[j], sf[j]);
      f[j] = silu ? silu_f(t) : t;
    }
    uint4 o;
    o.x = pack_bf16(f[0], f[1]);
    o.y = pack_bf16(f[2], f[3]);
    o.z = pack_bf16(f[4], f[5]);
    o.w = pack_bf16(f[6], f[7]);
    *reinterpret_cast<uint4*>(dst + pix * C) = o;
  };
#pragma unroll
  for (int k = 0; k < kGnRegVec; ++k) {
    const int pix = p0 + k * ppb;
    if (pix < p1) apply8(u[k], pix);
  }
  {
    int pix = p0 + kGnRegVec * ppb;
    for (; pix + 3 * ppb < p1; pix += 4 * ppb) {
      const uint4 w0 = ld_nc_v4(src + pix * cs), w1 = ld_nc_v4(src + (pix + ppb) * cs);
      const uint4 w2 = ld_nc_v4(src + (pix + 2 * ppb) * cs), w3 = ld_nc_v4(src + (pix + 3 * ppb) * cs);
      apply8(w0, pix); apply8(w1, pix + ppb); apply8(w2, pix + 2 * ppb); apply8(w3, pix + 3 * ppb);
    }
    for (; pix < p1; pix += ppb) apply8(ld_nc_v4(src + pix * cs), pix);
  }
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// LayerNorm: one warp per row, row cached in registers (C <= 32*8*kMaxVec).
constexpr int kMaxVec = 6;  // up to C = 1536
__global__ void layernorm_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, __nv_bfloat16* __restrict__ out, int rows, int C,
                                 float eps) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nvec = C / 8;
  for (long long row = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < rows;
       row += (long long)gridDim.x * warps_per_block) {
    const __nv_bfloat16* src = x + row * C;
    float f[kMaxVec][8];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxVec; ++k) {
      const int v = lane + k * 32;
      if (v < nvec) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + v * 8));
        const float2 a = unpack_bf16(u.x), b2 = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
        f[k][0] = a.x; f[k][1] = a.y; f[k][2] = b2.x; f[k][3] = b2.y;
        f[k][4] = c.x; f[k][5] = c.y; f[k][6] = d.x; f[k][7] = d.y;
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += f[k][j];
      }
    }
    const float mean = warp_sum(sum) / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxVec; ++k) {
      const int v = lane + k * 32;
      if (v < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = f[k][j] - mean;
          sq += d * d;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) / (float)C + eps);
#pragma unroll
    for (int k = 0; k < kMaxVec; ++k) {
      const int v = lane + k * 32;
      if (v < nvec) {
        const int c0 = v * 8;
        const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c0));
        const float4 gb = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
        const float4 ba = __ldg(reinterpret_cast<const float4*>(beta + c0));
        const float4 bb = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
        const float gm[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
        const float bt[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (f[k][j] - mean) * rstd * gm[j] + bt[j];
        uint4 u;
        u.x = pack_bf16(o[0], o[1]);
        u.y = pack_bf16(o[2], o[3]);
        u.z = pack_bf16(o[4], o[5]);
        u.w = pack_bf16(o[6], o[7]);
        *reinterpret_cast<uint4*>(out + row * C + c0) = u;
      }
    }
  }
}

// LayerNorm for C = LPR * 40 (320 / 640 / 1280): LPR lanes share a row, five 16-byte vectors per lane, 32/LPR rows per
// warp -- every lane busy, five independent loads in flight per lane, sub-warp xor-shuffle reductions.
template <int LPR>
__global__ void layernorm_rows_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, __nv_bfloat16* __restrict__ out, int rows, float eps) {
  constexpr int C = LPR * 40;
  constexpr int RPW = 32 / LPR;  // rows per warp
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, sl = lane % LPR;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long row = warp_global * RPW + sub;
  if (row >= rows) return;  // whole sub-groups leave together; the shuffles below stay inside a sub-group
  const __nv_bfloat16* src = x + row * C;
  uint4 u[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) u[k] = ld_nc_v4(src + (sl + k * LPR) * 8);
  float f[5][8];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const float2 a = unpack_bf16(u[k].x), b2 = unpack_bf16(u[k].y), c = unpack_bf16(u[k].z), d = unpack_bf16(u[k].w);
    f[k][0] = a.x; f[k][1] = a.y; f[k][2] = b2.x; f[k][3] = b2.y;
    f[k][4] = c.x; f[k][5] = c.y; f[k][6] = d.x; f[k][7] = d.y;
#pragma unroll
    for (int j = 0; j < 8; ++j) sum += f[k][j];
  }
  const unsigned mask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(mask, sum, o);
  const float mean = sum / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < 5; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = f[k][j] - mean;
      sq += d * d;
    }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(mask, sq, o);
  const float rstd = rsqrtf(sq / (float)C + eps);
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int c0 = (sl + k * LPR) * 8;
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c0));
    const float4 gb = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
    const float4 ba = __ldg(reinterpret_cast<const float4*>(beta + c0));
    const float4 bb = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
    const float gm[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
    const float bt[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (f[k][j] - mean) * rstd * gm[j] + bt[j];
    uint4 w;
    w.x = pack_bf16(o[0], o[1]);
    w.y = pack_bf16(o[2], o[3]);
    w.z = pack_bf16(o[4], o[5]);
    w.w = pack_bf16(o[6], o[7]);
    *reinterpret_cast<uint4*>(out + row * C + c0) = w;
  }
}

template <int LPR>
static int launch_layernorm_rows(const void* x, const float* gamma, const float* beta, void* out, int rows, float eps,
                                 cudaStream_t s) {
  constexpr int RPW = 32 / LPR;
  const int wpb = 8;
  const long long warps = ((long long)rows + RPW - 1) / RPW;
  const int grid = (int)((warps + wpb - 1) / wpb);
  cudaError_t le = ldm_host::launch_pdl(layernorm_rows_kernel<LPR>, dim3(grid), dim3(wpb * 32), (size_t)0, s, 1,
                                        reinterpret_cast<const __nv_bfloat16*>(x), gamma, beta,
                                        reinterpret_cast<__nv_bfloat16*>(out), rows, eps);
  if (le != cudaSuccess) return ldm_host::set_error(LDM_ERR_CUDA, "layernorm_rows_kernel launch: %s", cudaGetErrorString(le));
  ldm_host::count_launch();
  return ldm_host::check_launch("layernorm_rows_kernel");
}

// Row softmax of the unfused attention (RGB VAE mid block, one 512-wide head): one CTA per row, three sweeps over the
// row (maximum, sum of exponentials, normalised bf16 result); sweeps two and three hit L1 / L2. fp32 throughout, exp2 of
// log2(e)-scaled differences like the fused kernels.
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ p, int cols, long long ld_s, long long ld_p,
                    float scale) {
  __shared__ float red[8];
  const float* row = s + (long long)blockIdx.x * ld_s;
  __nv_bfloat16* dst = p + (long long)blockIdx.x * ld_p;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const bool vec = (cols % 4 == 0) && (ld_s % 4 == 0) && (ld_p % 4 == 0);
  float m = -INFINITY;
  if (vec) {
    for (int c = tid * 4; c < cols; c += 1024) {
      const float4 v = *reinterpret_cast<const float4*>(row + c);
      m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
    }
  } else {
    for (int c = tid; c < cols; c += 256) m = fmaxf(m, row[c]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[wid] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  // scale > 0: max(scale * s) = scale * max(s)
  const float k = scale * 1.4426950408889634f;
  const float mk = m * k;
  float sum = 0.f;
  if (vec) {
    for (int c = tid * 4; c < cols; c += 1024) {
      const float4 v = *reinterpret_cast<const float4*>(row + c);
      sum += exp2f(fmaf(v.x, k, -mk)) + exp2f(fmaf(v.y, k, -mk)) + exp2f(fmaf(v.z, k, -mk)) + exp2f(fmaf(v.w, k, -mk));
    }
  } else {
    for (int c = tid; c < cols; c += 256) sum += exp2f(fmaf(row[c], k, -mk));
  }
  sum = warp_sum(sum);
  if (lane == 0) red[wid] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += red[i];
  const float inv = 1.f / sum;
  if (vec) {
    for (int c = tid * 4; c < cols; c += 1024) {
      const float4 v = *reinterpret_cast<const float4*>(row + c);
      uint2 o;
      o.x = pack_bf16(exp2f(fmaf(v.x, k, -mk)) * inv, exp2f(fmaf(v.y, k, -mk)) * inv);
      o.y = pack_bf16(exp2f(fmaf(v.z, k, -mk)) * inv, exp2f(fmaf(v.w, k, -mk)) * inv);
      *reinterpret_cast<uint2*>(dst + c) = o;
    }
  } else {
    for (int c = tid; c < cols; c += 256) dst[c] = __float2bfloat16_rn(exp2f(fmaf(row[c], k, -mk)) * inv);
  }
}

}  // namespace

static int gn_chunks(int B, int HW, int ppb) {
  int chunks = (2 * ldm_host::num_sms() + B - 1) / B;
  const int max_chunks = (HW + ppb - 1) / ppb;
  if (chunks > max_chunks) chunks = max_chunks;
  return chunks < 1 ? 1 : chunks;
}

static size_t gn_partial_bytes(int B, int groups) {
  return sizeof(float) * 2 * (size_t)groups * (size_t)B * (size_t)(2 * ldm_host::num_sms());
}

extern "C" size_t ldm_groupnorm_scratch_bytes(int32_t B, int32_t groups) {
  // per-chunk partial moments (upper bound of chunks over all shapes: 2 * SM count per image) + the barrier counters
  return gn_partial_bytes(B, groups) + sizeof(unsigned int) * 2 * (size_t)B;
}

extern "C" int ldm_groupnorm_silu(const ldm_groupnorm_desc* d, ldm_stream_t stream) {
  using namespace ldm_host;
  LDM_REQUIRE(d && d->x1 && d->gamma && d->beta && d->out && d->stats, LDM_ERR_BAD_ARG, "ldm_groupnorm_silu: null arg");
  const int c2 = d->x2 ? d->c2 : 0;
  const int C = d->c1 + c2;
  LDM_REQUIRE(d->B > 0 && d->HW > 0 && d->groups > 0 && C % d->groups == 0, LDM_ERR_BAD_SHAPE,
              "ldm_groupnorm_silu: bad shape B=%d HW=%d C=%d groups=%d", d->B, d->HW, C, d->groups);
  LDM_REQUIRE(d->c1 % 8 == 0 && c2 % 8 == 0 && C / 8 <= 1024, LDM_ERR_ALIGNMENT,
              "ldm_groupnorm_silu: channels must be multiples of 8 and <= 8192");
  cudaStream_t s = as_stream(stream);
  const int vpp = C / 8;
  int ppb = 512 / vpp;
  if (ppb < 1) ppb = 1;
  if (ppb > d->HW) ppb = d->HW;
  const int chunks = gn_chunks(d->B, d->HW, ppb);
  // Small images: the one-pass cluster kernel. Choose (pixel slices per cluster, channel sets of whole groups and whole
  // 16-byte vectors): as many CTAs as fit one wave, the smallest cluster among equals (no exchange at all when a CTA
  // can own whole groups of an image), at most max_vec vectors per thread (beyond that the two-sweep kernel wins).
  {
    const int max_vec = diag_env("LDM_GN_CLUSTER_MAXVEC", 18);  // (0 in a diagnostic build: never this path)
    const int cpg = C / d->groups;
    int best_clu = 0, best_sets = 0, best_ctas = 0;
    for (int clu = 1; clu <= kGnCluMax; clu *= 2) {
      // co-resident clusters at one CTA per SM (measured: 15 clusters of 8 on B200; the GPCs do not all hold 16 free SMs)
      const int max_clusters = clu == 1 ? num_sms() : (clu == 8 ? 14 : (clu == 4 ? 32 : 68));
      for (int n_sets = d->groups; n_sets >= 1; n_sets /= 2) {
        if (d->groups % n_sets) continue;
        const int slab_c = C / n_sets;
        if (slab_c % 8 || slab_c % cpg || slab_c / 8 > 512) continue;
        const int ctas = d->B * clu * n_sets;
        if (ctas > num_sms() || d->B * n_sets > max_clusters) continue;
        const int per = (d->HW + clu - 1) / clu;
        int ppbc = 512 / (slab_c / 8);
        if (ppbc > per) ppbc = per;
        if ((slab_c / 8) * ppbc < 32 || (per + ppbc - 1) / ppbc > max_vec) continue;
        if (ctas > best_ctas) {
          best_ctas = ctas; best_clu = clu; best_sets = n_sets;
        }
      }
    }
    if (best_ctas > 0 && (long long)d->HW * (d->c1 > c2 ? d->c1 : c2) < (1LL << 31)) {
      const int clu = best_clu, slab_c = C / best_sets, cvps = slab_c / 8, gps = slab_c / cpg;
      const int per = (d->HW + clu - 1) / clu;
      const int cppb = 512 / cvps < per ? 512 / cvps : per;
      const int vec = (per + cppb - 1) / cppb;
      const size_t shc = sizeof(float) * 2 * (size_t)cppb * slab_c + sizeof(float) * 4 * gps;
      LDM_REQUIRE(shc <= 100 * 1024, LDM_ERR_BAD_SHAPE, "ldm_groupnorm_silu: cluster path smem %zu", shc);
      auto go = [&](auto kern) -> int {
        cudaError_t ae = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (ae != cudaSuccess) return set_error(LDM_ERR_CUDA, "cudaFuncSetAttribute(gn_cluster): %s", cudaGetErrorString(ae));
        cudaError_t le = launch_pdl(kern, dim3(clu, best_sets, d->B), dim3(cvps, cppb), shc, s, clu,
                                    reinterpret_cast<const __nv_bfloat16*>(d->x1), reinterpret_cast<const __nv_bfloat16*>(d->x2),
                                    d->c1, c2, d->HW, d->groups, slab_c, 1.0 / ((double)cpg * (double)d->HW), d->gamma,
                                    d->beta, d->eps, d->silu, reinterpret_cast<__nv_bfloat16*>(d->out));
        if (le != cudaSuccess) return set_error(LDM_ERR_CUDA, "gn_cluster_kernel launch: %s", cudaGetErrorString(le));
        count_launch();
        return check_launch("gn_cluster_kernel");
      };
      if (vec <= 4) return go(gn_cluster_kernel<4>);
      if (vec <= 8) return go(gn_cluster_kernel<8>);
      return go(gn_cluster_kernel<12>);
    }
  }
  float* partial = reinterpret_cast<float*>(d->stats);
  const size_t sh1 = sizeof(float) * 2 * (size_t)ppb * C;
  cudaFuncSetAttribute(gn_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);  // (per device)
  LDM_REQUIRE(sh1 <= 96 * 1024, LDM_ERR_BAD_SHAPE, "ldm_groupnorm_silu: C=%d too large", C);
  const size_t sh2 = sizeof(float) * 2 * d->groups + sizeof(double) * 2 * 8 * d->groups;
  const int fused = diag_env("LDM_GN_FUSED", 1);  // 0 (diagnostic builds): the two-launch path (A/B timing)
  if (fused && (long long)chunks * d->B <= 2LL * num_sms() && vpp * ppb <= 512) {
    cudaFuncSetAttribute(gn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    unsigned int* counters = reinterpret_cast<unsigned int*>(reinterpret_cast<uint8_t*>(d->stats) +
                                                             gn_partial_bytes(d->B, d->groups));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(chunks, d->B);
    cfg.blockDim = dim3(vpp, ppb);
    cfg.dynamicSmemBytes = sh1 > sh2 ? sh1 : sh2;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;  // the runtime checks that the whole grid is co-resident
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, gn_fused_kernel, reinterpret_cast<const __nv_bfloat16*>(d->x1),
                                        reinterpret_cast<const __nv_bfloat16*>(d->x2), d->c1, c2, d->HW, d->groups, partial,
                                        counters, d->gamma, d->beta, d->eps, d->silu,
                                        reinterpret_cast<__nv_bfloat16*>(d->out));
    if (le != cudaSuccess) return set_error(LDM_ERR_CUDA, "gn_fused_kernel launch: %s", cudaGetErrorString(le));
    count_launch();
    return check_launch("gn_fused_kernel");
  }
  cudaError_t le = launch_pdl(gn_stats_kernel, dim3(chunks, d->B), dim3(vpp, ppb), sh1, s, 1,
                              reinterpret_cast<const __nv_bfloat16*>(d->x1), reinterpret_cast<const __nv_bfloat16*>(d->x2),
                              d->c1, c2, d->HW, d->groups, partial);
  if (le != cudaSuccess) return set_error(LDM_ERR_CUDA, "gn_stats_kernel launch: %s", cudaGetErrorString(le));
  count_launch();
  int rc = check_launch("gn_stats_kernel");
  if (rc) return rc;
  le = launch_pdl(gn_apply_kernel, dim3(chunks, d->B), dim3(vpp, ppb),
                  sh2, s, 1,
                  reinterpret_cast<const __nv_bfloat16*>(d->x1), reinterpret_cast<const __nv_bfloat16*>(d->x2), d->c1, c2,
                  d->HW, d->groups, (const float*)partial, chunks, d->gamma, d->beta, d->eps, d->silu,
                  reinterpret_cast<__nv_bfloat16*>(d->out));
  if (le != cudaSuccess) return set_error(LDM_ERR_CUDA, "gn_apply_kernel launch: %s", cudaGetErrorString(le));
  count_launch();
  return check_launch("gn_apply_kernel");
}

extern "C" int ldm_layernorm(const void* x, const float* gamma, const float* beta, void* out, int32_t rows, int32_t C,
                             float eps, ldm_stream_t stream) {
  using namespace ldm_host;
  LDM_REQUIRE(x && gamma && beta && out, LDM_ERR_BAD_ARG, "ldm_layernorm: null arg");
  LDM_REQUIRE(rows > 0 && C > 0 && C % 8 == 0 && C <= 32 * 8 * kMaxVec, LDM_ERR_BAD_SHAPE,
              "ldm_layernorm: rows=%d C=%d unsupported (C %% 8 == 0, C <= %d)", rows, C, 32 * 8 * kMaxVec);
  if (C == 320) return launch_layernorm_rows<8>(x, gamma, beta, out, rows, eps, as_stream(stream));
  if (C == 640) return launch_layernorm_rows<16>(x, gamma, beta, out, rows, eps, as_stream(stream));
  if (C == 1280) return launch_layernorm_rows<32>(x, gamma, beta, out, rows, eps, as_stream(stream));
  const int wpb = 8;
  int grid = (rows + wpb - 1) / wpb;
  const int cap = num_sms() * 16;
  if (grid > cap) grid = cap;
  layernorm_kernel<<<grid, wpb * 32, 0, as_stream(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x), gamma, beta,
                                                             reinterpret_cast<__nv_bfloat16*>(out), rows, C, eps);
  count_launch();
  return check_launch("layernorm_kernel");
}

extern "C" int ldm_softmax_rows(const float* s, void* p, int64_t rows, int32_t cols, int64_t ld_s, int64_t ld_p,
                                float scale, ldm_stream_t stream) {
  using namespace ldm_host;
  LDM_REQUIRE(s && p && rows > 0 && rows < (1LL << 31) && cols > 0 && ld_s >= cols && ld_p >= cols && scale > 0.f,
              LDM_ERR_BAD_ARG, "ldm_softmax_rows: bad arg (rows=%lld cols=%d)", (long long)rows, cols);
  softmax_rows_kernel<<<(unsigned int)rows, 256, 0, as_stream(stream)>>>(s, reinterpret_cast<__nv_bfloat16*>(p), cols,
                                                                        (long long)ld_s, (long long)ld_p, scale);
  count_launch();
  return check_launch("softmax_rows_kernel");
}
