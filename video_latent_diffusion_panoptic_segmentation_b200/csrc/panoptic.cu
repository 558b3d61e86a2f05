// Integer tail of the sampler path: logits -> panoptic ids (fused bilinear up-sampling, argmax, softmax-max
// threshold, per-class pixel counts), the count/overlap merge filter, the bit codec, 4-connected component
// labelling with scipy numbering, and the joint (gt, pred) id histogram that vpq_eval / the Cityscapes PQ
// evaluator reduce to. All HBM-bound byte/integer work: coalesced 16-byte loads, warp-aggregated atomics.
#include "common.cuh"
#include "host_util.h"

namespace {
using namespace ldm;

// torch's CPU bilinear kernel (UpSampleKernel.cpp, Interpolate<2,...>::eval) evaluates
//   out = (x00*wx0 + x01*wx1) * wy0 + (x10*wx0 + x11*wx1) * wy1
// with "output = t*w0; output += t1*w1" which the vectorised build contracts to one multiply + one FMA.
__device__ __forceinline__ float lerp2(float a, float wa, float b, float wb) { return fmaf(b, wb, __fmul_rn(a, wa)); }

struct Axis {
  int i0, i1;
  float w0, w1;
};
// align_corners=False source index with an explicit scale_factor `up` (scale = 1/up), torch area_pixel_compute_source_index
__device__ __forceinline__ Axis axis_for(int dst, int in_size, int up) {
  Axis a;
  if (up == 1) {
    a.i0 = dst; a.i1 = dst < in_size - 1 ? dst + 1 : dst; a.w0 = 1.f; a.w1 = 0.f;
    return a;
  }
  const float scale = 1.0f / (float)up;
  float src = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
  if (src < 0.f) src = 0.f;
  a.i0 = (int)src;
  if (a.i0 > in_size - 1) a.i0 = in_size - 1;
  a.i1 = a.i0 < in_size - 1 ? a.i0 + 1 : a.i0;
  a.w1 = __fsub_rn(src, (float)a.i0);
  a.w0 = __fsub_rn(1.f, a.w1);
  return a;
}

constexpr int kMaxVecPerLane = 4;  // C <= 512

// grid = (chunks, B); one warp per output pixel (looping); lane owns channels {4*(lane+32k) .. +3}.
__global__ void logits_to_ids_kernel(const float* __restrict__ logits, int32_t* __restrict__ ids,
                                     int32_t* __restrict__ counts, int h, int w, int C, int up, float mask_th,
                                     int ignore_label) {
  extern __shared__ int hist[];  // [C] argmax-label histogram of this CTA
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int H = h * up, W = w * up;
  const int nvec = C / 4;
  for (int i = threadIdx.x; i < C; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  int over[kMaxVecPerLane * 4];
#pragma unroll
  for (int i = 0; i < kMaxVecPerLane * 4; ++i) over[i] = 0;
  const float* img = logits + (long long)b * h * w * C;
  const long long npix = (long long)H * W;
  for (long long pix = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); pix < npix;
       pix += (long long)gridDim.x * wpb) {
    const int oy = (int)(pix / W), ox = (int)(pix - (long long)oy * W);
    const Axis ay = axis_for(oy, h, up), ax = axis_for(ox, w, up);
    const float* p00 = img + ((long long)ay.i0 * w + ax.i0) * C;
    const float* p01 = img + ((long long)ay.i0 * w + ax.i1) * C;
    const float* p10 = img + ((long long)ay.i1 * w + ax.i0) * C;
    const float* p11 = img + ((long long)ay.i1 * w + ax.i1) * C;
    float val[kMaxVecPerLane * 4];
    float best = -INFINITY;
    int besti = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < kMaxVecPerLane; ++k) {
      const int v = lane + 32 * k;
      if (v < nvec) {
        float4 a = __ldg(reinterpret_cast<const float4*>(p00) + v);
        if (up != 1) {
          const float4 bq = __ldg(reinterpret_cast<const float4*>(p01) + v);
          const float4 c = __ldg(reinterpret_cast<const float4*>(p10) + v);
          const float4 d = __ldg(reinterpret_cast<const float4*>(p11) + v);
          a.x = lerp2(lerp2(a.x, ax.w0, bq.x, ax.w1), ay.w0, lerp2(c.x, ax.w0, d.x, ax.w1), ay.w1);
          a.y = lerp2(lerp2(a.y, ax.w0, bq.y, ax.w1), ay.w0, lerp2(c.y, ax.w0, d.y, ax.w1), ay.w1);
          a.z = lerp2(lerp2(a.z, ax.w0, bq.z, ax.w1), ay.w0, lerp2(c.z, ax.w0, d.z, ax.w1), ay.w1);
          a.w = lerp2(lerp2(a.w, ax.w0, bq.w, ax.w1), ay.w0, lerp2(c.w, ax.w0, d.w, ax.w1), ay.w1);
        }
        val[k * 4 + 0] = a.x; val[k * 4 + 1] = a.y; val[k * 4 + 2] = a.z; val[k * 4 + 3] = a.w;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (val[k * 4 + j] > best) {  // strict: first index wins (torch.argmax)
            best = val[k * 4 + j];
            besti = v * 4 + j;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) val[k * 4 + j] = -INFINITY;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ov > best || (ov == best && oi < besti)) {
        best = ov;
        besti = oi;
      }
    }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxVecPerLane; ++k) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float x = val[k * 4 + j];
        if (lane + 32 * k < nvec) {
          sum += expf(x - best);
          // sigmoid(x) >= mask_th, evaluated like torch: 1 / (1 + exp(-x))
          const float sg = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x)));
          over[k * 4 + j] += (sg >= mask_th) ? 1 : 0;
        }
      }
    }
    sum = warp_sum(sum);
    if (lane == 0) {
      const float pmax = __fdiv_rn(1.f, sum);
      int id = besti;
      if (pmax < mask_th) id = ignore_label;
      ids[(long long)b * npix + pix] = id;
      if (id >= 0 && id < C) atomicAdd(&hist[id], 1);
    }
  }
  __syncthreads();
  int32_t* cb = counts + (long long)b * 2 * C;
  for (int i = threadIdx.x; i < C; i += blockDim.x)
    if (hist[i]) atomicAdd(&cb[i], hist[i]);
  // per-class sigmoid-threshold areas: reduce the per-lane counters across the CTA's warps through smem
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) hist[i] = 0;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kMaxVecPerLane; ++k) {
    const int v = lane + 32 * k;
    if (v < nvec) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (over[k * 4 + j]) atomicAdd(&hist[v * 4 + j], over[k * 4 + j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x)
    if (hist[i]) atomicAdd(&cb[C + i], hist[i]);
}

// Materialising bilinear x`up`: NHWC f32 [B,h,w,C] -> NCHW f32 [B,C,H,W] (parity checks of the interpolation).
__global__ void bilinear_up_nchw_kernel(const float* __restrict__ logits, float* __restrict__ out, int B, int h, int w,
                                        int C, int up) {
  const int H = h * up, W = w * up;
  const long long total = (long long)B * C * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % W);
    long long t = i / W;
    const int oy = (int)(t % H);
    t /= H;
    const int c = (int)(t % C);
    const int b = (int)(t / C);
    const Axis ay = axis_for(oy, h, up), ax = axis_for(ox, w, up);
    const float* img = logits + (long long)b * h * w * C + c;
    const float x00 = img[((long long)ay.i0 * w + ax.i0) * C], x01 = img[((long long)ay.i0 * w + ax.i1) * C];
    const float x10 = img[((long long)ay.i1 * w + ax.i0) * C], x11 = img[((long long)ay.i1 * w + ax.i1) * C];
    out[i] = (up == 1) ? x00
                       : lerp2(lerp2(x00, ax.w0, x01, ax.w1), ay.w0, lerp2(x10, ax.w0, x11, ax.w1), ay.w1);
  }
}

// General bilinear resize (F.interpolate(size=..., mode="bilinear", align_corners=False): scale = in/out) of a crop
// window of NHWC f32 logits, NHWC out. One thread per (output pixel, 4 channels). Used for the resizes that are the
// identity on the benchmark shapes: to the RGB size, crop_padding + resize to meta.im_size
// (trainers_ldm_cond.py:1264-1284).
__device__ __forceinline__ Axis axis_scaled(int dst, int in_size, float scale) {
  Axis a;
  float src = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
  if (src < 0.f) src = 0.f;
  a.i0 = (int)src;
  if (a.i0 > in_size - 1) a.i0 = in_size - 1;
  a.i1 = a.i0 < in_size - 1 ? a.i0 + 1 : a.i0;
  a.w1 = __fsub_rn(src, (float)a.i0);
  a.w0 = __fsub_rn(1.f, a.w1);
  return a;
}

__global__ void resize_bilinear_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int h, int w,
                                            int C, int y0, int x0, int ch, int cw, int oh, int ow) {
  const int vec = C / 4;
  const float sy = (float)ch / (float)oh, sx = (float)cw / (float)ow;
  const long long total = (long long)B * oh * ow * vec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vec);
    long long t = i / vec;
    const int ox = (int)(t % ow);
    t /= ow;
    const int oy = (int)(t % oh);
    const int b = (int)(t / oh);
    const Axis ay = axis_scaled(oy, ch, sy), ax = axis_scaled(ox, cw, sx);
    const float* img = in + ((long long)b * h * w) * C + 4 * v;
    auto px = [&](int yy, int xx) {
      return __ldg(reinterpret_cast<const float4*>(img + ((long long)(y0 + yy) * w + (x0 + xx)) * C));
    };
    const float4 a = px(ay.i0, ax.i0), bq = px(ay.i0, ax.i1), c = px(ay.i1, ax.i0), d = px(ay.i1, ax.i1);
    float4 o;
    o.x = lerp2(lerp2(a.x, ax.w0, bq.x, ax.w1), ay.w0, lerp2(c.x, ax.w0, d.x, ax.w1), ay.w1);
    o.y = lerp2(lerp2(a.y, ax.w0, bq.y, ax.w1), ay.w0, lerp2(c.y, ax.w0, d.y, ax.w1), ay.w1);
    o.z = lerp2(lerp2(a.z, ax.w0, bq.z, ax.w1), ay.w0, lerp2(c.z, ax.w0, d.z, ax.w1), ay.w1);
    o.w = lerp2(lerp2(a.w, ax.w0, bq.w, ax.w1), ay.w0, lerp2(c.w, ax.w0, d.w, ax.w1), ay.w1);
    *reinterpret_cast<float4*>(out + (((long long)b * oh + oy) * ow + ox) * C + 4 * v) = o;
  }
}

// The same resize on planar maps [P, h, w] (NCHW images / latents of encode_inputs, trainers_ldm_cond.py:368-393).
__global__ void resize_bilinear_planar_kernel(const float* __restrict__ in, float* __restrict__ out, int P, int h, int w,
                                              int oh, int ow) {
  const float sy = (float)h / (float)oh, sx = (float)w / (float)ow;
  const long long total = (long long)P * oh * ow;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % ow);
    long long t = i / ow;
    const int oy = (int)(t % oh);
    const long long pl = t / oh;
    const Axis ay = axis_scaled(oy, h, sy), ax = axis_scaled(ox, w, sx);
    const float* img = in + pl * h * w;
    const float a = __ldg(img + (long long)ay.i0 * w + ax.i0), b = __ldg(img + (long long)ay.i0 * w + ax.i1);
    const float c = __ldg(img + (long long)ay.i1 * w + ax.i0), d = __ldg(img + (long long)ay.i1 * w + ax.i1);
    out[i] = lerp2(lerp2(a, ax.w0, b, ax.w1), ay.w0, lerp2(c, ax.w0, d, ax.w1), ay.w1);
  }
}

// Merge filter (trainers_ldm_cond.py:1307-1325). counts = [B][2][C] (argmax area, sigmoid>=th area).
__global__ void segment_filter_kernel(const int32_t* __restrict__ ids, const int32_t* __restrict__ counts,
                                      int32_t* __restrict__ cleaned, long long hw, int C, int count_th,
                                      double overlap_th, int ignore_label) {
  extern __shared__ int keep[];
  const int b = blockIdx.y;
  const int32_t* cb = counts + (long long)b * 2 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int cnt = cb[c], ov = cb[C + c];
    bool k = cnt >= count_th && c != ignore_label && cnt > 0;
    // numpy: int / int -> float64 (x/0 -> inf, which is not < overlap_th)
    if (k && ov > 0 && ((double)cnt / (double)ov) < overlap_th) k = false;
    keep[c] = k ? 1 : 0;
  }
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x) {
    const int id = ids[(long long)b * hw + i];
    cleaned[(long long)b * hw + i] = (id >= 0 && id < C && keep[id]) ? id : -1;
  }
}

__global__ void decode_bitmap_kernel(const float* __restrict__ x, int32_t* __restrict__ ids, int nbits, long long hw,
                                     int quirk31) {
  const int b = blockIdx.y;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x) {
    int id = 0;
    for (int k = 0; k < nbits; ++k) id |= (x[((long long)b * nbits + k) * hw + i] > 0.f) ? (1 << k) : 0;
    if (quirk31 && id == 31) id = 0;
    ids[(long long)b * hw + i] = id;
  }
}

__global__ void encode_bitmap_kernel(const int32_t* __restrict__ ids, float* __restrict__ x, int nbits, long long hw,
                                     int ignore_label, float fill) {
  const int b = blockIdx.y;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x) {
    const int id = ids[(long long)b * hw + i];
    for (int k = 0; k < nbits; ++k)
      x[((long long)b * nbits + k) * hw + i] = (id == ignore_label) ? fill : (float)((id >> k) & 1);
  }
}

// ---------------------------------------------------------------- connected components (4-connectivity)
__device__ __forceinline__ int cc_find(const int32_t* L, int i) {
  int p = L[i];
  while (p != i) {
    i = p;
    p = L[i];
  }
  return i;
}
__device__ __forceinline__ void cc_union(int32_t* L, int a, int b) {
  while (true) {
    a = cc_find(L, a);
    b = cc_find(L, b);
    if (a == b) return;
    if (a < b) {
      const int t = a; a = b; b = t;
    }
    // a > b: hook the larger root under the smaller so the root is the first pixel in raster order
    const int old = atomicMin(&L[a], b);
    if (old == a) return;
    a = old;
  }
}
__global__ void ccl_init_kernel(const int32_t* __restrict__ sem, int target, int32_t* __restrict__ L, long long n,
                                long long hw) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    L[i] = (sem[i] == target) ? (int)(i % hw) : -1;
}
__global__ void ccl_merge_kernel(int32_t* __restrict__ Lall, int H, int W) {
  const long long hw = (long long)H * W;
  int32_t* L = Lall + (long long)blockIdx.y * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x) {
    if (L[i] < 0) continue;
    const int x = (int)(i % W);
    if (x > 0 && L[i - 1] >= 0) cc_union(L, (int)i, (int)i - 1);
    if (i >= W && L[i - W] >= 0) cc_union(L, (int)i, (int)(i - W));
  }
}
__global__ void ccl_compress_count_kernel(int32_t* __restrict__ Lall, int32_t* __restrict__ block_counts, long long hw) {
  // blockDim = 1024 pixels per block; counts roots per block
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  int32_t* L = Lall + (long long)blockIdx.y * hw;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  bool root = false;
  if (i < hw && L[i] >= 0) {
    const int r = cc_find(L, (int)i);
    L[i] = r;
    root = (r == (int)i);
  }
  const unsigned m = __ballot_sync(0xffffffffu, root);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(&cnt, __popc(m));
  __syncthreads();
  if (threadIdx.x == 0) block_counts[(long long)blockIdx.y * gridDim.x + blockIdx.x] = cnt;
}
__global__ void ccl_scan_blocks_kernel(int32_t* __restrict__ block_counts, int32_t* __restrict__ ncomp, int nblocks) {
  // one CTA per image: exclusive scan of block_counts in place (sequential over chunks of blockDim)
  __shared__ int sh[1024];
  __shared__ int carry;
  int32_t* bc = block_counts + (long long)blockIdx.x * nblocks;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = i < nblocks ? bc[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < blockDim.x; o <<= 1) {
      const int t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    const int incl = sh[threadIdx.x];
    const int c = carry;
    __syncthreads();
    if (i < nblocks) bc[i] = c + incl - v;
    if (threadIdx.x == blockDim.x - 1) carry = c + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) ncomp[blockIdx.x] = carry;
}
__global__ void ccl_rank_kernel(const int32_t* __restrict__ Lall, const int32_t* __restrict__ block_offsets,
                                int32_t* __restrict__ rank_all, long long hw) {
  // rank[root pixel] = 1-based component number in raster order of the root
  __shared__ int warp_tot[32];
  const int32_t* L = Lall + (long long)blockIdx.y * hw;
  int32_t* rank = rank_all + (long long)blockIdx.y * hw;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool root = (i < hw) && (L[i] == (int)i);
  const unsigned m = __ballot_sync(0xffffffffu, root);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) warp_tot[wid] = __popc(m);
  __syncthreads();
  int prefix = block_offsets[(long long)blockIdx.y * gridDim.x + blockIdx.x];
  for (int k = 0; k < wid; ++k) prefix += warp_tot[k];
  prefix += __popc(m & ((1u << lane) - 1));
  if (root) rank[i] = prefix + 1;
}
__global__ void ccl_final_kernel(const int32_t* __restrict__ L, const int32_t* __restrict__ rank,
                                 int32_t* __restrict__ labels, long long n, long long hw) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = L[i];
    labels[i] = r < 0 ? 0 : rank[(i / hw) * hw + r];
  }
}

// ---------------------------------------------------------------- joint id histogram (open-addressing hash)
constexpr unsigned long long kEmptyKey = 0x8000000000000000ull;
__global__ void hash_clear_kernel(unsigned long long* keys, int32_t* counts, int cap, int32_t* overflow) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) {
    keys[i] = kEmptyKey;
    counts[i] = 0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *overflow = 0;
}
__global__ void joint_hist_kernel(const int32_t* __restrict__ a, const int32_t* __restrict__ b, long long n,
                                  unsigned long long* __restrict__ keys, int32_t* __restrict__ counts, int cap,
                                  int32_t* __restrict__ overflow) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long nround = ((n + stride - 1) / stride) * stride;  // keep warps converged for match_any
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
    const bool valid = i < n;
    const unsigned long long key =
        valid ? (((unsigned long long)(uint32_t)a[i] << 32) | (unsigned long long)(uint32_t)b[i]) : kEmptyKey;
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const int leader = __ffs(peers) - 1;
    if (valid && (int)(threadIdx.x & 31) == leader) {
      const int add = __popc(peers);
      unsigned long long hsh = key * 0x9E3779B97F4A7C15ull;
      int slot = (int)((hsh >> 32) & (unsigned)(cap - 1));
      int probes = 0;
      while (true) {
        const unsigned long long prev = atomicCAS(&keys[slot], kEmptyKey, key);
        if (prev == kEmptyKey || prev == key) {
          atomicAdd(&counts[slot], add);
          break;
        }
        slot = (slot + 1) & (cap - 1);
        if (++probes >= cap) {
          *overflow = 1;
          break;
        }
      }
    }
  }
}


// ---------------------------------------------------------------- batched panoptic maps of the Cityscapes evaluator
// One labelling pass for ALL thing classes of ALL images of a batch, prediction and ground truth together
// (cityscapes_pap_eval.py:66-110). A pixel takes part when slot[map][label] >= 0 (its label is a thing class; `map` is
// 0 for the B prediction maps, 1 for the B ground-truth maps); two 4-neighbours are connected when they carry the SAME
// label, so the components are exactly those of the per-class binary masks; scipy numbers the components of one class
// mask in raster order of their first pixel, i.e. a root's number is its rank among the roots of its own class.
constexpr int kMaxThings = 32;
__device__ __forceinline__ int thing_slot(const int8_t* __restrict__ slot, int label) {
  return (label >= 0 && label < 256) ? (int)slot[label] : -1;
}
__global__ void ccl_multi_init_kernel(const int32_t* __restrict__ pred, const int32_t* __restrict__ gt,
                                      const int8_t* __restrict__ slots, int32_t* __restrict__ L, long long hw, int B,
                                      int pred_void, int ignore_label) {
  const long long n = 2ll * B * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int img = (int)(i / hw);
    const bool is_gt = img >= B;
    int lab = is_gt ? gt[i - (long long)B * hw] : pred[i];
    if (!is_gt && lab == pred_void) lab = ignore_label;   // pred_seg[pred_seg == -1] = ignore_label (:66)
    L[i] = thing_slot(slots + (is_gt ? 256 : 0), lab) >= 0 ? (int)(i % hw) : -1;
  }
}
__device__ __forceinline__ int map_label(const int32_t* __restrict__ pred, const int32_t* __restrict__ gt, int img,
                                         long long px, long long hw, int B, int pred_void, int ignore_label) {
  if (img >= B) return gt[(long long)(img - B) * hw + px];
  const int lab = pred[(long long)img * hw + px];
  return lab == pred_void ? ignore_label : lab;
}
__global__ void ccl_multi_merge_kernel(int32_t* __restrict__ Lall, const int32_t* __restrict__ pred,
                                       const int32_t* __restrict__ gt, int H, int W, int B, int pred_void,
                                       int ignore_label) {
  const long long hw = (long long)H * W;
  const int img = blockIdx.y;
  int32_t* L = Lall + (long long)img * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x) {
    if (L[i] < 0) continue;
    const int lab = map_label(pred, gt, img, i, hw, B, pred_void, ignore_label);
    const int x = (int)(i % W);
    if (x > 0 && L[i - 1] >= 0 && map_label(pred, gt, img, i - 1, hw, B, pred_void, ignore_label) == lab)
      cc_union(L, (int)i, (int)i - 1);
    if (i >= W && L[i - W] >= 0 && map_label(pred, gt, img, i - W, hw, B, pred_void, ignore_label) == lab)
      cc_union(L, (int)i, (int)(i - W));
  }
}
// block_counts layout: [2B images][nslots][nblocks]
__global__ void ccl_multi_count_kernel(int32_t* __restrict__ Lall, const int32_t* __restrict__ pred,
                                       const int32_t* __restrict__ gt, const int8_t* __restrict__ slots,
                                       int32_t* __restrict__ block_counts, long long hw, int B, int nslots,
                                       int pred_void, int ignore_label) {
  __shared__ int cnt[kMaxThings];
  if (threadIdx.x < kMaxThings) cnt[threadIdx.x] = 0;
  __syncthreads();
  const int img = blockIdx.y;
  int32_t* L = Lall + (long long)img * hw;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < hw && L[i] >= 0) {
    const int r = cc_find(L, (int)i);
    L[i] = r;
    if (r == (int)i) {
      const int sl = thing_slot(slots + (img >= B ? 256 : 0), map_label(pred, gt, img, i, hw, B, pred_void, ignore_label));
      atomicAdd(&cnt[sl], 1);
    }
  }
  __syncthreads();
  if (threadIdx.x < nslots)
    block_counts[((long long)img * nslots + threadIdx.x) * gridDim.x + blockIdx.x] = cnt[threadIdx.x];
}
__global__ void ccl_multi_rank_kernel(const int32_t* __restrict__ Lall, const int32_t* __restrict__ pred,
                                      const int32_t* __restrict__ gt, const int8_t* __restrict__ slots,
                                      const int32_t* __restrict__ block_offsets, int32_t* __restrict__ rank_all,
                                      long long hw, int B, int nslots, int pred_void, int ignore_label) {
  // rank[root pixel] = 1-based number of the component among the components of its class, in raster order of the roots
  __shared__ int warp_tot[32][kMaxThings];
  const int img = blockIdx.y;
  const int32_t* L = Lall + (long long)img * hw;
  int32_t* rank = rank_all + (long long)img * hw;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int k = lane; k < kMaxThings; k += 32) warp_tot[wid][k] = 0;
  __syncwarp();
  const bool root = (i < hw) && (L[i] == (int)i);
  int sl = -1;
  if (root) sl = thing_slot(slots + (img >= B ? 256 : 0), map_label(pred, gt, img, i, hw, B, pred_void, ignore_label));
  const unsigned same = __match_any_sync(0xffffffffu, sl);     // lanes with the same slot (-1: not a root)
  if (root && lane == __ffs(same) - 1) warp_tot[wid][sl] = __popc(same);
  __syncthreads();
  if (root) {
    int prefix = block_offsets[((long long)img * nslots + sl) * gridDim.x + blockIdx.x];
    for (int k = 0; k < wid; ++k) prefix += warp_tot[k][sl];
    prefix += __popc(same & ((1u << lane) - 1));
    rank[i] = prefix + 1;
  }
}
// pan maps of :76-110 from the labelled roots, both maps of an image in one pass
__global__ void city_pan_final_kernel(const int32_t* __restrict__ Lall, const int32_t* __restrict__ rank_all,
                                      const int32_t* __restrict__ pred, const int32_t* __restrict__ gt,
                                      int32_t* __restrict__ pred_pan, int32_t* __restrict__ gt_pan, long long hw, int B,
                                      int pred_void, int ignore_label, int max_ins) {
  const long long n = (long long)B * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long img = i / hw;
    int p = pred[i];
    if (p == pred_void) p = ignore_label;
    const int g = gt[i];
    const int lp = Lall[i], lg = Lall[n + i];
    // ground truth: thing pixels -> sem * max_ins + component, ignore -> -1 (:46-49)
    int gp = lg >= 0 ? g * max_ins + rank_all[n + img * hw + lg] : g;
    if (g == ignore_label) gp = -1;
    // prediction: zeros, stuff labels copied, thing pixels -> label * max_ins + component (:90-105); ignore -> -1 (:108-110)
    int pp = lp >= 0 ? p * max_ins + rank_all[img * hw + lp] : p;
    if (p == ignore_label) pp = 0;
    if (g == ignore_label || p == ignore_label) pp = -1;
    gt_pan[i] = gp;
    pred_pan[i] = pp;
  }
}
// joint histograms of n_tables (possibly overlapping) windows of two id streams: table t counts the pairs
// (a[t * stride + j], b[t * stride + j]), j < n_per_table, in its own hash table keys/counts[t * cap ...]
__global__ void joint_hist_batch_kernel(const int32_t* __restrict__ a, const int32_t* __restrict__ b,
                                        long long n_per_table, long long tstride, unsigned long long* __restrict__ keys_all,
                                        int32_t* __restrict__ counts_all, int cap, int32_t* __restrict__ overflow) {
  const int t = blockIdx.y;
  const int32_t* at = a + (long long)t * tstride;
  const int32_t* bt = b + (long long)t * tstride;
  unsigned long long* keys = keys_all + (long long)t * cap;
  int32_t* counts = counts_all + (long long)t * cap;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long nround = ((n_per_table + stride - 1) / stride) * stride;  // keep warps converged for match_any
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
    const bool valid = i < n_per_table;
    const unsigned long long key =
        valid ? (((unsigned long long)(uint32_t)at[i] << 32) | (unsigned long long)(uint32_t)bt[i]) : kEmptyKey;
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const int leader = __ffs(peers) - 1;
    if (valid && (int)(threadIdx.x & 31) == leader) {
      const int add = __popc(peers);
      unsigned long long hsh = key * 0x9E3779B97F4A7C15ull;
      int slot = (int)((hsh >> 32) & (unsigned)(cap - 1));
      int probes = 0;
      while (true) {
        const unsigned long long prev = atomicCAS(&keys[slot], kEmptyKey, key);
        if (prev == kEmptyKey || prev == key) {
          atomicAdd(&counts[slot], add);
          break;
        }
        slot = (slot + 1) & (cap - 1);
        if (++probes >= cap) {
          overflow[t] = 1;
          break;
        }
      }
    }
  }
}
__global__ void hash_clear_batch_kernel(unsigned long long* keys, int32_t* counts, long long n, int32_t* overflow,
                                        int n_tables) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    keys[i] = kEmptyKey;
    counts[i] = 0;
    if (i < n_tables) overflow[i] = 0;
  }
}

// ---------------------------------------------------------------- small id-map helpers of the PQ evaluators
__global__ void pan_insert_kernel(const int32_t* __restrict__ sem, const int32_t* __restrict__ labels, int target,
                                  int max_ins, int32_t* __restrict__ pan, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (sem[i] == target) pan[i] = target * max_ins + labels[i];
}
__global__ void pan_combine_kernel(const int32_t* __restrict__ cat, const int32_t* __restrict__ ins, int max_ins,
                                   int32_t* __restrict__ pan, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    pan[i] = cat[i] * max_ins + ins[i];
}
__global__ void id_mask_kernel(int32_t* x, const int32_t* a, int va,
                               const int32_t* b, int vb, int fill, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (a[i] == va || (b != nullptr && b[i] == vb)) x[i] = fill;
}

int grid1d(long long work, int threads, int mult = 16) {
  long long g = (work + threads - 1) / threads;
  const long long cap = (long long)ldm_host::num_sms() * mult;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

extern "C" int ldm_logits_to_ids(const float* logits, int32_t* ids, int32_t* counts, int32_t B, int32_t h, int32_t w,
                                 int32_t C, int32_t up, float mask_th, int32_t ignore_label, ldm_stream_t stream) {
  using namespace ldm_host;
  LDM_REQUIRE(logits && ids && counts, LDM_ERR_BAD_ARG, "ldm_logits_to_ids: null arg");
  LDM_REQUIRE(B > 0 && h > 0 && w > 0 && C > 0 && C % 4 == 0 && C <= 128 * kMaxVecPerLane && (up == 1 || up == 2),
              LDM_ERR_BAD_SHAPE, "ldm_logits_to_ids: B=%d h=%d w=%d C=%d up=%d", B, h, w, C, up);
  cudaStream_t s = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(int32_t) * 2 * C * B, s);
  if (e != cudaSuccess) return set_error(LDM_ERR_CUDA, "memset counts: %s", cudaGetErrorString(e));
  int chunks = (num_sms() * 4 + B - 1) / B;
  logits_to_ids_kernel<<<dim3(chunks, B), 256, sizeof(int) * C, s>>>(logits, ids, counts, h, w, C, up, mask_th,
                                                                     ignore_label);
  count_launch();
  return check_launch("logits_to_ids_kernel");
}

extern "C" int ldm_bilinear_up_nchw(const float* logits, float* out, int32_t B, int32_t h, int32_t w, int32_t C,
                                    int32_t up, ldm_stream_t stream) {
  using namespace ldm_host;
  LDM_REQUIRE(logits && out && B > 0 && h > 0 && w > 0 && C > 0 && (up == 1 || up == 2), LDM_ERR_BAD_ARG,
              "ldm_bilinear_up_nchw: bad arg");
  const long long total = (long long)B * C * h * up * w * up;
  bilinear_up_nchw_kernel<<<grid1d(total, 256), 256, 0, as_stream(stream)>>>(logits, out, B, h, w, C, up);
  count_launch();
  return check_launch("bilinear_up_nchw_kernel");
}

extern "C" int ldm_resize_bilinear_nhwc(const float* in, float* out, int32_t B, int32_t h, int32_t w, int32_t C,
                                        int32_t y0, int32_t x0, int32_t ch, int32_t cw, int32_t oh, int32_t ow,
                                        ldm_stream_t stream) {
  using namespace ldm_host;
  LDM_REQUIRE(in && out && B > 0 && h > 0 && w > 0 && C > 0 && C % 4 == 0 && oh > 0 && ow > 0, LDM_ERR_BAD_ARG,
              "ldm_resize_bilinear_nhwc: bad arg");
  LDM_REQUIRE(y0 >= 0 && x0 >= 0 && ch > 0 && cw > 0 && y0 + ch <= h && x0 + cw <= w, LDM_ERR_BAD_SHAPE,
              "ldm_resize_bilinear_nhwc: crop window (%d,%d,%d,%d) outside %dx%d", y0, x0, ch, cw, h, w);
  const long long total = (long long)B * oh * ow * (C / 4);
  resize_bilinear_nhwc_kernel<<<grid1d(total, 256), 256, 0, as_stream(stream)>>>(in, out, B, h, w, C, y0, x0, ch, cw, oh,
                                                                                ow);
  count_launch();
  return check_launch("resize_bilinear_nhwc_kernel");
}

extern "C" int ldm_resize_bilinear_planar(const float* in, float* out, int32_t P, int32_t h, int32_t w, int32_t oh,
                                          int32_t ow, ldm_stream_t stream) {
  using namespace ldm_host;
  LDM_REQUIRE(in && out && P > 0 && h > 0 && w > 0 && oh > 0 && ow > 0, LDM_ERR_BAD_ARG,
              "ldm_resize_bilinear_planar: bad arg");
  const long long total = (long long)P * oh * ow;
  resize_bilinear_planar_kernel<<<grid1d(total, 256), 256, 0, as_stream(stream)>>>(in, out, P, h, w, oh, ow);
  count_launch();
  return check_launch("resize_bilinear_planar_kernel");
}

extern "C" int ldm_segment_filter(const int32_t* ids, const int32_t* counts, int32_t* cleaned, int32_t B, int64_t hw,
                                  int32_t C, int32_t count_th, double overlap_th, int32_t ignore_label,
                                  ldm_stream_t stream) {
  using namespace ldm_host;
  LDM_REQUIRE(ids && counts && cleaned && B > 0 && hw > 0 && C > 0 && C <= 8192, LDM_ERR_BAD_ARG,
              "ldm_segment_filter: bad arg");
  int chunks = grid1d(hw, 256, 4);
  segment_filter_kernel<<<dim3(chunks, B), 256, sizeof(int) * C, as_stream(stream)>>>(ids, counts, cleaned, hw, C,
                                                                                      count_th, overlap_th, ignore_label);
  count_launch();
  return check_launch("segment_filter_kernel");
}

extern "C" int ldm_decode_bitmap(const float* x, int32_t* ids, int32_t B, int32_t nbits, int64_t hw, int32_t quirk31,
                                 ldm_stream_t stream) {
  using namespace ldm_host;
  LDM_REQUIRE(x && ids && B > 0 && hw > 0 && nbits > 0 && nbits <= 24, LDM_ERR_BAD_ARG,
              "ldm_decode_bitmap: bad arg (nbits must be in [1,24])");
  decode_bitmap_kernel<<<dim3(grid1d(hw, 256, 4), B), 256, 0, as_stream(stream)>>>(x, ids, nbits, hw, quirk31);
  count_launch();
  return check_launch("decode_bitmap_kernel");
}

extern "C" int ldm_encode_bitmap(const int32_t* ids, float* x, int32_t B, int32_t nbits, int64_t hw,
                                 int32_t ignore_label, float fill, ldm_stream_t stream) {
  using namespace ldm_host;
  LDM_REQUIRE(x && ids && B > 0 && hw > 0 && nbits > 0 && nbits <= 31, LDM_ERR_BAD_ARG, "ldm_encode_bitmap: bad arg");
  encode_bitmap_kernel<<<dim3(grid1d(hw, 256, 4), B), 256, 0, as_stream(stream)>>>(ids, x, nbits, hw, ignore_label, fill);
  count_launch();
  return check_launch("encode_bitmap_kernel");
}

extern "C" size_t ldm_ccl_scratch_bytes(int32_t B, int32_t H, int32_t W) {
  const long long hw = (long long)H * W;
  const long long nblocks = (hw + 1023) / 1024;
  return (size_t)(sizeof(int32_t) * (2 * (long long)B * hw + (long long)B * nblocks));
}

extern "C" int ldm_ccl_label4(const int32_t* sem, int32_t target, int32_t* labels, int32_t* ncomp, int32_t* scratch,
                              int32_t B, int32_t H, int32_t W, ldm_stream_t stream) {
  using namespace ldm_host;
  LDM_REQUIRE(sem && labels && ncomp && scratch && B > 0 && H > 0 && W > 0, LDM_ERR_BAD_ARG, "ldm_ccl_label4: bad arg");
  const long long hw = (long long)H * W;
  LDM_REQUIRE(hw < (1ll << 31), LDM_ERR_BAD_SHAPE, "ldm_ccl_label4: image too large");
  const long long n = hw * B;
  const int nblocks = (int)((hw + 1023) / 1024);
  int32_t* L = scratch;
  int32_t* rank = scratch + n;
  int32_t* bcnt = scratch + 2 * n;
  cudaStream_t s = as_stream(stream);
  ccl_init_kernel<<<grid1d(n, 256), 256, 0, s>>>(sem, target, L, n, hw);
  ccl_merge_kernel<<<dim3(grid1d(hw, 256, 8), B), 256, 0, s>>>(L, H, W);
  ccl_compress_count_kernel<<<dim3(nblocks, B), 1024, 0, s>>>(L, bcnt, hw);
  ccl_scan_blocks_kernel<<<B, 1024, 0, s>>>(bcnt, ncomp, nblocks);
  ccl_rank_kernel<<<dim3(nblocks, B), 1024, 0, s>>>(L, bcnt, rank, hw);
  ccl_final_kernel<<<grid1d(n, 256), 256, 0, s>>>(L, rank, labels, n, hw);
  count_launch(6);
  return check_launch("ccl kernels");
}

extern "C" int ldm_joint_hist(const int32_t* a, const int32_t* b, int64_t n, unsigned long long* keys, int32_t* counts,
                              int32_t capacity, int32_t* overflow, ldm_stream_t stream) {
  using namespace ldm_host;
  LDM_REQUIRE(a && b && keys && counts && overflow && n > 0, LDM_ERR_BAD_ARG, "ldm_joint_hist: bad arg");
  LDM_REQUIRE(capacity >= 64 && (capacity & (capacity - 1)) == 0, LDM_ERR_BAD_SHAPE,
              "ldm_joint_hist: capacity must be a power of two >= 64");
  cudaStream_t s = as_stream(stream);
  hash_clear_kernel<<<grid1d(capacity, 256, 2), 256, 0, s>>>(keys, counts, capacity, overflow);
  joint_hist_kernel<<<grid1d(n, 256, 8), 256, 0, s>>>(a, b, n, keys, counts, capacity, overflow);
  count_launch(2);
  return check_launch("joint_hist_kernel");
}

extern "C" size_t ldm_city_pan_scratch_bytes(int32_t B, int32_t H, int32_t W, int32_t n_things) {
  const long long hw = (long long)H * W;
  const long long nblocks = (hw + 1023) / 1024;
  return (size_t)(sizeof(int32_t) * (4ll * B * hw + 2ll * B * n_things * (nblocks + 1)));
}

extern "C" int ldm_city_pan_maps(const int32_t* pred_seg, const int32_t* gt_sem, int32_t* pred_pan, int32_t* gt_pan,
                                 const int8_t* thing_slots, int32_t n_things, int32_t ignore_label, int32_t max_ins,
                                 int32_t* scratch, int32_t B, int32_t H, int32_t W, ldm_stream_t stream) {
  using namespace ldm_host;
  LDM_REQUIRE(pred_seg && gt_sem && pred_pan && gt_pan && thing_slots && scratch && B > 0 && H > 0 && W > 0,
              LDM_ERR_BAD_ARG, "ldm_city_pan_maps: bad arg");
  LDM_REQUIRE(n_things >= 0 && n_things <= kMaxThings, LDM_ERR_BAD_SHAPE, "ldm_city_pan_maps: at most %d thing classes",
              kMaxThings);
  const long long hw = (long long)H * W;
  LDM_REQUIRE(hw < (1ll << 31) && 255ll * max_ins + hw < (1ll << 31), LDM_ERR_BAD_SHAPE,
              "ldm_city_pan_maps: ids overflow int32");
  const long long n2 = 2ll * B * hw;
  const int nblocks = (int)((hw + 1023) / 1024);
  const int ns = n_things > 0 ? n_things : 1;
  int32_t* Lbuf = scratch;
  int32_t* rank = scratch + n2;
  int32_t* bcnt = scratch + 2 * n2;
  int32_t* ncomp = bcnt + 2ll * B * ns * nblocks;
  cudaStream_t s = as_stream(stream);
  const int pred_void = -1;
  ccl_multi_init_kernel<<<grid1d(n2, 256), 256, 0, s>>>(pred_seg, gt_sem, thing_slots, Lbuf, hw, B, pred_void, ignore_label);
  if (n_things > 0) {
    ccl_multi_merge_kernel<<<dim3(grid1d(hw, 256, 8), 2 * B), 256, 0, s>>>(Lbuf, pred_seg, gt_sem, H, W, B, pred_void,
                                                                         ignore_label);
    ccl_multi_count_kernel<<<dim3(nblocks, 2 * B), 1024, 0, s>>>(Lbuf, pred_seg, gt_sem, thing_slots, bcnt, hw, B, ns,
                                                                pred_void, ignore_label);
    ccl_scan_blocks_kernel<<<2 * B * ns, 1024, 0, s>>>(bcnt, ncomp, nblocks);
    ccl_multi_rank_kernel<<<dim3(nblocks, 2 * B), 1024, 0, s>>>(Lbuf, pred_seg, gt_sem, thing_slots, bcnt, rank, hw, B, ns,
                                                               pred_void, ignore_label);
  }
  city_pan_final_kernel<<<grid1d((long long)B * hw, 256), 256, 0, s>>>(Lbuf, rank, pred_seg, gt_sem, pred_pan, gt_pan, hw,
                                                                       B, pred_void, ignore_label, max_ins);
  count_launch(n_things > 0 ? 6 : 2);
  return check_launch("city_pan_maps kernels");
}

extern "C" int ldm_joint_hist_batch(const int32_t* a, const int32_t* b, int64_t n_per_table, int64_t stride,
                                    int32_t n_tables, unsigned long long* keys, int32_t* counts, int32_t capacity,
                                    int32_t* overflow, ldm_stream_t stream) {
  using namespace ldm_host;
  LDM_REQUIRE(a && b && keys && counts && overflow && n_per_table > 0 && stride >= 0 && n_tables > 0, LDM_ERR_BAD_ARG,
              "ldm_joint_hist_batch: bad arg");
  LDM_REQUIRE(capacity >= 64 && (capacity & (capacity - 1)) == 0 && n_tables <= 65535, LDM_ERR_BAD_SHAPE,
              "ldm_joint_hist_batch: capacity must be a power of two >= 64, at most 65535 tables");
  cudaStream_t s = as_stream(stream);
  const long long slots = (long long)capacity * n_tables;
  hash_clear_batch_kernel<<<grid1d(slots > n_tables ? slots : n_tables, 256, 2), 256, 0, s>>>(keys, counts, slots, overflow,
                                                                                               n_tables);
  int gx = grid1d(n_per_table, 256, 8) / n_tables;
  if (gx < 8) gx = 8;
  joint_hist_batch_kernel<<<dim3(gx, n_tables), 256, 0, s>>>(a, b, n_per_table, stride, keys, counts, capacity, overflow);
  count_launch(2);
  return check_launch("joint_hist_batch_kernel");
}

extern "C" int ldm_pan_insert(const int32_t* sem, const int32_t* labels, int32_t target, int32_t max_ins, int32_t* pan,
                              int64_t n, ldm_stream_t stream) {
  using namespace ldm_host;
  LDM_REQUIRE(sem && labels && pan && n > 0, LDM_ERR_BAD_ARG, "ldm_pan_insert: bad arg");
  LDM_REQUIRE((long long)target * max_ins + n < (1ll << 31), LDM_ERR_BAD_SHAPE, "ldm_pan_insert: id overflows int32");
  pan_insert_kernel<<<grid1d(n, 256), 256, 0, as_stream(stream)>>>(sem, labels, target, max_ins, pan, n);
  count_launch();
  return check_launch("pan_insert_kernel");
}

extern "C" int ldm_pan_combine(const int32_t* cat, const int32_t* ins, int32_t max_ins, int32_t* pan, int64_t n,
                               ldm_stream_t stream) {
  using namespace ldm_host;
  LDM_REQUIRE(cat && ins && pan && n > 0 && max_ins > 0, LDM_ERR_BAD_ARG, "ldm_pan_combine: bad arg");
  pan_combine_kernel<<<grid1d(n, 256), 256, 0, as_stream(stream)>>>(cat, ins, max_ins, pan, n);
  count_launch();
  return check_launch("pan_combine_kernel");
}

// Depth-aware masking of DVPQ (eval/eval_dvpq.py:123-145). One thread per depth pixel; the per-CTA (sum, count) of the
// abs-rel error go to partial[] in a fixed order (warp shuffles, then warp 0 over the warps), the host adds the CTAs.
__global__ void depth_mask_kernel(int32_t* __restrict__ pred, int pred_stride, const int32_t* __restrict__ dp,
                                  const int32_t* __restrict__ dg, int H, int Wd, uint32_t wrap_mask, double thres,
                                  int32_t fill, double* __restrict__ psum, unsigned long long* __restrict__ pcnt) {
  __shared__ double sh_s[32];
  __shared__ unsigned int sh_c[32];
  const long long n = (long long)H * Wd;
  double s = 0.0;
  unsigned int c = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int32_t g = dg[i];
    if (g > 0) {
      const int32_t d = dp[i] - g;
      // numpy: unsigned 8 / 16 bit PNG arrays subtract modulo 2^bits (np.abs is then the identity), int32 ones are signed
      const double diff = wrap_mask ? (double)((uint32_t)d & wrap_mask) : (double)(d < 0 ? -(long long)d : (long long)d);
      const double rel = diff / (double)g;
      s += rel;
      ++c;
      if (rel > thres) {
        const long long y = i / Wd, x = i - y * Wd;
        pred[y * pred_stride + x] = fill;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    sh_s[warp] = s;
    sh_c[warp] = c;
  }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    s = lane < nw ? sh_s[lane] : 0.0;
    c = lane < nw ? sh_c[lane] : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (lane == 0) {
      psum[blockIdx.x] = s;
      pcnt[blockIdx.x] = c;
    }
  }
}

extern "C" int ldm_depth_mask_pred(int32_t* pred, int32_t pred_stride, const int32_t* depth_pred, const int32_t* depth_gt,
                                   int32_t H, int32_t Wd, int32_t elem_bits, double thres, int32_t fill, double* partial_sum,
                                   unsigned long long* partial_cnt, int32_t nblocks, ldm_stream_t stream) {
  using namespace ldm_host;
  LDM_REQUIRE(pred && depth_pred && depth_gt && partial_sum && partial_cnt, LDM_ERR_BAD_ARG, "ldm_depth_mask_pred: null arg");
  LDM_REQUIRE(H > 0 && Wd > 0 && pred_stride >= Wd && nblocks > 0, LDM_ERR_BAD_SHAPE,
              "ldm_depth_mask_pred: H=%d Wd=%d pred_stride=%d nblocks=%d", H, Wd, pred_stride, nblocks);
  LDM_REQUIRE(elem_bits == 8 || elem_bits == 16 || elem_bits == 32, LDM_ERR_BAD_ARG,
              "ldm_depth_mask_pred: elem_bits=%d (8 / 16: unsigned PNG samples, 32: int32)", elem_bits);
  const uint32_t wrap = elem_bits == 8 ? 0xffu : (elem_bits == 16 ? 0xffffu : 0u);
  depth_mask_kernel<<<nblocks, 256, 0, as_stream(stream)>>>(pred, pred_stride, depth_pred, depth_gt, H, Wd, wrap, thres, fill,
                                                          partial_sum, partial_cnt);
  count_launch();
  return check_launch("depth_mask_kernel");
}

extern "C" int ldm_id_mask(int32_t* x, const int32_t* a, int32_t va, const int32_t* b, int32_t vb, int32_t fill,
                           int64_t n, ldm_stream_t stream) {
  using namespace ldm_host;
  LDM_REQUIRE(x && a && n > 0, LDM_ERR_BAD_ARG, "ldm_id_mask: bad arg");
  id_mask_kernel<<<grid1d(n, 256), 256, 0, as_stream(stream)>>>(x, a, va, b, vb, fill, n);
  count_launch();
  return check_launch("id_mask_kernel");
}
