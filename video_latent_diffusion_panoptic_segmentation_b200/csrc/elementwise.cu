// Small HBM-/latency-bound kernels of the sampler loop: DDIM update, timestep sinusoid + GEMV chain, the
// few-channel 3x3 convs at the ends of the UNet / seg-AE (fused with concat+cast), nearest upsample, stride-2 im2col.
#include "common.cuh"
#include "host_util.h"

namespace {
using namespace ldm;

// ---------------------------------------------------------------- DDIM update (ddim_scheduler.py:234-267)
// Explicit round-to-nearest mul/sub/div/add intrinsics keep the reference's op order (no FMA contraction) so the
// fp32 result is bit-identical to torch eager on the CPU.
// `prev` / `x0out` may alias `sample` (in-place update by the sampler loop): no __restrict__ on those.
// With `eps_text` the noise prediction is first combined as classifier-free guidance does (trainers_ldm_cond.py:1147-1149):
// eps = eps_uncond + g * (eps_text - eps_uncond), three roundings in the reference's order.
__global__ void ddim_step_kernel(const float* __restrict__ eps, const float* __restrict__ eps_text, float guidance,
                                 const float* sample, const float* __restrict__ coef,
                                 const int32_t* __restrict__ t_index, float* prev, float* x0out, long long n,
                                 float clip, int reclip_eps) {
  // clip > 0: pred_original_sample is clamped to [-clip, clip] (clip_sample, :253-257); reclip_eps: the noise is then
  // re-derived from the clamped x0 (use_clipped_model_output, :259-261), both in the reference's op order
  const int ti = t_index ? *t_index : 0;
  const float s1m_at = coef[ti * 4 + 0];   // sqrt(1 - alpha_t)
  const float s_at = coef[ti * 4 + 1];     // sqrt(alpha_t)
  const float s_ap = coef[ti * 4 + 2];     // sqrt(alpha_prev)
  const float s1m_ap = coef[ti * 4 + 3];   // sqrt(1 - alpha_prev)
  const long long nv = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    const float4 e = __ldg(reinterpret_cast<const float4*>(eps) + i);
    const float4 x = reinterpret_cast<const float4*>(sample)[i];
    float ev[4] = {e.x, e.y, e.z, e.w};
    if (eps_text) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(eps_text) + i);
      const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) ev[j] = __fadd_rn(ev[j], __fmul_rn(guidance, __fsub_rn(tv[j], ev[j])));
    }
    const float xv[4] = {x.x, x.y, x.z, x.w};
    float p0[4], pv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      p0[j] = __fdiv_rn(__fsub_rn(xv[j], __fmul_rn(s1m_at, ev[j])), s_at);
      if (clip > 0.f) p0[j] = fminf(fmaxf(p0[j], -clip), clip);
      if (reclip_eps) ev[j] = __fdiv_rn(__fsub_rn(xv[j], __fmul_rn(s_at, p0[j])), s1m_at);
      pv[j] = __fadd_rn(__fmul_rn(s_ap, p0[j]), __fmul_rn(s1m_ap, ev[j]));
    }
    if (prev) reinterpret_cast<float4*>(prev)[i] = make_float4(pv[0], pv[1], pv[2], pv[3]);
    if (x0out) reinterpret_cast<float4*>(x0out)[i] = make_float4(p0[0], p0[1], p0[2], p0[3]);
  }
  // tail
  for (long long i = nv * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float e = eps[i];
    if (eps_text) e = __fadd_rn(e, __fmul_rn(guidance, __fsub_rn(eps_text[i], e)));
    float p0 = __fdiv_rn(__fsub_rn(sample[i], __fmul_rn(s1m_at, e)), s_at);
    if (clip > 0.f) p0 = fminf(fmaxf(p0, -clip), clip);
    if (reclip_eps) e = __fdiv_rn(__fsub_rn(sample[i], __fmul_rn(s_at, p0)), s1m_at);
    if (prev) prev[i] = __fadd_rn(__fmul_rn(s_ap, p0), __fmul_rn(s1m_ap, e));
    if (x0out) x0out[i] = p0;
  }
}

// ---------------------------------------------------------------- timestep sinusoid (diffusers Timesteps)
// out[0:half] = cos(t * f_i), out[half:2*half] = sin(t * f_i)   (flip_sin_to_cos=True, freq_shift=0)
__global__ void timestep_sinusoid_kernel(const long long* __restrict__ timesteps, const int32_t* __restrict__ t_index,
                                         const float* __restrict__ freqs, float* __restrict__ out, int half) {
  const float t = (float)timesteps[t_index ? *t_index : 0];
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float a = __fmul_rn(t, freqs[i]);
    out[i] = cosf(a);
    out[half + i] = sinf(a);
  }
}

// ---------------------------------------------------------------- GEMV: out[n] = act(bias[n] + bias2[n] + w[n,:] . x)
// one warp per output row; x (fp32) staged in shared memory; w bf16 row-major [N,K], K % 8 == 0.
__global__ void gemv_bf16_kernel(const __nv_bfloat16* __restrict__ w, const float* __restrict__ bias,
                                 const float* __restrict__ bias2, const float* __restrict__ x, float* __restrict__ out,
                                 int N, int K, int silu_out) {
  extern __shared__ float xs[];
  for (int i = threadIdx.x; i < K; i += blockDim.x) xs[i] = x[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int n = blockIdx.x * wpb + (threadIdx.x >> 5); n < N; n += gridDim.x * wpb) {
    const __nv_bfloat16* row = w + (long long)n * K;
    float acc = 0.f;
    for (int k = lane * 8; k < K; k += 256) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(row + k));
      const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
      acc += a.x * xs[k] + a.y * xs[k + 1] + b.x * xs[k + 2] + b.y * xs[k + 3] + c.x * xs[k + 4] + c.y * xs[k + 5] +
             d.x * xs[k + 6] + d.y * xs[k + 7];
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      float v = acc + (bias ? bias[n] : 0.f) + (bias2 ? bias2[n] : 0.f);
      out[n] = silu_out ? silu_f(v) : v;
    }
  }
}

// ---------------------------------------------------------------- few-input-channel conv3x3 (conv_in / AE conv_in)
// Input: up to 3 NCHW fp32 sources of `cps` channels each (the latent concat, trainers_ldm_cond.py:1134-1141),
// scaled by `scale`; weights fp32 [cin, 3, 3, cout] (output channel innermost); output bf16 NHWC. One thread per output channel, a CTA per
// strip of pixels in one image row; the input patch is staged in shared memory (broadcast reads).
constexpr int kStrip = 32;
__global__ void conv3x3_small_cin_kernel(const float* __restrict__ s0, const float* __restrict__ s1,
                                         const float* __restrict__ s2, int nsrc, int cps, float scale, float shift,
                                         const float* __restrict__ w, const float* __restrict__ bias,
                                         __nv_bfloat16* __restrict__ out, int B, int h, int wd, int cout, int silu) {
  extern __shared__ float patch[];  // [cin][3][kStrip+2]
  const int cin = nsrc * cps;
  const int strips = (wd + kStrip - 1) / kStrip;
  const int bid = blockIdx.x;
  const int sx = bid % strips;
  const int y = (bid / strips) % h;
  const int b = bid / (strips * h);
  const int x0 = sx * kStrip;
  const int pw = kStrip + 2;
  for (int i = threadIdx.x; i < cin * 3 * pw; i += blockDim.x) {
    const int c = i / (3 * pw);
    const int r = (i / pw) % 3;
    const int px = i % pw;
    const int yy = y + r - 1, xx = x0 + px - 1;
    float v = 0.f;
    if (yy >= 0 && yy < h && xx >= 0 && xx < wd) {
      const int si = c / cps, cc = c % cps;
      const float* src = si == 0 ? s0 : (si == 1 ? s1 : s2);
      // (two roundings, as `2. * images - 1.` of encode_inputs; the conv's zero padding stays zero)
      v = __fadd_rn(__fmul_rn(src[(((long long)b * cps + cc) * h + yy) * wd + xx], scale), shift);
    }
    patch[i] = v;
  }
  __syncthreads();
  const int co = threadIdx.x;
  if (co >= cout) return;
  float acc[kStrip];
  const float bv = bias ? bias[co] : 0.f;
#pragma unroll
  for (int i = 0; i < kStrip; ++i) acc[i] = bv;
  for (int c = 0; c < cin; ++c) {
    float wk[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) wk[t] = __ldg(&w[((long long)c * 9 + t) * cout + co]);  // [cin,3,3,cout]: coalesced over co
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float* prow = patch + (c * 3 + r) * pw;
#pragma unroll
      for (int i = 0; i < kStrip; ++i)
        acc[i] += wk[r * 3] * prow[i] + wk[r * 3 + 1] * prow[i + 1] + wk[r * 3 + 2] * prow[i + 2];
    }
  }
  for (int i = 0; i < kStrip; ++i) {
    const int x = x0 + i;
    if (x < wd) out[(((long long)b * h + y) * wd + x) * cout + co] = __float2bfloat16_rn(silu ? silu_f(acc[i]) : acc[i]);
  }
}

// ---------------------------------------------------------------- conv_out: bf16 NHWC [B,h,w,cin] -> fp32 NCHW [B,cout,h,w]
// one warp per output pixel; lanes split the channels; weights staged in smem as [tap][cin][cout] fp32.
constexpr int kMaxCoutSmall = 8;
__global__ void conv_out_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                const float* __restrict__ bias, float* __restrict__ out, int B, int h, int wd,
                                int cin, int cout) {
  extern __shared__ float ws[];  // [9][cin][cout]
  for (int i = threadIdx.x; i < 9 * cin * cout; i += blockDim.x) {
    const int co = i % cout;
    const int c = (i / cout) % cin;
    const int t = i / (cout * cin);
    ws[i] = w[((long long)co * cin + c) * 9 + t];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const long long npix = (long long)B * h * wd;
  const int nvec = cin / 8;
  for (long long pix = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); pix < npix;
       pix += (long long)gridDim.x * wpb) {
    const int xx = (int)(pix % wd);
    const int yy = (int)((pix / wd) % h);
    const int b = (int)(pix / ((long long)wd * h));
    float acc[kMaxCoutSmall];
#pragma unroll
    for (int o = 0; o < kMaxCoutSmall; ++o) acc[o] = 0.f;
    for (int t = 0; t < 9; ++t) {
      const int sy = yy + t / 3 - 1, sx = xx + t % 3 - 1;
      if (sy < 0 || sy >= h || sx < 0 || sx >= wd) continue;
      const __nv_bfloat16* src = x + (((long long)b * h + sy) * wd + sx) * cin;
      for (int v = lane; v < nvec; v += 32) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + v * 8));
        const float2 a = unpack_bf16(u.x), b2 = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
        const float f[8] = {a.x, a.y, b2.x, b2.y, c.x, c.y, d.x, d.y};
        const float* wp = ws + ((long long)t * cin + v * 8) * cout;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
          for (int o = 0; o < kMaxCoutSmall; ++o)
            if (o < cout) acc[o] += f[j] * wp[j * cout + o];
        }
      }
    }
#pragma unroll
    for (int o = 0; o < kMaxCoutSmall; ++o) {
      if (o < cout) {
        const float s = warp_sum(acc[o]);
        if (lane == 0) out[(((long long)b * cout + o) * h + yy) * wd + xx] = s + (bias ? bias[o] : 0.f);
      }
    }
  }
}

// ---------------------------------------------------------------- nearest upsample (F.interpolate mode="nearest")
__global__ void upsample_nearest_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int B,
                                        int h, int w, int C, int oh, int ow, float sh_, float sw_) {
  const int vpp = C / 8;
  const long long total = (long long)B * oh * ow * vpp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vpp);
    long long p = i / vpp;
    const int ox = (int)(p % ow);
    p /= ow;
    const int oy = (int)(p % oh);
    const int b = (int)(p / oh);
    int sy = (int)floorf(__fmul_rn((float)oy, sh_));
    int sx = (int)floorf(__fmul_rn((float)ox, sw_));
    sy = sy < h - 1 ? sy : h - 1;
    sx = sx < w - 1 ? sx : w - 1;
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + (((long long)b * h + sy) * w + sx) * C) + v);
    reinterpret_cast<uint4*>(out + (((long long)b * oh + oy) * ow + ox) * C)[v] = u;
  }
}

// ---------------------------------------------------------------- im2col for the stride-2 Downsample2D conv
__global__ void im2col3x3_s2_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, int h,
                                    int w, int C, int oh, int ow, int pad) {
  const int vpp = C / 8;
  const long long total = (long long)B * oh * ow * 9 * vpp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vpp);
    long long p = i / vpp;
    const int t = (int)(p % 9);
    p /= 9;
    const int ox = (int)(p % ow);
    p /= ow;
    const int oy = (int)(p % oh);
    const int b = (int)(p / oh);
    const int sy = 2 * oy + t / 3 - pad, sx = 2 * ox + t % 3 - pad;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (sy >= 0 && sy < h && sx >= 0 && sx < w)
      u = __ldg(reinterpret_cast<const uint4*>(x + (((long long)b * h + sy) * w + sx) * C) + v);
    reinterpret_cast<uint4*>(out)[i] = u;
  }
}

int grid_for(long long work_items, int threads) {
  long long g = (work_items + threads - 1) / threads;
  const long long cap = (long long)ldm_host::num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

extern "C" int ldm_ddim_step(const float* eps, const float* sample, const float* coef, const int32_t* t_index,
                             float* prev_sample, float* pred_x0, int64_t n, ldm_stream_t stream) {
  return ldm_ddim_step_clip(eps, nullptr, 0.f, sample, coef, t_index, prev_sample, pred_x0, n, 0.f, 0, stream);
}

extern "C" int ldm_ddim_step_cfg(const float* eps_uncond, const float* eps_text, float guidance_scale,
                                 const float* sample, const float* coef, const int32_t* t_index, float* prev_sample,
                                 float* pred_x0, int64_t n, ldm_stream_t stream) {
  return ldm_ddim_step_clip(eps_uncond, eps_text, guidance_scale, sample, coef, t_index, prev_sample, pred_x0, n, 0.f, 0,
                            stream);
}

extern "C" int ldm_ddim_step_clip(const float* eps_uncond, const float* eps_text, float guidance_scale,
                                  const float* sample, const float* coef, const int32_t* t_index, float* prev_sample,
                                  float* pred_x0, int64_t n, float clip_sample_range, int32_t use_clipped_model_output,
                                  ldm_stream_t stream) {
  using namespace ldm_host;
  LDM_REQUIRE(clip_sample_range >= 0.f, LDM_ERR_BAD_ARG, "ldm_ddim_step: clip_sample_range=%g (0 = no clipping)",
              (double)clip_sample_range);
  LDM_REQUIRE(eps_uncond && sample && coef, LDM_ERR_BAD_ARG, "ldm_ddim_step: null arg");
  LDM_REQUIRE(n > 0, LDM_ERR_BAD_SHAPE, "ldm_ddim_step: n=%lld", (long long)n);
  LDM_REQUIRE(((uintptr_t)eps_uncond | (uintptr_t)eps_text | (uintptr_t)sample | (uintptr_t)prev_sample |
               (uintptr_t)pred_x0) % 16 == 0,
              LDM_ERR_ALIGNMENT, "ldm_ddim_step: pointers must be 16-byte aligned");
  ddim_step_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, as_stream(stream)>>>(eps_uncond, eps_text, guidance_scale, sample,
                                                                            coef, t_index, prev_sample, pred_x0, n,
                                                                            clip_sample_range,
                                                                            use_clipped_model_output != 0);
  count_launch();
  return check_launch("ddim_step_kernel");
}

extern "C" int ldm_timestep_sinusoid(const int64_t* timesteps, const int32_t* t_index, const float* freqs, float* out,
                                     int32_t half, ldm_stream_t stream) {
  using namespace ldm_host;
  LDM_REQUIRE(timesteps && freqs && out && half > 0, LDM_ERR_BAD_ARG, "ldm_timestep_sinusoid: bad arg");
  timestep_sinusoid_kernel<<<1, 256, 0, as_stream(stream)>>>(reinterpret_cast<const long long*>(timesteps), t_index,
                                                             freqs, out, half);
  count_launch();
  return check_launch("timestep_sinusoid_kernel");
}

extern "C" int ldm_gemv_bf16(const void* w, const float* bias, const float* bias2, const float* x, float* out,
                             int32_t N, int32_t K, int32_t silu_out, ldm_stream_t stream) {
  using namespace ldm_host;
  LDM_REQUIRE(w && x && out, LDM_ERR_BAD_ARG, "ldm_gemv_bf16: null arg");
  LDM_REQUIRE(N > 0 && K > 0 && K % 8 == 0 && K <= 8192, LDM_ERR_BAD_SHAPE, "ldm_gemv_bf16: N=%d K=%d", N, K);
  const int wpb = 8;
  int grid = (N + wpb - 1) / wpb;
  const int cap = num_sms() * 8;
  if (grid > cap) grid = cap;
  gemv_bf16_kernel<<<grid, wpb * 32, sizeof(float) * K, as_stream(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(w), bias, bias2, x, out, N, K, silu_out);
  count_launch();
  return check_launch("gemv_bf16_kernel");
}

extern "C" int ldm_conv3x3_small_cin(const float* s0, const float* s1, const float* s2, int32_t nsrc, int32_t cps,
                                     float scale, const float* w, const float* bias, void* out, int32_t B, int32_t h,
                                     int32_t wd, int32_t cout, ldm_stream_t stream) {
  return ldm_conv3x3_small_cin_act(s0, s1, s2, nsrc, cps, scale, w, bias, out, B, h, wd, cout, 0, stream);
}

extern "C" int ldm_conv3x3_small_cin_act(const float* s0, const float* s1, const float* s2, int32_t nsrc, int32_t cps,
                                         float scale, const float* w, const float* bias, void* out, int32_t B, int32_t h,
                                         int32_t wd, int32_t cout, int32_t silu, ldm_stream_t stream) {
  return ldm_conv3x3_small_cin_affine(s0, s1, s2, nsrc, cps, scale, 0.f, w, bias, out, B, h, wd, cout, silu, stream);
}

extern "C" int ldm_conv3x3_small_cin_affine(const float* s0, const float* s1, const float* s2, int32_t nsrc,
                                            int32_t cps, float scale, float shift, const float* w, const float* bias,
                                            void* out, int32_t B, int32_t h, int32_t wd, int32_t cout, int32_t silu,
                                            ldm_stream_t stream) {
  using namespace ldm_host;
  LDM_REQUIRE(s0 && w && out && nsrc >= 1 && nsrc <= 3 && (nsrc < 2 || s1) && (nsrc < 3 || s2), LDM_ERR_BAD_ARG,
              "ldm_conv3x3_small_cin: bad sources");
  LDM_REQUIRE(B > 0 && h > 0 && wd > 0 && cps > 0 && cout > 0 && cout <= 1024 && nsrc * cps <= 64, LDM_ERR_BAD_SHAPE,
              "ldm_conv3x3_small_cin: bad shape");
  const int threads = ((cout + 31) / 32) * 32;
  const int strips = (wd + kStrip - 1) / kStrip;
  const size_t shb = sizeof(float) * nsrc * cps * 3 * (kStrip + 2);
  conv3x3_small_cin_kernel<<<B * h * strips, threads, shb, as_stream(stream)>>>(
      s0, s1, s2, nsrc, cps, scale, shift, w, bias, reinterpret_cast<__nv_bfloat16*>(out), B, h, wd, cout, silu);
  count_launch();
  return check_launch("conv3x3_small_cin_kernel");
}

extern "C" int ldm_conv_out(const void* x, const float* w, const float* bias, float* out, int32_t B, int32_t h,
                            int32_t wd, int32_t cin, int32_t cout, ldm_stream_t stream) {
  using namespace ldm_host;
  LDM_REQUIRE(x && w && out, LDM_ERR_BAD_ARG, "ldm_conv_out: null arg");
  LDM_REQUIRE(cout > 0 && cout <= kMaxCoutSmall && cin % 8 == 0 && 9 * cin * cout * 4 <= 200 * 1024, LDM_ERR_BAD_SHAPE,
              "ldm_conv_out: cin=%d cout=%d unsupported", cin, cout);
  const size_t shb = sizeof(float) * 9 * cin * cout;
  cudaFuncSetAttribute(conv_out_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);  // (per device)
  const int grid = num_sms() * 2;
  conv_out_kernel<<<grid, 256, shb, as_stream(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x), w, bias, out, B, h,
                                                         wd, cin, cout);
  count_launch();
  return check_launch("conv_out_kernel");
}

extern "C" int ldm_upsample_nearest(const void* x, void* out, int32_t B, int32_t h, int32_t w, int32_t C, int32_t oh,
                                    int32_t ow, ldm_stream_t stream) {
  using namespace ldm_host;
  LDM_REQUIRE(x && out && C % 8 == 0 && B > 0 && h > 0 && w > 0 && oh > 0 && ow > 0, LDM_ERR_BAD_ARG,
              "ldm_upsample_nearest: bad arg");
  // torch: scale = in/out in float unless an explicit scale_factor was given; for exact 2x both give 0.5
  const float sh_ = (oh == 2 * h) ? 0.5f : (float)h / (float)oh;
  const float sw_ = (ow == 2 * w) ? 0.5f : (float)w / (float)ow;
  const long long total = (long long)B * oh * ow * (C / 8);
  upsample_nearest_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(out), B, h, w, C, oh, ow, sh_, sw_);
  count_launch();
  return check_launch("upsample_nearest_kernel");
}

extern "C" int ldm_im2col3x3_s2(const void* x, void* out, int32_t B, int32_t h, int32_t w, int32_t C, int32_t oh,
                                int32_t ow, ldm_stream_t stream) {
  return ldm_im2col3x3_s2_pad(x, out, B, h, w, C, oh, ow, 1, stream);
}

extern "C" int ldm_im2col3x3_s2_pad(const void* x, void* out, int32_t B, int32_t h, int32_t w, int32_t C, int32_t oh,
                                    int32_t ow, int32_t pad_lo, ldm_stream_t stream) {
  using namespace ldm_host;
  LDM_REQUIRE(x && out && C % 8 == 0 && B > 0 && (pad_lo == 0 || pad_lo == 1), LDM_ERR_BAD_ARG,
              "ldm_im2col3x3_s2: bad arg");
  // rows / columns before the image: pad_lo; after it: 1 (conv pad 1, or F.pad(x, (0, 1, 0, 1)) in front of a pad-0 conv)
  LDM_REQUIRE(oh == (h + pad_lo - 2) / 2 + 1 && ow == (w + pad_lo - 2) / 2 + 1, LDM_ERR_BAD_SHAPE,
              "ldm_im2col3x3_s2: output %dx%d does not match the stride-2 conv of %dx%d (pad %d before, 1 after)", oh, ow,
              h, w, pad_lo);
  const long long total = (long long)B * oh * ow * 9 * (C / 8);
  im2col3x3_s2_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(out), B, h, w, C, oh, ow, pad_lo);
  count_launch();
  return check_launch("im2col3x3_s2_kernel");
}
