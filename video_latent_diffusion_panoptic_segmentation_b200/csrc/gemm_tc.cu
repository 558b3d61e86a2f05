// Tensor-core contraction kernel for the LDMSeg UNet / seg-AE: plain GEMM (Linear, conv1x1) and implicit-GEMM
// conv3x3 (stride 1, pad 1) on NHWC bf16 activations.
//
//   persistent, warp-specialised, one CTA per SM (192 threads):
//     warp 0      TMA producer   : 4-D tiled tensor maps over [B,H,W,C]; the 3x3 taps are nine shifted box loads,
//                                  the zero padding is TMA out-of-bounds fill. Channel-concat inputs are two maps.
//     warp 1      MMA issuer     : tcgen05.mma cta_group::1 kind::f16, M=128 x N=block_n x K=16, fp32 accumulators
//                                  double-buffered in TMEM (2 x 256 columns) so the epilogue overlaps the next tile.
//     warps 2..5  epilogue       : tcgen05.ld 32x32b -> registers -> fused bias / time-embedding / residual / SiLU /
//                                  GEGLU / QKV head split / ConvT pixel-shuffle+LayerNorm2d+SiLU -> global.
//   smem ring of (A 128x64 | B block_n x 64) bf16 stages, SWIZZLE_128B, full/empty mbarriers.
#include "common.cuh"
#include "host_util.h"

namespace {

using namespace ldm;

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KiB
constexpr int kMaxStages = 8;
constexpr int kThreads = 192;
constexpr int kSmemBudget = 227 * 1024;

struct GemmParams {
  // tile geometry
  int tiles_x, tiles_y, B, bw, bh, W, H;
  int m_tiles, n_tiles, block_n;
  int N, taps, kblocks1, kblocks, ktap;
  int stages, stage_bytes;
  int flags;
  // epilogue
  const float* bias;
  const float* rowbias;
  const __nv_bfloat16* residual;
  void* out;
  __nv_bfloat16* q;
  __nv_bfloat16* k;
  __nv_bfloat16* vt;
  const float* ln_gamma;
  const float* ln_beta;
  float ln_eps;
  int heads, head_dim, dpad, seq, seq_pad, vt_rows;
  int n_store;      // OUT_NCHW_F32: leading output channels actually stored
  long long img_px; // pixels per image of the un-flattened problem (H*W)
};

__device__ __forceinline__ void store_bf16x8(__nv_bfloat16* dst, const float* v) {
  uint4 u;
  u.x = pack_bf16(v[0], v[1]);
  u.y = pack_bf16(v[2], v[3]);
  u.z = pack_bf16(v[4], v[5]);
  u.w = pack_bf16(v[6], v[7]);
  *reinterpret_cast<uint4*>(dst) = u;
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by SWIZZLE_128B tiles.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 128);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = p.m_tiles * p.n_tiles;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const uint32_t a_bytes = (uint32_t)(p.bw * p.bh) * (kBlockK * 2);
  const uint32_t b_bytes = (uint32_t)p.block_n * (kBlockK * 2);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_tile = tile / p.m_tiles;
        const int m_tile = tile - n_tile * p.m_tiles;
        const int b = m_tile / tiles_per_img;
        const int rem = m_tile - b * tiles_per_img;
        const int ty = rem / p.tiles_x;
        const int tx = rem - ty * p.tiles_x;
        const int x0 = tx * p.bw, y0 = ty * p.bh, n0 = n_tile * p.block_n;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dy = (p.taps == 9) ? tap / 3 - 1 : 0;
          const int dx = (p.taps == 9) ? tap % 3 - 1 : 0;
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * p.stage_bytes;
            uint8_t* sb = sa + kABytes;
            mbar_arrive_expect_tx(&full_bar[stage], a_bytes + b_bytes);
            if (kb < p.kblocks1)
              tma_load_4d(sa, &tmA1, &full_bar[stage], kb * kBlockK, x0 + dx, y0 + dy, b);
            else
              tma_load_4d(sa, &tmA2, &full_bar[stage], (kb - p.kblocks1) * kBlockK, x0 + dx, y0 + dy, b);
            tma_load_2d(sb, &tmB, &full_bar[stage], tap * p.ktap + kb * kBlockK, n0);
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(kBlockM, (uint32_t)p.block_n);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      const int ksteps = p.taps * p.kblocks;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)as * 256u;
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * p.stage_bytes);
          const uint64_t da = umma_desc_k_sw128(sa);
          const uint64_t db = umma_desc_k_sw128(sa + kABytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // advance 16 bf16 = 32 bytes inside the 128-byte swizzle atom: +2 in 16-byte units
            umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (ks | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[as]);
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int r = quad * 32 + lane;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n_tile = tile / p.m_tiles;
      const int m_tile = tile - n_tile * p.m_tiles;
      const int b = m_tile / tiles_per_img;
      const int rem = m_tile - b * tiles_per_img;
      const int ty = rem / p.tiles_x;
      const int tx = rem - ty * p.tiles_x;
      const int ly = r / p.bw, lx = r - ly * p.bw;
      const int y = ty * p.bh + ly, x = tx * p.bw + lx;
      const bool valid = (r < p.bw * p.bh) && (y < p.H) && (x < p.W);
      const long long grow = ((long long)b * p.H + y) * p.W + x;
      const int n0 = n_tile * p.block_n;

      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)as * 256u;

      if (p.flags & LDM_GEMM_CONVT_LN_SILU) {
        // One N tile == one (dy,dx) sub-pixel of ConvTranspose2d(k=2,s=2); LayerNorm2d over its block_n channels.
        const int cout = p.block_n;
        float mean = 0.f;
        for (int c = 0; c < cout; c += 32) {
          uint32_t v[32];
          tmem_ld32(t_addr + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) mean += __uint_as_float(v[j]) + __ldg(&p.bias[n0 + c + j]);
        }
        mean /= (float)cout;
        float var = 0.f;
        for (int c = 0; c < cout; c += 32) {
          uint32_t v[32];
          tmem_ld32(t_addr + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float d = __uint_as_float(v[j]) + __ldg(&p.bias[n0 + c + j]) - mean;
            var += d * d;
          }
        }
        const float rstd = 1.0f / sqrtf(var / (float)cout + p.ln_eps);
        const int sub = n_tile;  // dy*2+dx
        const long long orow = ((long long)b * (2 * p.H) + (2 * y + (sub >> 1))) * (2 * p.W) + (2 * x + (sub & 1));
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + orow * cout;
        for (int c = 0; c < cout; c += 32) {
          uint32_t v[32];
          tmem_ld32(t_addr + c, v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float o[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int cc = c + g * 8 + j;
                const float xn = (__uint_as_float(v[g * 8 + j]) + __ldg(&p.bias[n0 + cc]) - mean) * rstd;
                o[j] = silu_f(__ldg(&p.ln_gamma[cc]) * xn + __ldg(&p.ln_beta[cc]));
              }
              store_bf16x8(dst + c + g * 8, o);
            }
          }
        }
      } else {
        for (int c = 0; c < p.block_n; c += 32) {
          uint32_t v[32];
          tmem_ld32(t_addr + c, v);
          tmem_ld_wait();
          if (!valid) continue;
          const int nc = n0 + c;
          if (nc >= p.N) continue;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nc + j < p.N) f[j] += __ldg(&p.bias[nc + j]);
          }
          if (p.rowbias) {
            const float* rb = p.rowbias + (long long)b * p.N + nc;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nc + j < p.N) f[j] += __ldg(&rb[j]);
          }
          if (p.flags & LDM_GEMM_GEGLU) {
            // columns [0,16) value, [16,32) gate of the same 16 outputs
            float o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = f[j] * gelu_erf_f(f[16 + j]);
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + grow * (p.N / 2) + nc / 2;
            store_bf16x8(dst, o);
            store_bf16x8(dst + 8, o + 8);
            continue;
          }
          if (p.flags & LDM_GEMM_QKV_SPLIT) {
            const int C = p.heads * p.head_dim;
            const int bi = (int)(grow / p.seq);
            const int s = (int)(grow - (long long)bi * p.seq);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int n = nc + g * 8;
              if (n >= p.N) break;
              const int which = n / C;
              const int cc = n - which * C;
              const int head = cc / p.head_dim;
              const int e = cc - head * p.head_dim;
              const long long bh = (long long)bi * p.heads + head;
              if (which < 2) {
                __nv_bfloat16* dst = (which == 0 ? p.q : p.k) + (bh * p.seq + s) * p.dpad + e;
                store_bf16x8(dst, f + g * 8);
              } else {
                __nv_bfloat16* dst = p.vt + (bh * p.vt_rows + e) * p.seq_pad + s;
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[(long long)j * p.seq_pad] = __float2bfloat16_rn(f[g * 8 + j]);
              }
            }
            continue;
          }
          if (p.residual) {
            const __nv_bfloat16* rs = p.residual + grow * p.N + nc;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (nc + g * 8 >= p.N) break;
              const uint4 u = *reinterpret_cast<const uint4*>(rs + g * 8);
              const float2 a0 = unpack_bf16(u.x), a1 = unpack_bf16(u.y), a2 = unpack_bf16(u.z), a3 = unpack_bf16(u.w);
              f[g * 8 + 0] += a0.x; f[g * 8 + 1] += a0.y; f[g * 8 + 2] += a1.x; f[g * 8 + 3] += a1.y;
              f[g * 8 + 4] += a2.x; f[g * 8 + 5] += a2.y; f[g * 8 + 6] += a3.x; f[g * 8 + 7] += a3.y;
            }
          }
          if (p.flags & LDM_GEMM_SILU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = silu_f(f[j]);
          }
          if (p.flags & LDM_GEMM_OUT_NCHW_F32) {
            // planar fp32 output [B, n_store, H, W]: lanes hold consecutive pixels -> coalesced per channel
            const long long bi = grow / p.img_px, pix = grow - bi * p.img_px;
            float* dst = reinterpret_cast<float*>(p.out) + (bi * p.n_store) * p.img_px + pix;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nc + j < p.n_store) dst[(long long)(nc + j) * p.img_px] = f[j];
            continue;
          }
          if (p.flags & LDM_GEMM_OUT_F32) {
            float* dst = reinterpret_cast<float*>(p.out) + grow * p.N + nc;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              if (nc + g * 4 >= p.N) break;
              *reinterpret_cast<float4*>(dst + g * 4) = make_float4(f[g * 4], f[g * 4 + 1], f[g * 4 + 2], f[g * 4 + 3]);
            }
          } else {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + grow * p.N + nc;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (nc + g * 8 >= p.N) break;
              store_bf16x8(dst + g * 8, f + g * 8);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[as]);
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Choose the pixel box (bw x bh <= 128 rows) that covers HxW with the fewest tiles.
void pick_box(int H, int W, int taps, int* bw_out, int* bh_out) {
  (void)taps;
  if (H == 1) {
    *bw_out = W < 128 ? W : 128;
    *bh_out = 1;
    return;
  }
  long best_tiles = -1;
  int best_bw = 1, best_bh = 1;
  for (int bw = 1; bw <= 128 && bw <= W; ++bw) {
    int bh = 128 / bw;
    if (bh > H) bh = H;
    if (bh < 1) continue;
    const long tiles = (long)((W + bw - 1) / bw) * ((H + bh - 1) / bh);
    if (best_tiles < 0 || tiles < best_tiles || (tiles == best_tiles && bw > best_bw)) {
      best_tiles = tiles;
      best_bw = bw;
      best_bh = bh;
    }
  }
  *bw_out = best_bw;
  *bh_out = best_bh;
}

int pick_block_n(int N) {
  const int cands[] = {256, 192, 160, 128, 96, 64, 32};
  int best = 256;
  long best_pad = -1;
  for (int c : cands) {
    const long pad = (long)((N + c - 1) / c) * c;
    if (best_pad < 0 || pad < best_pad) {
      best_pad = pad;
      best = c;
    }
  }
  return best;
}

}  // namespace

extern "C" int ldm_gemm_bf16(const ldm_gemm_desc* d, ldm_stream_t stream) {
  using namespace ldm_host;
  LDM_REQUIRE(d != nullptr, LDM_ERR_BAD_ARG, "ldm_gemm_bf16: null descriptor");
  LDM_REQUIRE(d->a1 && d->w, LDM_ERR_BAD_ARG, "ldm_gemm_bf16: null a1/w");
  LDM_REQUIRE(d->taps == 1 || d->taps == 9, LDM_ERR_BAD_ARG, "ldm_gemm_bf16: taps must be 1 or 9 (got %d)", d->taps);
  LDM_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->N > 0 && d->c1 > 0, LDM_ERR_BAD_SHAPE,
              "ldm_gemm_bf16: non-positive extent B=%d H=%d W=%d N=%d c1=%d", d->B, d->H, d->W, d->N, d->c1);
  const int c2 = d->a2 ? d->c2 : 0;
  LDM_REQUIRE(d->c1 % 8 == 0 && c2 % 8 == 0, LDM_ERR_ALIGNMENT, "ldm_gemm_bf16: channels must be multiples of 8");
  if (d->taps == 9 || d->a2)
    LDM_REQUIRE(d->c1 % 64 == 0 && c2 % 64 == 0, LDM_ERR_BAD_SHAPE,
                "ldm_gemm_bf16: conv3x3 / concat sources need channels %% 64 == 0 (c1=%d c2=%d)", d->c1, c2);
  const int flags = d->flags;
  int block_n = d->block_n > 0 ? d->block_n : pick_block_n(d->N);
  LDM_REQUIRE(block_n % 32 == 0 && block_n >= 32 && block_n <= 256, LDM_ERR_BAD_ARG, "ldm_gemm_bf16: block_n=%d",
              block_n);
  if (flags & LDM_GEMM_GEGLU) LDM_REQUIRE(d->N % 32 == 0, LDM_ERR_BAD_SHAPE, "GEGLU needs N %% 32 == 0");
  if (flags & LDM_GEMM_QKV_SPLIT) {
    LDM_REQUIRE(d->q && d->k && d->vt && d->heads > 0 && d->head_dim % 8 == 0 && d->N == 3 * d->heads * d->head_dim &&
                    d->seq > 0 && d->seq_pad % 8 == 0 && d->dpad % 64 == 0 &&
                    (long long)d->B * d->H * d->W % d->seq == 0,
                LDM_ERR_BAD_SHAPE, "ldm_gemm_bf16: bad QKV split geometry");
  } else {
    LDM_REQUIRE(d->out != nullptr, LDM_ERR_BAD_ARG, "ldm_gemm_bf16: null out");
    LDM_REQUIRE(d->N % 8 == 0, LDM_ERR_ALIGNMENT, "ldm_gemm_bf16: N must be a multiple of 8 (got %d)", d->N);
  }
  if (flags & LDM_GEMM_CONVT_LN_SILU) {
    LDM_REQUIRE(d->taps == 1 && d->N == 4 * block_n && d->bias && d->ln_gamma && d->ln_beta, LDM_ERR_BAD_SHAPE,
                "ldm_gemm_bf16: CONVT_LN_SILU needs taps=1, N=4*block_n, bias, ln params");
  }

  GemmParams p{};
  p.B = d->B;
  p.H = d->H;
  p.W = d->W;
  // A pointwise GEMM has no spatial structure: flatten to one row of B*H*W pixels so tiles are always full,
  // except for modes whose epilogue needs (b, y, x).
  if (d->taps == 1 && !(flags & LDM_GEMM_CONVT_LN_SILU) && !d->rowbias) {
    p.W = d->B * d->H * d->W;
    p.H = 1;
    p.B = 1;
  }
  pick_box(p.H, p.W, d->taps, &p.bw, &p.bh);
  p.tiles_x = (p.W + p.bw - 1) / p.bw;
  p.tiles_y = (p.H + p.bh - 1) / p.bh;
  p.m_tiles = p.tiles_x * p.tiles_y * p.B;
  p.block_n = block_n;
  p.n_tiles = (d->N + block_n - 1) / block_n;
  p.N = d->N;
  p.taps = d->taps;
  p.kblocks1 = (d->c1 + kBlockK - 1) / kBlockK;
  p.kblocks = p.kblocks1 + (c2 + kBlockK - 1) / kBlockK;
  p.ktap = d->c1 + c2;
  p.stage_bytes = kABytes + block_n * kBlockK * 2;
  p.stages = (kSmemBudget - 2048) / p.stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  LDM_REQUIRE(p.stages >= 2, LDM_ERR_BAD_SHAPE, "ldm_gemm_bf16: not enough shared memory for 2 stages");
  p.flags = flags;
  p.bias = d->bias;
  p.rowbias = d->rowbias;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(d->residual);
  p.out = d->out;
  p.q = reinterpret_cast<__nv_bfloat16*>(d->q);
  p.k = reinterpret_cast<__nv_bfloat16*>(d->k);
  p.vt = reinterpret_cast<__nv_bfloat16*>(d->vt);
  p.ln_gamma = d->ln_gamma;
  p.ln_beta = d->ln_beta;
  p.ln_eps = d->ln_eps;
  p.heads = d->heads;
  p.head_dim = d->head_dim;
  p.dpad = d->dpad;
  p.seq = d->seq;
  p.seq_pad = d->seq_pad;
  p.vt_rows = d->vt_rows > 0 ? d->vt_rows : d->head_dim;
  p.n_store = d->n_store > 0 ? d->n_store : d->N;
  p.img_px = (long long)d->H * d->W;

  CUtensorMap tmA1, tmA2, tmB;
  {
    const uint64_t dims[4] = {(uint64_t)d->c1, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.B};
    const uint64_t str[3] = {(uint64_t)d->c1 * 2, (uint64_t)d->c1 * 2 * p.W, (uint64_t)d->c1 * 2 * p.W * p.H};
    const uint32_t box[4] = {kBlockK, (uint32_t)p.bw, (uint32_t)p.bh, 1};
    int rc = make_tmap(&tmA1, d->a1, 4, dims, str, box, 2, true);
    if (rc) return rc;
  }
  if (c2 > 0) {
    const uint64_t dims[4] = {(uint64_t)c2, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.B};
    const uint64_t str[3] = {(uint64_t)c2 * 2, (uint64_t)c2 * 2 * p.W, (uint64_t)c2 * 2 * p.W * p.H};
    const uint32_t box[4] = {kBlockK, (uint32_t)p.bw, (uint32_t)p.bh, 1};
    int rc = make_tmap(&tmA2, d->a2, 4, dims, str, box, 2, true);
    if (rc) return rc;
  } else {
    tmA2 = tmA1;
  }
  {
    const uint64_t ktot = (uint64_t)d->taps * p.ktap;
    const uint64_t dims[2] = {ktot, (uint64_t)d->N};
    const uint64_t str[1] = {ktot * 2};
    const uint32_t box[2] = {kBlockK, (uint32_t)block_n};
    int rc = make_tmap(&tmB, d->w, 2, dims, str, box, 2, true);
    if (rc) return rc;
  }

  const int smem_bytes = p.stages * p.stage_bytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) return set_error(LDM_ERR_CUDA, "cudaFuncSetAttribute(gemm): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int num_tiles = p.m_tiles * p.n_tiles;
  int grid = num_sms();
  if (grid > num_tiles) grid = num_tiles;
  gemm_tc_kernel<<<grid, kThreads, smem_bytes, as_stream(stream)>>>(tmA1, tmA2, tmB, p);
  count_launch();
  return check_launch("gemm_tc_kernel");
}
