// Tensor-core contraction kernel for the LDMSeg UNet / seg-AE: plain GEMM (Linear, conv1x1) and implicit-GEMM
// conv3x3 (stride 1, pad 1) on NHWC bf16 activations.
//
//   persistent, warp-specialised, one CTA per SM (192 threads):
//     warp 0      TMA producer   : 4-D tiled tensor maps over [B,H,W,C]; the 3x3 taps are nine shifted box loads,
//                                  the zero padding is TMA out-of-bounds fill. Channel-concat inputs are two maps.
//     warp 1      MMA issuer     : tcgen05.mma cta_group::1 kind::f16, M=128 x N=block_n x K=16, fp32 accumulators
//                                  double-buffered in TMEM (2 x 256 columns) so the epilogue overlaps the next tile.
//     warps 2..9  epilogue       : tcgen05.ld 32x32b -> registers -> fused bias / time-embedding / residual / SiLU /
//                                  GEGLU / QKV head split / ConvT pixel-shuffle+LayerNorm2d+SiLU -> global. Two warps
//                                  per TMEM lane quadrant (even / odd 32-column chunks); the residual rows of a tile
//                                  are prefetched into registers before the accumulator is waited for.
//   smem ring of (A 128x64 | B block_n x 64) bf16 stages, SWIZZLE_128B, full/empty mbarriers.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "host_util.h"

namespace {

using namespace ldm;

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KiB
constexpr int kMaxStages = 8;
constexpr int kEpiWarps = 8;  // two warps per TMEM lane quadrant, each takes every other 32-column chunk
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kSmemBudget = 227 * 1024;
constexpr int kEpiStageBytes = 2 * 16384 + 1024;  // epilogue output staging for the TMA-store path + per-half bias slice
// Diagnostic switches (env LDM_GEMM_DEBUG, never set by the product): isolate the two halves of the main loop.
constexpr int kPrefetchNext = 1 << 21; // internal: producer prefetches the next tile's A boxes / residual rows into L2
constexpr int kStagedStore = 1 << 20; // internal: epilogue writes the output through shared memory + TMA store
constexpr int kDbgNoTma = 1 << 29;  // producer signals the stages without loading them
constexpr int kDbgNoMma = 1 << 30;
constexpr int kDbgNoFence = 1 << 27; // skip tcgen05.fence::after_thread_sync after the full-barrier wait
constexpr int kDbgKeepCommit = 1 << 26; // (with nowait) keep the per-k-block tcgen05.commit
constexpr int kDbgAltAcc = 1 << 25;  // (timing only, wrong results) alternate two accumulators between consecutive MMAs
constexpr int kDbgNoWait = 1 << 28; // (with notma) MMA thread neither waits for nor releases stages: a pure MMA stream  // MMA thread releases the stages without issuing MMAs

struct GemmParams {
  // tile geometry
  int tiles_x, tiles_y, B, bw, bh, W, H;
  int m_tiles, n_tiles, block_n;
  int N, taps, kblocks1, kblocks, ktap;
  int stages, stage_bytes;
  int res_kblocks;  // > 0: the residual is accumulated by res_kblocks extra K blocks (residual box x identity band)
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

// kPair = false: one CTA per 128 x block_n tile (tcgen05 cta_group::1).
// kPair = true : a cluster of two CTAs (one TPC) per 256 x block_n tile (cta_group::2, UMMA M = 256). Each CTA loads
//                its own 128-row A box and HALF of the B tile (block_n/2 weight rows) -- the tensor core reads the other
//                half from the peer's shared memory -- so the per-SM operand ingest from L2, which bounds the
//                single-CTA kernel, drops from 16 KiB + block_n*128 B to 16 KiB + block_n*64 B per k-block. The leader
//                (cluster rank 0) issues every MMA; both CTAs' TMA loads complete on the leader's full barrier; MMA
//                completion is multicast to both CTAs' empty / accumulator-full barriers; both epilogues release
//                the accumulator on the leader's barrier.
template <bool kPair>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO,
               const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmE, const GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by SWIZZLE_128B tiles (the dynamic smem base is the same in both CTAs of a pair).
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_smem = smem + p.stages * p.stage_bytes;  // 2 x 16 KiB output staging (one 128 x 64 bf16 block per half)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + kEpiStageBytes);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  // warp index through a shuffle so that the compiler knows it is warp-uniform: the producer / MMA warps then run
  // their loops on all 32 lanes with uniform control flow (descriptors and barrier addresses live in uniform
  // registers) and only the TMA / MMA / commit instructions themselves are issued by one elected lane. Running the
  // loops on a single lane inside a divergent branch made ptxas wrap every UTCHMMA in an ELECT / R2UR.BROADCAST loop.
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;   // 0 = leader
  const int worker = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int num_workers = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], (kPair ? 2 : 1) * kEpiWarps);  // one arrival per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    if (kPair) {
      tmem_alloc_pair(tmem_slot, 512);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work items: (m_unit, n_tile) with m_unit = one M tile, or a pair of consecutive M tiles
  const int m_units = kPair ? (p.m_tiles + 1) / 2 : p.m_tiles;
  const int num_tiles = m_units * p.n_tiles;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const uint32_t a_bytes = (uint32_t)(p.bw * p.bh) * (kBlockK * 2);
  const int b_rows = kPair ? p.block_n / 2 : p.block_n;   // weight rows this CTA loads per k-block
  const uint32_t b_bytes = (uint32_t)b_rows * (kBlockK * 2);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = worker; tile < num_tiles; tile += num_workers) {
        const int n_tile = tile / m_units;
        const int m_unit = tile - n_tile * m_units;
        const int m_tile = kPair ? 2 * m_unit + (int)rank : m_unit;  // == m_tiles for the odd tail: all-OOB box, zeros
        const int b = m_tile / tiles_per_img;
        const int rem = m_tile - b * tiles_per_img;
        const int ty = rem / p.tiles_x;
        const int tx = rem - ty * p.tiles_x;
        const int x0 = tx * p.bw, y0 = ty * p.bh;
        const int n0 = n_tile * p.block_n + (int)rank * b_rows;
        if ((p.flags & kPrefetchNext) && elect_one()) {
          // Short main loops: the smem ring holds barely one tile, so the next tile's loads can only be issued once
          // this tile's MMAs have drained it and would pay the full DRAM latency. Pull the next tile's A boxes and
          // the residual rows (this tile's on the first pass, then always one tile ahead) into L2 now.
          for (int pass = (tile == worker ? 0 : 1); pass < 2; ++pass) {
            const int t2 = tile + pass * num_workers;
            if (t2 >= num_tiles) break;
            const int nt2 = t2 / m_units;
            const int mu2 = t2 - nt2 * m_units;
            const int mt2 = kPair ? 2 * mu2 + (int)rank : mu2;
            const int b2 = mt2 / tiles_per_img;
            const int rem2 = mt2 - b2 * tiles_per_img;
            const int ty2 = rem2 / p.tiles_x;
            const int tx2 = rem2 - ty2 * p.tiles_x;
            if (pass == 1)
              for (int kb = 0; kb < p.kblocks; ++kb)
                tma_prefetch_l2_4d(kb < p.kblocks1 ? &tmA1 : &tmA2, (kb < p.kblocks1 ? kb : kb - p.kblocks1) * kBlockK,
                                   tx2 * p.bw, ty2 * p.bh, b2);
            if (p.residual)
              for (int c = 0; c < p.block_n && nt2 * p.block_n + c < p.N; c += 64)
                tma_prefetch_l2_4d(&tmR, nt2 * p.block_n + c, tx2 * p.bw, ty2 * p.bh, b2);
          }
        }
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dy = (p.taps == 9) ? tap / 3 - 1 : 0;
          const int dx = (p.taps == 9) ? tap % 3 - 1 : 0;
          for (int kb = 0; kb < p.kblocks; ++kb) {
            if (p.flags & kDbgNoWait) continue;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * p.stage_bytes;
            uint8_t* sb = sa + kABytes;
            const CUtensorMap* tmA = (kb < p.kblocks1) ? &tmA1 : &tmA2;
            const int kc = (kb < p.kblocks1 ? kb : kb - p.kblocks1) * kBlockK;
            if (!elect_one()) {
            } else if (p.flags & kDbgNoTma) {
              if (rank == 0) mbar_arrive(&full_bar[stage]);
            } else if (kPair) {
              // the leader's barrier collects the bytes of both CTAs
              if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (a_bytes + b_bytes));
              tma_load_4d_pair(sa, tmA, &full_bar[stage], kc, x0 + dx, y0 + dy, b);
              tma_load_2d_pair(sb, &tmB, &full_bar[stage], tap * p.ktap + kb * kBlockK, n0);
            } else {
              mbar_arrive_expect_tx(&full_bar[stage], a_bytes + b_bytes);
              tma_load_4d(sa, tmA, &full_bar[stage], kc, x0 + dx, y0 + dy, b);
              tma_load_2d(sb, &tmB, &full_bar[stage], tap * p.ktap + kb * kBlockK, n0);
            }
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        // residual as extra K blocks: A = residual[rows of this tile, 64 output columns], B = the matching band of the
        // identity matrix (row n of the tile x column n), so that D += R on the tensor core
        for (int j = 0; j < p.res_kblocks; ++j) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * p.stage_bytes;
          uint8_t* sb = sa + kABytes;
          if (elect_one()) {
            if (kPair) {
              if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (a_bytes + b_bytes));
              tma_load_4d_pair(sa, &tmR, &full_bar[stage], n_tile * p.block_n + j * kBlockK, x0, y0, b);
              tma_load_2d_pair(sb, &tmE, &full_bar[stage], j * kBlockK, (int)rank * b_rows);
            } else {
              mbar_arrive_expect_tx(&full_bar[stage], a_bytes + b_bytes);
              tma_load_4d(sa, &tmR, &full_bar[stage], n_tile * p.block_n + j * kBlockK, x0, y0, b);
              tma_load_2d(sb, &tmE, &full_bar[stage], j * kBlockK, 0);
            }
          }
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (pair: leader CTA only)
    if (rank == 0) {
      const uint32_t idesc = umma_idesc_bf16(kPair ? 256 : kBlockM, (uint32_t)p.block_n);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      const int ksteps = p.taps * p.kblocks + p.res_kblocks;
      for (int tile = worker; tile < num_tiles; tile += num_workers) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)as * 256u;
        for (int ks = 0; ks < ksteps; ++ks) {
          if (!(p.flags & kDbgNoWait)) {
            mbar_wait(&full_bar[stage], phase);
            if (!(p.flags & kDbgNoFence)) tc_fence_after();
          }
          const uint32_t sa = smem_u32(smem + stage * p.stage_bytes);
          const uint64_t da = umma_desc_k_sw128(sa);
          const uint64_t db = umma_desc_k_sw128(sa + kABytes);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              if (p.flags & kDbgNoMma) break;
              // advance 16 bf16 = 32 bytes inside the 128-byte swizzle atom: +2 in 16-byte units
              const uint32_t dt = (p.flags & kDbgAltAcc) ? d_tmem + (uint32_t)(k & 1) * 128u : d_tmem;
              if (kPair)
                umma_bf16_pair(dt, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (ks | k) != 0);
              else
                umma_bf16(dt, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (ks | k) != 0);
            }
            if ((p.flags & kDbgNoWait) && !(p.flags & kDbgKeepCommit)) {
            } else if (kPair) umma_commit_pair(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) {
          if (kPair) umma_commit_pair(&tfull_bar[as]); else umma_commit(&tfull_bar[as]);
        }
        __syncwarp();
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    const int quad = warp & 3;            // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;     // 0: even 32-column chunks, 1: odd chunks
    const int r = quad * 32 + lane;
    const bool plain = !(p.flags & (LDM_GEMM_CONVT_LN_SILU | LDM_GEMM_GEGLU | LDM_GEMM_QKV_SPLIT));
    const bool geglu = (p.flags & LDM_GEMM_GEGLU) != 0;
    const bool staged = (p.flags & kStagedStore) != 0;  // bf16 [rows, N] (or GEGLU [rows, N/2]) through smem + TMA store
    const bool epi_leader = threadIdx.x == 64;  // issues the TMA stores of the staged path
    int sb = 0;                                  // staging buffer of the next output block (alternates)
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = worker; tile < num_tiles; tile += num_workers) {
      const int n_tile = tile / m_units;
      const int m_unit = tile - n_tile * m_units;
      const int m_tile = kPair ? 2 * m_unit + (int)rank : m_unit;
      const int b = m_tile / tiles_per_img;
      const int rem = m_tile - b * tiles_per_img;
      const int ty = rem / p.tiles_x;
      const int tx = rem - ty * p.tiles_x;
      const int ly = r / p.bw, lx = r - ly * p.bw;
      const int y = ty * p.bh + ly, x = tx * p.bw + lx;
      const bool valid = (m_tile < p.m_tiles) && (r < p.bw * p.bh) && (y < p.H) && (x < p.W);
      const long long grow = ((long long)b * p.H + y) * p.W + x;
      const int n0 = n_tile * p.block_n;

      // residual rows of this tile: issue every load now, so that they are in flight while the MMAs finish
      // staged path: both halves work on the same 64-column output block, half h on its 32-column chunk h, so chunk
      // ci of a thread covers columns min(ci*64, block_n-64) + h*32; direct path: a half owns every other chunk.
      uint4 rs[4][4];
      const bool use_res = plain && p.residual != nullptr && valid && p.res_kblocks == 0;
      if (use_res) {
        const __nv_bfloat16* rrow = p.residual + grow * p.N + n0;
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          // (the last 64-column block of a block_n that is not a multiple of 64 is shifted left to end at block_n)
          const int c = staged ? (ci * 64 < p.block_n ? min(ci * 64, p.block_n - 64) + half * 32 : p.block_n)
                               : (2 * ci + half) * 32;
#pragma unroll
          for (int g = 0; g < 4; ++g)
            if (c < p.block_n && n0 + c + g * 8 < p.N) rs[ci][g] = ld_nc_v4(rrow + c + g * 8);
        }
      }

      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)as * 256u;

      if (p.flags & LDM_GEMM_CONVT_LN_SILU) {
        if (half == 0) {
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
        }
      } else if (staged) {
        // 64 output columns (= one 128-byte swizzle atom per row) at a time: registers -> swizzled smem -> TMA store.
        // All eight warps work on the same block (half h converts accumulator chunk(s) h), the two staging buffers
        // alternate, and one thread issues the store. The TMA store clips pixels / columns outside the tensor, so
        // ragged tiles need no predication here.
        const int acc_w = geglu ? 128 : 64;  // accumulator columns behind 64 output columns
#pragma unroll
        for (int ob = 0; ob < 4; ++ob) {
          if (ob * acc_w >= p.block_n) break;
          // a ragged last block is shifted left so that it ends at block_n: the overlap is rewritten with the same values
          const int c0 = min(ob * acc_w, p.block_n - acc_w);
          uint8_t* sbuf = epi_smem + sb * 16384;
          float* sbias = reinterpret_cast<float*>(epi_smem + 2 * 16384) + sb * 128;
          if (epi_leader) bulk_wait_read1();  // the store that last used this buffer (two blocks ago) has been read
          {
            // this block's bias values (bias + per-image bias) once, through shared memory
            const int t = threadIdx.x - 64;
            if (t < acc_w) {
              const int n = n0 + c0 + t;
              float bv = 0.f;
              if (n < p.N) {
                if (p.bias) bv = __ldg(p.bias + n);
                if (p.rowbias) bv += __ldg(p.rowbias + (long long)b * p.N + n);
              }
              sbias[t] = bv;
            }
          }
          named_bar_sync(1, 32 * kEpiWarps);
          uint8_t* rowp = sbuf + r * 128;
#pragma unroll
          for (int cq = 0; cq < 2; ++cq) {
            if (!geglu && cq == 1) break;            // plain: one 32-column chunk per thread and block
            const int cc = geglu ? half * 64 + cq * 32 : half * 32;
            const int c = c0 + cc;
            uint32_t v[32];
            tmem_ld32(t_addr + c, v);
            tmem_ld_wait();
            const int nc = n0 + c;
            float f[32];
            {
              const float4* bp = reinterpret_cast<const float4*>(sbias + cc);
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float4 bv = bp[g];
                f[g * 4] = __uint_as_float(v[g * 4]) + bv.x;
                f[g * 4 + 1] = __uint_as_float(v[g * 4 + 1]) + bv.y;
                f[g * 4 + 2] = __uint_as_float(v[g * 4 + 2]) + bv.z;
                f[g * 4 + 3] = __uint_as_float(v[g * 4 + 3]) + bv.w;
              }
            }
            if (geglu) {
              // columns [0,16) value, [16,32) gate of the same 16 outputs -> output 16-byte chunks (cc/32)*2, +1
              float o[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] = f[j] * gelu_erf_fast(f[16 + j]);
#pragma unroll
              for (int g = 0; g < 2; ++g) {
                uint4 u;
                u.x = pack_bf16(o[g * 8 + 0], o[g * 8 + 1]); u.y = pack_bf16(o[g * 8 + 2], o[g * 8 + 3]);
                u.z = pack_bf16(o[g * 8 + 4], o[g * 8 + 5]); u.w = pack_bf16(o[g * 8 + 6], o[g * 8 + 7]);
                *reinterpret_cast<uint4*>(rowp + ((((cc >> 5) * 2 + g) ^ (r & 7)) << 4)) = u;
              }
            } else {
              if (use_res) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                  if (nc + g * 8 >= p.N) break;
                  const uint4 u = rs[ob][g];
                  const float2 a0 = unpack_bf16(u.x), a1 = unpack_bf16(u.y), a2 = unpack_bf16(u.z), a3 = unpack_bf16(u.w);
                  f[g * 8 + 0] += a0.x; f[g * 8 + 1] += a0.y; f[g * 8 + 2] += a1.x; f[g * 8 + 3] += a1.y;
                  f[g * 8 + 4] += a2.x; f[g * 8 + 5] += a2.y; f[g * 8 + 6] += a3.x; f[g * 8 + 7] += a3.y;
                }
              }
              if (p.flags & LDM_GEMM_SILU) {
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = silu_f(f[j]);
              }
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                uint4 u;
                u.x = pack_bf16(f[g * 8 + 0], f[g * 8 + 1]); u.y = pack_bf16(f[g * 8 + 2], f[g * 8 + 3]);
                u.z = pack_bf16(f[g * 8 + 4], f[g * 8 + 5]); u.w = pack_bf16(f[g * 8 + 6], f[g * 8 + 7]);
                *reinterpret_cast<uint4*>(rowp + (((half * 4 + g) ^ (r & 7)) << 4)) = u;
              }
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(1, 32 * kEpiWarps);
          if (epi_leader) {
            const int ncol = geglu ? (n0 + c0) / 2 : n0 + c0;
            tma_store_4d(&tmO, sbuf, ncol, tx * p.bw, ty * p.bh, b);
            bulk_commit();
          }
          sb ^= 1;
        }
      } else {
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c = (2 * ci + half) * 32;
          if (c >= p.block_n) break;
          uint32_t v[32];
          tmem_ld32(t_addr + c, v);
          tmem_ld_wait();
          if (!valid) continue;
          const int nc = n0 + c;
          if (nc >= p.N) continue;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (p.bias) {  // N % 8 == 0 (GEGLU: % 32): 4-wide groups are all-in or all-out
            const float4* bp = reinterpret_cast<const float4*>(p.bias + nc);
#pragma unroll
            for (int g = 0; g < 8; ++g)
              if (nc + g * 4 < p.N) {
                const float4 bv = __ldg(bp + g);
                f[g * 4] += bv.x; f[g * 4 + 1] += bv.y; f[g * 4 + 2] += bv.z; f[g * 4 + 3] += bv.w;
              }
          }
          if (p.rowbias) {
            const float4* bp = reinterpret_cast<const float4*>(p.rowbias + (long long)b * p.N + nc);
#pragma unroll
            for (int g = 0; g < 8; ++g)
              if (nc + g * 4 < p.N) {
                const float4 bv = __ldg(bp + g);
                f[g * 4] += bv.x; f[g * 4 + 1] += bv.y; f[g * 4 + 2] += bv.z; f[g * 4 + 3] += bv.w;
              }
          }
          if (p.flags & LDM_GEMM_GEGLU) {
            // columns [0,16) value, [16,32) gate of the same 16 outputs
            float o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = f[j] * gelu_erf_fast(f[16 + j]);
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
          if (use_res) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (nc + g * 8 >= p.N) break;
              const uint4 u = rs[ci][g];
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
      // the accumulator buffer is free again: every lane's TMEM loads are complete (tcgen05.wait::ld), one lane per
      // warp tells the MMA warp (of the leader CTA)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kPair) mbar_arrive_cluster(&tempty_bar[as], 0); else mbar_arrive(&tempty_bar[as]);
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  }

  if (threadIdx.x == 64) bulk_wait_all();  // outstanding TMA stores
  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();  // pair: neither CTA may exit while its peer can still signal it
  if (warp == 1) {
    tc_fence_after();
    if (kPair) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
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

// Choose block_n (and single CTA vs CTA pair) with a cost model of the persistent schedule, fitted to block_n sweeps on
// B200 (tools/profile_kernels.py --sweep; profiles/r01d_sweep_block_n*.json): work items run in waves over the SMs (or
// SM pairs); an item costs kblocks * 4 MMAs of a measured per-MMA time plus a per-item epilogue / pipeline term.
// Measured per-MMA cycles (M = 128 per SM, K = 16): roughly a 100-cycle floor for N <= 128 and ~60 + operand bytes /
// 114 B/clk above it -- the tensor pipe's shared-memory operand read, not its math, is what binds, which is why the
// CTA pair (each SM reads only half of B) wins on the long-K convolutions, and why N = 160 beats N = 256 for 320-wide
// layers. The pair's cross-CTA accumulator hand-off costs more per item, so short-K GEMMs stay on single CTAs.
// Avoids the two failure modes of "largest tile that divides N": a second, nearly empty wave (150 tiles on 148 SMs) and a
// handful of huge tiles when M is small (M = 960 at the 6x20 level).
int pick_block_n(int N, long m_tiles, long kblocks, int sms, bool pair, int taps, double* cost_out) {
  static const int cands[] = {256, 224, 192, 160, 128, 96};
  static const double mma_single[] = {167, 158, 151, 141, 111, 106};
  static const double mma_pair[] = {160, 149, 136, 124, 99, 92};
  int best = 256;
  double best_cost = -1.0;
  const long m_units = pair ? (m_tiles + 1) / 2 : m_tiles;
  const long workers = pair ? sms / 2 : sms;
  for (int i = 0; i < 6; ++i) {
    const int c = cands[i];
    if (c == 224 && taps != 9) continue;  // 224 only pays on the long-K convolutions (and is erratic on short K)
    const long n_tiles = (N + c - 1) / c;
    const long waves = (m_units * n_tiles + workers - 1) / workers;
    const double item = (double)kblocks * 4.0 * (pair ? mma_pair[i] : mma_single[i]) +
                        (pair ? 3300.0 + 11.0 * c : 1000.0 + 17.0 * c);
    const double cost = (double)waves * item;
    if (best_cost < 0 || cost < best_cost * 0.999) {
      best_cost = cost;
      best = c;
    }
  }
  if (cost_out) *cost_out = best_cost;
  return best;
}

int debug_flags() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LDM_GEMM_DEBUG");
    v = 0;
    if (e && strstr(e, "notma")) v |= kDbgNoTma;
    if (e && strstr(e, "nomma")) v |= kDbgNoMma;
    if (e && strstr(e, "nowait")) v |= kDbgNoWait | kDbgNoTma;
    if (e && strstr(e, "nofence")) v |= kDbgNoFence;
    if (e && strstr(e, "altacc")) v |= kDbgAltAcc;
    if (e && strstr(e, "keepcommit")) v |= kDbgKeepCommit;
  }
  return v;
}

// LDM_GEMM_PREFETCH=0 disables the one-tile-ahead L2 prefetch (A/B timing)
bool prefetch_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LDM_GEMM_PREFETCH");
    v = e ? atoi(e) : 0;
  }
  return v != 0;
}

// Residual on the tensor core for main loops of up to this many K blocks (LDM_GEMM_RESMMA overrides, 0 = never)
int resmma_max_kblocks() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LDM_GEMM_RESMMA");
    v = e ? atoi(e) : 1 << 30;
  }
  return v;
}

// LDM_GEMM_STAGED=0 keeps the direct (row-per-thread) global stores (A/B timing)
bool staged_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LDM_GEMM_STAGED");
    v = e ? atoi(e) : 1;
  }
  return v != 0;
}

// LDM_GEMM_PAIR=0 / 1 forces the single-CTA / CTA-pair kernel (A/B timing); default: the cost model decides.
int pair_override() {
  static int v = -2;
  if (v == -2) {
    const char* e = getenv("LDM_GEMM_PAIR");
    v = e ? atoi(e) : -1;
  }
  return v;
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
  // tile geometry first: the block_n choice depends on the number of M tiles
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
  p.kblocks1 = (d->c1 + kBlockK - 1) / kBlockK;
  p.kblocks = p.kblocks1 + (c2 + kBlockK - 1) / kBlockK;
  const long kblocks_total = (long)d->taps * p.kblocks;
  double cost1 = 0.0, cost2 = 0.0;
  const int bn1 = pick_block_n(d->N, p.m_tiles, kblocks_total, num_sms(), false, d->taps, &cost1);
  const int bn2 = pick_block_n(d->N, p.m_tiles, kblocks_total, num_sms(), true, d->taps, &cost2);
  // an explicit block_n keeps the single-CTA kernel (unless the A/B override forces pairs on a pairable block_n)
  const bool pair_ok = p.m_tiles >= 2 && !(flags & LDM_GEMM_CONVT_LN_SILU) &&
                       (d->block_n <= 0 || (pair_override() == 1 && d->block_n >= 64 && d->block_n % 32 == 0));
  bool pair = pair_ok && cost2 < cost1 && kblocks_total >= 32;  // the pair's hand-offs only pay on long main loops
  if (pair_override() >= 0) pair = pair_ok && pair_override() != 0;  // LDM_GEMM_PAIR=-1: model decides
  int block_n = d->block_n > 0 ? d->block_n : (pair ? bn2 : bn1);
  if ((flags & LDM_GEMM_GEGLU) && d->block_n <= 0 && block_n < 128) block_n = 128;  // staged GEGLU blocks span 128 columns
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

  p.block_n = block_n;
  p.n_tiles = (d->N + block_n - 1) / block_n;
  p.N = d->N;
  p.taps = d->taps;
  p.ktap = d->c1 + c2;
  p.stage_bytes = kABytes + (pair ? block_n / 2 : block_n) * kBlockK * 2;
  p.stages = (kSmemBudget - 2048 - kEpiStageBytes) / p.stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  LDM_REQUIRE(p.stages >= 2, LDM_ERR_BAD_SHAPE, "ldm_gemm_bf16: not enough shared memory for 2 stages");
  // bf16 [rows, N] outputs (and GEGLU's [rows, N/2]) leave through shared memory + TMA store
  const int n_out = (flags & LDM_GEMM_GEGLU) ? d->N / 2 : d->N;
  const bool staged = !(flags & (LDM_GEMM_OUT_F32 | LDM_GEMM_OUT_NCHW_F32 | LDM_GEMM_QKV_SPLIT | LDM_GEMM_CONVT_LN_SILU)) &&
                      n_out >= 64 && n_out % 8 == 0 && block_n >= ((flags & LDM_GEMM_GEGLU) ? 128 : 64) && staged_enabled();
  // pointwise GEMMs with a short K loop are latency-bound on the next tile's first loads: prefetch one tile ahead
  const bool prefetch_next = d->taps == 1 && kblocks_total <= 32 && prefetch_enabled();  // measured: no gain, off by default
  p.flags = flags | debug_flags() | (staged ? kStagedStore : 0) | (prefetch_next ? kPrefetchNext : 0);
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
    const uint32_t box[2] = {kBlockK, (uint32_t)(pair ? block_n / 2 : block_n)};
    int rc = make_tmap(&tmB, d->w, 2, dims, str, box, 2, true);
    if (rc) return rc;
  }

  CUtensorMap tmO = tmB;
  if (staged) {
    const uint64_t dims[4] = {(uint64_t)n_out, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.B};
    const uint64_t str[3] = {(uint64_t)n_out * 2, (uint64_t)n_out * 2 * p.W, (uint64_t)n_out * 2 * p.W * p.H};
    const uint32_t box[4] = {64, (uint32_t)p.bw, (uint32_t)p.bh, 1};
    int rc = make_tmap(&tmO, d->out, 4, dims, str, box, 2, true);
    if (rc) return rc;
  }

  // short-K pointwise GEMMs: the residual goes through the tensor core (see ldm_gemm_desc.identity)
  const bool plain_mode = !(flags & (LDM_GEMM_CONVT_LN_SILU | LDM_GEMM_GEGLU | LDM_GEMM_QKV_SPLIT));
  const bool res_mma = d->residual && d->identity && plain_mode && kblocks_total <= resmma_max_kblocks() && d->N % 8 == 0;
  p.res_kblocks = res_mma ? (block_n + kBlockK - 1) / kBlockK : 0;
  CUtensorMap tmE = tmB;
  if (res_mma) {
    const uint64_t dims[2] = {256, 256};
    const uint64_t str[1] = {512};
    const uint32_t box[2] = {kBlockK, (uint32_t)(pair ? block_n / 2 : block_n)};
    int rc = make_tmap(&tmE, d->identity, 2, dims, str, box, 2, true);
    if (rc) return rc;
  }
  CUtensorMap tmR = tmB;
  if ((prefetch_next || res_mma) && d->residual) {
    const uint64_t dims[4] = {(uint64_t)d->N, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.B};
    const uint64_t str[3] = {(uint64_t)d->N * 2, (uint64_t)d->N * 2 * p.W, (uint64_t)d->N * 2 * p.W * p.H};
    const uint32_t box[4] = {64, (uint32_t)p.bw, (uint32_t)p.bh, 1};
    int rc = make_tmap(&tmR, d->residual, 4, dims, str, box, 2, true);
    if (rc) return rc;
  }

  const int smem_bytes = p.stages * p.stage_bytes + kEpiStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) return set_error(LDM_ERR_CUDA, "cudaFuncSetAttribute(gemm): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  if (pair) {
    const int units = ((p.m_tiles + 1) / 2) * p.n_tiles;
    int clusters = num_sms() / 2;
    if (clusters > units) clusters = units;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = as_stream(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<true>, tmA1, tmA2, tmB, tmO, tmR, tmE, p);
    if (e != cudaSuccess) return set_error(LDM_ERR_CUDA, "gemm_tc_kernel<pair> launch: %s", cudaGetErrorString(e));
    count_launch();
    return check_launch("gemm_tc_kernel<pair>");
  }
  const int num_tiles = p.m_tiles * p.n_tiles;
  int grid = num_sms();
  if (grid > num_tiles) grid = num_tiles;
  gemm_tc_kernel<false><<<grid, kThreads, smem_bytes, as_stream(stream)>>>(tmA1, tmA2, tmB, tmO, tmR, tmE, p);
  count_launch();
  return check_launch("gemm_tc_kernel");
}
