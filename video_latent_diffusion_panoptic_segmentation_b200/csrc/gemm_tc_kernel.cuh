// Kernel template of the tensor-core contraction (see gemm_tc.cu for the host side and the design notes).
#pragma once
#include "common.cuh"
#include "host_util.h"

namespace ldm_gemm {

using namespace ldm;

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KiB
constexpr int kMaxStages = 8;
constexpr int kEpiWarps = 8;  // two warps per TMEM lane quadrant, each takes every other 32-column chunk
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kSmemBudget = 227 * 1024;
constexpr int kEpiStageBytes = 2 * 16384 + 1280;  // epilogue output staging for the TMA-store path + bias and LN column-sum slices (128 floats each; QKV_SPLIT with the LN fold: 160 each)
// Diagnostic switches (env LDM_GEMM_DEBUG, never set by the product): isolate the two halves of the main loop.
constexpr int kStagedStore = 1 << 20; // internal: epilogue writes the output through shared memory + TMA store
constexpr int kQkvStaged = 1 << 21;   // internal (QKV_SPLIT): q / k tiles are whole heads and leave through TMA stores
constexpr int kDbgNoTma = 1 << 29;  // producer signals the stages without loading them
constexpr int kDbgNoMma = 1 << 30;
constexpr int kDbgNoFence = 1 << 27; // skip tcgen05.fence::after_thread_sync after the full-barrier wait
constexpr int kDbgKeepCommit = 1 << 26; // (with nowait) keep the per-k-block tcgen05.commit
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
  int qkv_part0;    // QKV_SPLIT: first of the q|k|v parts the N columns hold (0: q.., 1: k.., 2: v only)
  unsigned long long magic_seq, magic_c, magic_d;  // ceil(2^40 / divisor): x / d == (x * magic) >> 40 for x * d < 2^40
  // strided / sub-pixel convolutions (ldm_gemm_desc.a_stride, up2): element-strided tensor maps, no gather pass
  int a_stride;     // 1, or 2: Conv2d(3x3, stride 2), the A boxes start at (2 x0 + kx - a_pad, 2 y0 + ky - a_pad)
  int a_pad;        // taps == 9: tap (ky, kx) reads (y + ky - a_pad, x + kx - a_pad); 1 for the dense convolution
  int up2_ntiles;   // > 0: nearest-upsample x2 + conv3x3 as four 2x2 convolutions; N tiles per output parity class
  int up2_cout;     // output channels of one class
  int n_store;      // OUT_NCHW_F32: leading output channels actually stored
  long long img_px; // pixels per image of the un-flattened problem (H*W)
  // LayerNorm folded into the GEMMs around it (no normalisation pass, see ldm_gemm_desc.ln_stats):
  float* row_stats;        // producer (staged epilogue): per row and 32-column chunk (sum, sum of squares) of the bf16 output
  const float* ln_stats;   // consumer: the producer's partials of ITS INPUT rows, [ln_parts][rows][2]
  const float* ln_colsum;  // consumer: g[n] = sum_c W'[n, c] (W' = gamma-scaled weight as stored)
  int ln_parts;
  long long rows_total;    // B * H * W: the partials are part-major, [parts][rows_total] (sum, sum of squares)
  float ln_invc, ln_fold_eps;
  // split-K (kEpiSplit): every tile is computed by split_k work items, each over a contiguous range of the K blocks
  int split_k;
  float* ws;                  // fp32 partial tiles [split_k][m_tiles_alloc][n_tiles][block_n][128 rows]
  long long ws_slice_stride;  // floats per slice
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
// kEpi selects the epilogue at compile time (one instantiation per mode keeps each kernel's hot loop small: the
// all-modes-in-one kernel was 170 KB of SASS and its epilogue warps stalled on instruction fetch):
// kEpiSplit = split-K: when M is small and K long (the 6x20 level: 8 x 14 tiles for 148 SMs, 180 - 360 K blocks each;
// at B = 1 a single row of tiles) the K range of a tile is cut into split_k work items. Each item leaves its fp32
// partial tile in an L2-resident workspace; splitk_fixup_kernel (gemm_tc.cu) then adds the partials IN SLICE ORDER
// (bit-reproducible) and applies the plain epilogue on all SMs. (A first version let the tile's last-arriving item do
// that inside this kernel: one SM pulling split_k x 48 - 128 KB through ~10 rounds of L2 latency cost more than the
// main loop it saved -- 29 us instead of 45 at M = 120, no gain at M = 960.)
// kEpiGegluLn / kEpiQkvLn = the GEGLU / QKV_SPLIT epilogues with the LayerNorm fold (ldm_gemm_desc.ln_stats) compiled
// in. (As a runtime branch inside the plain instantiations the fold grew both kernels by a third and slowed them -- with
// or without fold -- by 30 - 50 us per launch at the 48x156 level.)
enum { kEpiStaged = 0, kEpiGeglu = 1, kEpiQkv = 2, kEpiDirect = 3, kEpiConvT = 4, kEpiSplit = 5, kEpiGegluLn = 6, kEpiQkvLn = 7 };

template <bool kPair, int kEpiT>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO,
               const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmE, const GemmParams p) {
  constexpr bool kLn = kEpiT == kEpiGegluLn || kEpiT == kEpiQkvLn;                                     // LayerNorm fold
  constexpr int kEpi = kEpiT == kEpiGegluLn ? (int)kEpiGeglu : kEpiT == kEpiQkvLn ? (int)kEpiQkv : kEpiT;  // epilogue mode
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by SWIZZLE_128B tiles (the dynamic smem base is the same in both CTAs of a pair).
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // an offset, not an integer round trip: the compiler keeps the shared address space (LDS / STS instead of generic LD / ST)
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
  // the set-up above overlapped the previous kernel's tail; from here on its results are needed
  pdl_launch_dependents();
  pdl_wait();

  // work items: (m_unit, n_tile) with m_unit = one M tile, or a pair of consecutive M tiles
  const int m_units = kPair ? (p.m_tiles + 1) / 2 : p.m_tiles;
  const int num_tiles = m_units * p.n_tiles;
  constexpr bool kSplit = kEpi == kEpiSplit;
  const int split = kSplit ? p.split_k : 1;
  const int num_items = num_tiles * split;   // item = tile * split + slice: the slices of a tile run on different SMs
  const int kmain = p.taps * p.kblocks;      // K blocks of a tile without the residual blocks
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const uint32_t a_bytes = (uint32_t)(p.bw * p.bh) * (kBlockK * 2);
  const int b_rows = kPair ? p.block_n / 2 : p.block_n;   // weight rows this CTA loads per k-block
  const uint32_t b_bytes = (uint32_t)b_rows * (kBlockK * 2);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = worker; item < num_items; item += num_workers) {
        const int tile = kSplit ? item / split : item;
        const int slice = kSplit ? item - tile * split : 0;
        const int k_lo = kSplit ? (int)((long long)kmain * slice / split) : 0;
        const int k_hi = kSplit ? (int)((long long)kmain * (slice + 1) / split) : kmain;
        const int n_tile = tile / m_units;
        const int m_unit = tile - n_tile * m_units;
        const int m_tile = kPair ? 2 * m_unit + (int)rank : m_unit;  // == m_tiles for the odd tail: all-OOB box, zeros
        const int b = m_tile / tiles_per_img;
        const int rem = m_tile - b * tiles_per_img;
        const int ty = rem / p.tiles_x;
        const int tx = rem - ty * p.tiles_x;
        const int x0 = tx * p.bw, y0 = ty * p.bh;
        const int n0 = n_tile * p.block_n + (int)rank * b_rows;
        int tap = k_lo / p.kblocks, kb = k_lo - tap * p.kblocks;
        // up2: output parity class of this N tile (row parity, column parity) shifts its 2x2 window
        const int cls = p.up2_ntiles > 0 ? n_tile / p.up2_ntiles : 0;
        const int ax0 = x0 * p.a_stride, ay0 = y0 * p.a_stride;
        for (int ki = k_lo; ki < k_hi; ++ki) {
          const int dy = (p.taps == 9) ? tap / 3 - p.a_pad : (p.taps == 4) ? (tap >> 1) - 1 + (cls >> 1) : 0;
          const int dx = (p.taps == 9) ? tap % 3 - p.a_pad : (p.taps == 4) ? (tap & 1) - 1 + (cls & 1) : 0;
          {
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
              tma_load_4d_pair(sa, tmA, &full_bar[stage], kc, ax0 + dx, ay0 + dy, b);
              tma_load_2d_pair(sb, &tmB, &full_bar[stage], tap * p.ktap + kb * kBlockK, n0);
            } else {
              mbar_arrive_expect_tx(&full_bar[stage], a_bytes + b_bytes);
              tma_load_4d(sa, tmA, &full_bar[stage], kc, ax0 + dx, ay0 + dy, b);
              tma_load_2d(sb, &tmB, &full_bar[stage], tap * p.ktap + kb * kBlockK, n0);
            }
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
          if (++kb == p.kblocks) {
            kb = 0;
            ++tap;
          }
        }
        // residual as extra K blocks: A = residual[rows of this tile, 64 output columns], B = the matching band of the
        // identity matrix (row n of the tile x column n), so that D += R on the tensor core (split-K: the last slice)
        const int nres = (slice == split - 1) ? p.res_kblocks : 0;
        for (int j = 0; j < nres; ++j) {
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
      for (int item = worker; item < num_items; item += num_workers) {
        int ksteps = kmain + p.res_kblocks;
        if (kSplit) {
          const int slice = item % split;
          ksteps = (int)((long long)kmain * (slice + 1) / split) - (int)((long long)kmain * slice / split) +
                   (slice == split - 1 ? p.res_kblocks : 0);
        }
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
              const uint32_t dt = d_tmem;
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
    constexpr bool geglu = kEpi == kEpiGeglu;
    constexpr bool staged = kEpi == kEpiStaged || kEpi == kEpiGeglu;  // bf16 [rows, N] (GEGLU: [rows, N/2]) via smem + TMA store
    // kEpiDirect also serves the rare GEGLU shapes that are too narrow for the staged path (runtime flag there)
    const bool plain = kEpi == kEpiStaged || (kEpi == kEpiDirect && !(p.flags & LDM_GEMM_GEGLU));
    const bool epi_leader = threadIdx.x == 64;  // issues the TMA stores of the staged path
    int sb = 0;                                  // staging buffer of the next output block (alternates)
    int as = 0;
    uint32_t aphase = 0;
    // LayerNorm fold: the row moments of a tile are loaded one tile ahead. This kernel is epilogue-bound at K = 320, all
    // eight epilogue warps reach the loads at the same moment, and an L2 round trip in front of every tile was the whole
    // cost of the fold (ncu: long_scoreboard on the first use; 110 -> 153 us for the GEGLU launch of the 48x156 level).
    // part-major partials [ln_parts][rows]: the 32 lanes of a warp (consecutive rows) read 256 contiguous bytes per part.
    constexpr int kLnPre = 20;  // partials held in registers across a tile (C = 320 / 640: all of them; 1280: half)
    [[maybe_unused]] float2 ln_pre[kLnPre];
    [[maybe_unused]] auto ln_issue = [&](int item2) {
      const int tile2 = kSplit ? item2 / split : item2;
      const int n_tile2 = tile2 / m_units;
      const int m_unit2 = tile2 - n_tile2 * m_units;
      const int m_tile2 = kPair ? 2 * m_unit2 + (int)rank : m_unit2;
      const int b2 = m_tile2 / tiles_per_img;
      const int rem2 = m_tile2 - b2 * tiles_per_img;
      const int ty2 = rem2 / p.tiles_x;
      const int tx2 = rem2 - ty2 * p.tiles_x;
      const int ly2 = r / p.bw, lx2 = r - ly2 * p.bw;
      const int y2 = ty2 * p.bh + ly2, x2 = tx2 * p.bw + lx2;
      const bool valid2 = (m_tile2 < p.m_tiles) && (r < p.bw * p.bh) && (y2 < p.H) && (x2 < p.W);
      const float2* sp = reinterpret_cast<const float2*>(p.ln_stats) + (((long long)b2 * p.H + y2) * p.W + x2);
#pragma unroll
      for (int j = 0; j < kLnPre; ++j)
        ln_pre[j] = (valid2 && j < p.ln_parts) ? __ldg(sp + (long long)j * p.rows_total) : make_float2(0.f, 0.f);
    };
    if constexpr (kLn) {
      if (worker < num_items) ln_issue(worker);
    }
    for (int item = worker; item < num_items; item += num_workers) {
      const int tile = kSplit ? item / split : item;
      const int slice = kSplit ? item - tile * split : 0;
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
      int qkv_bi = 0, qkv_s = 0;  // QKV_SPLIT: image and token of this thread's row
      // vectorised V^T stores need 8-token groups that never straddle an image and rows that are consecutive tokens
      const bool qkv_vec = kEpi == kEpiQkv && (p.seq % 8 == 0) && p.bh == 1 && p.bw == kBlockM;
      if (kEpi == kEpiQkv) {
        qkv_bi = (int)(((unsigned long long)grow * p.magic_seq) >> 40);
        qkv_s = (int)(grow - (long long)qkv_bi * p.seq);
      }

      // LayerNorm fold: out = rstd (acc - mean g[n]) + b'[n] = ln_a acc + ln_c g[n] + b'[n] with this row's moments
      float ln_a = 1.f, ln_c = 0.f;
      constexpr bool ln_fold = kLn;
      if constexpr (ln_fold) {
        float su = 0.f, sq = 0.f;
#pragma unroll
        for (int j = 0; j < kLnPre; ++j) {  // fixed order: bit-reproducible (the slots past ln_parts hold zeros)
          su += ln_pre[j].x;
          sq += ln_pre[j].y;
        }
        if (valid) {
          const float2* sp = reinterpret_cast<const float2*>(p.ln_stats) + grow;
          for (int i = kLnPre; i < p.ln_parts; i += 10) {  // C = 1280: the second half, ten loads in flight at once
            float2 t[10];
#pragma unroll
            for (int j = 0; j < 10; ++j)
              t[j] = i + j < p.ln_parts ? __ldg(sp + (long long)(i + j) * p.rows_total) : make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 10; ++j) {
              su += t[j].x;
              sq += t[j].y;
            }
          }
        }
        const float mean = su * p.ln_invc;
        ln_a = rsqrtf(fmaxf(fmaf(-mean, mean, sq * p.ln_invc), 0.f) + p.ln_fold_eps);
        ln_c = -mean * ln_a;
        if (item + num_workers < num_items) ln_issue(item + num_workers);  // in flight during this tile's epilogue
      }
      // QKV_SPLIT with the fold: the tile's folded bias and column sums go through shared memory once (the epilogue
      // bounds this kernel at K = 320; sixteen broadcast loads per 32-column chunk behind the TMEM wait were not free)
      const bool ln_smem = kEpi == kEpiQkv && ln_fold && p.block_n <= 160;  // (wider tiles: broadcast loads)
      float* sLn = reinterpret_cast<float*>(epi_smem + 2 * 16384);  // [0, 160) bias', [160, 320) g
      if (ln_smem) {
        named_bar_sync(1, 32 * kEpiWarps);  // the previous tile's readers are done
        const int t = threadIdx.x - 64;
        if (t < p.block_n) {
          const int n = n0 + t;
          sLn[t] = (p.bias && n < p.N) ? __ldg(p.bias + n) : 0.f;
          sLn[160 + t] = n < p.N ? __ldg(p.ln_colsum + n) : 0.f;
        }
        named_bar_sync(1, 32 * kEpiWarps);
      }
      // residual rows of this tile: issue every load now, so that they are in flight while the MMAs finish
      // staged path: both halves work on the same 64-column output block, half h on its 32-column chunk h, so chunk
      // ci of a thread covers columns min(ci*64, block_n-64) + h*32; direct path: a half owns every other chunk.
      uint4 rs[4][4];
      const bool use_res = !kSplit && plain && p.residual != nullptr && valid && p.res_kblocks == 0;
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

      // staged epilogues without the fold: the tile's bias (+ per-image bias) goes to shared memory ONCE per tile and
      // before the accumulator wait (per 64-column block it cost a global round trip behind every block's barrier).
      // Safe without a barrier of its own: every read of the previous tile's values lies before that tile's last
      // block barrier, and the first block barrier below publishes these.
      constexpr bool tile_bias = staged && !kSplit && !kLn;
      if constexpr (tile_bias) {
        float* sbias = reinterpret_cast<float*>(epi_smem + 2 * 16384);
        const int t = threadIdx.x - 64;
        if (t < p.block_n) {
          const int n = n0 + t;
          float bv = 0.f;
          if (n < p.N) {
            if (p.bias) bv = __ldg(p.bias + n);
            if (p.rowbias) bv += __ldg(p.rowbias + (long long)b * p.N + n);
          }
          sbias[t] = bv;
        }
      }
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)as * 256u;

      // split-K: this item's fp32 partial tile goes to the workspace, column-major inside the tile (the 32 lanes of a
      // warp = 32 consecutive rows write 128 contiguous bytes per column); the fix-up kernel does the rest
      if constexpr (kSplit) {
        const long long tile_off = ((long long)m_tile * p.n_tiles + n_tile) * (long long)(p.block_n * kBlockM);
        float* wdst = p.ws + (long long)slice * p.ws_slice_stride + tile_off + r;
        for (int c = half * 32; c < p.block_n; c += 64) {
          uint32_t v[32];
          tmem_ld32(t_addr + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) wdst[(c + j) * kBlockM] = __uint_as_float(v[j]);
        }
      }

      if constexpr (kSplit) {
      } else if constexpr (kEpi == kEpiConvT) {
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
      } else if constexpr (staged) {
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
          // (single-buffered: the next block's values are written after this block's second barrier, behind every read)
          float* sbias = reinterpret_cast<float*>(epi_smem + 2 * 16384);
          float* sgsum = sbias + 128;  // LN fold: g[n] of the same columns
          if (epi_leader) bulk_wait_read1();  // the store that last used this buffer (two blocks ago) has been read
          if constexpr (!tile_bias) {
            // (fold: folded bias and column sums of this block; 2 x block_n floats do not fit the slice)
            const int t = threadIdx.x - 64;
            if (t < acc_w) {
              const int n = n0 + c0 + t;
              float bv = 0.f;
              if (n < p.N) {
                if (p.bias) bv = __ldg(p.bias + n);
                if (p.rowbias) bv += __ldg(p.rowbias + (long long)b * p.N + n);
              }
              sbias[t] = bv;
              if (ln_fold) sgsum[t] = n < p.N ? __ldg(p.ln_colsum + n) : 0.f;
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
            const int nc = n0 + c;
            float f[32];
            {
              const float4* bp = reinterpret_cast<const float4*>(sbias + (tile_bias ? c : cc));
              if (ln_fold) {
                tmem_ld_wait();
                const float4* gp = reinterpret_cast<const float4*>(sgsum + cc);
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                  const float4 bv = bp[g], gv = gp[g];
                  f[g * 4] = fmaf(ln_a, __uint_as_float(v[g * 4]), fmaf(ln_c, gv.x, bv.x));
                  f[g * 4 + 1] = fmaf(ln_a, __uint_as_float(v[g * 4 + 1]), fmaf(ln_c, gv.y, bv.y));
                  f[g * 4 + 2] = fmaf(ln_a, __uint_as_float(v[g * 4 + 2]), fmaf(ln_c, gv.z, bv.z));
                  f[g * 4 + 3] = fmaf(ln_a, __uint_as_float(v[g * 4 + 3]), fmaf(ln_c, gv.w, bv.w));
                }
              } else {
                float4 bv[8];  // shared-memory reads under the TMEM load
#pragma unroll
                for (int g = 0; g < 8; ++g) bv[g] = bp[g];
                tmem_ld_wait();
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                  f[g * 4] = __uint_as_float(v[g * 4]) + bv[g].x;
                  f[g * 4 + 1] = __uint_as_float(v[g * 4 + 1]) + bv[g].y;
                  f[g * 4 + 2] = __uint_as_float(v[g * 4 + 2]) + bv[g].z;
                  f[g * 4 + 3] = __uint_as_float(v[g * 4 + 3]) + bv[g].w;
                }
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
              float st_s = 0.f, st_q = 0.f;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                uint4 u;
                u.x = pack_bf16(f[g * 8 + 0], f[g * 8 + 1]); u.y = pack_bf16(f[g * 8 + 2], f[g * 8 + 3]);
                u.z = pack_bf16(f[g * 8 + 4], f[g * 8 + 5]); u.w = pack_bf16(f[g * 8 + 6], f[g * 8 + 7]);
                *reinterpret_cast<uint4*>(rowp + (((half * 4 + g) ^ (r & 7)) << 4)) = u;
                if (kEpi == kEpiStaged && p.row_stats) {  // moments of the values as the consumer will read them (bf16)
                  const float2 a0 = unpack_bf16(u.x), a1 = unpack_bf16(u.y), a2 = unpack_bf16(u.z), a3 = unpack_bf16(u.w);
                  st_s += ((a0.x + a0.y) + (a1.x + a1.y)) + ((a2.x + a2.y) + (a3.x + a3.y));
                  st_q = fmaf(a0.x, a0.x, st_q); st_q = fmaf(a0.y, a0.y, st_q); st_q = fmaf(a1.x, a1.x, st_q);
                  st_q = fmaf(a1.y, a1.y, st_q); st_q = fmaf(a2.x, a2.x, st_q); st_q = fmaf(a2.y, a2.y, st_q);
                  st_q = fmaf(a3.x, a3.x, st_q); st_q = fmaf(a3.y, a3.y, st_q);
                }
              }
              // one partial per 32-column chunk and row (part = column / 32: independent of block_n; the overlap
              // columns of a shifted last block are written twice with the same value)
              if (kEpi == kEpiStaged && p.row_stats && valid && nc < p.N)
                *reinterpret_cast<float2*>(p.row_stats + ((long long)(nc >> 5) * p.rows_total + grow) * 2) =
                    make_float2(st_s, st_q);
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(1, 32 * kEpiWarps);
          if (epi_leader) {
            if (p.up2_ntiles > 0) {
              // this tile's pixels (y, x) are the output pixels (2y + a, 2x + b) of its parity class: tmO steps by two
              const int cls = n_tile / p.up2_ntiles;
              tma_store_4d(&tmO, sbuf, n0 + c0 - cls * p.up2_cout, 2 * tx * p.bw + (cls & 1), 2 * ty * p.bh + (cls >> 1), b);
            } else {
              const int ncol = geglu ? (n0 + c0) / 2 : n0 + c0;
              tma_store_4d(&tmO, sbuf, ncol, tx * p.bw, ty * p.bh, b);
            }
            bulk_commit();
          }
          sb ^= 1;
        }
      } else {
        bool qkv_done = false;
        if constexpr (kEpi == kEpiQkv) {
          // q / k tiles through shared memory + TMA stores. block_n is a multiple of head_dim and divides C here, so
          // the tile holds whole heads of ONE part. Per head and 64-column block of its padded row (head_dim data
          // columns, then zeros up to dpad) the eight warps stage a [128 tokens x 64] swizzled block and one thread
          // stores it into q / k [B*heads, seq, dpad] (map {dpad, seq, heads, B}). Full 128-byte lines instead of the
          // row-per-thread 16-byte stores (26.7 sectors per request, epilogue 3.5x longer than the K = 320 main loop).
          const int C = p.heads * p.head_dim;
          const int part_in = n0 / C;
          const int which_t = part_in + p.qkv_part0;
          const long long grow0 = (long long)m_tile * kBlockM;  // flattened rows: bh == 1, bw == kBlockM
          const int bi0 = (int)(((unsigned long long)grow0 * p.magic_seq) >> 40);
          const int s0 = (int)(grow0 - (long long)bi0 * p.seq);
          const int nimg = (int)(((long long)p.W * p.H * p.B) / p.seq);
          // (a tile whose 128 rows run into the next image keeps the row-per-thread stores below: one tile per image)
          if ((p.flags & kQkvStaged) && which_t < 2 && (s0 + kBlockM <= p.seq || bi0 + 1 >= nimg)) {
            const int head0 = (n0 - part_in * C) / p.head_dim;
            const int heads_in_tile = p.block_n / p.head_dim;
            const int blocks_per_head = p.dpad >> 6;
            for (int hd = 0; hd < heads_in_tile; ++hd) {
              for (int kb = 0; kb < blocks_per_head; ++kb) {
                uint8_t* sbuf = epi_smem + sb * 16384;
                if (epi_leader) bulk_wait_read1();  // the store that last used this buffer has been read
                named_bar_sync(1, 32 * kEpiWarps);
                const int cbase = hd * p.head_dim + kb * 64 + half * 32;  // accumulator column of this thread's chunk
                int ng = ((hd + 1) * p.head_dim - cbase) >> 3;             // 8-column groups that hold data (warp-uniform)
                ng = ng < 0 ? 0 : (ng > 4 ? 4 : ng);
                uint32_t v[32];
#pragma unroll
                for (int g = 0; g < 4; ++g)
                  if (g < ng) tmem_ld8(t_addr + cbase + g * 8, v + g * 8);
                tmem_ld_wait();
                if (ln_fold) {
#pragma unroll
                  for (int g = 0; g < 4; ++g)
                    if (g < ng) {
                      const int cl = cbase + g * 8;  // column inside the tile
                      float4 g0, g1, b0, b1;
                      if (ln_smem) {
                        b0 = *reinterpret_cast<const float4*>(sLn + cl);
                        b1 = *reinterpret_cast<const float4*>(sLn + cl + 4);
                        g0 = *reinterpret_cast<const float4*>(sLn + 160 + cl);
                        g1 = *reinterpret_cast<const float4*>(sLn + 160 + cl + 4);
                      } else {
                        const int n = n0 + cl;
                        g0 = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + n));
                        g1 = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + n + 4));
                        b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
                        if (p.bias) {
                          b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
                          b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4));
                        }
                      }
                      const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                      for (int j = 0; j < 8; ++j)
                        v[g * 8 + j] = __float_as_uint(fmaf(ln_a, __uint_as_float(v[g * 8 + j]), fmaf(ln_c, gv[j], bv[j])));
                    }
                }
                uint8_t* rowp = sbuf + r * 128;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                  uint4 u = make_uint4(0, 0, 0, 0);
                  if (g < ng) {
                    u.x = pack_bf16(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1]));
                    u.y = pack_bf16(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3]));
                    u.z = pack_bf16(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5]));
                    u.w = pack_bf16(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7]));
                  }
                  *reinterpret_cast<uint4*>(rowp + (((half * 4 + g) ^ (r & 7)) << 4)) = u;
                }
                fence_proxy_async_smem();
                named_bar_sync(1, 32 * kEpiWarps);
                if (epi_leader) {  // rows past the end of the last image are clipped by the TMA unit
                  if (which_t == 0) tma_store_4d(&tmO, sbuf, kb * 64, s0, head0 + hd, bi0);
                  else tma_store_4d(&tmR, sbuf, kb * 64, s0, head0 + hd, bi0);
                  bulk_commit();
                }
                sb ^= 1;
              }
            }
            qkv_done = true;
          }
          if ((p.flags & kQkvStaged) && !qkv_done) {
            // This tile takes the direct path, whose V^T transpose scratch (sT below) lies inside staging buffer 0: a
            // TMA store of the previous (q / k) tile may still be reading that buffer. Wait for the reads, all warps.
            if (epi_leader) bulk_wait_read0();
            named_bar_sync(1, 32 * kEpiWarps);
          }
        }
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          if (qkv_done) break;
          const int c = (2 * ci + half) * 32;
          if (c >= p.block_n) break;
          uint32_t v[32];
          tmem_ld32(t_addr + c, v);
          tmem_ld_wait();
          if (!valid && kEpi != kEpiQkv) continue;  // (QKV: the V transpose below needs every lane of the warp)
          const int nc = n0 + c;
          if (nc >= p.N) continue;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (ln_smem) {  // rstd (acc - mean g[n]) + b'[n], both column vectors from shared memory
            const float4* bp = reinterpret_cast<const float4*>(sLn + c);
            const float4* gp = reinterpret_cast<const float4*>(sLn + 160 + c);
#pragma unroll
            for (int g = 0; g < 8; ++g)
              if (nc + g * 4 < p.N) {
                const float4 gv = gp[g], bv = bp[g];
                f[g * 4] = fmaf(ln_a, f[g * 4], fmaf(ln_c, gv.x, bv.x));
                f[g * 4 + 1] = fmaf(ln_a, f[g * 4 + 1], fmaf(ln_c, gv.y, bv.y));
                f[g * 4 + 2] = fmaf(ln_a, f[g * 4 + 2], fmaf(ln_c, gv.z, bv.z));
                f[g * 4 + 3] = fmaf(ln_a, f[g * 4 + 3], fmaf(ln_c, gv.w, bv.w));
              }
          } else if (ln_fold) {  // rstd (acc - mean g[n]); the folded bias b' follows below
            const float4* gp = reinterpret_cast<const float4*>(p.ln_colsum + nc);
#pragma unroll
            for (int g = 0; g < 8; ++g)
              if (nc + g * 4 < p.N) {
                const float4 gv = __ldg(gp + g);
                f[g * 4] = fmaf(ln_a, f[g * 4], ln_c * gv.x); f[g * 4 + 1] = fmaf(ln_a, f[g * 4 + 1], ln_c * gv.y);
                f[g * 4 + 2] = fmaf(ln_a, f[g * 4 + 2], ln_c * gv.z); f[g * 4 + 3] = fmaf(ln_a, f[g * 4 + 3], ln_c * gv.w);
              }
          }
          if (p.bias && !ln_smem) {  // N % 8 == 0 (GEGLU: % 32): 4-wide groups are all-in or all-out
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
          if constexpr (kEpi == kEpiQkv) {
            // column -> (q|k|v, head, e) and row -> (image, token) with multiply-shift divisions (exact for the
            // operand ranges checked on the host); the column part is warp-uniform, the row part is per tile
            const int C = p.heads * p.head_dim;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int n = nc + g * 8;
              if (n >= p.N) break;
              const int part = (int)(((unsigned long long)(unsigned)n * p.magic_c) >> 40);
              const int cc = n - part * C;
              const int which = part + p.qkv_part0;
              const int head = (int)(((unsigned long long)(unsigned)cc * p.magic_d) >> 40);
              const int e = cc - head * p.head_dim;
              const long long bh = (long long)qkv_bi * p.heads + head;
              if (which < 2) {
                if (valid) {
                  __nv_bfloat16* dst = (which == 0 ? p.q : p.k) + (bh * p.seq + qkv_s) * p.dpad + e;
                  store_bf16x8(dst, f + g * 8);
                }
              } else if (qkv_vec) {
                // V^T through a per-warp 8 x 32 transpose in shared memory: lane (c, tq) then stores channel e + c of
                // the 8 tokens [8 tq, 8 tq + 8) of this warp's rows with ONE 16-byte store. (Eight 2-byte stores per lane
                // at a row stride of seq_pad cost ~17 cycles each and made this epilogue 4x longer than its main loop.)
                __nv_bfloat16* sT = reinterpret_cast<__nv_bfloat16*>(epi_smem) + (warp - 2) * 256;
#pragma unroll
                for (int j = 0; j < 8; ++j) sT[j * 32 + lane] = __float2bfloat16_rn(f[g * 8 + j]);
                __syncwarp();
                const int cch = lane >> 2, tq = lane & 3;
                const uint4 u = *reinterpret_cast<const uint4*>(sT + cch * 32 + tq * 8);
                __syncwarp();
                // the 8-token group starts at a row that is a multiple of 8 and seq % 8 == 0: one image, all valid or none
                const int g_bi = __shfl_sync(0xffffffffu, qkv_bi, tq * 8), g_s = __shfl_sync(0xffffffffu, qkv_s, tq * 8);
                const int g_ok = __shfl_sync(0xffffffffu, (int)valid, tq * 8);
                if (g_ok) {
                  __nv_bfloat16* dst =
                      p.vt + (((long long)g_bi * p.heads + head) * p.vt_rows + e + cch) * p.seq_pad + g_s;
                  *reinterpret_cast<uint4*>(dst) = u;
                }
              } else if (valid) {
                __nv_bfloat16* dst = p.vt + (bh * p.vt_rows + e) * p.seq_pad + qkv_s;
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

}  // namespace ldm_gemm
