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

#include "gemm_tc_kernel.cuh"

namespace ldm_gemm {
#define LDM_GEMM_DECLARE(NAME)                                                                                      \
  cudaError_t NAME(bool pair, int grid, int smem_bytes, cudaStream_t stream, const CUtensorMap& tmA1,               \
                   const CUtensorMap& tmA2, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmR, \
                   const CUtensorMap& tmE, const GemmParams& p);
LDM_GEMM_DECLARE(launch_gemm_staged)
LDM_GEMM_DECLARE(launch_gemm_geglu)
LDM_GEMM_DECLARE(launch_gemm_qkv)
LDM_GEMM_DECLARE(launch_gemm_direct)
LDM_GEMM_DECLARE(launch_gemm_convt)
LDM_GEMM_DECLARE(launch_gemm_split)
LDM_GEMM_DECLARE(launch_gemm_geglu_ln)
LDM_GEMM_DECLARE(launch_gemm_qkv_ln)
}  // namespace ldm_gemm

namespace {

using namespace ldm;
using namespace ldm_gemm;

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
// Split-K (max_split > 1: the caller has a workspace and a plain bf16 epilogue): each tile becomes `split` work items
// over contiguous K ranges, which fills the SMs when there are few tiles and lets the wide, cheap-per-column block_n
// run at small M. An item then writes its 128 x block_n fp32 partial (~8 cycles per column) and a second, small launch
// (splitk_fixup_kernel: ~3 us of launch + L2 latency, then the partials at L2 bandwidth) finishes the tiles. Adopted
// only when the model sees a clear gain (12 %): it was fitted on few shapes (tools/bench_gemm_shapes.py).
int pick_block_n(int N, long m_tiles, long kblocks, int sms, bool pair, int taps, double* cost_out, int max_split = 1,
                 long ws_floats = 0, int* split_out = nullptr) {
  static const int cands[] = {256, 224, 192, 160, 128, 96};
  static const double mma_single[] = {167, 158, 151, 141, 111, 106};
  static const double mma_pair[] = {160, 149, 136, 124, 99, 92};
  static const int splits[] = {1, 2, 3, 4, 5, 6, 8, 10, 12, 16};
  int best = 256, best_s = 256, best_split = 1;
  double best_cost = -1.0, best_cost_s = -1.0;
  const long m_units = pair ? (m_tiles + 1) / 2 : m_tiles;
  const long m_alloc = pair ? 2 * m_units : m_units;
  const long workers = pair ? sms / 2 : sms;
  for (int i = 0; i < 6; ++i) {
    const int c = cands[i];
    if (c == 224 && taps != 9) continue;  // 224 only pays on the long-K convolutions (and is erratic on short K)
    const long n_tiles = (N + c - 1) / c;
    for (int si = 0; si < 10; ++si) {
      const int sp = splits[si];
      if (sp > max_split) break;
      if (sp > 1) {
        if (kblocks / sp < (pair ? 24 : 8)) break;  // an item should still be a pipeline-filling main loop
        if ((long)sp * m_alloc * n_tiles * c * 128 > ws_floats) break;
      }
      const long waves = (m_units * n_tiles * sp + workers - 1) / workers;
      const double item = (double)((kblocks + sp - 1) / sp) * 4.0 * (pair ? mma_pair[i] : mma_single[i]) +
                          (pair ? 3300.0 + 11.0 * c : 1000.0 + 17.0 * c) + (sp > 1 ? 8.0 * c : 0.0);
      double cost = (double)waves * item;
      if (sp > 1) cost += 6000.0 + (double)sp * m_alloc * n_tiles * c * 128.0 * 4.0 / 2000.0;
      if (sp == 1) {
        if (best_cost < 0 || cost < best_cost * 0.999) {
          best_cost = cost;
          best = c;
        }
      } else if (best_cost_s < 0 || cost < best_cost_s * 0.999) {
        best_cost_s = cost;
        best_s = c;
        best_split = sp;
      }
    }
  }
  if (best_cost_s > 0 && best_cost_s < 0.88 * best_cost) {
    best_cost = best_cost_s;
    best = best_s;
  } else {
    best_split = 1;
  }
  if (cost_out) *cost_out = best_cost;
  if (split_out) *split_out = best_split;
  return best;
}

int debug_flags() {
  using ldm_host::diag_env_has;
  int v = 0;
  if (diag_env_has("LDM_GEMM_DEBUG", "notma")) v |= kDbgNoTma;
  if (diag_env_has("LDM_GEMM_DEBUG", "nomma")) v |= kDbgNoMma;
  if (diag_env_has("LDM_GEMM_DEBUG", "nowait")) v |= kDbgNoWait | kDbgNoTma;
  if (diag_env_has("LDM_GEMM_DEBUG", "nofence")) v |= kDbgNoFence;
  if (diag_env_has("LDM_GEMM_DEBUG", "keepcommit")) v |= kDbgKeepCommit;
  return v;
}

// Residual on the tensor core for main loops of up to this many K blocks (LDM_GEMM_RESMMA overrides, 0 = never)
int resmma_max_kblocks() { return ldm_host::diag_env("LDM_GEMM_RESMMA", 1 << 30); }

// LDM_GEMM_STAGED=0 keeps the direct (row-per-thread) global stores (A/B timing)
bool staged_enabled() { return ldm_host::diag_env("LDM_GEMM_STAGED", 1) != 0; }

// LDM_GEMM_SPLITK=0 keeps every tile on one work item (A/B timing)
bool splitk_enabled() { return ldm_host::diag_env("LDM_GEMM_SPLITK", 1) != 0; }

thread_local int g_last_cfg[3] = {0, 0, 1};  // block_n, pair, split_k of this thread's last ldm_gemm_bf16 launch

constexpr size_t kSplitCounterBytes = 0;

// Second half of a split-K launch: out = epilogue(sum over the slices, in slice order, of the fp32 partial tiles).
// One CTA per (tile, 16 output columns): the partials are tile-column-major (lanes = rows: coalesced), the sums cross a
// small shared-memory transpose, and every row leaves as two 16-byte vectors (one full 32-byte sector).
struct FixupParams {
  const float* ws;
  long long slice_stride;
  int split, tiles_x, tiles_y, bw, bh, W, H, m_tiles, n_tiles, block_n, N, silu;
  const float* bias;
  const float* rowbias;
  const __nv_bfloat16* residual;
  __nv_bfloat16* out;
};

__global__ void __launch_bounds__(256) splitk_fixup_kernel(const FixupParams p) {
  __shared__ float tile[kBlockM][17];
  pdl_launch_dependents();
  pdl_wait();
  const int cb = blockIdx.x, n_tile = blockIdx.y, m_tile = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long tile_off = ((long long)m_tile * p.n_tiles + n_tile) * (long long)(p.block_n * kBlockM);
  {
    // thread: columns cb*16 + warp*2 + {0, 1}, rows lane + 32 i. Four slices of loads in flight, added in slice order.
    float a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = 0.f;
    const float* src = p.ws + tile_off + (long long)(cb * 16 + warp * 2) * kBlockM + lane;
    for (int s0 = 0; s0 < p.split; s0 += 4) {
      float v[4][8];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int k = 0; k < 8; ++k)
          v[q][k] = s0 + q < p.split ? __ldcg(src + (long long)(s0 + q) * p.slice_stride + (k >> 2) * kBlockM + (k & 3) * 32) : 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] += v[q][k];  // (+ 0.0f past the last slice: exact)
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) tile[(k & 3) * 32 + lane][warp * 2 + (k >> 2)] = a[k];
  }
  __syncthreads();
  const int r = threadIdx.x >> 1, vec = threadIdx.x & 1;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int b = m_tile / tiles_per_img;
  const int rem = m_tile - b * tiles_per_img;
  const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
  const int ly = r / p.bw, lx = r - ly * p.bw;
  const int y = ty * p.bh + ly, x = tx * p.bw + lx;
  const int n = n_tile * p.block_n + cb * 16 + vec * 8;
  if (r >= p.bw * p.bh || y >= p.H || x >= p.W || n >= p.N) return;
  const long long grow = ((long long)b * p.H + y) * p.W + x;
  float f[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = tile[r][vec * 8 + j];
  if (p.bias) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n)), b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4));
    f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w; f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
  }
  if (p.rowbias) {
    const float* rb = p.rowbias + (long long)b * p.N + n;
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(rb)), b1 = __ldg(reinterpret_cast<const float4*>(rb + 4));
    f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w; f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
  }
  if (p.residual) {
    const uint4 u = ld_nc_v4(p.residual + grow * p.N + n);
    const float2 a0 = unpack_bf16(u.x), a1 = unpack_bf16(u.y), a2 = unpack_bf16(u.z), a3 = unpack_bf16(u.w);
    f[0] += a0.x; f[1] += a0.y; f[2] += a1.x; f[3] += a1.y; f[4] += a2.x; f[5] += a2.y; f[6] += a3.x; f[7] += a3.y;
  }
  if (p.silu) {
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = silu_f(f[j]);
  }
  store_bf16x8(p.out + grow * p.N + n, f);
}

// LDM_GEMM_PAIR=0 / 1 forces the single-CTA / CTA-pair kernel (A/B timing); default: the cost model decides.
int pair_override() { return ldm_host::diag_env("LDM_GEMM_PAIR", -1); }

}  // namespace

extern "C" int ldm_gemm_bf16(const ldm_gemm_desc* d, ldm_stream_t stream) {
  using namespace ldm_host;
  LDM_REQUIRE(d != nullptr, LDM_ERR_BAD_ARG, "ldm_gemm_bf16: null descriptor");
  LDM_REQUIRE(d->a1 && d->w, LDM_ERR_BAD_ARG, "ldm_gemm_bf16: null a1/w");
  const bool up2 = d->up2 != 0;
  const int a_stride = d->a_stride > 1 ? d->a_stride : 1;
  LDM_REQUIRE(d->taps == 1 || d->taps == 9 || (d->taps == 4 && up2), LDM_ERR_BAD_ARG,
              "ldm_gemm_bf16: taps must be 1 or 9, or 4 with up2 (got %d)", d->taps);
  LDM_REQUIRE(!up2 || (d->taps == 4 && a_stride == 1), LDM_ERR_BAD_ARG, "ldm_gemm_bf16: up2 needs taps = 4 and a dense A");
  LDM_REQUIRE(a_stride == 1 || (a_stride == 2 && d->taps == 9 && (d->a_pad == 0 || d->a_pad == 1) && d->a_H > 0 && d->a_W > 0),
              LDM_ERR_BAD_ARG, "ldm_gemm_bf16: a_stride = 2 needs taps = 9, a_pad in {0, 1} and the input extents a_H, a_W");
  LDM_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->N > 0 && d->c1 > 0, LDM_ERR_BAD_SHAPE,
              "ldm_gemm_bf16: non-positive extent B=%d H=%d W=%d N=%d c1=%d", d->B, d->H, d->W, d->N, d->c1);
  const int c2 = d->a2 ? d->c2 : 0;
  LDM_REQUIRE(d->c1 % 8 == 0 && c2 % 8 == 0, LDM_ERR_ALIGNMENT, "ldm_gemm_bf16: channels must be multiples of 8");
  if (d->taps != 1 || d->a2)
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
  // split-K needs the caller's workspace and the plain bf16 [rows, N] epilogue (checked again below)
  const bool split_ok = d->splitk_ws && d->splitk_ws_bytes > 0 && d->block_n <= 0 && !d->row_stats_out && !up2 &&
                        !(flags & (LDM_GEMM_OUT_F32 | LDM_GEMM_OUT_NCHW_F32 | LDM_GEMM_QKV_SPLIT |
                                   LDM_GEMM_CONVT_LN_SILU | LDM_GEMM_GEGLU)) &&
                        d->N >= 64 && d->N % 8 == 0 && staged_enabled() && splitk_enabled() &&
                        (reinterpret_cast<uintptr_t>(d->splitk_ws) & 15) == 0;
  const long ws_floats = split_ok ? (long)((d->splitk_ws_bytes - kSplitCounterBytes) / 4) : 0;
  int sp1 = 1, sp2 = 1;
  const int bn1 = pick_block_n(d->N, p.m_tiles, kblocks_total, num_sms(), false, d->taps, &cost1, split_ok ? 16 : 1,
                               ws_floats, &sp1);
  const int bn2 = pick_block_n(d->N, p.m_tiles, kblocks_total, num_sms(), true, d->taps, &cost2, split_ok ? 16 : 1,
                               ws_floats, &sp2);
  // an explicit block_n keeps the single-CTA kernel (unless the A/B override forces pairs on a pairable block_n)
  const bool pair_ok = p.m_tiles >= 2 && !(flags & LDM_GEMM_CONVT_LN_SILU) &&
                       (d->block_n <= 0 || (pair_override() == 1 && d->block_n >= 64 && d->block_n % 32 == 0));
  bool pair = pair_ok && cost2 < cost1 && kblocks_total >= 32;  // the pair's hand-offs only pay on long main loops
  if (pair_override() >= 0) pair = pair_ok && pair_override() != 0;  // LDM_GEMM_PAIR=-1: model decides
  // QKV head split: with block_n = 160 (the only multiple of the 32-column TMEM chunk that 40-, 80- and 160-wide heads
  // all divide) every tile holds whole heads of one part, and the q / k tiles can leave through TMA stores
  // (LDM_GEMM_QKV_TMA=0: the row-per-thread stores, A/B timing)
  const bool qkv_tma_enabled = diag_env("LDM_GEMM_QKV_TMA", 1) != 0;
  bool qkv_tma = false;
  if ((flags & LDM_GEMM_QKV_SPLIT) && qkv_tma_enabled && d->block_n <= 0 && (!d->bias || d->ln_stats) && !d->rowbias &&
      d->head_dim > 0 &&
      160 % d->head_dim == 0 && (d->heads * d->head_dim) % 160 == 0 && p.H == 1 && p.bw == kBlockM && p.bh == 1 &&
      d->dpad == ((d->head_dim + 63) / 64) * 64 && d->seq >= kBlockM && kblocks_total <= 6) {
    // only where the epilogue, not the main loop, bounds the tile (K <= 384: the 48x156 level). Measured: 105 -> 73 us
    // there, but 46 -> 50 us at K = 640, where block_n = 256 has the cheaper main loop per column.
    qkv_tma = true;
    pair = false;
  }
  int block_n = d->block_n > 0 ? d->block_n : (qkv_tma ? 160 : (pair ? bn2 : bn1));
  int split_k = d->block_n > 0 || qkv_tma ? 1 : (pair ? sp2 : sp1);
  if ((flags & LDM_GEMM_GEGLU) && d->block_n <= 0 && block_n < 128) block_n = 128;  // staged GEGLU blocks span 128 columns
  int up2_cout = 0;
  if (up2) {
    // an N tile lies inside one parity class: block_n must divide cout (the widest that does, at most the model's choice)
    LDM_REQUIRE(d->N % 4 == 0 && (d->N / 4) % 64 == 0 && d->out_H == 2 * d->H && d->out_W == 2 * d->W && !d->residual &&
                    !d->rowbias && !d->row_stats_out && !d->a2 &&
                    !(flags & ~(LDM_GEMM_SILU)),
                LDM_ERR_BAD_ARG, "ldm_gemm_bf16: up2 needs N = 4 * cout (cout %% 64 == 0), out_H = 2H, out_W = 2W, "
                "one source and the plain bf16 epilogue");
    up2_cout = d->N / 4;
    if (d->block_n <= 0) {
      int bn = block_n;
      while (bn > 64 && up2_cout % bn != 0) bn -= 32;
      block_n = bn;
    }
    LDM_REQUIRE(up2_cout % block_n == 0 && block_n >= 64, LDM_ERR_BAD_ARG, "ldm_gemm_bf16: up2 block_n=%d does not divide cout=%d",
                block_n, up2_cout);
  }
  LDM_REQUIRE(block_n % 32 == 0 && block_n >= 32 && block_n <= 256, LDM_ERR_BAD_ARG, "ldm_gemm_bf16: block_n=%d",
              block_n);
  if (flags & LDM_GEMM_GEGLU) LDM_REQUIRE(d->N % 32 == 0, LDM_ERR_BAD_SHAPE, "GEGLU needs N %% 32 == 0");
  if (flags & LDM_GEMM_QKV_SPLIT) {
    const int Cqkv = d->heads * d->head_dim;
    const int nparts = Cqkv > 0 ? d->N / Cqkv : 0;
    const int part0 = d->qkv_part0;
    LDM_REQUIRE(d->heads > 0 && d->head_dim % 8 == 0 && d->N == nparts * Cqkv && nparts >= 1 && part0 >= 0 &&
                    part0 + nparts <= 3 && d->seq > 0 && d->seq_pad % 8 == 0 && d->dpad % 64 == 0 &&
                    (long long)d->B * d->H * d->W % d->seq == 0,
                LDM_ERR_BAD_SHAPE, "ldm_gemm_bf16: bad QKV split geometry");
    LDM_REQUIRE((part0 > 0 || d->q) && (part0 > 1 || part0 + nparts < 2 || d->k) && (part0 + nparts < 3 || d->vt),
                LDM_ERR_BAD_ARG, "ldm_gemm_bf16: QKV split needs the q / k / vt buffers of the parts it writes");
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
  p.stages = (kSmemBudget - (1024 + 256) /* alignment slack + barriers, see smem_bytes below */ - kEpiStageBytes) / p.stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  LDM_REQUIRE(p.stages >= 2, LDM_ERR_BAD_SHAPE, "ldm_gemm_bf16: not enough shared memory for 2 stages");
  // bf16 [rows, N] outputs (and GEGLU's [rows, N/2]) leave through shared memory + TMA store
  const int n_out = (flags & LDM_GEMM_GEGLU) ? d->N / 2 : d->N;
  const bool staged = !(flags & (LDM_GEMM_OUT_F32 | LDM_GEMM_OUT_NCHW_F32 | LDM_GEMM_QKV_SPLIT | LDM_GEMM_CONVT_LN_SILU)) &&
                      n_out >= 64 && n_out % 8 == 0 && block_n >= ((flags & LDM_GEMM_GEGLU) ? 128 : 64) && staged_enabled();
  p.flags = flags | debug_flags() | (staged ? kStagedStore : 0) | (qkv_tma ? kQkvStaged : 0);
  if (!staged || block_n < 64) split_k = 1;
  p.split_k = split_k;
  if (split_k > 1) {
    const long m_alloc = pair ? 2 * ((p.m_tiles + 1) / 2) : p.m_tiles;
    p.ws = reinterpret_cast<float*>(d->splitk_ws);
    p.ws_slice_stride = (long long)m_alloc * p.n_tiles * block_n * kBlockM;
  }
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
  p.qkv_part0 = d->qkv_part0;
  if (flags & LDM_GEMM_QKV_SPLIT) {
    auto magic = [](long long dv) { return (unsigned long long)(((1ULL << 40) + (unsigned long long)dv - 1) / (unsigned long long)dv); };
    p.magic_seq = magic(d->seq);
    p.magic_c = magic((long long)d->heads * d->head_dim);
    p.magic_d = magic(d->head_dim);
    LDM_REQUIRE((long long)d->B * d->H * d->W * d->seq < (1LL << 40) && (long long)d->N * d->heads * d->head_dim < (1LL << 40),
                LDM_ERR_BAD_SHAPE, "ldm_gemm_bf16: QKV split extents too large for the multiply-shift division");
  }
  // LayerNorm fold (see ldm_gemm_desc.ln_stats)
  if (d->row_stats_out)
    LDM_REQUIRE(staged && !(flags & LDM_GEMM_GEGLU), LDM_ERR_BAD_ARG,
                "ldm_gemm_bf16: row_stats_out needs the plain bf16 [rows, N] epilogue (N >= 64, N %% 8 == 0)");
  if (d->ln_stats) {
    LDM_REQUIRE((flags & (LDM_GEMM_GEGLU | LDM_GEMM_QKV_SPLIT)) && d->ln_colsum && d->taps == 1 && !d->a2 && !d->rowbias &&
                    d->ln_parts == (d->c1 + 31) / 32 && d->N % 8 == 0,
                LDM_ERR_BAD_ARG, "ldm_gemm_bf16: ln_stats needs a GEGLU / QKV_SPLIT pointwise GEMM, ln_colsum and "
                "ln_parts == ceil(c1 / 32) (got %d for c1 = %d)", d->ln_parts, d->c1);
    LDM_REQUIRE((flags & LDM_GEMM_QKV_SPLIT) || staged, LDM_ERR_BAD_SHAPE,
                "ldm_gemm_bf16: the folded GEGLU needs the staged epilogue (N >= 128, N %% 32 == 0)");
  }
  p.row_stats = reinterpret_cast<float*>(d->row_stats_out);
  p.ln_stats = reinterpret_cast<const float*>(d->ln_stats);
  p.ln_colsum = d->ln_colsum;
  p.ln_parts = d->ln_parts;
  p.rows_total = (long long)d->B * d->H * d->W;
  p.ln_invc = 1.0f / (float)d->c1;
  p.ln_fold_eps = d->ln_fold_eps;
  p.a_stride = a_stride;
  p.a_pad = a_stride == 2 ? d->a_pad : 1;
  p.up2_cout = up2_cout;
  p.up2_ntiles = up2 ? up2_cout / block_n : 0;
  if (up2) LDM_REQUIRE(staged && split_k == 1, LDM_ERR_BAD_SHAPE, "ldm_gemm_bf16: up2 needs the staged epilogue");
  p.n_store = d->n_store > 0 ? d->n_store : d->N;
  p.img_px = (long long)d->H * d->W;

  CUtensorMap tmA1, tmA2, tmB;
  // A: [B, aH, aW, c] with boxes of bw x bh pixels; stride 2: every other pixel of a (2 bw) x (2 bh) window
  const uint64_t aW = a_stride == 2 ? (uint64_t)d->a_W : (uint64_t)p.W, aH = a_stride == 2 ? (uint64_t)d->a_H : (uint64_t)p.H;
  const uint32_t a_es[4] = {1, (uint32_t)a_stride, (uint32_t)a_stride, 1};
  if (a_stride == 2) {
    // padding 1 on both sides, or (a_pad = 0) one zero row / column on the high side only
    const long long pad_total = d->a_pad == 1 ? 2 : 1;
    LDM_REQUIRE(((long long)aH + pad_total - 3) / 2 + 1 == d->H && ((long long)aW + pad_total - 3) / 2 + 1 == d->W,
                LDM_ERR_BAD_SHAPE, "ldm_gemm_bf16: stride-2 extents: input %llux%llu with a_pad=%d does not give %dx%d",
                (unsigned long long)aH, (unsigned long long)aW, d->a_pad, d->H, d->W);
  }
  {
    const uint64_t dims[4] = {(uint64_t)d->c1, aW, aH, (uint64_t)p.B};
    const uint64_t str[3] = {(uint64_t)d->c1 * 2, (uint64_t)d->c1 * 2 * aW, (uint64_t)d->c1 * 2 * aW * aH};
    const uint32_t box[4] = {kBlockK, (uint32_t)(p.bw * a_stride), (uint32_t)(p.bh * a_stride), 1};
    int rc = make_tmap(&tmA1, d->a1, 4, dims, str, box, 2, true, a_es);
    if (rc) return rc;
  }
  if (c2 > 0) {
    const uint64_t dims[4] = {(uint64_t)c2, aW, aH, (uint64_t)p.B};
    const uint64_t str[3] = {(uint64_t)c2 * 2, (uint64_t)c2 * 2 * aW, (uint64_t)c2 * 2 * aW * aH};
    const uint32_t box[4] = {kBlockK, (uint32_t)(p.bw * a_stride), (uint32_t)(p.bh * a_stride), 1};
    int rc = make_tmap(&tmA2, d->a2, 4, dims, str, box, 2, true, a_es);
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
  if (staged && up2) {
    // the dense [B, out_H, out_W, cout] output, written through a map that steps by two pixels: a tile's (y, x) lands on
    // (2y + a, 2x + b)
    const uint64_t dims[4] = {(uint64_t)up2_cout, (uint64_t)d->out_W, (uint64_t)d->out_H, (uint64_t)p.B};
    const uint64_t str[3] = {(uint64_t)up2_cout * 2, (uint64_t)up2_cout * 2 * d->out_W,
                             (uint64_t)up2_cout * 2 * d->out_W * d->out_H};
    const uint32_t box[4] = {64, (uint32_t)(2 * p.bw), (uint32_t)(2 * p.bh), 1};
    const uint32_t es[4] = {1, 2, 2, 1};
    int rc = make_tmap(&tmO, d->out, 4, dims, str, box, 2, true, es);
    if (rc) return rc;
  } else if (staged) {
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
  if (res_mma) {
    const uint64_t dims[4] = {(uint64_t)d->N, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.B};
    const uint64_t str[3] = {(uint64_t)d->N * 2, (uint64_t)d->N * 2 * p.W, (uint64_t)d->N * 2 * p.W * p.H};
    const uint32_t box[4] = {64, (uint32_t)p.bw, (uint32_t)p.bh, 1};
    int rc = make_tmap(&tmR, d->residual, 4, dims, str, box, 2, true);
    if (rc) return rc;
  }

  if (qkv_tma) {
    // tmO / tmR (unused by this epilogue otherwise) become the q / k maps: {dpad, seq, heads, images}, box 64 x 128 tokens
    const uint64_t nimg = (uint64_t)((long long)d->B * d->H * d->W / d->seq);
    const uint64_t dims[4] = {(uint64_t)d->dpad, (uint64_t)d->seq, (uint64_t)d->heads, nimg};
    const uint64_t str[3] = {(uint64_t)d->dpad * 2, (uint64_t)d->dpad * 2 * d->seq, (uint64_t)d->dpad * 2 * d->seq * d->heads};
    const uint32_t box[4] = {64, (uint32_t)kBlockM, 1, 1};
    if (d->q) {
      int rc = make_tmap(&tmO, d->q, 4, dims, str, box, 2, true);
      if (rc) return rc;
    }
    if (d->k) {
      int rc = make_tmap(&tmR, d->k, 4, dims, str, box, 2, true);
      if (rc) return rc;
    }
  }

  const int smem_bytes = p.stages * p.stage_bytes + kEpiStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  int grid;
  if (pair) {
    const int units = ((p.m_tiles + 1) / 2) * p.n_tiles * split_k;
    int clusters = num_sms() / 2;
    if (clusters > units) clusters = units;
    grid = 2 * clusters;
  } else {
    const int num_tiles = p.m_tiles * p.n_tiles * split_k;
    grid = num_sms();
    if (grid > num_tiles) grid = num_tiles;
  }
  cudaError_t e;
  const cudaStream_t st = as_stream(stream);
  if (flags & LDM_GEMM_CONVT_LN_SILU)
    e = launch_gemm_convt(pair, grid, smem_bytes, st, tmA1, tmA2, tmB, tmO, tmR, tmE, p);
  else if ((flags & LDM_GEMM_QKV_SPLIT) && d->ln_stats)
    e = launch_gemm_qkv_ln(pair, grid, smem_bytes, st, tmA1, tmA2, tmB, tmO, tmR, tmE, p);
  else if (flags & LDM_GEMM_QKV_SPLIT)
    e = launch_gemm_qkv(pair, grid, smem_bytes, st, tmA1, tmA2, tmB, tmO, tmR, tmE, p);
  else if (staged && (flags & LDM_GEMM_GEGLU) && d->ln_stats)
    e = launch_gemm_geglu_ln(pair, grid, smem_bytes, st, tmA1, tmA2, tmB, tmO, tmR, tmE, p);
  else if (staged && (flags & LDM_GEMM_GEGLU))
    e = launch_gemm_geglu(pair, grid, smem_bytes, st, tmA1, tmA2, tmB, tmO, tmR, tmE, p);
  else if (staged && split_k > 1)
    e = launch_gemm_split(pair, grid, smem_bytes, st, tmA1, tmA2, tmB, tmO, tmR, tmE, p);
  else if (staged)
    e = launch_gemm_staged(pair, grid, smem_bytes, st, tmA1, tmA2, tmB, tmO, tmR, tmE, p);
  else
    e = launch_gemm_direct(pair, grid, smem_bytes, st, tmA1, tmA2, tmB, tmO, tmR, tmE, p);
  if (e != cudaSuccess) return set_error(LDM_ERR_CUDA, "gemm_tc_kernel launch: %s", cudaGetErrorString(e));
  if (split_k > 1) {
    FixupParams f{};
    f.ws = p.ws;
    f.slice_stride = p.ws_slice_stride;
    f.split = split_k;
    f.tiles_x = p.tiles_x; f.tiles_y = p.tiles_y; f.bw = p.bw; f.bh = p.bh; f.W = p.W; f.H = p.H;
    f.m_tiles = p.m_tiles; f.n_tiles = p.n_tiles; f.block_n = block_n; f.N = d->N;
    f.silu = (flags & LDM_GEMM_SILU) ? 1 : 0;
    f.bias = d->bias;
    f.rowbias = d->rowbias;
    f.residual = res_mma ? nullptr : reinterpret_cast<const __nv_bfloat16*>(d->residual);  // (else: on the tensor core)
    f.out = reinterpret_cast<__nv_bfloat16*>(d->out);
    cudaError_t fe = launch_pdl(splitk_fixup_kernel, dim3(block_n / 16, p.n_tiles, p.m_tiles), dim3(256), (size_t)0, st, 1, f);
    if (fe != cudaSuccess) return set_error(LDM_ERR_CUDA, "splitk_fixup_kernel launch: %s", cudaGetErrorString(fe));
    count_launch();
  }
  g_last_cfg[0] = block_n;
  g_last_cfg[1] = pair ? 1 : 0;
  g_last_cfg[2] = split_k;
  count_launch();
  return check_launch("gemm_tc_kernel");
}

extern "C" int ldm_gemm_last_config(int32_t* block_n, int32_t* pair, int32_t* split_k) {
  if (block_n) *block_n = g_last_cfg[0];
  if (pair) *pair = g_last_cfg[1];
  if (split_k) *split_k = g_last_cfg[2];
  return LDM_OK;
}
