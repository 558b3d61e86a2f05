// Fused flash-style self-attention for the UNet's BasicTransformerBlock.attn1 (softmax(QK^T * d^-0.5) V).
//
//   grid = (ceil(seq / (128*NQ)), B*heads); one CTA owns NQ tiles of 128 queries of one (image, head).
//     warp 0              TMA producer : Q tiles once, then a ring of (K block | V^T block) stages
//     warp 1              MMA issuer   : S = Q K^T and O_part = P V with tcgen05.mma (fp32 in TMEM)
//     warps 4..4+4*NQ-1   softmax      : one query row per thread. The whole score row of a KV block is read from TMEM
//                                        into registers in one pass and the S columns are handed back at once
//                                        (s_free), so the tensor core computes S of the NEXT block while this block's
//                                        exp2 / bf16 P are produced. P goes back into TMEM (tcgen05.st) and is the
//                                        A operand of the PV MMA straight from there: an MMA whose A tile comes from
//                                        shared memory costs 43 + N/2 cycles, from TMEM 10 + N/2
//                                        (tools/microbench/tmem_bench.cu), and PV has N = 48.
//                                        O accumulates in TMEM across blocks; it is rescaled (TMEM ld/st) only when a
//                                        row maximum grew by more than 2^8 since the last rescale, which is rare after
//                                        the first blocks, so the steady-state loop is max -> exp2 -> pack -> store.
//   TMEM columns of one query group: a ring of 1.5*BKV columns shared by S and P, then O. S of an even block sits at
//   [0, BKV), its P (bf16 pairs) at [0, BKV/2); S of an odd block at [BKV/2, 3BKV/2), its P at [BKV, 3BKV/2). A thread
//   has its whole score row in registers before it writes P over it, S of block j+1 never overlaps P of block j, and
//   S of block j+2 is issued after PV of block j on the in-order tensor pipe -- so the softmax threads never wait for
//   a PV MMA (only the rare rescale of O does).
//   Layouts: q,k [B*heads, seq, dpad] (dpad = 64*ceil(d/64), zero padded), vt [B*heads, vt_rows, seq_pad] (V transposed
//   so that both MMAs see K-major operands; vt_rows = 16*ceil(d/16); when d % 16 != 0 the caller keeps row d of every
//   head at 1.0 so the PV MMA also produces the softmax row sums), out [B*seq, heads*d]; all bf16.
#include <stdlib.h>

#include "attn_common.cuh"
#include "host_util.h"

namespace {
using namespace ldm;
using namespace ldm_attn;

template <int D, int NQ, int BKV, int STAGES>
struct AttnCfg {
  static constexpr int kAtoms = (D + 63) / 64;         // 64-wide K atoms of Q / K tiles
  static constexpr int kSteps = (D + 15) / 16;         // UMMA k-steps for S = Q K^T
  static constexpr int kDN = ((D + 15) / 16) * 16;     // UMMA N for O = P V
  static constexpr bool kOnesRow = (D % 16) != 0;      // spare V^T row D holds ones -> O_part[:, D] = row sum of P
  static constexpr int kQBytes = kAtoms * 128 * 128;   // per group
  static constexpr int kKBytes = kAtoms * BKV * 128;
  static constexpr int kVAtoms = BKV / 64;
  static constexpr int kVBytes = kVAtoms * kDN * 128;
  static constexpr int kVBytesPad = ((kVBytes + 1023) / 1024) * 1024;
  static constexpr int kStageBytes = kKBytes + kVBytesPad;
  static constexpr int kSmem = NQ * kQBytes + STAGES * kStageBytes + 1024 + 512;
  static constexpr int kThreads = 128 + 128 * NQ;  // warpgroup 0: producer, MMA issuer (+2 idle warps); then NQ softmax warpgroups
  static constexpr int kTmemGroupStride = 256;
  static constexpr int kRing = BKV + BKV / 2;          // S / P ring of one group (columns)
  static constexpr int kTmemCols = (NQ == 2) ? 512 : ((kRing + kDN <= 256) ? 256 : 512);
  static_assert(NQ == 1 || kRing + kDN <= 256, "TMEM budget");
  static_assert(kRing + kDN <= 512, "TMEM budget");
  static_assert(kSmem <= 227 * 1024, "smem budget");
};

template <int D, int NQ, int BKV, int STAGES, int kPoly>
__global__ void __launch_bounds__(AttnCfg<D, NQ, BKV, STAGES>::kThreads, 1)
flash_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using Cfg = AttnCfg<D, NQ, BKV, STAGES>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // an offset, not an integer round trip: the compiler keeps the shared address space (LDS / STS instead of generic LD / ST)
  uint8_t* sQ = smem;                                  // NQ * kQBytes
  uint8_t* sKV = sQ + NQ * Cfg::kQBytes;               // STAGES * kStageBytes
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + STAGES * Cfg::kStageBytes);
  uint64_t* q_full = bars;               // [NQ]
  uint64_t* kv_full = q_full + 2;        // [STAGES]
  uint64_t* kv_empty = kv_full + 8;      // [STAGES]
  uint64_t* s_full = kv_empty + 8;       // [NQ]
  uint64_t* p_full = s_full + 2;         // [NQ][2]: one barrier per (group, block parity), see the MMA warp's wait
  uint64_t* o_full = p_full + 4;         // [NQ]
  uint64_t* s_free = o_full + 2;         // [NQ]
  uint64_t* o_done = s_free + 2;         // [NQ] completes ONCE, when the last block's PV MMAs have landed (see the final wait)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  // warp-uniform by construction, so that the control warps' loops run with uniform control flow (see gemm_tc.cu)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int q_base = blockIdx.x * 128 * NQ;
  const int nblk = (p.kv_seq + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int g = 0; g < NQ; ++g) {
      mbar_init(&q_full[g], 1);
      mbar_init(&s_full[g], 1);
      mbar_init(&p_full[2 * g], 128);
      mbar_init(&p_full[2 * g + 1], 128);
      mbar_init(&o_full[g], 1);
      mbar_init(&o_done[g], 1);
      mbar_init(&s_free[g], 128);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();  // set-up done: the next kernel may start its own; q / k / v are needed from here on
  pdl_wait();

  // Register rebalancing (NQ = 2 launches 384 threads at 168 registers): the control warpgroup gives registers back,
  // the softmax warpgroups (a 128-wide score row per thread) take them.
  if (warp < 4) {
    if (NQ == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    {
      if (elect_one()) {
        for (int g = 0; g < NQ; ++g) {
          mbar_arrive_expect_tx(&q_full[g], Cfg::kQBytes);
          for (int a = 0; a < Cfg::kAtoms; ++a)
            tma_load_3d(sQ + g * Cfg::kQBytes + a * 128 * 128, &tmQ, &q_full[g], a * 64, q_base + g * 128, bh);
        }
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nblk; ++j) {
        mbar_wait(&kv_empty[stage], phase ^ 1);
        uint8_t* sk = sKV + stage * Cfg::kStageBytes;
        uint8_t* sv = sk + Cfg::kKBytes;
        if (elect_one()) {
          mbar_arrive_expect_tx(&kv_full[stage], Cfg::kKBytes + Cfg::kVBytes);
          for (int a = 0; a < Cfg::kAtoms; ++a)
            tma_load_3d(sk + a * BKV * 128, &tmK, &kv_full[stage], a * 64, j * BKV, bh);
          for (int a = 0; a < Cfg::kVAtoms; ++a)
            tma_load_3d(sv + a * Cfg::kDN * 128, &tmV, &kv_full[stage], j * BKV + a * 64, 0, bh);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (all lanes wait, one lane issues)
    {
      const uint32_t idesc_s = umma_idesc_bf16(128, BKV);
      const uint32_t idesc_o = umma_idesc_bf16(128, Cfg::kDN);
      auto issue_s = [&](int g, int stage, int blk) {
        const uint32_t qa = smem_u32(sQ + g * Cfg::kQBytes);
        const uint32_t ka = smem_u32(sKV + stage * Cfg::kStageBytes);
        const uint32_t d_tmem = tmem_base + g * Cfg::kTmemGroupStride + (blk & 1) * (BKV / 2);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < Cfg::kSteps; ++kk) {
            const uint64_t da = umma_desc_k_sw128(qa + (kk >> 2) * (128 * 128) + (kk & 3) * 32);
            const uint64_t db = umma_desc_k_sw128(ka + (kk >> 2) * (BKV * 128) + (kk & 3) * 32);
            umma_bf16(d_tmem, da, db, idesc_s, kk != 0);
          }
          umma_commit(&s_full[g]);
        }
        __syncwarp();
      };
      auto issue_o = [&](int g, int stage, int blk) {
        const uint32_t pa = tmem_base + g * Cfg::kTmemGroupStride + (blk & 1) * BKV;  // 8 columns per k-step
        const uint32_t va = smem_u32(sKV + stage * Cfg::kStageBytes + Cfg::kKBytes);
        const uint32_t d_tmem = tmem_base + g * Cfg::kTmemGroupStride + Cfg::kRing;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < BKV / 16; ++kk) {
            const uint64_t db = umma_desc_k_sw128(va + (kk >> 2) * (Cfg::kDN * 128) + (kk & 3) * 32);
            umma_bf16_ts(d_tmem, pa + kk * 8, db, idesc_o, blk > 0 || kk != 0);
          }
          umma_commit(&o_full[g]);
          if (blk == nblk - 1) umma_commit(&o_done[g]);
        }
        __syncwarp();
      };
      for (int g = 0; g < NQ; ++g) mbar_wait(&q_full[g], 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      for (int g = 0; g < NQ; ++g) issue_s(g, 0, 0);
      // Fixed order S(g0, j+1) S(g1, j+1) PV(g0, j) PV(g1, j). (An event loop that serves the groups in the order their
      // barriers complete was measured slower: 1.17 ms against 1.08 ms at d = 40, 7 488 tokens.)
      for (int j = 0; j < nblk; ++j) {
        const int stage = j % STAGES, nstage = (j + 1) % STAGES;
        if (j + 1 < nblk) {
          // S of the next block as soon as the softmax threads have pulled this block's scores out of TMEM
          mbar_wait(&kv_full[nstage], ((j + 1) / STAGES) & 1);
          for (int g = 0; g < NQ; ++g) {
            mbar_wait(&s_free[g], j & 1);
            tc_fence_after();
            TRACE(2 + g, j + 1);
            issue_s(g, nstage, j + 1);
          }
        }
        for (int g = 0; g < NQ; ++g) {
          // Two barriers per group, alternating with the block parity. With one, a group whose softmax threads finish
          // blocks j AND j + 1 (S of j + 1 is issued above, before this wait) while this warp is still held up by the
          // other group's s_free would complete two phases before the wait looks: a parity wait then never returns.
          // Arrivals for block j + 2 need S of j + 2, which is only issued after this wait has returned.
          mbar_wait(&p_full[2 * g + (j & 1)], (j >> 1) & 1);
          tc_fence_after();
          TRACE(4 + g, j);
          issue_o(g, stage, j);
        }
        if (elect_one()) umma_commit(&kv_empty[stage]);
        __syncwarp();
      }
    }
    __syncwarp();
  }
  } else {
    // ------------------------------------------------------------------ softmax / output (one row per thread)
    if (NQ == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int g = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const int qrow = q_base + g * 128 + r;
    const uint32_t t_g = tmem_base + ((uint32_t)(quad * 32) << 16) + g * Cfg::kTmemGroupStride;
    const uint32_t t_o = t_g + Cfg::kRing;
    float m = -INFINITY, l = 0.f;      // m: reference maximum the stored P / O are relative to (log2 domain)
    float m_seen = -INFINITY;          // running row maximum including the previous block (m lags behind it)
    constexpr bool kOnes = Cfg::kOnesRow;  // row sums come out of the PV MMA (ones row of V^T at index D)
    constexpr float kLazy = 8.0f;       // rescale O only when the row maximum grew by more than 2^8 ...
    constexpr float kRedo = 64.0f;      // ... and redo a block's exponentials if it alone jumps by more than 2^64

    for (int j = 0; j < nblk; ++j) {
      const int nvalid = p.kv_seq - j * BKV;  // keys of this block inside the sequence (>= 1)
      const uint32_t t_s = t_g + (j & 1) * (BKV / 2);  // this block's scores ...
      const uint32_t t_p = t_g + (j & 1) * BKV;        // ... and where its P goes (see the column ring above)
      TRACE(0, j);
      mbar_wait(&s_full[g], j & 1);
      tc_fence_after();
      TRACE(1, j);
      uint32_t sv[BKV];  // the whole score row of this block stays in registers: one TMEM pass
#pragma unroll
      for (int c = 0; c < BKV; c += 32) tmem_ld32(t_s + c, sv + c);
      tmem_ld_wait();
      TRACE(2, j);
      tc_fence_before();
      mbar_arrive(&s_free[g]);  // the S columns may be overwritten by the next block's QK^T now
      if (nvalid < BKV) {  // only the last block of a ragged sequence
#pragma unroll
        for (int i = 0; i < BKV; ++i)
          if (i >= nvalid) sv[i] = 0xff800000u;  // -inf
      }

      // O (TMEM) and l move to the new reference max(m, target) for the rows that need it
      auto rescale = [&](bool need, float target) {
        if (j > 0) mbar_wait(&o_full[g], (j - 1) & 1);  // PV of the previous block has landed in O
        tc_fence_after();
        const float m_new = need ? target : m;
        const float alpha = ex2(m - m_new);  // 1 for the rows that keep their reference
        m = m_new;
        if (!kOnes) l *= alpha;
#pragma unroll
        for (int c = 0; c < Cfg::kDN; c += 16) {
          uint32_t v[16];
          tmem_ld16(t_o + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st16(t_o + c, v);
        }
        tmem_st_wait();
      };
      // P = exp2(s*scale - m) as bf16 pairs into TMEM (the A operand of the PV MMA); the block maximum is tracked in
      // the same pass (the FMNMX issue between the MUFUs instead of in a MUFU-idle phase of their own)
      auto exp_pass = [&](float& bmax, float& bsum) {
        float mx0 = -INFINITY, mx1 = -INFINITY, l0 = 0.f, l1 = 0.f;
        const uint64_t scale2 = pack2(p.scale_log2, p.scale_log2), negm2 = pack2(-m, -m);
#pragma unroll
        for (int c = 0; c < BKV; c += 16) {
          uint32_t u[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {  // pairs of keys; kPoly of every 8 pairs take the FMA-pipe exp2
            const float s0 = __uint_as_float(sv[c + 2 * q]), s1 = __uint_as_float(sv[c + 2 * q + 1]);
            if (q & 1) mx1 = fmax3(mx1, s0, s1); else mx0 = fmax3(mx0, s0, s1);
            const uint64_t x = ffma2(pack2(s0, s1), scale2, negm2);
            float e0, e1;
            if (q % 8 < 8 - kPoly) {
              float x0, x1;
              unpack2(x, x0, x1);
              e0 = ex2(x0);
              e1 = ex2(x1);
            } else {
              ex2_poly2(x, e0, e1);
            }
            if (!kOnes) {
              l0 += e0;
              l1 += e1;
            }
            u[q] = pack_bf16(e0, e1);
          }
          tmem_st8(t_p + (c >> 1), u);  // keys c .. c+15 of this thread's row = 8 columns of bf16 pairs
        }
        bmax = fmaxf(mx0, mx1) * p.scale_log2;  // scale > 0
        bsum = l0 + l1;
      };

      if (j == 0) {
        // the first block needs its true maximum up front (nothing to be relative to yet)
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int i = 0; i < BKV; i += 4) {
          mx0 = fmax3(mx0, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
          mx1 = fmax3(mx1, __uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
        }
        m = fmaxf(mx0, mx1) * p.scale_log2;
        m_seen = m;
      } else {
        const bool need = m_seen - m > kLazy;
        if (__any_sync(0xffffffffu, need)) rescale(need, m_seen);
      }
      float bmax, bsum;
      exp_pass(bmax, bsum);
      if (j > 0) {
        const bool redo = bmax - m > kRedo;  // would leave the comfortable fp32 / bf16 range: never on sane inputs
        if (__any_sync(0xffffffffu, redo)) {
          rescale(redo, bmax);
          exp_pass(bmax, bsum);
        }
      }
      m_seen = fmaxf(m_seen, bmax);
      if (!kOnes) l += bsum;
      TRACE(3, j);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[2 * g + (j & 1)]);
      TRACE(4, j);
    }
    // NOT o_full with the parity of the last block: a softmax thread never waits for a PV MMA inside the loop, so here
    // it can be TWO commits ahead of the tensor pipe (s_full of block j only implies PV of block j-2), and a parity wait
    // cannot tell "phase nblk-3 complete" from "phase nblk-1 complete" -- a fast warp then read O without the last two
    // blocks (seen as a run-to-run difference of one warp's rows). o_done has a single phase.
    mbar_wait(&o_done[g], 0);
    tc_fence_after();
    float o_acc[Cfg::kDN];
    {
      uint32_t v[Cfg::kDN];
#pragma unroll
      for (int c = 0; c < Cfg::kDN; c += 16) tmem_ld16(t_o + c, v + c);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < Cfg::kDN; ++i) o_acc[i] = __uint_as_float(v[i]);
    }
    if (kOnes) l = o_acc[D];
    if (qrow < p.seq) {
      const float inv = 1.0f / l;
      const int b = bh / p.heads, head = bh - b * p.heads;
      __nv_bfloat16* dst = p.out + ((long long)b * p.seq + qrow) * (p.heads * p.head_dim) + head * p.head_dim;
#pragma unroll
      for (int c = 0; c < D; c += 8) {
        uint4 u;
        u.x = pack_bf16(o_acc[c + 0] * inv, o_acc[c + 1] * inv);
        u.y = pack_bf16(o_acc[c + 2] * inv, o_acc[c + 3] * inv);
        u.z = pack_bf16(o_acc[c + 4] * inv, o_acc[c + 5] * inv);
        u.w = pack_bf16(o_acc[c + 6] * inv, o_acc[c + 7] * inv);
        *reinterpret_cast<uint4*>(dst + c) = u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int D, int NQ, int BKV, int STAGES, int kPoly>
int launch_attn(const ldm_attn_desc* d, cudaStream_t s) {
  using namespace ldm_host;
  using Cfg = AttnCfg<D, NQ, BKV, STAGES>;
  const int BH = d->B * d->heads;
  const int kv_seq = d->kv_seq > 0 ? d->kv_seq : d->seq;
  CUtensorMap tmQ, tmK, tmV;
  {
    const uint64_t dims[3] = {(uint64_t)d->dpad, (uint64_t)d->seq, (uint64_t)BH};
    const uint64_t str[2] = {(uint64_t)d->dpad * 2, (uint64_t)d->dpad * 2 * d->seq};
    const uint64_t dimsk[3] = {(uint64_t)d->dpad, (uint64_t)kv_seq, (uint64_t)BH};
    const uint64_t strk[2] = {(uint64_t)d->dpad * 2, (uint64_t)d->dpad * 2 * kv_seq};
    const uint32_t boxq[3] = {64, 128, 1};
    const uint32_t boxk[3] = {64, (uint32_t)BKV, 1};
    int rc = make_tmap(&tmQ, d->q, 3, dims, str, boxq, 2, true);
    if (rc) return rc;
    rc = make_tmap(&tmK, d->k, 3, dimsk, strk, boxk, 2, true);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)kv_seq, (uint64_t)d->vt_rows, (uint64_t)BH};
    const uint64_t str[2] = {(uint64_t)d->seq_pad * 2, (uint64_t)d->seq_pad * 2 * d->vt_rows};
    const uint32_t box[3] = {64, (uint32_t)Cfg::kDN, 1};
    int rc = make_tmap(&tmV, d->vt, 3, dims, str, box, 2, true);
    if (rc) return rc;
  }
  AttnParams p;
  p.seq = d->seq;
  p.kv_seq = kv_seq;
  p.heads = d->heads;
  p.head_dim = d->head_dim;
  p.scale_log2 = d->scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__nv_bfloat16*>(d->out);
  auto kern = flash_attn_kernel<D, NQ, BKV, STAGES, kPoly>;
  {  // per launch (the attribute is per device; see gemm_tc_inst.cuh)
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
    if (e != cudaSuccess) return set_error(LDM_ERR_CUDA, "cudaFuncSetAttribute(attn): %s", cudaGetErrorString(e));
  }
  dim3 grid((d->seq + 128 * NQ - 1) / (128 * NQ), BH);
  cudaError_t le = launch_pdl(kern, grid, dim3(Cfg::kThreads), (size_t)Cfg::kSmem, s, 1, tmQ, tmK, tmV, p);
  if (le != cudaSuccess) return set_error(LDM_ERR_CUDA, "flash_attn_kernel launch: %s", cudaGetErrorString(le));
  count_launch();
  return check_launch("flash_attn_kernel");
}

}  // namespace

extern "C" int ldm_attn_vt_rows(int head_dim) { return ((head_dim + 15) / 16) * 16; }

extern "C" int ldm_flash_attn_fwd(const ldm_attn_desc* d, ldm_stream_t stream) {
  using namespace ldm_host;
  LDM_REQUIRE(d && d->q && d->k && d->vt && d->out, LDM_ERR_BAD_ARG, "ldm_flash_attn_fwd: null arg");
  LDM_REQUIRE(d->B > 0 && d->heads > 0 && d->seq > 0, LDM_ERR_BAD_SHAPE, "ldm_flash_attn_fwd: bad B/heads/seq");
  LDM_REQUIRE(d->vt_rows == ldm_attn_vt_rows(d->head_dim), LDM_ERR_BAD_SHAPE,
              "ldm_flash_attn_fwd: vt_rows=%d, expected ldm_attn_vt_rows(%d)=%d", d->vt_rows, d->head_dim,
              ldm_attn_vt_rows(d->head_dim));
  const int kv_seq = d->kv_seq > 0 ? d->kv_seq : d->seq;
  LDM_REQUIRE(d->kv_seq >= 0 && d->dpad == ((d->head_dim + 63) / 64) * 64 && d->seq_pad % 8 == 0 && d->seq_pad >= kv_seq,
              LDM_ERR_BAD_SHAPE, "ldm_flash_attn_fwd: dpad=%d seq_pad=%d inconsistent with head_dim=%d kv_seq=%d", d->dpad,
              d->seq_pad, d->head_dim, kv_seq);
  cudaStream_t s = as_stream(stream);
  switch (d->head_dim) {
    case 40: {
      // diagnostic builds: LDM_ATTN40=0 runs the generic kernel below instead of attn40_tc.cu, LDM_ATTN_POLY=0..4 sets
      // the FMA-pipe share of the exponentials (n of 8) -- A/B timing
      if (diag_env("LDM_ATTN40", 1)) return ldm_launch_attn40(d, s);
      const int poly = diag_env("LDM_ATTN_POLY", 2);
      if (poly == 0) return launch_attn<40, 2, 128, 4, 0>(d, s);
      if (poly == 1) return launch_attn<40, 2, 128, 4, 1>(d, s);
      if (poly == 3) return launch_attn<40, 2, 128, 4, 3>(d, s);
      if (poly == 4) return launch_attn<40, 2, 128, 4, 4>(d, s);
      return launch_attn<40, 2, 128, 4, 2>(d, s);
    }
    case 80: {
      // LDM_ATTN_D80=0 (diagnostic builds): one query group, 128-key blocks; default: two groups, 64-key blocks
      if (diag_env("LDM_ATTN_D80", 1) == 0) return launch_attn<80, 1, 128, 3, 0>(d, s);
      return launch_attn<80, 2, 64, 4, 0>(d, s);
    }
    case 160:
      return launch_attn<160, 1, 64, 3, 0>(d, s);
    default:
      return set_error(LDM_ERR_BAD_SHAPE, "ldm_flash_attn_fwd: head_dim %d not built (40, 80, 160)", d->head_dim);
  }
}
