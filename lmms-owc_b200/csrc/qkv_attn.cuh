// qkv_attn.cuh - wide encoder path (H = 768, head_dim 64): the QKV projection with the self-attention in its epilogue.
//
// Why: as separate kernels (k_gemm_wide<FOLD> -> k_attn_grp) the bf16 q|k|v rows cross HBM twice - 9.2 KB of the 43 KB of
// DRAM traffic per token and layer - and the attention is a launch of its own worth 12 % of the c3 step. Sequences on this
// path are short (labels: 3-16 tokens, at most 64 here), so a 128-row block made of WHOLE sequences is a closed system for
// attention. This kernel is the persistent CTA-pair GEMM of gemm_wide.cuh with
//   * work items = pairs of such row blocks (host table, as for the block-resident encoder); the A tile of a CTA is the 128
//     packed rows starting at its block's first row - TMA takes any row coordinate, and rows past the block's last sequence
//     are simply the next block's rows, computed and ignored (nothing is ever stored for them);
//   * column tiles of 192 = one head's [q | k | v] (the gamma-folded weight rows are regrouped per head at load time), so an
//     accumulator holds everything one head needs for the block;
//   * an epilogue that applies the deferred LayerNorm (out = rstd_r (acc - mu_r c1[n]) + c2[n]), writes the head's bf16
//     q|k|v rows into shared memory (XOR-swizzled 16-byte chunks: every fragment access below is bank-conflict free), runs
//     the attention of the block's sequences on the warp-level tensor path (several short sequences per 16 x 16 tile, as in
//     k_attn_grp) and writes ONLY the attention output rows to global memory (the A operand of the O-projection).
// The MMAs of head h+1 run while the epilogue warps work on head h (two TMEM accumulators of 192 columns).
// Reference semantics: modeling_mpnet.py:116-186 / modeling_bert.py:114-207 (scaled dot-product attention with the
// additive relative-position bias of MPNet; packed sequences carry no padding, so the padding mask is a no-op).
#pragma once

#include "attn_mma.cuh"
#include "common.cuh"
#include "gemm_tc.cuh"

namespace owc {

struct QkvAttnArgs {
  int n_blk;                 // row blocks of whole sequences (<= 128 tokens each) in this chunk
  const int32_t* blk_seq;    // [n_blk + 1] absolute index of the first sequence of each block (device)
  const int32_t* cu;         // device cu_seqlens of the whole call
  int tok0;                  // cu[first sequence of the chunk]: rows are chunk-relative
  int M;                     // token rows of the chunk
  int K;                     // hidden size (K % 64 == 0)
  int heads;
  const float* c1;           // [heads * 192] row sums of the folded, per-head regrouped weights
  const float* c2;           // [heads * 192] W beta + bias, regrouped
  const float* stats_in;     // partial (sum, sumsq) slots of the A operand's pre-LayerNorm rows (gemm_wide.cuh)
  int slots_in;
  long long stat_stride;
  float inv_h, ln_eps;
  __nv_bfloat16* ao;         // [M, ldo] attention output
  int ldo;
  float scale_log2;          // head_dim^-0.5 * log2(e)
  const float* rel_lut;      // MPNet: [heads, 2 * max_pos - 1] bias by (k - q), or nullptr
  int max_pos;
};

namespace qa {
constexpr int D = 64;                       // head dim
constexpr int BN = 3 * D;                   // 192: one head's q | k | v
constexpr int BM = 128, BK = 64;
constexpr int A_BYTES = BM * BK * 2;        // 16 KB
constexpr int BH_BYTES = (BN / 2) * BK * 2; // 12 KB: this CTA's 96 rows of the head's weight tile
constexpr int STAGE_BYTES = A_BYTES + BH_BYTES;
constexpr int STAGES = 6;
constexpr int QROW = 3 * D * 2;             // 384 B per q|k|v row
constexpr int QKV_BYTES = BM * QROW;        // 48 KB
constexpr int OFF_QKV = STAGES * STAGE_BYTES;
constexpr int OFF_TAIL = OFF_QKV + QKV_BYTES;
constexpr int TAIL_BYTES = 512 /*barriers*/ + 544 /*seq_row*/ + 160 /*groups*/ + 2 * 2 * BN * 4 /*c1|c2, two parities*/;
constexpr int SMEM_BYTES = 1024 + OFF_TAIL + TAIL_BYTES;
static_assert(SMEM_BYTES <= kSmemLimit, "shared memory budget");
}  // namespace qa

// 16-byte chunk c (0-7 q, 8-15 k, 16-23 v) of q|k|v row `row`: the chunk index is XORed inside its 128-byte third with
// (row & 7), so 8 consecutive rows put the same logical chunk into 8 different bank groups.
__device__ __forceinline__ uint8_t* qa_chunk(uint8_t* qkv, int row, int c) {
  return qkv + row * qa::QROW + (((c & ~7) | ((c & 7) ^ (row & 7))) << 4);
}

// One attention unit out of the shared q|k|v rows: a GROUP of consecutive sequences (<= 16 rows together, block-diagonal
// mask; TP == 16) or one longer sequence (TP == 32 / 64). Rows [beg, beg + T) of the block. The output overwrites the unit's
// own (dead) q chunks and is then copied out with 16-byte row-contiguous stores.
template <int TP>
__device__ __forceinline__ void qa_attend(uint8_t* qkv, int beg, int T, const int* seq_row, int s_first, int s_last, int lane,
                                          float scale_log2, const float* lut, __nv_bfloat16* out, int ldo) {
  constexpr int NT = TP / 8, MT = TP / 16, DN = qa::D / 8;
  const int g = lane >> 2, t4 = lane & 3;
  auto row_of = [&](int i) { return min(beg + i, 127); };   // rows past the block are never used unmasked
  int lo = 0, hi = T;
  if (TP == 16) {   // key range of every tile row: lane r finds the sequence that owns row r
    lo = 0; hi = 0;
    for (int sq = s_first; sq < s_last; ++sq) {
      const int b = seq_row[sq] - beg, e = seq_row[sq + 1] - beg;
      if (lane >= b && lane < e) { lo = b; hi = e; }
    }
  }
#pragma unroll 1
  for (int mt = 0; mt < MT; ++mt) {
    if (mt * 16 >= T) break;
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    int lo0 = 0, hi0 = T, lo1 = 0, hi1 = T;
    if (TP == 16) {
      lo0 = __shfl_sync(0xffffffffu, lo, r0); hi0 = __shfl_sync(0xffffffffu, hi, r0);
      lo1 = __shfl_sync(0xffffffffu, lo, r1); hi1 = __shfl_sync(0xffffffffu, hi, r1);
    } else {
      hi0 = (r0 < T) ? T : 0;   // rows past the sequence see no key
      hi1 = (r1 < T) ? T : 0;
    }
    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) { s[nt][0] = 0.f; s[nt][1] = 0.f; s[nt][2] = 0.f; s[nt][3] = 0.f; }
    const int qr0 = row_of(r0), qr1 = row_of(r1);
#pragma unroll
    for (int ks = 0; ks < qa::D / 16; ++ks) {
      uint32_t af[4];
      af[0] = *reinterpret_cast<const uint32_t*>(qa_chunk(qkv, qr0, 2 * ks) + t4 * 4);
      af[1] = *reinterpret_cast<const uint32_t*>(qa_chunk(qkv, qr1, 2 * ks) + t4 * 4);
      af[2] = *reinterpret_cast<const uint32_t*>(qa_chunk(qkv, qr0, 2 * ks + 1) + t4 * 4);
      af[3] = *reinterpret_cast<const uint32_t*>(qa_chunk(qkv, qr1, 2 * ks + 1) + t4 * 4);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        if (nt * 8 < T) {   // warp-uniform
          const int kr = row_of(nt * 8 + g);
          uint32_t bf[2];
          bf[0] = *reinterpret_cast<const uint32_t*>(qa_chunk(qkv, kr, 8 + 2 * ks) + t4 * 4);
          bf[1] = *reinterpret_cast<const uint32_t*>(qa_chunk(qkv, kr, 9 + 2 * ks) + t4 * 4);
          mma_bf16_16816(s[nt], af, bf);
        }
      }
    }
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (nt * 8 < T) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = nt * 8 + t4 * 2 + (e & 1);
          const int i = (e < 2) ? r0 : r1;
          const bool ok = (e < 2) ? (j >= lo0 && j < hi0) : (j >= lo1 && j < hi1);
          float v = s[nt][e] * scale_log2;
          if (lut != nullptr && ok) v += __ldg(lut + (j - i)) * 1.4426950408889634f;
          v = ok ? v : -INFINITY;
          s[nt][e] = v;
          if (e < 2) m0 = fmaxf(m0, v); else m1 = fmaxf(m1, v);
        }
      }
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    m0 = (hi0 > lo0) ? m0 : 0.f;   // tile rows without a key: keep (-inf) - (-inf) out of the exponent
    m1 = (hi1 > lo1) ? m1 : 0.f;
    uint32_t pk[NT][2];
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (nt * 8 < T) {
        float e0, e1, e2, e3;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(s[nt][0] - m0));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(s[nt][1] - m0));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(s[nt][2] - m1));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e3) : "f"(s[nt][3] - m1));
        const __nv_bfloat162 p01 = __floats2bfloat162_rn(e0, e1);
        const __nv_bfloat162 p23 = __floats2bfloat162_rn(e2, e3);
        sum0 += __low2float(p01) + __high2float(p01);
        sum1 += __low2float(p23) + __high2float(p23);
        pk[nt][0] = *reinterpret_cast<const uint32_t*>(&p01);
        pk[nt][1] = *reinterpret_cast<const uint32_t*>(&p23);
      } else {
        pk[nt][0] = 0u;
        pk[nt][1] = 0u;
      }
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    float o[DN][4];
#pragma unroll
    for (int dn = 0; dn < DN; ++dn) { o[dn][0] = 0.f; o[dn][1] = 0.f; o[dn][2] = 0.f; o[dn][3] = 0.f; }
#pragma unroll
    for (int kt = 0; kt < MT; ++kt) {
      if (kt * 16 < T) {   // warp-uniform
        const uint32_t af[4] = {pk[2 * kt][0], pk[2 * kt][1], pk[2 * kt + 1][0], pk[2 * kt + 1][1]};
        const int vr = row_of(kt * 16 + (lane & 15));
#pragma unroll
        for (int dn = 0; dn < DN; ++dn) {
          uint32_t bf[2];
          ldmatrix_x2_trans(bf, qa_chunk(qkv, vr, 16 + dn));
          mma_bf16_16816(o[dn], af, bf);
        }
      }
    }
    float inv0, inv1;   // real rows sum to >= 1 (the row maximum contributes 2^0); rows without keys are not stored
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv0) : "f"(fmaxf(sum0, 1.0f)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv1) : "f"(fmaxf(sum1, 1.0f)));
    __syncwarp();   // every lane has read this m-tile's q rows
#pragma unroll
    for (int dn = 0; dn < DN; ++dn) {
      if (r0 < T) *reinterpret_cast<uint32_t*>(qa_chunk(qkv, beg + r0, dn) + t4 * 4) = pack_bf16x2(o[dn][0] * inv0, o[dn][1] * inv0);
      if (r1 < T) *reinterpret_cast<uint32_t*>(qa_chunk(qkv, beg + r1, dn) + t4 * 4) = pack_bf16x2(o[dn][2] * inv1, o[dn][3] * inv1);
    }
  }
  __syncwarp();
  // 16-byte, row-contiguous write-back: 8 lanes cover one 128-byte output row segment
  for (int idx = lane; idx < T * 8; idx += 32) {
    const int c = idx & 7, t = idx >> 3;
    *reinterpret_cast<uint4*>(out + static_cast<size_t>(beg + t) * ldo + c * 8) = *reinterpret_cast<const uint4*>(qa_chunk(qkv, beg + t, c));
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
k_qkv_attn(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const QkvAttnArgs g) {
  using namespace qa;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* qkv_s = tiles + OFF_QKV;
  uint8_t* tail = tiles + OFF_TAIL;
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  int* seq_row = reinterpret_cast<int*>(tail + 512);            // [<= 130] first row (block-relative) of each sequence
  uint8_t* grp_first = tail + 512 + 544;                        // [n_grp + 1] first sequence of each attention group
  int* n_grp_s = reinterpret_cast<int*>(tail + 512 + 544 + 144);
  float* param_s = reinterpret_cast<float*>(tail + 512 + 544 + 160);   // [2][c1 | c2][BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  const int n_items = (g.n_blk + 1) >> 1;
  const int num_kb = g.K / BK;

  if (warp == 1 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 2 * kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc_2sm(tmem_slot, 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  // first token row of this CTA's block of item `it`; a missing odd tail block re-reads the leader's rows (never stored)
  auto block_of = [&](int it) { return min(2 * it + static_cast<int>(rank), g.n_blk - 1); };

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      int stage = 0;
      uint32_t phase = 0;
      for (int it = cluster_id; it < n_items; it += n_clusters) {
        const int m0 = __ldg(g.cu + __ldg(g.blk_seq + block_of(it))) - g.tok0;
        for (int h = 0; h < g.heads; ++h) {
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = tiles + stage * STAGE_BYTES;
            const uint32_t bar = mapa_u32(smem_u32(&full[stage]), 0);
            if (leader) mbar_expect_tx(&full[stage], 2u * STAGE_BYTES);
            tma_load_2d_2sm(sa, &tmA, bar, kb * BK, m0);
            tma_load_2d_2sm(sa + A_BYTES, &tmB, bar, kb * BK, h * BN + static_cast<int>(rank) * (BN / 2));
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ---------------- MMA issuer: UMMA 256 x 192 x 16 across the pair ----------------
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN);
      const uint32_t tiles_lo = sw128_desc_lo(smem_u32(tiles));
      int stage = 0;
      uint32_t phase = 0;
      uint32_t t = 0;
      for (int it = cluster_id; it < n_items; it += n_clusters) {
        for (int h = 0; h < g.heads; ++h, ++t) {
          const uint32_t buf = t & 1u;
          mbar_wait(&tempty[buf], ((t >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * BN;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t a_lo = tiles_lo + static_cast<uint32_t>(stage * (STAGE_BYTES >> 4));
            if (elect_one_sync()) {
              umma_kblock_2sm(d_tmem, a_lo, a_lo + (A_BYTES >> 4), idesc, static_cast<uint32_t>(kb != 0));
              umma_commit_2sm(&empty[stage]);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (elect_one_sync()) umma_commit_2sm(&tfull[buf]);
          __syncwarp();
        }
      }
    }
  } else {
    // ---------------- epilogue warps: deferred LayerNorm, q|k|v rows to shared memory, attention, output rows ----------------
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int we = warp - 2;                    // 0..7
    const int et = threadIdx.x - 64;
    const int r = q * 32 + lane;                // token row of this thread inside the block
    uint32_t t = 0;
    for (int it = cluster_id; it < n_items; it += n_clusters) {
      const int blk = 2 * it + static_cast<int>(rank);
      const bool have = blk < g.n_blk;
      const int s0 = __ldg(g.blk_seq + block_of(it));
      const int nseq = have ? __ldg(g.blk_seq + blk + 1) - s0 : 0;
      const int row0 = __ldg(g.cu + s0) - g.tok0;
      named_bar_sync(1, kEpiThreads);           // everyone is done with the previous block's tables
      for (int i = et; i <= nseq; i += kEpiThreads) seq_row[i] = __ldg(g.cu + s0 + i) - g.tok0 - row0;
      named_bar_sync(1, kEpiThreads);
      if (et == 0) {
        int ng = 0, rows = 0;
        grp_first[0] = 0;
        for (int i = 0; i < nseq; ++i) {
          const int len = seq_row[i + 1] - seq_row[i];
          if (rows > 0 && rows + len > 16) {
            grp_first[++ng] = static_cast<uint8_t>(i);
            rows = 0;
          }
          rows += len;
        }
        if (nseq > 0) grp_first[++ng] = static_cast<uint8_t>(nseq);
        *n_grp_s = ng;
      }
      // per-row LayerNorm of the incoming stream: x = y * ra + rb
      float ra = 0.f, rb = 0.f;
      {
        const int row = row0 + r;
        if (row < g.M) {
          const float2* st = reinterpret_cast<const float2*>(g.stats_in);
          float2 s2 = __ldg(st + row);
          for (int sl = 1; sl < g.slots_in; ++sl) {
            const float2 p = __ldg(st + sl * g.stat_stride + row);
            s2.x += p.x;
            s2.y += p.y;
          }
          const float mu = s2.x * g.inv_h;
          ra = rsqrtf(fmaxf(s2.y * g.inv_h - mu * mu, 0.f) + g.ln_eps);
          rb = -mu * ra;
        }
      }
      for (int h = 0; h < g.heads; ++h, ++t) {
        const uint32_t buf = t & 1u;
        float* p0 = param_s + buf * 2 * BN;
        if (et < BN) {
          p0[et] = __ldg(g.c1 + h * BN + et);
          p0[BN + et] = __ldg(g.c2 + h * BN + et);
        }
        named_bar_sync(1, kEpiThreads);         // parameters staged; the attention of the previous head has left the q|k|v rows
        mbar_wait(&tfull[buf], (t >> 1) & 1u);
        tc_fence_after();
        const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN;
        // this warp's three 32-column chunks of the head tile: chunk c = 2 * cw + half (q: 0,1  k: 2,3  v: 4,5)
        uint32_t v[2][32];
        tmem_ld_32x32(tbase + static_cast<uint32_t>(half * 32), v[0]);
#pragma unroll
        for (int cw = 0; cw < 3; ++cw) {
          const int c = 2 * cw + half;
          tmem_ld_wait_dep(v[cw & 1]);
          if (cw + 1 < 3) tmem_ld_32x32(tbase + static_cast<uint32_t>((c + 2) * 32), v[(cw + 1) & 1]);
          const uint32_t (&vc)[32] = v[cw & 1];
#pragma unroll
          for (int j = 0; j < 4; ++j) {         // four 16-byte chunks of 8 columns
            const float4 k1a = *reinterpret_cast<const float4*>(p0 + c * 32 + 8 * j);
            const float4 k1b = *reinterpret_cast<const float4*>(p0 + c * 32 + 8 * j + 4);
            const float4 k2a = *reinterpret_cast<const float4*>(p0 + BN + c * 32 + 8 * j);
            const float4 k2b = *reinterpret_cast<const float4*>(p0 + BN + c * 32 + 8 * j + 4);
            uint4 w;
            w.x = pack_bf16x2(fmaf(__uint_as_float(vc[8 * j + 0]), ra, fmaf(rb, k1a.x, k2a.x)),
                              fmaf(__uint_as_float(vc[8 * j + 1]), ra, fmaf(rb, k1a.y, k2a.y)));
            w.y = pack_bf16x2(fmaf(__uint_as_float(vc[8 * j + 2]), ra, fmaf(rb, k1a.z, k2a.z)),
                              fmaf(__uint_as_float(vc[8 * j + 3]), ra, fmaf(rb, k1a.w, k2a.w)));
            w.z = pack_bf16x2(fmaf(__uint_as_float(vc[8 * j + 4]), ra, fmaf(rb, k1b.x, k2b.x)),
                              fmaf(__uint_as_float(vc[8 * j + 5]), ra, fmaf(rb, k1b.y, k2b.y)));
            w.w = pack_bf16x2(fmaf(__uint_as_float(vc[8 * j + 6]), ra, fmaf(rb, k1b.z, k2b.z)),
                              fmaf(__uint_as_float(vc[8 * j + 7]), ra, fmaf(rb, k1b.w, k2b.w)));
            *reinterpret_cast<uint4*>(qa_chunk(qkv_s, r, c * 4 + j)) = w;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[buf]), 0));   // accumulator drained: the MMAs of head h+2 may start
        named_bar_sync(2, kEpiThreads);         // every row of the head is in shared memory
        // ---- attention of the block's sequences for this head (unit u -> epilogue warp u % 8) ----
        const int n_grp = *n_grp_s;
        const float* lut = g.rel_lut ? g.rel_lut + static_cast<size_t>(h) * (2 * g.max_pos - 1) + g.max_pos - 1 : nullptr;
        __nv_bfloat16* out = g.ao + static_cast<size_t>(row0) * g.ldo + h * D;
        for (int u = we; u < n_grp; u += kEpiWarps) {
          const int s_first = grp_first[u], s_last = grp_first[u + 1];
          const int beg = seq_row[s_first];
          const int T = seq_row[s_last] - beg;
          if (T <= 0) continue;
          if (T <= 16) qa_attend<16>(qkv_s, beg, T, seq_row, s_first, s_last, lane, g.scale_log2, lut, out, g.ldo);
          else if (T <= 32) qa_attend<32>(qkv_s, beg, T, seq_row, s_first, s_last, lane, g.scale_log2, lut, out, g.ldo);
          else qa_attend<64>(qkv_s, beg, T, seq_row, s_first, s_last, lane, g.scale_log2, lut, out, g.ldo);
        }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

// Gamma-fold of Wqkv with the rows regrouped per head: destination row h * 192 + p * 64 + d  <-  source row p * H + h * 64 + d
// (p = 0 q, 1 k, 2 v), so a 192-row tile is one head's projection. Same arithmetic as k_fold_weight (gemm_wide.cuh).
__global__ void __launch_bounds__(256) k_fold_weight_heads(const float* __restrict__ W, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, const float* __restrict__ bias,
                                                           int H, int heads, __nv_bfloat16* __restrict__ Wf,
                                                           float* __restrict__ c1, float* __restrict__ c2) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // destination row
  if (n >= 3 * H) return;
  const int D = H / heads;
  const int h = n / (3 * D), p = (n / D) % 3, d = n % D;
  const int src = p * H + h * D + d;
  const float* w = W + static_cast<size_t>(src) * H;
  __nv_bfloat16* o = Wf + static_cast<size_t>(n) * H;
  float s1 = 0.f, s2 = 0.f;
  for (int k = lane; k < H; k += 32) {
    const float wv = w[k];
    const __nv_bfloat16 wf = __float2bfloat16_rn(wv * gamma[k]);
    o[k] = wf;
    s1 += __bfloat162float(wf);
    s2 = fmaf(wv, beta[k], s2);
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane == 0) { c1[n] = s1; c2[n] = s2 + bias[src]; }
}

}  // namespace owc
