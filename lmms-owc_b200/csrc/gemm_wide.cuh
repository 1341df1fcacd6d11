// gemm_wide.cuh - the contraction kernel of the WIDE encoder path (H = 768: mpnet-base), where a token row
// (768 fp32 columns) does not fit the 512 TMEM columns and the block-resident encoder of enc_fused.cuh cannot be used.
//
//   C[M, N] = A[M, K] . W[N, K]^T      A, W bf16 K-major; CTA PAIRS (cluster of 2, cta_group::2, UMMA 256 x 256 x 16)
//
// Two things set it apart from k_gemm_tc2 (gemm_tc.cuh):
//
// 1. PERSISTENT: the grid is one cluster per SM pair; every cluster walks a list of work items (row-block pair x column
//    split) with its three pipelines (TMA ring, TMEM accumulator pair, epilogue staging) running straight through the
//    item boundaries. A row block's last epilogue overlaps the next block's first MMAs, and the per-CTA prologue (TMEM
//    allocation, barrier init, cluster sync, first TMA round trip) is paid once per launch instead of once per 128 rows -
//    it was 30-40 % of the O-projection (3 column tiles per row block) in the one-block-per-CTA kernel.
//
// 2. DEFERRED LayerNorm: a post-LN encoder layer is  x1 = LN1(x + Wo a + bo),  x2 = LN2(x1 + W2 gelu(W1 x1 + b1) + b2)
//    (modeling_mpnet.py:276-322, modeling_bert.py:287-356). With 768 columns the normalisation cannot happen in the
//    epilogue that produces the sum (the row statistics need all three 256-column tiles, TMEM holds two), and a separate
//    LayerNorm kernel costs 10.5 KB of L2/HBM traffic per token and LayerNorm. Here the residual stream is carried as the
//    PRE-LayerNorm sum y (fp32, + a bf16 operand copy) together with the row statistics (sum, sum of squares), and every
//    consumer applies the LayerNorm on the fly:
//      * the GEMM that reads LN(y) as its A operand multiplies bf16(y) with gamma-folded weights W' = bf16(W diag(gamma))
//        and finishes in its epilogue:   LN(y) W^T + b  =  rstd * (acc - mu * c1[n]) + c2[n],
//        c1[n] = sum_k W'[n,k],  c2[n] = sum_k W[n,k] beta[k] + b[n]          (WEPI_FOLD_BF16 / WEPI_FOLD_GELU_BF16)
//      * the GEMM that adds the residual normalises it while adding:   v = acc + bias + (y - mu) * rstd * gamma + beta,
//        writes v over y (+ bf16 copy) and leaves the new row statistics                           (WEPI_RES_STATS)
//    Row statistics are kept as partial (sum, sumsq) pairs in SLOTS - one per (column split, chunk parity) of the
//    producing kernel, each written with a plain store by exactly one thread - and summed in slot order by the
//    consumer: no atomics, so the result does not depend on the order in which CTAs finish (bit-reproducible).
//    The numerics are those of the straightforward scheme (tools/lnfold_emulate.py: max |d cos| 1.15e-3 vs 1.12e-3
//    against the fp32 oracle; the bf16 rounding moves from LN(y) to y, whose row mean is small against its spread).
//
// Roles per CTA as in gemm_tc.cuh: warp 0 lane 0 TMA producer, warp 1 of the leader MMA issuer for the pair, warps 2-9
// epilogue (thread = output row of a TMEM lane quadrant, two warps per quadrant on alternate 32-column chunks).
#pragma once

#include "common.cuh"
#include "gemm_tc.cuh"

namespace owc {

enum : int {
  WEPI_FOLD_BF16 = 0,        // out bf16 = rstd_r * (acc - mu_r * c1[n]) + c2[n]
  WEPI_FOLD_GELU_BF16 = 1,   // out bf16 = gelu_erf(the same)
  WEPI_RES_STATS = 2,        // y = acc + bias[n] + LN_prev(y) in place, y as a bf16 pair hi | lo (tmD | tmC); row (sum, sumsq) left in slots
  WEPI_RES_STATS_DEEP = 3    // the same for long K (FFN2, K = 3072): a 5-stage operand ring paid for with a single residual
                             // buffer per warp (the residual fetch is exposed, which a 24 k clk column tile hides easily)
};

struct WideArgs {
  int M, N, K;               // K % 64 == 0, N % 256 == 0
  int n_pairs;               // ceil(M / 256)
  int splits;                // column splits per row-block pair (1 unless there are fewer pairs than clusters)
  int tiles_per_item;        // 256-column tiles per work item
  float inv_h;               // 1 / (row width the statistics were taken over)
  float ln_eps;
  // WEPI_FOLD_*
  const float* c1;           // [N] row sums of the gamma-folded bf16 weights
  const float* c2;           // [N] W beta + bias
  const float* stats_in;     // [slots_in][stat_stride][2] partial (sum, sumsq) of the A operand's pre-LayerNorm rows
  // WEPI_RES_STATS
  const float* bias;         // [N]
  const float* stats_prev;   // partial statistics of y on entry (same layout, slots_in slots)
  const float* gamma_prev;   // [N] LayerNorm that turns y (on entry) into the residual
  const float* beta_prev;
  float* stats_out;          // [2 * splits][stat_stride][2]: slot (split, chunk parity) written by this kernel
  // any epilogue
  int slots_in;              // partial-statistics slots of the incoming stream (1 after the embedding kernel)
  long long stat_stride;     // rows between two slots of a statistics buffer
  long long* dbg;            // DBG builds: [4 epilogue kinds][8] cycle counters, accumulated with atomics; else unused
  int rotate;                // cluster c walks the column tiles of its items starting at tile (c mod tiles): at any moment the
                             // clusters stream DIFFERENT weight tiles instead of all hitting the same L2 lines in lockstep
  int exp;                   // DBG builds only, WRONG results: bit 0 = load the A block for the first column tile of an item only
                             // (what the ring would carry if A stayed on chip: is the kernel bound by L2 -> SM traffic?)
};

template <int EPI>
struct WideCfg {
  static constexpr int BN = 256, BM = 128, BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int BH_BYTES = (BN / 2) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + BH_BYTES;          // 32 KB
  // per epilogue warp. fold: 2 x 2 KB bf16 output tiles. res: 2 x 4 KB = the hi | lo bf16 tiles of a residual chunk (they land
  // here by TMA, are updated in place and TMA-stored back)
  // O-projection: 2 buffers + 4 ring stages. Measured on one box: 1 buffer (and 5 stages) 45.7 ms per 300 k rows, 2 buffers
  // 35.8 ms, 3 buffers (which leave 3 ring stages) 36.6 ms.
  static constexpr int RES_BUFS = EPI == WEPI_RES_STATS ? 2 : 1;
  static constexpr int STG_WARP = (EPI == WEPI_RES_STATS || EPI == WEPI_RES_STATS_DEEP) ? RES_BUFS * 4096 : 2 * 2048;
  static constexpr int STAGES = EPI == WEPI_RES_STATS ? 4 : 5;
  static constexpr int PARAM_BYTES = 2 * 3 * BN * 4;              // [tile parity][3 column vectors][BN]
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + kEpiWarps * STG_WARP + 512 + PARAM_BYTES;
  static_assert(SMEM_BYTES <= kSmemLimit, "shared memory budget");
};

// DBG: the leader's MMA warp splits its time into waiting for operands (TMA-bound), waiting for a free accumulator
// (epilogue-bound) and the rest (issue), and one epilogue warp its time waiting for accumulators; summed over CTAs into
// g.dbg[EPI * 8 + {0: ring wait, 1: accumulator wait, 2: MMA warp total, 3: epilogue tfull wait, 4: epilogue total, 5: CTAs}].
template <int EPI, bool DBG = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
k_gemm_wide(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmD, const WideArgs g) {
  using Cfg = WideCfg<EPI>;
  constexpr int BN = Cfg::BN, BM = Cfg::BM, BK = Cfg::BK, STAGES = Cfg::STAGES;
  constexpr bool RES = EPI == WEPI_RES_STATS || EPI == WEPI_RES_STATS_DEEP;
  constexpr int RB = Cfg::RES_BUFS;             // res: fp32 residual / output buffers per epilogue warp

  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stg_all = tiles + STAGES * Cfg::STAGE_BYTES;
  uint8_t* tail = stg_all + kEpiWarps * Cfg::STG_WARP;
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint64_t* rbar = reinterpret_cast<uint64_t*>(tail + 256);   // [kEpiWarps][3] res: residual chunk (hi | lo tiles) landed in one of the warp's buffers
  float* param_s = reinterpret_cast<float*>(tail + 512);   // [2][3][BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  const int tiles_total = g.N / BN;
  const int n_items = g.n_pairs * g.splits;
  const int num_kb = g.K / BK;

  if (warp == 1 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    if constexpr (RES) tma_prefetch_desc(&tmD);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 2 * kEpiWarps); }
    if constexpr (RES) {
      for (int i = 0; i < 3 * kEpiWarps; ++i) mbar_init(&rbar[i], 1);
    }
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
  pdl_launch_dependents();   // the next kernel of the layer may be scheduled as CTAs of this one retire
  pdl_wait();                // everything below touches what the previous kernel wrote

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer: this CTA's 128 A rows and its half of every W tile ----------------
      int stage = 0;
      uint32_t phase = 0;
      for (int it = cluster_id; it < n_items; it += n_clusters) {
        const int bp = it / g.splits, sp = it - bp * g.splits;
        const int m0 = bp * 2 * BM + static_cast<int>(rank) * BM;
        const int tile0 = sp * g.tiles_per_item, tile1 = min(tile0 + g.tiles_per_item, tiles_total);
        const int nt = tile1 - tile0, rot = g.rotate ? cluster_id % nt : 0;
        for (int j = 0; j < nt; ++j) {
          const int tile = tile0 + (j + rot) % nt;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = tiles + stage * Cfg::STAGE_BYTES;
            const uint32_t bar = mapa_u32(smem_u32(&full[stage]), 0);
            const bool skip_a = DBG && (g.exp & 1) && j != 0;
            if (leader) mbar_expect_tx(&full[stage], skip_a ? 2u * Cfg::BH_BYTES : 2u * Cfg::STAGE_BYTES);
            if (!skip_a) tma_load_2d_2sm(sa, &tmA, bar, kb * BK, m0);
            tma_load_2d_2sm(sa + Cfg::A_BYTES, &tmB, bar, kb * BK, tile * BN + static_cast<int>(rank) * (BN / 2));
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ---------------- MMA issuer (leader; whole warp runs the loop, one elected lane issues) ----------------
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN);
      const uint32_t tiles_lo = sw128_desc_lo(smem_u32(tiles));
      int stage = 0;
      uint32_t phase = 0;
      uint32_t t = 0;   // accumulator tiles issued so far (across items)
      long long w_ring = 0, w_acc = 0, t_begin = 0;
      if constexpr (DBG) t_begin = clock64();
      for (int it = cluster_id; it < n_items; it += n_clusters) {
        const int sp = it % g.splits;
        const int tile0 = sp * g.tiles_per_item, tile1 = min(tile0 + g.tiles_per_item, tiles_total);
        for (int j = tile0; j < tile1; ++j, ++t) {   // the tile order (rotation) only matters to the producer and the epilogue
          const uint32_t buf = t & 1u;
          long long c0 = 0;
          if constexpr (DBG) c0 = clock64();
          mbar_wait(&tempty[buf], ((t >> 1) & 1u) ^ 1u);   // both CTAs' epilogues have drained this accumulator
          if constexpr (DBG) w_acc += clock64() - c0;
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * BN;
          for (int kb = 0; kb < num_kb; ++kb) {
            if constexpr (DBG) c0 = clock64();
            mbar_wait(&full[stage], phase);
            if constexpr (DBG) w_ring += clock64() - c0;
            tc_fence_after();
            const uint32_t a_lo = tiles_lo + static_cast<uint32_t>(stage * (Cfg::STAGE_BYTES >> 4));
            if (elect_one_sync()) {
              umma_kblock_2sm(d_tmem, a_lo, a_lo + (Cfg::A_BYTES >> 4), idesc, static_cast<uint32_t>(kb != 0));
              umma_commit_2sm(&empty[stage]);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (elect_one_sync()) umma_commit_2sm(&tfull[buf]);
          __syncwarp();
        }
      }
      if constexpr (DBG) {
        if (lane == 0 && g.dbg != nullptr) {
          atomicAdd(reinterpret_cast<unsigned long long*>(g.dbg + EPI * 8 + 0), static_cast<unsigned long long>(w_ring));
          atomicAdd(reinterpret_cast<unsigned long long*>(g.dbg + EPI * 8 + 1), static_cast<unsigned long long>(w_acc));
          atomicAdd(reinterpret_cast<unsigned long long*>(g.dbg + EPI * 8 + 2), static_cast<unsigned long long>(clock64() - t_begin));
          atomicAdd(reinterpret_cast<unsigned long long*>(g.dbg + EPI * 8 + 5), 1ull);
        }
      }
    }
  } else {
    // ---------------- epilogue (warps 2..9 of both CTAs, each CTA on its own 128 rows) ----------------
    constexpr int NCW = BN / 64;   // 32-column chunks per warp and tile
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;
    uint8_t* stg = stg_all + (warp - 2) * Cfg::STG_WARP;
    uint32_t t = 0;
    uint32_t store_count = 0;
    uint32_t res_n = 0;   // res: residual chunks fetched / consumed so far by this warp (buffer = parity, barrier phase = bit 1)
    long long w_tfull = 0, e_begin = 0;
    if constexpr (DBG) e_begin = clock64();
    for (int it = cluster_id; it < n_items; it += n_clusters) {
      const int bp = it / g.splits, sp = it - bp * g.splits;
      const int m0 = bp * 2 * BM + static_cast<int>(rank) * BM;
      const int row0 = m0 + q * 32;
      const int row = row0 + lane;
      const bool live = row < g.M;
      const int tile0 = sp * g.tiles_per_item, tile1 = min(tile0 + g.tiles_per_item, tiles_total);
      // per-row LayerNorm of the incoming stream: x = y * ra + rb
      float ra = 0.f, rb_ = 0.f;
      {
        const float2* st = reinterpret_cast<const float2*>(RES ? g.stats_prev : g.stats_in);
        if (live) {
          float2 s2 = __ldg(st + row);
          for (int sl = 1; sl < g.slots_in; ++sl) {   // fixed slot order: bit-reproducible
            const float2 p = __ldg(st + sl * g.stat_stride + row);
            s2.x += p.x;
            s2.y += p.y;
          }
          const float mu = s2.x * g.inv_h;
          ra = rsqrtf(fmaxf(s2.y * g.inv_h - mu * mu, 0.f) + g.ln_eps);
          rb_ = -mu * ra;
        }
      }
      float sum = 0.f, sq = 0.f;
      const int nt = tile1 - tile0, rot = g.rotate ? cluster_id % nt : 0;
      for (int j = 0; j < nt; ++j, ++t) {
        const int tile = tile0 + (j + rot) % nt;
        const int tile_next = tile0 + (j + 1 + rot) % nt;
        const uint32_t buf = t & 1u;
        float* p0 = param_s + buf * 3 * BN;   // fold: c1 | c2     res: bias | gamma | beta
        {
          const int n = tile * BN + et;       // 256 epilogue threads, one column each
          if constexpr (RES) {
            p0[et] = g.bias ? __ldg(g.bias + n) : 0.f;
            p0[BN + et] = __ldg(g.gamma_prev + n);
            p0[2 * BN + et] = __ldg(g.beta_prev + n);
          } else {
            p0[et] = __ldg(g.c1 + n);
            p0[BN + et] = __ldg(g.c2 + n);
          }
        }
        // res: the residual chunk (32 rows x 32 fp32 of y) is brought in by TMA - coalesced, no LSU wavefronts - into one of
        // the warp's two fp32 buffers, one chunk ahead; the thread updates its row in place and the buffer is TMA-stored back
        // over y. (Row-per-thread global loads cost 32 LSU wavefronts per instruction: 8 warps x 4 chunks x 256 wavefronts per
        // column tile exceeded the 6.1 k clk the O-projection's MMAs take and held that kernel at 0.46 PFLOP/s.)
        // Buffers are used round-robin by the warp's running chunk count (seq); the fetch runs LA = RB - 1 chunks ahead of the
        // arithmetic (one chunk ahead is less than a DRAM round trip: with a single buffer the O-projection takes 30 % longer).
        constexpr int LA = RB - 1;
        const int n_chunks = nt * NCW;
        auto chunk_col = [&](int k) {            // first column of chunk k of this item (k = tile step * NCW + chunk step)
          const int jj = k / NCW, cc = k - jj * NCW;
          return (tile0 + (jj + rot) % nt) * BN + (2 * cc + half) * 32;
        };
        auto fetch_seq = [&](uint32_t seq, int n0) {   // lane 0; the buffer's previous TMA stores have finished reading it
          uint64_t* rb = &rbar[(warp - 2) * 3 + (seq % RB)];
          uint8_t* b = stg + (seq % RB) * 4096;
          mbar_expect_tx(rb, 4096u);
          tma_load_2d(b, &tmD, rb, n0, row0);          // hi tile (32 x 32 bf16, SWIZZLE_64B)
          tma_load_2d(b + 2048, &tmC, rb, n0, row0);   // lo tile
        };
        if constexpr (RES) {
          if (j == 0) {
            if (lane == 0) {
              bulk_wait_group_read<0>();   // the previous item's stores have read every buffer
              for (int i = 0; i < (LA > 0 ? LA : 1) && i < n_chunks; ++i) fetch_seq(res_n + i, chunk_col(i));
            }
            __syncwarp();
          }
        }
        named_bar_sync(1, kEpiThreads);       // column parameters of this tile staged (the previous use of this
                                              // parity was two tiles ago: every warp has passed that tile's barrier)
        long long c0 = 0;
        if constexpr (DBG) c0 = clock64();
        mbar_wait(&tfull[buf], (t >> 1) & 1u);
        if constexpr (DBG) w_tfull += clock64() - c0;
        tc_fence_after();
        const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN;
        uint32_t v[2][32];
        tmem_ld_32x32(tbase + static_cast<uint32_t>(half * 32), v[0]);
#pragma unroll
        for (int cw = 0; cw < NCW; ++cw) {
          const int c = 2 * cw + half;
          const int n0 = tile * BN + c * 32;
          tmem_ld_wait_dep(v[cw & 1]);
          if (cw + 1 < NCW) tmem_ld_32x32(tbase + static_cast<uint32_t>((c + 2) * 32), v[(cw + 1) & 1]);
          const uint32_t (&vc)[32] = v[cw & 1];
          float f[32];
          if constexpr (RES) {
            uint8_t* fbuf = stg + (res_n % RB) * 4096;
            uint64_t* rb = &rbar[(warp - 2) * 3 + (res_n % RB)];
            const uint32_t rpar = (res_n / RB) & 1u;
            const uint32_t seq = res_n;
            ++res_n;
            const int k = j * NCW + cw;        // this chunk's index within the item
            if (lane == 0) {
              bulk_wait_group_read<0>();   // the stores of the chunk before this one have read their buffer: it takes chunk k + LA
              if (LA > 0 && k + LA < n_chunks) fetch_seq(seq + LA, chunk_col(k + LA));
            }
            __syncwarp();
            mbar_wait(rb, rpar);
            uint8_t* hrow = fbuf + lane * 64;           // this thread's row of the hi tile, of the lo tile
            uint8_t* lrow = fbuf + 2048 + lane * 64;
            const int sw = (lane >> 1) & 3;             // SWIZZLE_64B, as TMA wrote the tiles
#pragma unroll
            for (int j = 0; j < 4; ++j) {               // 16-byte units of 8 columns
              uint4* hc = reinterpret_cast<uint4*>(hrow + ((j ^ sw) << 4));
              uint4* lc = reinterpret_cast<uint4*>(lrow + ((j ^ sw) << 4));
              const uint4 hv = *hc, lv = *lc;
              float r[8];
              join_bf16x2(hv.x, lv.x, r[0], r[1]);
              join_bf16x2(hv.y, lv.y, r[2], r[3]);
              join_bf16x2(hv.z, lv.z, r[4], r[5]);
              join_bf16x2(hv.w, lv.w, r[6], r[7]);
              float x[8];
#pragma unroll
              for (int q4 = 0; q4 < 2; ++q4) {
                const float4 b = *reinterpret_cast<const float4*>(p0 + c * 32 + 8 * j + 4 * q4);
                const float4 gm = *reinterpret_cast<const float4*>(p0 + BN + c * 32 + 8 * j + 4 * q4);
                const float4 bt = *reinterpret_cast<const float4*>(p0 + 2 * BN + c * 32 + 8 * j + 4 * q4);
                x[4 * q4 + 0] = __uint_as_float(vc[8 * j + 4 * q4 + 0]) + b.x + fmaf(fmaf(r[4 * q4 + 0], ra, rb_), gm.x, bt.x);
                x[4 * q4 + 1] = __uint_as_float(vc[8 * j + 4 * q4 + 1]) + b.y + fmaf(fmaf(r[4 * q4 + 1], ra, rb_), gm.y, bt.y);
                x[4 * q4 + 2] = __uint_as_float(vc[8 * j + 4 * q4 + 2]) + b.z + fmaf(fmaf(r[4 * q4 + 2], ra, rb_), gm.z, bt.z);
                x[4 * q4 + 3] = __uint_as_float(vc[8 * j + 4 * q4 + 3]) + b.w + fmaf(fmaf(r[4 * q4 + 3], ra, rb_), gm.w, bt.w);
              }
#pragma unroll
              for (int k = 0; k < 8; k += 4) {
                sum += (x[k] + x[k + 1]) + (x[k + 2] + x[k + 3]);
                sq = fmaf(x[k], x[k], sq); sq = fmaf(x[k + 1], x[k + 1], sq); sq = fmaf(x[k + 2], x[k + 2], sq); sq = fmaf(x[k + 3], x[k + 3], sq);
              }
              uint4 ho, lo;
              split_bf16x2(x[0], x[1], ho.x, lo.x);
              split_bf16x2(x[2], x[3], ho.y, lo.y);
              split_bf16x2(x[4], x[5], ho.z, lo.z);
              split_bf16x2(x[6], x[7], ho.w, lo.w);
              *hc = ho;                                  // in place: both tiles go back from the buffer they came into
              *lc = lo;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmD, fbuf, n0, row0);
              tma_store_2d(&tmC, fbuf + 2048, n0, row0);
              bulk_commit_group();
              if (RB == 1 && k + 1 < n_chunks) {   // one buffer: the next chunk can only come in once this store has read it
                bulk_wait_group_read<0>();
                fetch_seq(seq + 1, chunk_col(k + 1));
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 k1 = *reinterpret_cast<const float4*>(p0 + c * 32 + 4 * j);
              const float4 k2 = *reinterpret_cast<const float4*>(p0 + BN + c * 32 + 4 * j);
              f[4 * j + 0] = fmaf(__uint_as_float(vc[4 * j + 0]), ra, fmaf(rb_, k1.x, k2.x));
              f[4 * j + 1] = fmaf(__uint_as_float(vc[4 * j + 1]), ra, fmaf(rb_, k1.y, k2.y));
              f[4 * j + 2] = fmaf(__uint_as_float(vc[4 * j + 2]), ra, fmaf(rb_, k1.z, k2.z));
              f[4 * j + 3] = fmaf(__uint_as_float(vc[4 * j + 3]), ra, fmaf(rb_, k1.w, k2.w));
            }
            if constexpr (EPI == WEPI_FOLD_GELU_BF16) {
#pragma unroll
              for (int j = 0; j < 32; j += 2) gelu_fast2(f[j], f[j + 1]);
            }
            uint8_t* buf_s = stg + (store_count & 1u) * 2048;
            if (lane == 0) bulk_wait_group_read<1>();
            __syncwarp();
            stage_chunk<EPI_BIAS_BF16>(buf_s, lane, f);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmC, buf_s, n0, row0);
              bulk_commit_group();
            }
            ++store_count;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[buf]), 0));
      }
      if constexpr (RES) {
        if (live)   // this thread is the only writer of slot (split, chunk parity) of its row
          reinterpret_cast<float2*>(g.stats_out)[(2 * sp + half) * g.stat_stride + row] = make_float2(sum, sq);
      }
    }
    if (lane == 0) bulk_wait_group<0>();   // staged tiles fully written before the CTA retires
    if constexpr (DBG) {
      if (warp == 2 && lane == 0 && leader && g.dbg != nullptr) {
        atomicAdd(reinterpret_cast<unsigned long long*>(g.dbg + EPI * 8 + 3), static_cast<unsigned long long>(w_tfull));
        atomicAdd(reinterpret_cast<unsigned long long*>(g.dbg + EPI * 8 + 4), static_cast<unsigned long long>(clock64() - e_begin));
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();   // the peer's shared memory and TMEM must outlive the leader's last MMA
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

// Gamma-fold of one weight matrix at load time: W [N, K] fp32, LayerNorm (gamma, beta) [K] of the operand, bias [N] ->
//   Wf = bf16(W * gamma),  c1[n] = sum_k float(Wf[n,k]),  c2[n] = sum_k W[n,k] * beta[k] + bias[n].   Warp per row.
__global__ void __launch_bounds__(256) k_fold_weight(const float* __restrict__ W, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, const float* __restrict__ bias, int N,
                                                     int K, __nv_bfloat16* __restrict__ Wf, float* __restrict__ c1,
                                                     float* __restrict__ c2) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  const float* w = W + static_cast<size_t>(n) * K;
  __nv_bfloat16* o = Wf + static_cast<size_t>(n) * K;
  float s1 = 0.f, s2 = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float wv = w[k];
    const __nv_bfloat16 wf = __float2bfloat16_rn(wv * gamma[k]);
    o[k] = wf;
    s1 += __bfloat162float(wf);
    s2 = fmaf(wv, beta[k], s2);
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane == 0) { c1[n] = s1; c2[n] = s2 + bias[n]; }
}

}  // namespace owc
