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
//        writes v over y (+ bf16 copy) and accumulates the new row statistics                      (WEPI_RES_STATS)
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
  WEPI_RES_STATS = 2         // y = acc + bias[n] + LN_prev(y) in place (fp32) + bf16 copy; row (sum, sumsq) accumulated
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
  const float* stats_in;     // [M][2] (sum, sumsq) of the A operand's pre-LayerNorm rows
  // WEPI_RES_STATS
  const float* bias;         // [N]
  float* y;                  // [M, N] fp32 residual stream, updated in place
  const float* stats_prev;   // [M][2] statistics of y on entry
  const float* gamma_prev;   // [N] LayerNorm that turns y (on entry) into the residual
  const float* beta_prev;
  float* stats_out;          // [M][2] accumulated atomically (zeroed by an earlier kernel of the layer)
  // any epilogue
  float* stats_zero;         // [M][2] or nullptr: zeroed here for a later kernel of the layer
};

struct WideCfg {
  static constexpr int BN = 256, BM = 128, BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int BH_BYTES = (BN / 2) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + BH_BYTES;          // 32 KB
  static constexpr int STAGES = 5;
  static constexpr int STG_WARP = 6144;                           // per epilogue warp: 4 KB fp32 tile | 2 KB bf16 tile (or 2 x 2 KB bf16)
  static constexpr int PARAM_BYTES = 2 * 3 * BN * 4;              // [tile parity][3 column vectors][BN]
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + kEpiWarps * STG_WARP + 512 + PARAM_BYTES;
  static_assert(SMEM_BYTES <= kSmemLimit, "shared memory budget");
};

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
k_gemm_wide(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmD, const WideArgs g) {
  using Cfg = WideCfg;
  constexpr int BN = Cfg::BN, BM = Cfg::BM, BK = Cfg::BK, STAGES = Cfg::STAGES;
  constexpr bool RES = EPI == WEPI_RES_STATS;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stg_all = tiles + STAGES * Cfg::STAGE_BYTES;
  uint8_t* tail = stg_all + kEpiWarps * Cfg::STG_WARP;
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
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
        for (int tile = tile0; tile < tile1; ++tile) {
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = tiles + stage * Cfg::STAGE_BYTES;
            const uint32_t bar = mapa_u32(smem_u32(&full[stage]), 0);
            if (leader) mbar_expect_tx(&full[stage], 2u * Cfg::STAGE_BYTES);
            tma_load_2d_2sm(sa, &tmA, bar, kb * BK, m0);
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
      for (int it = cluster_id; it < n_items; it += n_clusters) {
        const int sp = it % g.splits;
        const int tile0 = sp * g.tiles_per_item, tile1 = min(tile0 + g.tiles_per_item, tiles_total);
        for (int tile = tile0; tile < tile1; ++tile, ++t) {
          const uint32_t buf = t & 1u;
          mbar_wait(&tempty[buf], ((t >> 1) & 1u) ^ 1u);   // both CTAs' epilogues have drained this accumulator
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * BN;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&full[stage], phase);
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
    for (int it = cluster_id; it < n_items; it += n_clusters) {
      const int bp = it / g.splits, sp = it - bp * g.splits;
      const int m0 = bp * 2 * BM + static_cast<int>(rank) * BM;
      const int row0 = m0 + q * 32;
      const int row = row0 + lane;
      const bool live = row < g.M;
      const int tile0 = sp * g.tiles_per_item, tile1 = min(tile0 + g.tiles_per_item, tiles_total);
      // per-row LayerNorm of the incoming stream: x = y * ra + rb
      float ra = 0.f, rb = 0.f;
      {
        const float* st = RES ? g.stats_prev : g.stats_in;
        if (live) {
          const float2 s2 = __ldg(reinterpret_cast<const float2*>(st) + row);
          const float mu = s2.x * g.inv_h;
          ra = rsqrtf(fmaxf(s2.y * g.inv_h - mu * mu, 0.f) + g.ln_eps);
          rb = -mu * ra;
        }
      }
      if (g.stats_zero != nullptr && live && sp == 0 && half == 0)
        reinterpret_cast<float2*>(g.stats_zero)[row] = make_float2(0.f, 0.f);
      float sum = 0.f, sq = 0.f;
      for (int tile = tile0; tile < tile1; ++tile, ++t) {
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
        // residual chunk of the first own column chunk, fetched before the accumulator is ready
        float4 res[8];
        auto fetch_res = [&](int n0) {
          if constexpr (RES) {
            if (live) {
              const float4* r4 = reinterpret_cast<const float4*>(g.y + static_cast<size_t>(row) * g.N + n0);
#pragma unroll
              for (int j = 0; j < 8; ++j) res[j] = r4[j];
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) res[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
        };
        fetch_res(tile * BN + half * 32);
        named_bar_sync(1, kEpiThreads);       // column parameters of this tile staged (the previous use of this
                                              // parity was two tiles ago: every warp has passed that tile's barrier)
        mbar_wait(&tfull[buf], (t >> 1) & 1u);
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
            float4 cur[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) cur[j] = res[j];
            if (cw + 1 < NCW) fetch_res(n0 + 64);   // next own chunk's residual flies under this chunk's math
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = *reinterpret_cast<const float4*>(p0 + c * 32 + 4 * j);
              const float4 gm = *reinterpret_cast<const float4*>(p0 + BN + c * 32 + 4 * j);
              const float4 bt = *reinterpret_cast<const float4*>(p0 + 2 * BN + c * 32 + 4 * j);
              const float x0 = __uint_as_float(vc[4 * j + 0]) + b.x + fmaf(fmaf(cur[j].x, ra, rb), gm.x, bt.x);
              const float x1 = __uint_as_float(vc[4 * j + 1]) + b.y + fmaf(fmaf(cur[j].y, ra, rb), gm.y, bt.y);
              const float x2 = __uint_as_float(vc[4 * j + 2]) + b.z + fmaf(fmaf(cur[j].z, ra, rb), gm.z, bt.z);
              const float x3 = __uint_as_float(vc[4 * j + 3]) + b.w + fmaf(fmaf(cur[j].w, ra, rb), gm.w, bt.w);
              sum += (x0 + x1) + (x2 + x3);
              sq = fmaf(x0, x0, sq); sq = fmaf(x1, x1, sq); sq = fmaf(x2, x2, sq); sq = fmaf(x3, x3, sq);
              f[4 * j + 0] = x0; f[4 * j + 1] = x1; f[4 * j + 2] = x2; f[4 * j + 3] = x3;
            }
            if (lane == 0) bulk_wait_group_read<0>();   // the stores that last used the staging buffers have read them
            __syncwarp();
            stage_chunk<EPI_BIAS_F32>(stg, lane, f);
            stage_chunk<EPI_BIAS_BF16>(stg + 4096, lane, f);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmC, stg, n0, row0);
              tma_store_2d(&tmD, stg + 4096, n0, row0);
              bulk_commit_group();
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 k1 = *reinterpret_cast<const float4*>(p0 + c * 32 + 4 * j);
              const float4 k2 = *reinterpret_cast<const float4*>(p0 + BN + c * 32 + 4 * j);
              f[4 * j + 0] = fmaf(__uint_as_float(vc[4 * j + 0]), ra, fmaf(rb, k1.x, k2.x));
              f[4 * j + 1] = fmaf(__uint_as_float(vc[4 * j + 1]), ra, fmaf(rb, k1.y, k2.y));
              f[4 * j + 2] = fmaf(__uint_as_float(vc[4 * j + 2]), ra, fmaf(rb, k1.z, k2.z));
              f[4 * j + 3] = fmaf(__uint_as_float(vc[4 * j + 3]), ra, fmaf(rb, k1.w, k2.w));
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
        if (live) {
          atomicAdd(g.stats_out + 2 * static_cast<size_t>(row), sum);
          atomicAdd(g.stats_out + 2 * static_cast<size_t>(row) + 1, sq);
        }
      }
    }
    if (lane == 0) bulk_wait_group<0>();   // staged tiles fully written before the CTA retires
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
