// gemm_tc.cuh - the tcgen05 / TMEM / TMA contraction kernels of the scoring path.
//
//   C[M, N] = A[M, K] . W[N, K]^T        A, W bf16, K-major (row-major with K innermost)
//
// One skeleton serves every dense contraction on the path: a CTA owns
// one 128-row block of A and sweeps a range of BN-wide column tiles of W. Roles:
//   warp 0 (one lane)  TMA producer: per k-block of 64, a 128x64 A box and a W box into a
//                      SWIZZLE_128B stage ring, completion on `full[stage]` (expect_tx)
//   warp 1 (one lane)  MMA issuer: 4 x tcgen05.mma (K=16 each) per stage into one of TWO TMEM
//                      accumulator buffers; tcgen05.commit frees the stage and, after the last
//                      k-block, publishes the accumulator on `tfull[buf]`
//   warps 2-9          epilogue: thread r of TMEM lane quadrant (warp & 3) owns output row r, the two
//                      warps of a quadrant take alternate 32-column chunks (tcgen05.ld, software-
//                      pipelined); fused bias / erf-GELU / residual add, results staged in swizzled
//                      shared memory and written with TMA stores (a row-per-thread global store
//                      costs 32 LSU wavefronts per instruction and bounded the first version);
//                      or the per-row top-k candidate list kept in registers across the whole class
//                      sweep (the similarity matrix never leaves the SM)
// The two accumulator buffers let the epilogue of tile t overlap the MMAs of tile t+1.
//
// Two kernels share the epilogue:
//   k_gemm_tc   one CTA per 128 rows (cta_group::1)
//   k_gemm_tc2  a CTA PAIR per 256 rows (cluster of 2, cta_group::2, UMMA M=256): each CTA stages
//               only HALF of every W tile and the pair's tensor cores read both halves. With ARES the
//               CTA's whole A block (K <= 384) stays resident in shared memory for the sweep.
#pragma once

#include "common.cuh"

namespace owc {

enum : int {
  EPI_BIAS_BF16 = 0,       // out bf16 = acc + bias
  EPI_BIAS_GELU_BF16 = 1,  // out bf16 = gelu_erf(acc + bias)
  EPI_BIAS_RES_F32 = 2,    // out fp32 = acc + bias + residual   (pre-LayerNorm sum)
  EPI_BIAS_F32 = 3,        // out fp32 = acc (+ bias)            (debug / check of the mainloop)
  EPI_TOPK = 4,            // per-row top-KC candidates over all columns handled by this CTA
  EPI_RES_LN = 5           // x = LayerNorm(acc + bias + x) in place (fp32) + bf16 copy; N == 2*BN (whole rows in TMEM)
};

struct GemmArgs {
  int M, N, K;             // K % 64 == 0
  int tiles_per_cta;       // column tiles swept by one CTA (gridDim.y = ceil(tiles / tiles_per_cta))
  const float* bias;       // [N] or nullptr
  const float* residual;   // [M, ldo] fp32 (EPI_BIAS_RES_F32)
  int ldo;                 // leading dimension of `residual`
  float* cand_score;       // EPI_TOPK: [2 * gridDim.y, M, KC] (two chunk parities per column split)
  int* cand_idx;
  const float* gamma;      // EPI_RES_LN: LayerNorm weight / bias [N] and epsilon
  const float* beta;
  float ln_eps;
  long long* dbg;          // optional per-CTA timeline stamps [cta][8] (clock64), nullptr in production
};

constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kGemmThreads = 64 + kEpiThreads;
constexpr int kSmemLimit = 232448;   // 227 KB of dynamic shared memory per CTA

// bytes of one per-warp output staging buffer: 32 rows x 32 columns
__host__ __device__ constexpr int epi_stage_bytes(int epi) {
  return epi == EPI_TOPK ? 0 : ((epi == EPI_BIAS_BF16 || epi == EPI_BIAS_GELU_BF16) ? 32 * 64 : 32 * 128);
}
__host__ __device__ constexpr int epi_fixed_bytes(int bn, int epi) {
  // EPI_RES_LN: 4 epilogue warps x 2 residual/staging buffers of 4 KB, and bias|gamma|beta of a whole row (2*bn each)
  return epi == EPI_RES_LN ? 1024 + 512 + 3 * 2 * bn * 4 + (kEpiWarps * 4096 + 2048)
                           : 1024 /*align slack*/ + 512 /*barriers*/ + 2 * bn * 4 /*bias*/ + kEpiWarps * 2 * epi_stage_bytes(epi);
}
__host__ __device__ constexpr int imin(int a, int b) { return a < b ? a : b; }

// ARES: the CTA's whole A block (128 x K, K <= 384) is loaded once and stays resident, so a stage only
// carries the W tile. The kernel is shared-memory-bandwidth bound (every k-block is written by TMA and
// read by the tensor core: 2 x 40 KB at BN = 192 against ~128 B/clk), so dropping the A refills of
// every column tile is a direct 20 % cut of the per-k-block traffic.
template <int BN, int EPI, bool ARES = false>
struct GemmCfg {
  static constexpr int BM = 128;
  static constexpr int BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = (ARES ? 0 : A_BYTES) + B_BYTES;
  static constexpr int ARES_BYTES = ARES ? 6 * A_BYTES : 0;
  static constexpr int STAGES = imin(8, (kSmemLimit - epi_fixed_bytes(BN, EPI) - ARES_BYTES) / STAGE_BYTES);
  static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
  static constexpr int STG_BYTES = epi_stage_bytes(EPI);
  // EPI_RES_LN: a 4 KB transpose scratch per epilogue warp + the 2 KB row-statistics exchange area
  static constexpr int STG_TOTAL = (EPI == EPI_RES_LN) ? kEpiWarps * 4096 + 2048 : kEpiWarps * 2 * STG_BYTES;
  static constexpr int SMEM_BYTES = ARES_BYTES + STAGES * STAGE_BYTES + epi_fixed_bytes(BN, EPI);
  static_assert(STAGES >= 3, "pipeline too shallow");
};

// CTA-pair kernel: per CTA a stage holds (optionally) a 128x64 A box and HALF a W tile (BN/2 rows).
template <int BN, int EPI, bool ARES>
struct Gemm2Cfg {
  static constexpr int BM = 128;          // rows per CTA (256 per pair)
  static constexpr int BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int BH_BYTES = (BN / 2) * BK * 2;
  static constexpr int STAGE_BYTES = (ARES ? 0 : A_BYTES) + BH_BYTES;
  static constexpr int MAX_KB = 6;        // resident A block: K <= 384
  static constexpr int ARES_BYTES = ARES ? MAX_KB * A_BYTES : 0;
  static constexpr int STAGES = imin(8, (kSmemLimit - epi_fixed_bytes(BN, EPI) - ARES_BYTES) / STAGE_BYTES);
  static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
  static constexpr int STG_BYTES = epi_stage_bytes(EPI);
  static constexpr int SMEM_BYTES = ARES_BYTES + STAGES * STAGE_BYTES + epi_fixed_bytes(BN, EPI);
  static_assert(STAGES >= 3, "pipeline too shallow");
};
constexpr int kAresMaxK = 384;

template <int KC>
struct TopK {
  float s[KC];
  int id[KC];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < KC; ++i) { s[i] = -INFINITY; id[i] = -1; }
  }
  // Keeps s[] descending; a value equal to an entry goes after it (stable), so equal scores keep the
  // order in which they are inserted. Caller guarantees v > s[KC-1].
  __device__ __forceinline__ void insert(float v, int col) {
    float cv = v;
    int ci = col;
#pragma unroll
    for (int i = 0; i < KC; ++i) {
      const bool gt = cv > s[i];
      const float ts = s[i];
      const int ti = id[i];
      if (gt) { s[i] = cv; id[i] = ci; cv = ts; ci = ti; }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// shared epilogue pieces (executed by the 8 epilogue warps: threadIdx.x in [64, 320))
// ------------------------------------------------------------------------------------------------
template <int BN, int EPI>
__device__ __forceinline__ void epi_stage_bias(const GemmArgs& g, int tile, float* bias_s) {
  if constexpr (EPI != EPI_TOPK) {
    // stage this tile's bias slice in shared memory while the MMAs of the tile are still running
    const int et = threadIdx.x - 64;
    for (int i = et; i < BN; i += kEpiThreads) {
      const int n = tile * BN + i;
      bias_s[i] = (g.bias != nullptr && n < g.N) ? __ldg(g.bias + n) : 0.f;
    }
    named_bar_sync(1, kEpiThreads);
  }
}

// Top-k over one 32-column chunk held in registers. Branch-free in the common case: one max-reduce
// and one compare against the current k-th best. Only when some element qualifies does the (compact,
// single-site) extraction loop run: take the maximum, insert it, knock it out, repeat. Equal values
// are extracted lowest column first and inserted after their equals, so ties resolve to the lower
// class index exactly as a column-order scan would.
template <int KC>
__device__ __forceinline__ void topk_chunk(uint32_t (&v)[32], int n0, int n_valid, TopK<KC>& top) {
  if (n_valid < 32) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j >= n_valid) v[j] = 0xff800000u;   // -inf
  }
  float m = __uint_as_float(v[0]);
#pragma unroll
  for (int j = 1; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[j]));
  while (m > top.s[KC - 1]) {
    int jj = 31;
#pragma unroll
    for (int j = 30; j >= 0; --j) jj = (__uint_as_float(v[j]) == m) ? j : jj;
    top.insert(m, n0 + jj);
    float m2 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      v[j] = (j == jj) ? 0xff800000u : v[j];
      m2 = fmaxf(m2, __uint_as_float(v[j]));
    }
    m = m2;
  }
}

// Write one 32x32 chunk (this thread: 32 consecutive values of row `lane`) into a per-warp staging
// buffer laid out for a TMA store: bf16 rows of 64 B under SWIZZLE_64B, fp32 rows of 128 B under
// SWIZZLE_128B (16-byte unit index XORed with the row bits TMA expects) - conflict-free stores.
template <int EPI>
__device__ __forceinline__ void stage_chunk(uint8_t* stg, int lane, const float (&f)[32]) {
  if constexpr (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16) {
    uint8_t* rowp = stg + lane * 64;
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 w;
      w.x = pack_bf16x2(f[8 * j + 0], f[8 * j + 1]);
      w.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
      w.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
      w.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
      *reinterpret_cast<uint4*>(rowp + ((j ^ sw) << 4)) = w;
    }
  } else {
    uint8_t* rowp = stg + lane * 128;
    const int sw = lane & 7;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<float4*>(rowp + ((j ^ sw) << 4)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
  }
}

// Drain one accumulator buffer. tbase = TMEM address of (this warp's lane quadrant, buffer column 0);
// stg = this warp's two staging buffers; row0 = first output row of this warp's lane quadrant.
template <int BN, int EPI, int KC>
__device__ __forceinline__ void epi_process(const GemmArgs& g, const CUtensorMap* tmC, int tile, uint32_t tbase,
                                            const float* bias_s, uint8_t* stg, int half, int lane, int row0,
                                            uint32_t& store_count, TopK<KC>& top) {
  constexpr int NC = BN / 32;   // 32-column chunks per tile
  constexpr int NCW = NC / 2;   // chunks per warp (two warps per lane quadrant)
  if constexpr (EPI == EPI_TOPK) {
#pragma unroll 1
    for (int cw = 0; cw < NCW; ++cw) {
      const int c = 2 * cw + half;
      const int n0 = tile * BN + c * 32;
      if (n0 >= g.N) break;  // warp-uniform
      uint32_t v[32];
      tmem_ld_32x32(tbase + static_cast<uint32_t>(c * 32), v);
      tmem_ld_wait_dep(v);
      topk_chunk<KC>(v, n0, g.N - n0, top);
    }
  } else {
    constexpr int STG = epi_stage_bytes(EPI);
    const int row = row0 + lane;
    uint32_t v[2][32];
    tmem_ld_32x32(tbase + static_cast<uint32_t>(half * 32), v[0]);
#pragma unroll
    for (int cw = 0; cw < NCW; ++cw) {
      const int c = 2 * cw + half;
      const int n0 = tile * BN + c * 32;
      tmem_ld_wait_dep(v[cw & 1]);
      if (cw + 1 < NCW) tmem_ld_32x32(tbase + static_cast<uint32_t>((c + 2) * 32), v[(cw + 1) & 1]);  // overlaps the math below
      if (n0 >= g.N) continue;  // warp-uniform
      const uint32_t (&vc)[32] = v[cw & 1];
      float f[32];
      const float4* b4 = reinterpret_cast<const float4*>(bias_s + c * 32);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b = b4[j];
        f[4 * j + 0] = __uint_as_float(vc[4 * j + 0]) + b.x;
        f[4 * j + 1] = __uint_as_float(vc[4 * j + 1]) + b.y;
        f[4 * j + 2] = __uint_as_float(vc[4 * j + 2]) + b.z;
        f[4 * j + 3] = __uint_as_float(vc[4 * j + 3]) + b.w;
      }
      if constexpr (EPI == EPI_BIAS_GELU_BF16) {
#pragma unroll
        for (int j = 0; j < 32; j += 2) gelu_fast2(f[j], f[j + 1]);
      }
      if constexpr (EPI == EPI_BIAS_RES_F32) {
        if (row < g.M) {
          const float4* r4 = reinterpret_cast<const float4*>(g.residual + static_cast<size_t>(row) * g.ldo + n0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 r = __ldg(r4 + j);
            f[4 * j + 0] += r.x; f[4 * j + 1] += r.y; f[4 * j + 2] += r.z; f[4 * j + 3] += r.w;
          }
        }
      }
      // stage in swizzled shared memory, then one TMA store per 32x32 chunk (clipped at the edges)
      uint8_t* buf = stg + (store_count & 1u) * STG;
      if (lane == 0) bulk_wait_group_read<1>();   // the store that last used this buffer has read it
      __syncwarp();
      stage_chunk<EPI>(buf, lane, f);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(tmC, buf, n0, row0);
        bulk_commit_group();
      }
      ++store_count;
    }
  }
}

// Fused residual + LayerNorm epilogue (EPI_RES_LN): N == 2*BN, so a row block's whole rows sit in the
// two TMEM accumulator buffers and thread r owns row r. All 8 epilogue warps take part; the two warps
// of a lane quadrant split the row's 32-column chunks (even / odd) and exchange their partial row
// statistics through shared memory.
//   pass A    as each accumulator completes: v = acc + bias + residual for the own chunks. The 32x32
//             fp32 residual tiles are fetched with coalesced 16-byte loads, prefetched one chunk ahead
//             in registers (the first one before the accumulator is even ready), and transposed to
//             row-per-thread through a 4 KB per-warp scratch; v goes back to TMEM (tcgen05.st) and
//             feeds the row sum / sum of squares. Tile 0's pass A overlaps tile 1's MMAs.
//   pass B    y = (v - mean) * rstd * gamma + beta -> fp32 tile TMA-stored over the residual (x is
//             updated in place) and bf16 tile TMA-stored to the operand copy; staging lives in the
//             operand memory, which is idle by then.
template <int BN>
__device__ __forceinline__ void ln_finish(const GemmArgs& g, const CUtensorMap* tmX, const CUtensorMap* tmXb,
                                          uint32_t tmem_q, uint64_t* tfull, float* stat_s, uint8_t* scratch,
                                          uint8_t* stg, const float* bias_s, const float* gamma_s,
                                          const float* beta_s, int warp_e, int q, int half, int lane, int row0) {
  constexpr int NCW = 2 * BN / 64;   // chunks per warp: NCW/2 in each accumulator buffer
  const float inv_n = 1.0f / static_cast<float>(2 * BN);
  const int lr = lane >> 3, lc = lane & 7;
  uint32_t v[2][32];
  float4 pre[8];
  auto fetch = [&](int k) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = row0 + lr + 4 * i;
      pre[i] = (row < g.M) ? __ldg(reinterpret_cast<const float4*>(g.residual + static_cast<size_t>(row) * g.ldo + k * 32) + lc)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  // ---- pass A ----
  float sum = 0.f, sq = 0.f;
  fetch(half);
#pragma unroll 1
  for (int cw = 0; cw < NCW; ++cw) {
    const int k = 2 * cw + half;
    if (cw == 0) { mbar_wait(&tfull[0], 0); tc_fence_after(); }
    if (cw == NCW / 2) { mbar_wait(&tfull[1], 0); tc_fence_after(); }
    tmem_ld_32x32(tmem_q + static_cast<uint32_t>(k * 32), v[0]);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = lr + 4 * i;
      *reinterpret_cast<float4*>(scratch + r * 128 + ((lc ^ (r & 7)) << 4)) = pre[i];
    }
    __syncwarp();
    if (cw + 1 < NCW) fetch(k + 2);   // next tile's residual flies while this chunk is processed
    const uint8_t* rrow = scratch + lane * 128;
    tmem_ld_wait_dep(v[0]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 r = *reinterpret_cast<const float4*>(rrow + ((j ^ (lane & 7)) << 4));
      const float4 b = *reinterpret_cast<const float4*>(bias_s + k * 32 + 4 * j);
      const float x0 = __uint_as_float(v[0][4 * j + 0]) + b.x + r.x;
      const float x1 = __uint_as_float(v[0][4 * j + 1]) + b.y + r.y;
      const float x2 = __uint_as_float(v[0][4 * j + 2]) + b.z + r.z;
      const float x3 = __uint_as_float(v[0][4 * j + 3]) + b.w + r.w;
      sum += (x0 + x1) + (x2 + x3);
      sq = fmaf(x0, x0, sq); sq = fmaf(x1, x1, sq); sq = fmaf(x2, x2, sq); sq = fmaf(x3, x3, sq);
      v[0][4 * j + 0] = __float_as_uint(x0); v[0][4 * j + 1] = __float_as_uint(x1);
      v[0][4 * j + 2] = __float_as_uint(x2); v[0][4 * j + 3] = __float_as_uint(x3);
    }
    tmem_st_32x32(tmem_q + static_cast<uint32_t>(k * 32), v[0]);
    __syncwarp();   // scratch is rewritten by the next iteration
  }
  tmem_st_wait();
  stat_s[(warp_e * 32 + lane) * 2 + 0] = sum;
  stat_s[(warp_e * 32 + lane) * 2 + 1] = sq;
  named_bar_sync(2 + q, 64);   // the two warps that share this lane quadrant
  sum += stat_s[((warp_e ^ 4) * 32 + lane) * 2 + 0];
  sq += stat_s[((warp_e ^ 4) * 32 + lane) * 2 + 1];
  const float mean = sum * inv_n;
  const float rstd = rsqrtf(fmaxf(sq * inv_n - mean * mean, 0.f) + g.ln_eps);
  // ---- pass B: normalise, stage, TMA-store (fp32 over the residual, bf16 operand copy) ----
  tmem_ld_32x32(tmem_q + static_cast<uint32_t>(half * 32), v[0]);
#pragma unroll
  for (int cw = 0; cw < NCW; ++cw) {
    const int k = 2 * cw + half;
    tmem_ld_wait_dep(v[cw & 1]);
    if (cw + 1 < NCW) tmem_ld_32x32(tmem_q + static_cast<uint32_t>((k + 2) * 32), v[(cw + 1) & 1]);
    float f[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 gm = *reinterpret_cast<const float4*>(gamma_s + k * 32 + 4 * j);
      const float4 bt = *reinterpret_cast<const float4*>(beta_s + k * 32 + 4 * j);
      f[4 * j + 0] = (__uint_as_float(v[cw & 1][4 * j + 0]) - mean) * rstd * gm.x + bt.x;
      f[4 * j + 1] = (__uint_as_float(v[cw & 1][4 * j + 1]) - mean) * rstd * gm.y + bt.y;
      f[4 * j + 2] = (__uint_as_float(v[cw & 1][4 * j + 2]) - mean) * rstd * gm.z + bt.z;
      f[4 * j + 3] = (__uint_as_float(v[cw & 1][4 * j + 3]) - mean) * rstd * gm.w + bt.w;
    }
    uint8_t* fbuf = stg + (cw & 1) * 4096;          // fp32 tile
    uint8_t* bbuf = stg + 8192 + (cw & 1) * 2048;   // bf16 tile
    if (lane == 0) bulk_wait_group_read<1>();       // the stores that last used this buffer pair have read it
    __syncwarp();
    stage_chunk<EPI_BIAS_F32>(fbuf, lane, f);
    stage_chunk<EPI_BIAS_BF16>(bbuf, lane, f);
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(tmX, fbuf, k * 32, row0);
      tma_store_2d(tmXb, bbuf, k * 32, row0);
      bulk_commit_group();
    }
  }
  if (lane == 0) bulk_wait_group<0>();
}

template <int KC>
__device__ __forceinline__ void epi_store_candidates(const GemmArgs& g, int half, int row, const TopK<KC>& top) {
  if (row < g.M) {
    // each (column split, chunk parity) pair contributes its own candidate list
    const size_t base = (static_cast<size_t>(blockIdx.y * 2 + half) * g.M + row) * KC;
#pragma unroll
    for (int i = 0; i < KC; ++i) { g.cand_score[base + i] = top.s[i]; g.cand_idx[base + i] = top.id[i]; }
  }
}

// ------------------------------------------------------------------------------------------------
// one CTA per 128 rows
// ------------------------------------------------------------------------------------------------
template <int BN, int EPI, int KC, bool ARES>
__global__ void __launch_bounds__(kGemmThreads, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
          const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmD, const GemmArgs g) {
  using Cfg = GemmCfg<BN, EPI, ARES>;
  constexpr int BM = Cfg::BM, BK = Cfg::BK, STAGES = Cfg::STAGES;

  extern __shared__ uint8_t smem_raw[];
  // keep the pointer arithmetic on the __shared__ array itself so accesses stay in the shared window
  // (LDS/STS) instead of degrading to generic loads; SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* ares = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // resident A block (ARES)
  uint8_t* tiles = ares + Cfg::ARES_BYTES;
  uint8_t* stg_all = tiles + STAGES * Cfg::STAGE_BYTES;                    // per-warp staging buffers, 1024-aligned
  uint8_t* tail = stg_all + Cfg::STG_TOTAL;
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* afull = tempty + 2;
  uint64_t* tinit = afull + 1;                               // EPI_RES_LN: accumulators pre-loaded with residual + bias
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tinit + 1);
  float* bias_smem = reinterpret_cast<float*>(tail + 512);   // [2][BN]  (EPI_RES_LN: bias | gamma | beta, 2*BN each)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_blk = blockIdx.x;
  const int tiles_total = (g.N + BN - 1) / BN;
  const int tile0 = blockIdx.y * g.tiles_per_cta;
  const int tile1 = min(tile0 + g.tiles_per_cta, tiles_total);
  const int num_kb = g.K / BK;
  long long* dbg = g.dbg ? g.dbg + (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 8 : nullptr;
#define OWC_STAMP(i) do { if (dbg) dbg[i] = clock64(); } while (0)
  if (threadIdx.x == 0 && dbg) {
    dbg[0] = clock64();
    long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    dbg[7] = gt;
  }

  if (warp == 1 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if constexpr (EPI != EPI_TOPK) tma_prefetch_desc(&tmC);
    if constexpr (EPI == EPI_RES_LN) tma_prefetch_desc(&tmD);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], kEpiWarps); }
    mbar_init(tinit, kEpiWarps);
    mbar_init(afull, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) OWC_STAMP(1);
  pdl_launch_dependents();   // the next kernel may begin its own prologue
  pdl_wait();                // everything below reads what the previous kernel wrote

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      if constexpr (ARES) {   // the row block's A operand, all k-blocks, once
        mbar_expect_tx(afull, static_cast<uint32_t>(num_kb) * Cfg::A_BYTES);
        for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(ares + kb * Cfg::A_BYTES, &tmA, afull, kb * BK, m_blk * BM);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < tile1; ++tile) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = tiles + stage * Cfg::STAGE_BYTES;
          mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
          if constexpr (!ARES) {
            tma_load_2d(sa, &tmA, &full[stage], kb * BK, m_blk * BM);
            sa += Cfg::A_BYTES;
          }
          tma_load_2d(sa, &tmB, &full[stage], kb * BK, tile * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    {
      // ---------------- MMA issuer: the whole warp runs the loop (uniform operands), one elected lane issues ----------------
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      if constexpr (ARES) {
        mbar_wait(afull, 0);
        tc_fence_after();
      }
      const uint32_t tiles_lo = sw128_desc_lo(smem_u32(tiles));
      const uint32_t ares_lo = sw128_desc_lo(smem_u32(ares));
      int stage = 0;
      uint32_t phase = 0;
      int t = 0;
      for (int tile = tile0; tile < tile1; ++tile, ++t) {
        const int buf = t & 1;
        const uint32_t use = static_cast<uint32_t>(t >> 1);
        mbar_wait(&tempty[buf], (use & 1) ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          if (t == 0 && kb == 0 && lane == 0) OWC_STAMP(2);
          tc_fence_after();
          const uint32_t s_lo = tiles_lo + static_cast<uint32_t>(stage * (Cfg::STAGE_BYTES >> 4));
          const uint32_t a_lo = ARES ? ares_lo + static_cast<uint32_t>(kb * (Cfg::A_BYTES >> 4)) : s_lo;
          const uint32_t b_lo = ARES ? s_lo : s_lo + (Cfg::A_BYTES >> 4);
          if (elect_one_sync()) {
            umma_kblock_1sm(d_tmem, a_lo, b_lo, idesc, static_cast<uint32_t>(kb != 0));
            umma_commit(&empty[stage]);  // stage reusable once these MMAs have read it
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one_sync()) umma_commit(&tfull[buf]);  // accumulator complete
        __syncwarp();
      }
      if (lane == 0) OWC_STAMP(3);
    }
  } else if constexpr (EPI == EPI_RES_LN) {
    // ---------------- fused residual + LayerNorm epilogue (all 8 epilogue warps) ----------------
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int warp_e = warp - 2;
    const int row0 = m_blk * BM + q * 32;
    const uint32_t tmem_q = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float* gamma_s = bias_smem + 2 * BN;
    float* beta_s = bias_smem + 4 * BN;
    float* stat_s = reinterpret_cast<float*>(stg_all + kEpiWarps * 4096);   // [8 warps][32 rows][sum, sumsq]
    for (int i = threadIdx.x - 64; i < 2 * BN; i += kEpiThreads) {
      bias_smem[i] = g.bias ? __ldg(g.bias + i) : 0.f;
      gamma_s[i] = __ldg(g.gamma + i);
      beta_s[i] = __ldg(g.beta + i);
    }
    named_bar_sync(1, kEpiThreads);
    // pass-B staging lives at the start of the operand memory (resident A block / stage ring), idle by then
    ln_finish<BN>(g, &tmC, &tmD, tmem_q, tfull, stat_s, stg_all + warp_e * 4096, ares + warp_e * 12288, bias_smem,
                  gamma_s, beta_s, warp_e, q, half, lane, row0);
    if (threadIdx.x == 64) OWC_STAMP(5);
  } else {
    // ---------------- epilogue (warps 2..9): two warps per TMEM lane quadrant ----------------
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;   // 0 | 1
    const int row0 = m_blk * BM + q * 32;
    uint8_t* stg = stg_all + (warp - 2) * 2 * Cfg::STG_BYTES;
    uint32_t store_count = 0;
    TopK<KC> top;
    if constexpr (EPI == EPI_TOPK) top.init();
    int t = 0;
    for (int tile = tile0; tile < tile1; ++tile, ++t) {
      const int buf = t & 1;
      const uint32_t use = static_cast<uint32_t>(t >> 1);
      float* bias_s = bias_smem + buf * BN;
      epi_stage_bias<BN, EPI>(g, tile, bias_s);
      mbar_wait(&tfull[buf], use & 1);
      if (t == 0 && threadIdx.x == 64) OWC_STAMP(4);
      tc_fence_after();
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * BN);
      epi_process<BN, EPI, KC>(g, &tmC, tile, tbase, bias_s, stg, half, lane, row0, store_count, top);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);
    }
    if (threadIdx.x == 64) OWC_STAMP(5);
    if constexpr (EPI == EPI_TOPK) epi_store_candidates<KC>(g, half, row0 + lane, top);
    else if (lane == 0) bulk_wait_group<0>();   // staged tiles fully written before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) OWC_STAMP(6);
#undef OWC_STAMP
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// CTA pair per 256 rows (cluster of 2, cta_group::2)
// ------------------------------------------------------------------------------------------------
// Protocol (rank 0 = leader, issues every MMA for the pair):
//   full[s]   lives in the LEADER: expect_tx covers the bytes of BOTH CTAs' TMA loads of stage s;
//             each CTA's loads complete_tx on the leader's barrier (cta_group::2 TMA)
//   empty[s]  one per CTA: the leader's tcgen05.commit is multicast to both, freeing the stage in each
//   tfull[b]  one per CTA (multicast commit): accumulator b complete in this CTA's TMEM
//   tempty[b] lives in the LEADER: all 16 epilogue warps of the pair arrive (the peer's remotely)
//   afull     (ARES) lives in the LEADER: both CTAs' resident A blocks have landed
template <int BN, int EPI, int KC, bool ARES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
k_gemm_tc2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
           const __grid_constant__ CUtensorMap tmC, const GemmArgs g) {
  using Cfg = Gemm2Cfg<BN, EPI, ARES>;
  constexpr int BM = Cfg::BM, BK = Cfg::BK, STAGES = Cfg::STAGES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* ares = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // resident A block (ARES)
  uint8_t* tiles = ares + Cfg::ARES_BYTES;
  uint8_t* stg_all = tiles + STAGES * Cfg::STAGE_BYTES;
  uint8_t* tail = stg_all + kEpiWarps * 2 * Cfg::STG_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* afull = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(afull + 1);
  float* bias_smem = reinterpret_cast<float*>(tail + 512);   // [2][BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int m_blk = blockIdx.x;                       // blocks 2p, 2p+1 form pair p: rows [128*m_blk, +128)
  const int tiles_total = (g.N + BN - 1) / BN;
  const int tile0 = blockIdx.y * g.tiles_per_cta;
  const int tile1 = min(tile0 + g.tiles_per_cta, tiles_total);
  const int num_kb = g.K / BK;

  if (warp == 1 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if constexpr (EPI != EPI_TOPK) tma_prefetch_desc(&tmC);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 2 * kEpiWarps); }
    mbar_init(afull, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();          // barriers of both CTAs initialised, TMEM allocated in both
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer (each CTA loads its own A rows and its half of W) ----------------
      if constexpr (ARES) {
        const uint32_t bar = mapa_u32(smem_u32(afull), 0);
        if (leader) mbar_expect_tx(afull, 2u * static_cast<uint32_t>(num_kb) * Cfg::A_BYTES);
        for (int kb = 0; kb < num_kb; ++kb) tma_load_2d_2sm(ares + kb * Cfg::A_BYTES, &tmA, bar, kb * BK, m_blk * BM);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < tile1; ++tile) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = tiles + stage * Cfg::STAGE_BYTES;
          const uint32_t bar = mapa_u32(smem_u32(&full[stage]), 0);
          if (leader) mbar_expect_tx(&full[stage], 2u * Cfg::STAGE_BYTES);
          if constexpr (!ARES) {
            tma_load_2d_2sm(sa, &tmA, bar, kb * BK, m_blk * BM);
            sa += Cfg::A_BYTES;
          }
          tma_load_2d_2sm(sa, &tmB, bar, kb * BK, tile * BN + static_cast<int>(rank) * (BN / 2));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ---------------- MMA issuer (leader only, whole warp, one elected lane): UMMA 256 x BN x 16 across the pair ----------------
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN);
      if constexpr (ARES) {
        mbar_wait(afull, 0);
        tc_fence_after();
      }
      const uint32_t tiles_lo = sw128_desc_lo(smem_u32(tiles));
      const uint32_t ares_lo = sw128_desc_lo(smem_u32(ares));
      int stage = 0;
      uint32_t phase = 0;
      int t = 0;
      for (int tile = tile0; tile < tile1; ++tile, ++t) {
        const int buf = t & 1;
        const uint32_t use = static_cast<uint32_t>(t >> 1);
        mbar_wait(&tempty[buf], (use & 1) ^ 1);  // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t s_lo = tiles_lo + static_cast<uint32_t>(stage * (Cfg::STAGE_BYTES >> 4));
          const uint32_t a_lo = ARES ? ares_lo + static_cast<uint32_t>(kb * (Cfg::A_BYTES >> 4)) : s_lo;
          const uint32_t b_lo = ARES ? s_lo : s_lo + (Cfg::A_BYTES >> 4);
          if (elect_one_sync()) {
            umma_kblock_2sm(d_tmem, a_lo, b_lo, idesc, static_cast<uint32_t>(kb != 0));
            umma_commit_2sm(&empty[stage]);  // multicast: frees the stage in both CTAs
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one_sync()) umma_commit_2sm(&tfull[buf]);      // multicast: accumulator complete in both CTAs
        __syncwarp();
      }
    }
  } else {
    // ---------------- epilogue (warps 2..9 of BOTH CTAs, each on its own 128 rows) ----------------
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row0 = m_blk * BM + q * 32;
    uint8_t* stg = stg_all + (warp - 2) * 2 * Cfg::STG_BYTES;
    uint32_t store_count = 0;
    TopK<KC> top;
    if constexpr (EPI == EPI_TOPK) top.init();
    int t = 0;
    for (int tile = tile0; tile < tile1; ++tile, ++t) {
      const int buf = t & 1;
      const uint32_t use = static_cast<uint32_t>(t >> 1);
      float* bias_s = bias_smem + buf * BN;
      epi_stage_bias<BN, EPI>(g, tile, bias_s);
      mbar_wait(&tfull[buf], use & 1);
      tc_fence_after();
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * BN);
      epi_process<BN, EPI, KC>(g, &tmC, tile, tbase, bias_s, stg, half, lane, row0, store_count, top);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[buf]), 0));
    }
    if constexpr (EPI == EPI_TOPK) epi_store_candidates<KC>(g, half, row0 + lane, top);
    else if (lane == 0) bulk_wait_group<0>();
  }

  tc_fence_before();
  cluster_sync_all();   // the peer's shared memory and TMEM must outlive the leader's last MMA
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace owc
