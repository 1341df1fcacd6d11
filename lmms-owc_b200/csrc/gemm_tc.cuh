// gemm_tc.cuh - the tcgen05 / TMEM / TMA contraction kernel of the scoring path.
//
//   C[M, N] = A[M, K] . W[N, K]^T        A, W bf16, K-major (row-major with K innermost)
//
// One kernel skeleton serves every dense contraction on the path:
// a CTA owns one 128-row block of A and sweeps a range of BN-wide column tiles of W. Roles:
//   warp 0 (one lane)  TMA producer: per k-block of 64, one 128x64 A box + one BNx64 W box into a
//                      SWIZZLE_128B stage ring, completion on `full[stage]` (expect_tx)
//   warp 1 (one lane)  MMA issuer: 4 x tcgen05.mma (128 x BN x 16) per stage into one of TWO TMEM
//                      accumulator buffers; tcgen05.commit frees the stage and, after the last
//                      k-block, publishes the accumulator on `tfull[buf]`
//   warps 2-9          epilogue: thread r of TMEM lane quadrant (warp & 3) owns output row r, the two
//                      warps of a quadrant take alternate 32-column chunks (tcgen05.ld, software-
//                      pipelined); fused bias / erf-GELU / residual add, or
//                      the per-row top-k candidate list kept in registers across the whole class
//                      sweep (the similarity matrix never leaves the SM)
// The two accumulator buffers let the epilogue of tile t overlap the MMAs of tile t+1.
#pragma once

#include "common.cuh"

namespace owc {

enum : int {
  EPI_BIAS_BF16 = 0,       // out bf16 = acc + bias
  EPI_BIAS_GELU_BF16 = 1,  // out bf16 = gelu_erf(acc + bias)
  EPI_BIAS_RES_F32 = 2,    // out fp32 = acc + bias + residual   (pre-LayerNorm sum)
  EPI_BIAS_F32 = 3,        // out fp32 = acc (+ bias)            (debug / check of the mainloop)
  EPI_TOPK = 4             // per-row top-KC candidates over all columns handled by this CTA
};

struct GemmArgs {
  int M, N, K;             // K % 64 == 0; N % 32 == 0 unless EPI_TOPK
  int tiles_per_cta;       // column tiles swept by one CTA (gridDim.y = ceil(tiles / tiles_per_cta))
  const float* bias;       // [N] or nullptr
  const float* residual;   // [M, ldo] fp32 (EPI_BIAS_RES_F32)
  void* out;               // bf16 or fp32 [M, ldo]
  int ldo;
  float* cand_score;       // EPI_TOPK: [2 * gridDim.y, M, KC] (two chunk parities per column split)
  int* cand_idx;
};

template <int BN>
struct GemmCfg {
  static constexpr int BM = 128;
  static constexpr int BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 192 ? 5 : 6);
  static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
  static constexpr int EPI_THREADS = 256;                 // 8 epilogue warps
  static constexpr int THREADS = 64 + EPI_THREADS;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 2 * BN * 4 /*bias*/;
};

template <int KC>
struct TopK {
  float s[KC];
  int id[KC];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < KC; ++i) { s[i] = -INFINITY; id[i] = -1; }
  }
  // Keeps s[] descending; a value equal to an entry goes after it, so with columns visited in
  // increasing order ties resolve to the lower class index.
  __device__ __forceinline__ void offer(float v, int col) {
    if (v > s[KC - 1]) {
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
  }
};

template <int BN, int EPI, int KC>
__global__ void __launch_bounds__(GemmCfg<BN>::THREADS, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
  using Cfg = GemmCfg<BN>;
  constexpr int BM = Cfg::BM, BK = Cfg::BK, STAGES = Cfg::STAGES;

  extern __shared__ uint8_t smem_raw[];
  // keep the pointer arithmetic on the __shared__ array itself so accesses stay in the shared window
  // (LDS/STS) instead of degrading to generic loads; SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* bias_smem = reinterpret_cast<float*>(tiles + STAGES * Cfg::STAGE_BYTES + 256);   // [2][BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_blk = blockIdx.x;
  const int tiles_total = (g.N + BN - 1) / BN;
  const int tile0 = blockIdx.y * g.tiles_per_cta;
  const int tile1 = min(tile0 + g.tiles_per_cta, tiles_total);
  const int num_kb = g.K / BK;

  if (warp == 1 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], Cfg::EPI_THREADS / 32); }
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

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < tile1; ++tile) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = tiles + stage * Cfg::STAGE_BYTES;
          mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
          tma_load_2d(sa, &tmA, &full[stage], kb * BK, m_blk * BM);
          tma_load_2d(sa + Cfg::A_BYTES, &tmB, &full[stage], kb * BK, tile * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer ----------------
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
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
          tc_fence_after();
          const uint32_t a_addr = smem_u32(tiles + stage * Cfg::STAGE_BYTES);
          const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            umma_f16(d_tmem, make_sw128_kmajor_desc(a_addr + k * 32), make_sw128_kmajor_desc(b_addr + k * 32),
                     idesc, static_cast<uint32_t>((kb | k) != 0));
          }
          umma_commit(&empty[stage]);  // stage reusable once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[buf]);  // accumulator complete
      }
    }
  } else {
    // ---------------- epilogue (warps 2..9): two warps per TMEM lane quadrant ----------------
    // Warp w may only touch TMEM lanes [32*(w&3), +32); the two warps of a quadrant take alternate
    // 32-column chunks, so one warp's tcgen05.ld latency hides behind the other's math and stores.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;   // 0 | 1
    const int row = m_blk * BM + q * 32 + lane;
    const bool row_ok = row < g.M;
    TopK<KC> top;
    if constexpr (EPI == EPI_TOPK) top.init();
    const int et = threadIdx.x - 64;    // 0..255 within the epilogue group
    constexpr int NC = BN / 32;         // chunks per tile
    constexpr int NCW = NC / 2;         // chunks per warp
    int t = 0;
    for (int tile = tile0; tile < tile1; ++tile, ++t) {
      const int buf = t & 1;
      const uint32_t use = static_cast<uint32_t>(t >> 1);
      float* bias_s = bias_smem + buf * BN;
      if constexpr (EPI != EPI_TOPK) {
        // stage this tile's bias slice in shared memory while the MMAs of the tile are still running
        for (int i = et; i < BN; i += Cfg::EPI_THREADS) {
          const int n = tile * BN + i;
          bias_s[i] = (g.bias != nullptr && n < g.N) ? __ldg(g.bias + n) : 0.f;
        }
        named_bar_sync(1, Cfg::EPI_THREADS);
      }
      mbar_wait(&tfull[buf], use & 1);
      tc_fence_after();
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * BN);
      if constexpr (EPI == EPI_TOPK) {
        // code kept small on purpose (32 inlined offers): the sweep runs hundreds of tiles
#pragma unroll 1
        for (int cw = 0; cw < NCW; ++cw) {
          const int c = 2 * cw + half;
          const int n0 = tile * BN + c * 32;
          if (n0 >= g.N) break;  // warp-uniform
          uint32_t v[32];
          tmem_ld_32x32(tbase + static_cast<uint32_t>(c * 32), v);
          tmem_ld_wait_dep(v);
          if (n0 + 32 <= g.N) {
#pragma unroll
            for (int j = 0; j < 32; ++j) top.offer(__uint_as_float(v[j]), n0 + j);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < g.N) top.offer(__uint_as_float(v[j]), n0 + j);
          }
        }
      } else {
        uint32_t v[2][32];
        tmem_ld_32x32(tbase + static_cast<uint32_t>(half * 32), v[0]);
#pragma unroll
        for (int cw = 0; cw < NCW; ++cw) {
          const int c = 2 * cw + half;
          const int n0 = tile * BN + c * 32;
          tmem_ld_wait_dep(v[cw & 1]);
          if (cw + 1 < NCW) tmem_ld_32x32(tbase + static_cast<uint32_t>((c + 2) * 32), v[(cw + 1) & 1]);  // overlaps the math below
          if (n0 >= g.N || !row_ok) continue;
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
          const size_t off = static_cast<size_t>(row) * g.ldo + n0;
          if constexpr (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16) {
            if constexpr (EPI == EPI_BIAS_GELU_BF16) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = gelu_fast(f[j]);
            }
            uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(g.out) + off);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 w;
              w.x = pack_bf16x2(f[8 * j + 0], f[8 * j + 1]);
              w.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
              w.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
              w.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
              o[j] = w;
            }
          } else {
            if constexpr (EPI == EPI_BIAS_RES_F32) {
              const float4* r4 = reinterpret_cast<const float4*>(g.residual + off);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 r = __ldg(r4 + j);
                f[4 * j + 0] += r.x; f[4 * j + 1] += r.y; f[4 * j + 2] += r.z; f[4 * j + 3] += r.w;
              }
            }
            float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(g.out) + off);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);
    }
    if constexpr (EPI == EPI_TOPK) {
      if (row_ok) {
        // each (column split, chunk parity) pair contributes its own candidate list
        const size_t base = (static_cast<size_t>(blockIdx.y * 2 + half) * g.M + row) * KC;
#pragma unroll
        for (int i = 0; i < KC; ++i) { g.cand_score[base + i] = top.s[i]; g.cand_idx[base + i] = top.id[i]; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace owc
