// enc_fused.cuh - the block-resident encoder: ALL transformer layers of a MiniLM-width BERT encoder
// (H = 384, 12 heads of 32, FFN width a multiple of 128) in ONE persistent kernel.
//
// Why: the per-GEMM kernels of gemm_tc.cuh stream W from L2 once per 128-row block (64 B/clk/SM at
// BN = 192 against the ~43 B/clk/SM the L2 delivers chip-wide) and bounce every activation through
// L2/HBM between launches. Sequences on this path are short (3-64 tokens), so a 128-token row block
// made of WHOLE sequences is a closed system: attention never leaves it. One CTA therefore carries
// its block through every layer with the activations pinned on chip:
//   TMEM  cols [0, 384)    the fp32 residual stream x of the block (thread r <-> token row r). The
//                          O-projection and FFN2 MMAs accumulate straight onto it (the residual add
//                          is free); LayerNorm rewrites it in place.
//         cols [384, 512)  T1: the transient accumulator - one head's [q|k|v] (96 cols) or one
//                          128-wide slice of the FFN intermediate.
//   SMEM  XB  96 KB        bf16 copy of x in UMMA K-major SWIZZLE_128B layout (A operand of QKV / FFN1)
//         QKV 2 x 24 KB    one head's q|k|v rows (bf16, 192 B rows, XOR-swizzled 16-byte chunks) for the mma.sync
//                          attention - one buffer per SIMT group (the two groups alternate heads)
//         AO  16 KB        attention output of a head PAIR as A operand (K = 64) of the O-projection
//         FC  2 x 32 KB    GELU(FFN1 slice) as A operand (K = 128) of FFN2 (aliases QKV/AO)
//         RING 6 x 8 KB    weight tiles streamed by TMA: 48/64 rows x 64 K per slot (T1 tiles) or 2 x 96 rows x 64 K
//                          over three slots (the N = 384 accumulate GEMMs run as two N = 192 UMMA groups)
// CTAs work in PAIRS (cluster of 2, tcgen05 cta_group::2, UMMA M = 256): each CTA stages only half of
// every weight tile, so the pair reads each weight byte once per 256 token rows (~28 B/clk/SM).
// The only global traffic is the residual stream in and out (1.5 KB per token each way) + weights.
//
// Roles per CTA: warps 0..15 = SIMT warps (TMEM drains, bias/GELU/LayerNorm, attention), warp 16 lane 0 = TMA
// producer (its half of every tile), warp 17 of the LEADER = MMA issuer for the pair (uniform loop, elected lane).
// In the attention block the SIMT warps form two groups of 8 that alternate heads: while group g runs the attention
// of head h out of its QKV buffer, the other group drains and converts head h+1.
// MMA issue order keeps the tensor pipe busy while the SIMT warps work on the previous tiles:
//   attention block:  Q(0) Q(1) Q(2) Q(3) O(0) Q(4) Q(5) O(1) ...   (O(p) = heads 2p, 2p+1, K = 64)
//   feed-forward:     F1(0) F1(1) F2(0) F1(2) F2(1) ...
// Reference semantics: transformers modeling_bert.py:114-207 (self-attention + output LayerNorm),
// :287-356 (intermediate GELU + output LayerNorm), post-LN residual blocks.
#pragma once

#include "attn_mma.cuh"
#include "common.cuh"

namespace owc {

struct FusedArgs {
  float* x;                  // [rows(+128 slack), 384] fp32 residual stream, chunk-relative rows (in: embeddings, out: last layer)
  const int32_t* cu;         // device cu_seqlens of the whole call
  const int32_t* blk_seq;    // [n_blk + 1] absolute index of the first sequence of each row block
  int n_blk;
  int tok0;                  // cu[first sequence of the chunk]
  int layers;
  int n_chunks;              // FFN width / 128
  long long layer_stride;    // floats between the same parameter of consecutive layers in the blob
  const float *bqkv, *bo, *ln1g, *ln1b, *b1, *b2, *ln2g, *ln2b;   // layer 0
  float ln_eps;
  float scale_log2;          // head_dim^-0.5 * log2(e)
  int no_groups;             // A/B: 1 = one sequence per attention tile (OWC_ATTN_GRP=0)
  int ablate;                // debug: bit0 skip attention math, bit1 skip GELU, bit2 skip LayerNorm math, bit3 skip T1 drains (wrong results)
  long long* dbg;            // optional timeline stamps of cluster 0 / leader (clock64), nullptr in production
};

namespace fz {
constexpr int H = 384, HEADS = 12, D = 32;
constexpr int NW = 4;                       // SIMT warps per TMEM lane quadrant
constexpr int NSIMT = 4 * NW;               // SIMT warps per CTA
constexpr int THREADS = 64 + NSIMT * 32;
// The warp scheduler favours the highest warp id of an SMSP: the two single-lane issue roles take the top ids so
// that 16 busy SIMT warps cannot starve the TMA / MMA issue loops.
constexpr int W_TMA = NSIMT, W_MMA = NSIMT + 1;
constexpr int SLOT = 8192;                  // ring slot: up to 64 rows x 128 B
constexpr int SLOTS = 6;                    // T1 tiles take one slot per k-block, x-accumulate k-blocks three (2 x 96 rows)
constexpr int QROW = 192;                   // bytes of one QKV row: q | k | v of one head, 16-byte chunks XOR-swizzled inside each 64 B third
constexpr int OFF_XB = 0;                   // 6 k-block tiles of 16 KB
constexpr int OFF_U = 98304;                // union: attention phase AO[2] + QKV | FFN phase FC[2]
constexpr int OFF_AO = OFF_U;               // one head-pair tile of 16384
constexpr int OFF_QKV = OFF_U + 16384;      // 2 x (128 x 192) = 2 x 24576: one buffer per SIMT group
constexpr int OFF_FC = OFF_U;               // 2 x 32768
constexpr int OFF_RING = OFF_U + 65536;     // 163840
constexpr int OFF_TAIL = OFF_RING + SLOTS * SLOT;   // 221184
constexpr int OFF_GRP = 1024 + 4096 + 3 * 384 * 4 + 3 * 384 * 4;      // attention groups of the block: uint8 [<= 130] + count
constexpr int TAIL_BYTES = OFF_GRP + 256;   // barriers / tables, LayerNorm statistics, LayerNorm bias|gamma|beta, QKV bias, groups
constexpr int SMEM_BYTES = OFF_TAIL + TAIL_BYTES + 1024;   // + alignment slack
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
constexpr uint32_t TM_X = 0, TM_T1 = 384;   // TMEM column offsets
}  // namespace fz

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[N]) {   // N consecutive fp32 columns of this thread's lane
  static_assert(N % 8 == 0, "column count");
#pragma unroll
  for (int i = 0; i + 16 <= N; i += 16) tmem_ld_32x32b_x16(taddr + static_cast<uint32_t>(i), v + i);
  if constexpr (N % 16 != 0) tmem_ld_32x32b_x8(taddr + static_cast<uint32_t>(N - 8), v + (N - 8));
}
// Wait for outstanding tcgen05.ld, then pin every destination register behind the wait (volatile asms keep
// their order), so no use of v[] can be scheduled above it.
template <int N>
__device__ __forceinline__ void tmem_ld_wait_regs(uint32_t (&v)[N]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < N; ++i) asm volatile("" : "+r"(v[i]));
}

// 32 fp32 values of row `r`, columns [32k, 32k+32) -> bf16 into the XB-style A-operand layout: k-block tile
// (k >> 1) of 16 KB, row r at r*128 B, 16-byte chunk j stored at j ^ (r & 7)  (SWIZZLE_128B).
__device__ __forceinline__ void store_a_chunk32(uint8_t* tiles, int r, int k, const float (&f)[32]) {
  uint8_t* rowp = tiles + (k >> 1) * 16384 + r * 128;
  const int sw = r & 7;
  const int c0 = (k & 1) * 4;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 w;
    w.x = pack_bf16x2(f[8 * j + 0], f[8 * j + 1]);
    w.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
    w.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
    w.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
    *reinterpret_cast<uint4*>(rowp + (((c0 + j) ^ sw) << 4)) = w;
  }
}

// Address of 16-byte chunk `c` (0-3 q, 4-7 k, 8-11 v) of QKV row `row`: rows are 192 B apart and the chunk index is
// XORed inside its 64-byte third with (row >> 1) & 3, which keeps every access pattern of the attention body
// (32-bit fragment loads of 8 rows x 4 lanes, ldmatrix rows, 16-byte row stores) free of bank conflicts.
__device__ __forceinline__ const uint8_t* qkv_chunk(const uint8_t* qkv, int row, int c) {
  return qkv + row * fz::QROW + (((c & 12) | ((c & 3) ^ ((row >> 1) & 3))) << 4);
}

// One (sequence, head) of the block: q|k|v rows in `qkv`, output into the AO tile (A-operand layout, 16-byte chunks
// cb..cb+3 of each 128-byte row). Same arithmetic as attn_pair (attn_mma.cuh). `pre_store` runs before the first
// write to the AO tile (it waits until the O-projection that last read the tile has completed).
template <int TP, typename PreStore>
__device__ __forceinline__ void attn_head_smem(const uint8_t* qkv, uint8_t* ao_tile, int beg, int T, int lane, int cb,
                                               float scale_log2, PreStore pre_store) {
  constexpr int NT = TP / 8, MT = TP / 16, DN = 4;
  const int g = lane >> 2, t4 = lane & 3;
  auto row_of = [&](int i) { return min(beg + i, 127); };   // rows past the block are never used unmasked
#pragma unroll 1
  for (int mt = 0; mt < MT; ++mt) {
    if (mt * 16 >= T) break;
    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) { s[nt][0] = 0.f; s[nt][1] = 0.f; s[nt][2] = 0.f; s[nt][3] = 0.f; }
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    const int qr0 = row_of(r0), qr1 = row_of(r1);
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t af[4];
      af[0] = *reinterpret_cast<const uint32_t*>(qkv_chunk(qkv, qr0, 2 * ks) + t4 * 4);
      af[1] = *reinterpret_cast<const uint32_t*>(qkv_chunk(qkv, qr1, 2 * ks) + t4 * 4);
      af[2] = *reinterpret_cast<const uint32_t*>(qkv_chunk(qkv, qr0, 2 * ks + 1) + t4 * 4);
      af[3] = *reinterpret_cast<const uint32_t*>(qkv_chunk(qkv, qr1, 2 * ks + 1) + t4 * 4);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        if (nt * 8 < T) {   // warp-uniform
          const int kr = row_of(nt * 8 + g);
          uint32_t bf[2];
          bf[0] = *reinterpret_cast<const uint32_t*>(qkv_chunk(qkv, kr, 4 + 2 * ks) + t4 * 4);
          bf[1] = *reinterpret_cast<const uint32_t*>(qkv_chunk(qkv, kr, 5 + 2 * ks) + t4 * 4);
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
          float v = s[nt][e] * scale_log2;
          v = (j < T) ? v : -INFINITY;
          s[nt][e] = v;
          if (e < 2) m0 = fmaxf(m0, v); else m1 = fmaxf(m1, v);
        }
      }
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
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
          ldmatrix_x2_trans(bf, qkv_chunk(qkv, vr, 8 + dn));
          mma_bf16_16816(o[dn], af, bf);
        }
      }
    }
    float inv0, inv1;   // sums are >= 1 (the row maximum contributes 2^0): the approximate reciprocal is exact enough
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv0) : "f"(sum0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv1) : "f"(sum1));
    pre_store();
    const int gr0 = beg + r0, gr1 = beg + r1;
    if (r0 < T) {
      uint8_t* p0 = ao_tile + gr0 * 128 + t4 * 4;
      const int sw0 = gr0 & 7;
#pragma unroll
      for (int dn = 0; dn < DN; ++dn)
        *reinterpret_cast<uint32_t*>(p0 + (((cb + dn) ^ sw0) << 4)) = pack_bf16x2(o[dn][0] * inv0, o[dn][1] * inv0);
    }
    if (r1 < T) {
      uint8_t* p1 = ao_tile + gr1 * 128 + t4 * 4;
      const int sw1 = gr1 & 7;
#pragma unroll
      for (int dn = 0; dn < DN; ++dn)
        *reinterpret_cast<uint32_t*>(p1 + (((cb + dn) ^ sw1) << 4)) = pack_bf16x2(o[dn][2] * inv1, o[dn][3] * inv1);
    }
  }
}

// Several short sequences of the block in ONE 16 x 16 tile (see attn_group16 in attn_mma.cuh): a group is a run of
// consecutive sequences with at most 16 tokens together, rows [beg, beg + T) of the block; seq_row[s_first .. s_last] are
// the sequences' first rows. Block-diagonal mask: tile row i only sees the keys [lo_i, hi_i) of its own sequence. The
// instruction count is that of ONE sequence in attn_head_smem<16>, so label-length inputs (2-3 sequences per tile) spend
// less than half the SIMT time per head.
template <typename PreStore>
__device__ __forceinline__ void attn_group_smem(const uint8_t* qkv, uint8_t* ao_tile, int beg, int T, const int* seq_row,
                                                int s_first, int s_last, int lane, int cb, float scale_log2,
                                                PreStore pre_store) {
  constexpr int NT = 2, DN = 4;
  const int g = lane >> 2, t4 = lane & 3;
  auto row_of = [&](int i) { return min(beg + i, 127); };
  int lo = 0, hi = 0;
  for (int sq = s_first; sq < s_last; ++sq) {
    const int b = seq_row[sq] - beg, e = seq_row[sq + 1] - beg;
    if (lane >= b && lane < e) { lo = b; hi = e; }
  }
  const int r0 = g, r1 = g + 8;
  const int lo0 = __shfl_sync(0xffffffffu, lo, r0), hi0 = __shfl_sync(0xffffffffu, hi, r0);
  const int lo1 = __shfl_sync(0xffffffffu, lo, r1), hi1 = __shfl_sync(0xffffffffu, hi, r1);
  float s[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) { s[nt][0] = 0.f; s[nt][1] = 0.f; s[nt][2] = 0.f; s[nt][3] = 0.f; }
  const int qr0 = row_of(r0), qr1 = row_of(r1);
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    uint32_t af[4];
    af[0] = *reinterpret_cast<const uint32_t*>(qkv_chunk(qkv, qr0, 2 * ks) + t4 * 4);
    af[1] = *reinterpret_cast<const uint32_t*>(qkv_chunk(qkv, qr1, 2 * ks) + t4 * 4);
    af[2] = *reinterpret_cast<const uint32_t*>(qkv_chunk(qkv, qr0, 2 * ks + 1) + t4 * 4);
    af[3] = *reinterpret_cast<const uint32_t*>(qkv_chunk(qkv, qr1, 2 * ks + 1) + t4 * 4);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (nt * 8 < T) {   // warp-uniform
        const int kr = row_of(nt * 8 + g);
        uint32_t bf[2];
        bf[0] = *reinterpret_cast<const uint32_t*>(qkv_chunk(qkv, kr, 4 + 2 * ks) + t4 * 4);
        bf[1] = *reinterpret_cast<const uint32_t*>(qkv_chunk(qkv, kr, 5 + 2 * ks) + t4 * 4);
        mma_bf16_16816(s[nt], af, bf);
      }
    }
  }
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = nt * 8 + t4 * 2 + (e & 1);
      const bool ok = (e < 2) ? (j >= lo0 && j < hi0) : (j >= lo1 && j < hi1);
      const float v = ok ? s[nt][e] * scale_log2 : -INFINITY;
      s[nt][e] = v;
      if (e < 2) m0 = fmaxf(m0, v); else m1 = fmaxf(m1, v);
    }
  }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  m0 = (hi0 > lo0) ? m0 : 0.f;   // tile rows past the group have no key: keep (-inf) - (-inf) out of the exponent
  m1 = (hi1 > lo1) ? m1 : 0.f;
  uint32_t pk[NT][2];
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
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
  }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
  float o[DN][4];
#pragma unroll
  for (int dn = 0; dn < DN; ++dn) { o[dn][0] = 0.f; o[dn][1] = 0.f; o[dn][2] = 0.f; o[dn][3] = 0.f; }
  {
    const uint32_t af[4] = {pk[0][0], pk[0][1], pk[1][0], pk[1][1]};
    const int vr = row_of(lane & 15);
#pragma unroll
    for (int dn = 0; dn < DN; ++dn) {
      uint32_t bf[2];
      ldmatrix_x2_trans(bf, qkv_chunk(qkv, vr, 8 + dn));
      mma_bf16_16816(o[dn], af, bf);
    }
  }
  float inv0, inv1;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv0) : "f"(fmaxf(sum0, 1.0f)));   // real rows sum to >= 1; rows past the group are not stored
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv1) : "f"(fmaxf(sum1, 1.0f)));
  pre_store();
  const int gr0 = beg + r0, gr1 = beg + r1;
  if (r0 < T) {
    uint8_t* p0 = ao_tile + gr0 * 128 + t4 * 4;
    const int sw0 = gr0 & 7;
#pragma unroll
    for (int dn = 0; dn < DN; ++dn)
      *reinterpret_cast<uint32_t*>(p0 + (((cb + dn) ^ sw0) << 4)) = pack_bf16x2(o[dn][0] * inv0, o[dn][1] * inv0);
  }
  if (r1 < T) {
    uint8_t* p1 = ao_tile + gr1 * 128 + t4 * 4;
    const int sw1 = gr1 & 7;
#pragma unroll
    for (int dn = 0; dn < DN; ++dn)
      *reinterpret_cast<uint32_t*>(p1 + (((cb + dn) ^ sw1) << 4)) = pack_bf16x2(o[dn][2] * inv1, o[dn][3] * inv1);
  }
}

// Debug-only phase accounting: time since the previous mark is charged to bucket i.
template <bool ON>
struct PhaseAcc {
  long long t = 0, a[16] = {};
  long long* ev = nullptr;   // event trace: (id, clock) pairs while `tracing`
  int n_ev = 0;
  bool tracing = false;
  __device__ __forceinline__ void event(int id) {
    if constexpr (ON) {
      if (tracing && ev != nullptr && n_ev < 220) { ev[2 * n_ev] = id; ev[2 * n_ev + 1] = clock64(); ++n_ev; }
    }
  }
  __device__ __forceinline__ void start() { if constexpr (ON) t = clock64(); }
  __device__ __forceinline__ void mark(int i) {
    if constexpr (ON) { const long long n = clock64(); a[i] += n - t; t = n; }
  }
};

// MODE 0: production. MODE 1: phase accounting + event traces into a.dbg (tools/fused_check.py). MODE 2: role ablation
// by a.ablate (tools/fused_ablate.py; wrong results by construction). The debug hooks compile away in MODE 0.
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(fz::THREADS, 1)
k_encoder_fused(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmWO,
                const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2, const FusedArgs a) {
  constexpr bool DBG = MODE == 1;
  const int ablate = (MODE == 2) ? a.ablate : 0;
  using namespace fz;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* XB = base + OFF_XB;
  uint8_t* AO = base + OFF_AO;
  __nv_bfloat16* QKV = reinterpret_cast<__nv_bfloat16*>(base + OFF_QKV);
  uint8_t* FC = base + OFF_FC;
  uint8_t* RING = base + OFF_RING;
  uint8_t* tail = base + OFF_TAIL;
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);   // [SLOTS]  leader's: both CTAs' TMA bytes of the slot landed
  uint64_t* empty = full + SLOTS;                       // [SLOTS]  per CTA (multicast commit): slot consumed
  uint64_t* t1_full = empty + SLOTS;                    // [2] per CTA (multicast commit): T1 tile n complete -> t1_full[n & 1]
  uint64_t* t1_free_a = t1_full + 2;                    // leader's: the draining GROUP (8 warps per CTA) has read T1 (attention phase)
  uint64_t* t1_free_f = t1_free_a + 1;                  // leader's: all 2*NSIMT warps have read T1 (feed-forward phase)
  uint64_t* ao_full = t1_free_f + 1;                    // leader's: both heads of a pair written to AO by all warps of the pair
  uint64_t* ao_free = ao_full + 1;                      // per CTA (commit): the O-projection has read AO
  uint64_t* f_full = ao_free + 1;                       // [2] leader's
  uint64_t* f_free = f_full + 2;                        // [2] per CTA (commit)
  uint64_t* accx_full = f_free + 2;                     // per CTA (commit): residual accumulator complete
  uint64_t* xb_ready = accx_full + 1;                   // leader's: x (TMEM) and XB rewritten by all warps of the pair
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 256);
  int* seq_row = reinterpret_cast<int*>(tail + 320);    // [<= 130] first row of each sequence of this CTA's block
  long long* issue_clk = reinterpret_cast<long long*>(tail + 896);   // [SLOTS] debug: clock of the TMA issue per slot
  float* stat_s = reinterpret_cast<float*>(tail + 1024);   // [NSIMT][32][2]
  float* lnp = reinterpret_cast<float*>(tail + 1024 + 4096);   // [3][384]: bias | gamma | beta of the next LayerNorm
  float* bqkv_s = lnp + 3 * 384;                               // [3][384]: q | k | v projection bias of the current layer
  uint8_t* grp_first = tail + OFF_GRP;                         // [n_grp + 1] first sequence (block-relative) of each attention group
  int* n_grp_s = reinterpret_cast<int*>(tail + OFF_GRP + 192);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_pairs = (a.n_blk + 1) >> 1;
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;

  if (warp == W_MMA && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmWO);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    for (int s = 0; s < SLOTS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&t1_full[0], 1);
    mbar_init(&t1_full[1], 1);
    mbar_init(t1_free_a, NSIMT);            // 2 CTAs x NSIMT/2 warps of the draining group
    mbar_init(t1_free_f, 2 * NSIMT);
    mbar_init(ao_full, 2 * NSIMT);
    mbar_init(ao_free, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&f_full[b], 2 * NSIMT);
      mbar_init(&f_free[b], 1);
    }
    mbar_init(accx_full, 1);
    mbar_init(xb_ready, 2 * NSIMT);
    fence_barrier_init();
  }
  if (warp == W_TMA) {
    tmem_alloc_2sm(tmem_slot, 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == W_TMA) {
    if (lane == 0) {
      // ================= TMA producer: this CTA's half of every weight tile, in MMA order =================
      int stage = 0;
      uint32_t ebits = 0;   // per slot: parity of its next `empty` wait (slots are not all used in every pass)
      auto acquire = [&](int s) {
        mbar_wait(&empty[s], ((ebits >> s) & 1u) ^ 1u);
        ebits ^= 1u << s;
      };
      // one k-block of a T1 tile: this CTA's half (48 or 64 rows x 64 K) into one slot
      auto load1 = [&](const CUtensorMap* tm, int c0, int c1, uint32_t bytes) {
        acquire(stage);
        if (leader) mbar_expect_tx(&full[stage], 2u * bytes);
        if constexpr (DBG) issue_clk[stage] = clock64();
        tma_load_2d_2sm(RING + stage * SLOT, tm, mapa_u32(smem_u32(&full[stage]), 0), c0, c1);
        if (++stage == SLOTS) stage = 0;
      };
      // one k-block of an x-accumulate GEMM (N = 384 as two UMMA groups of 192): this CTA's 96 rows of each group
      // into three consecutive slots (stage is 0 or 3 here: every tile takes a multiple of three slots)
      auto load3 = [&](const CUtensorMap* tm, int c0, int c1) {
        acquire(stage); acquire(stage + 1); acquire(stage + 2);
        const uint32_t bar = mapa_u32(smem_u32(&full[stage]), 0);
        if (leader) mbar_expect_tx(&full[stage], 2u * 2u * 96u * 128u);
        if constexpr (DBG) issue_clk[stage] = clock64();
        tma_load_2d_2sm(RING + stage * SLOT, tm, bar, c0, c1);
        tma_load_2d_2sm(RING + stage * SLOT + 96 * 128, tm, bar, c0, c1 + 192);
        stage += 3;
        if (stage == SLOTS) stage = 0;
      };
      const int r = static_cast<int>(rank);
      for (int bp = cluster_id; bp < n_pairs; bp += n_clusters) {
        for (int l = 0; l < a.layers; ++l) {
          auto emitQ = [&](int h) {
            for (int kb = 0; kb < 6; ++kb) load1(&tmQKV, kb * 64, l * 1152 + h * 96 + r * 48, 48 * 128);
          };
          auto emitO = [&](int hp) { load3(&tmWO, hp * 64, l * 384 + r * 96); };
          auto emitF1 = [&](int c) {
            for (int kb = 0; kb < 6; ++kb) load1(&tmW1, kb * 64, l * (a.n_chunks * 128) + c * 128 + r * 64, 64 * 128);
          };
          auto emitF2 = [&](int c) {
            for (int kb = 0; kb < 2; ++kb) load3(&tmW2, c * 128 + kb * 64, l * 384 + r * 96);
          };
          emitQ(0);
          for (int h = 0; h < HEADS; ++h) {
            if (h + 1 < HEADS) emitQ(h + 1);
            if (h >= 2 && (h & 1) == 0) emitO((h - 2) >> 1);
          }
          emitO(HEADS / 2 - 1);
          emitF1(0);
          for (int c = 0; c < a.n_chunks; ++c) {
            if (c + 1 < a.n_chunks) emitF1(c + 1);
            emitF2(c);
          }
        }
      }
    }
  } else if (warp == W_MMA) {
    if (leader) {
      // ================= MMA issuer (leader only; the WHOLE warp runs the loop, one elected lane issues) =================
      constexpr uint32_t idesc_q = make_idesc_bf16(256, 96);
      constexpr uint32_t idesc_n128 = make_idesc_bf16(256, 128);
      constexpr uint32_t idesc_n192 = make_idesc_bf16(256, 192);
      const uint32_t xb_lo = sw128_desc_lo(smem_u32(XB));      // descriptor low words; +1024 per 16 KB tile
      const uint32_t ao_lo = sw128_desc_lo(smem_u32(AO));
      const uint32_t fc_lo = sw128_desc_lo(smem_u32(FC));
      const uint32_t ring_lo = sw128_desc_lo(smem_u32(RING));
      const uint32_t d_x = tmem_base + TM_X;
      const uint32_t d_t1 = tmem_base + TM_T1;
      int stage = 0;
      uint32_t fbits = 0;   // per slot: parity of its next `full` wait
      uint32_t n_t1 = 0, n_xb = 0;
      uint32_t n_ao = 0, f_use[2] = {0, 0};
      PhaseAcc<DBG> pa;
      pa.start();
      if constexpr (DBG) { if (a.dbg != nullptr && blockIdx.x == 0 && lane == 0) pa.ev = a.dbg + 512; }
      long long lat_sum = 0, lat_n = 0, lat_max = 0, lat_all = 0, n_all = 0;
      auto slot_wait = [&]() -> uint32_t {
        pa.mark(3);
        bool waited = false;
        const uint32_t par = (fbits >> stage) & 1u;
        if constexpr (DBG) waited = !mbar_test_wait(&full[stage], par);
        mbar_wait(&full[stage], par);
        fbits ^= 1u << stage;
        if constexpr (DBG) {   // latency issue -> landed, exact only when this thread had to wait for the slot
          const long long lat = clock64() - *reinterpret_cast<volatile long long*>(&issue_clk[stage]);
          lat_all += lat; ++n_all;
          if (waited) { lat_sum += lat; ++lat_n; lat_max = lat > lat_max ? lat : lat_max; }
        }
        pa.mark(2);
        tc_fence_after();
        return ring_lo + stage * (SLOT >> 4);
      };
      auto slot_release = [&](int n) {   // n consecutive slots become free once the MMAs issued so far have read them
        if (elect_one_sync()) {
          for (int i = 0; i < n; ++i) umma_commit_2sm(&empty[stage + i]);
        }
        __syncwarp();
        stage += n;
        if (stage == SLOTS) stage = 0;
      };
      // T1 = XB[128 x 384] . Wtile^T (N = 96 or 128), overwriting T1. Tile n completes on t1_full[n & 1]; before it is
      // issued the previous tile must have been drained: by one SIMT group (attention phase) or by all warps (FFN)
      uint32_t n_free_a = 0, n_free_f = 0;
      int pending = 0;   // 0 none, 1 attention tile, 2 feed-forward tile
      auto issue_T1 = [&](uint32_t idesc, int kind) {
        pa.mark(3);
        if (pending == 1) { mbar_wait(t1_free_a, n_free_a & 1); ++n_free_a; }
        else if (pending == 2) { mbar_wait(t1_free_f, n_free_f & 1); ++n_free_f; }
        pending = kind;
        pa.mark(1);
        pa.event(100);   // t1_free observed
        tc_fence_after();
        for (int kb = 0; kb < 6; ++kb) {
          const uint32_t b_lo = slot_wait();
          if (elect_one_sync()) umma_kblock_2sm(d_t1, xb_lo + kb * 1024, b_lo, idesc, static_cast<uint32_t>(kb != 0));
          slot_release(1);
        }
        if (elect_one_sync()) umma_commit_2sm(&t1_full[n_t1 & 1]);
        __syncwarp();
        ++n_t1;
        pa.event(101);   // T1 tile issued
      };
      // x[:, 0:384] += Atile[128 x 64] . W[384 x 64]^T as two N = 192 column groups (B rows 0-95 | 96-191 of the 3-slot group)
      auto issue_X = [&](uint32_t a_lo) {
        const uint32_t b_lo = slot_wait();
        if (elect_one_sync()) {
          umma_kblock_2sm(d_x, a_lo, b_lo, idesc_n192, 1u);
          umma_kblock_2sm(d_x + 192, a_lo, b_lo + ((96 * 128) >> 4), idesc_n192, 1u);
        }
        slot_release(3);
        pa.event(103);   // x-accumulate k-block issued
      };
      for (int bp = cluster_id; bp < n_pairs; bp += n_clusters) {
        for (int l = 0; l < a.layers; ++l) {
          if constexpr (DBG) pa.tracing = (bp == cluster_id && l == 1);
          // ---- attention block ----
          pa.mark(3);
          mbar_wait(xb_ready, n_xb & 1); ++n_xb;
          pa.mark(0);
          tc_fence_after();
          // O-projection of head pair p (heads 2p by SIMT group 0, 2p+1 by group 1) goes out behind Q(2p+3): by then
          // group 1 is finishing head 2p+1 and the issuer never sits on ao_full while a Q tile could be issued
          auto issue_O = [&]() {
            pa.mark(3);
            mbar_wait(ao_full, n_ao & 1); ++n_ao;
            pa.mark(4);
            pa.event(102);   // ao_full observed
            tc_fence_after();
            issue_X(ao_lo);
            if (elect_one_sync()) umma_commit_2sm(ao_free);
            __syncwarp();
          };
          issue_T1(idesc_q, 1);
          for (int h = 0; h < HEADS; ++h) {
            if (h + 1 < HEADS) issue_T1(idesc_q, 1);
            if (h >= 2 && (h & 1) == 0) issue_O();
          }
          issue_O();
          if (elect_one_sync()) umma_commit_2sm(accx_full);
          __syncwarp();
          // ---- feed-forward block ----
          pa.mark(3);
          mbar_wait(xb_ready, n_xb & 1); ++n_xb;
          pa.mark(0);
          tc_fence_after();
          issue_T1(idesc_n128, 2);
          for (int c = 0; c < a.n_chunks; ++c) {
            if (c + 1 < a.n_chunks) issue_T1(idesc_n128, 2);
            const int b = c & 1;
            pa.mark(3);
            mbar_wait(&f_full[b], f_use[b] & 1); ++f_use[b];
            pa.mark(5);
            pa.event(104);   // f_full observed
            tc_fence_after();
            issue_X(fc_lo + b * 2048);
            issue_X(fc_lo + b * 2048 + 1024);
            if (elect_one_sync()) umma_commit_2sm(&f_free[b]);
            __syncwarp();
          }
          if (elect_one_sync()) umma_commit_2sm(accx_full);
          __syncwarp();
        }
      }
      if constexpr (DBG) {
        pa.mark(3);
        if (a.dbg != nullptr && blockIdx.x == 0 && lane == 0)
          for (int i = 0; i < 16; ++i) a.dbg[48 + i] = pa.a[i];
        if (a.dbg != nullptr && blockIdx.x == 0 && lane == 0) {
          a.dbg[54] = lat_sum; a.dbg[55] = lat_n; a.dbg[56] = lat_max; a.dbg[57] = lat_all; a.dbg[58] = n_all;
        }
      }
    }
  } else {
    // ================= SIMT warps: TMEM drains, bias / GELU / LayerNorm, attention =================
    const int warp_e = warp;
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int part = warp_e >> 2;           // 0..NW-1: which share of the columns
    const int r = q * 32 + lane;            // token row of this thread inside the block
    const uint32_t tq = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    // barriers that live in the leader: its own warps arrive locally (cta scope), the peer's through the cluster
    auto arrive_leader = [&](uint64_t* bar) {
      if (leader) mbar_arrive(bar);
      else mbar_arrive_cluster(mapa_u32(smem_u32(bar), 0));
    };
    uint32_t n_t1 = 0, n_accx = 0;
    uint32_t n_pair = 0, f_use[2] = {0, 0};   // head pairs this warp has contributed to; uses of the two FC buffers
    const bool stamp = DBG && a.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    int n_stamp = 0;
    PhaseAcc<DBG> pa;
    pa.start();
    if constexpr (DBG) { if (stamp) pa.ev = a.dbg + 64; }
#define FZ_STAMP() do { if (stamp && n_stamp < 30) { a.dbg[n_stamp++] = clock64(); long long gt_; \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_)); a.dbg[n_stamp == 1 ? 30 : 31] = gt_; } } while (0)

    // x (TMEM) and XB from 32 fp32 values of this thread's row, chunk k
    auto put_x_chunk = [&](int k, const float (&f)[32]) {
      uint32_t u[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) u[j] = __float_as_uint(f[j]);
      tmem_st_32x32(tq + TM_X + static_cast<uint32_t>(k * 32), u);
      store_a_chunk32(XB, r, k, f);
    };
    auto publish_x = [&]() {   // x / XB of this warp's rows are final: tell the MMA issuer
      tmem_st_wait();
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) arrive_leader(xb_ready);
    };
    // LayerNorm of (x + bias) over the row, in place (x in TMEM + XB); the block's last one writes the fp32 row
    // to global instead. Two passes over TMEM (statistics, then normalise) in packed fp32x2 arithmetic; the
    // NW warps of a lane quadrant each own every NW-th 32-column chunk and exchange partial statistics.
    // Shared memory left to the L1 by this kernel is ~0, so every __ldg is an L2 round trip: the parameters of the
    // next LayerNorm are staged in shared memory at the start of the phase that ends with it.
    auto stage_ln_params = [&](const float* bias, const float* gamma, const float* beta, const float* qkv_bias) {
      named_bar_sync(1, NSIMT * 32);   // the previous LayerNorm's readers are done
      const int i = threadIdx.x;
      if (i < 3 * (H / 4)) {
        const float* src = (i < H / 4) ? bias : (i < 2 * (H / 4) ? gamma : beta);
        reinterpret_cast<float4*>(lnp)[i] = __ldg(reinterpret_cast<const float4*>(src) + (i % (H / 4)));
        if (qkv_bias != nullptr) reinterpret_cast<float4*>(bqkv_s)[i] = __ldg(reinterpret_cast<const float4*>(qkv_bias) + i);
      }
    };
    auto layer_norm = [&](bool to_global, float* gout) {
      pa.mark(15);
      mbar_wait(accx_full, n_accx & 1); ++n_accx;
      pa.mark(6);
      pa.event(20);   // accx_full observed
      tc_fence_after();
      if (ablate & 4) return;
      named_bar_sync(1, NSIMT * 32);   // the staged parameters (stage_ln_params) are visible to every warp
      uint64_t s2 = 0ull, q2 = 0ull;   // (0.f, 0.f)
      uint32_t v[32];
#pragma unroll 1
      for (int k = part; k < H / 32; k += NW) {
        tmem_ld_32x32(tq + TM_X + static_cast<uint32_t>(k * 32), v);
        tmem_ld_wait_dep(v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = *reinterpret_cast<const float4*>(lnp + k * 32 + 4 * j);
          const uint64_t u0 = f2_add(f2_pack(__uint_as_float(v[4 * j + 0]), __uint_as_float(v[4 * j + 1])), f2_pack(b.x, b.y));
          const uint64_t u1 = f2_add(f2_pack(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), f2_pack(b.z, b.w));
          s2 = f2_add(s2, f2_add(u0, u1));
          q2 = f2_fma(u0, u0, q2);
          q2 = f2_fma(u1, u1, q2);
        }
      }
      pa.mark(7);
      float sum, sq, t0, t1;
      f2_unpack(s2, t0, t1); sum = t0 + t1;
      f2_unpack(q2, t0, t1); sq = t0 + t1;
      float* st = stat_s;
      st[(warp_e * 32 + lane) * 2 + 0] = sum;
      st[(warp_e * 32 + lane) * 2 + 1] = sq;
      named_bar_sync(2 + q, NW * 32);
      pa.mark(8);
      sum = 0.f; sq = 0.f;
#pragma unroll
      for (int p = 0; p < NW; ++p) {
        const float2 t = *reinterpret_cast<const float2*>(st + ((p * 4 + (warp_e & 3)) * 32 + lane) * 2);
        sum += t.x;
        sq += t.y;
      }
      const float mean = sum * (1.0f / H);
      const float rstd = rsqrtf(fmaxf(sq * (1.0f / H) - mean * mean, 0.f) + a.ln_eps);
      const uint64_t rstd2 = f2_pack(rstd, rstd);
      const uint64_t nmr2 = f2_pack(-mean * rstd, -mean * rstd);
#pragma unroll 1
      for (int k = part; k < H / 32; k += NW) {
        tmem_ld_32x32(tq + TM_X + static_cast<uint32_t>(k * 32), v);
        tmem_ld_wait_dep(v);
        float f[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = *reinterpret_cast<const float4*>(lnp + k * 32 + 4 * j);
          const float4 gm = *reinterpret_cast<const float4*>(lnp + H + k * 32 + 4 * j);
          const float4 bt = *reinterpret_cast<const float4*>(lnp + 2 * H + k * 32 + 4 * j);
          uint64_t u0 = f2_add(f2_pack(__uint_as_float(v[4 * j + 0]), __uint_as_float(v[4 * j + 1])), f2_pack(b.x, b.y));
          uint64_t u1 = f2_add(f2_pack(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), f2_pack(b.z, b.w));
          u0 = f2_fma(f2_fma(u0, rstd2, nmr2), f2_pack(gm.x, gm.y), f2_pack(bt.x, bt.y));
          u1 = f2_fma(f2_fma(u1, rstd2, nmr2), f2_pack(gm.z, gm.w), f2_pack(bt.z, bt.w));
          f2_unpack(u0, f[4 * j + 0], f[4 * j + 1]);
          f2_unpack(u1, f[4 * j + 2], f[4 * j + 3]);
        }
        if (to_global) {   // warp-uniform: the block's last LayerNorm leaves the chip
          if (gout != nullptr) {
            float4* g4 = reinterpret_cast<float4*>(gout + k * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) g4[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          }
        } else {
          put_x_chunk(k, f);
        }
      }
      named_bar_sync(2 + q, NW * 32);   // stat_s is rewritten by the next LayerNorm
      pa.mark(9);
      pa.event(21);   // LayerNorm done
    };

    // Rows of the AO tiles that no sequence covers are read by the O-projection as padding rows: start finite.
    for (int i = threadIdx.x; i < 65536 / 16; i += NSIMT * 32) reinterpret_cast<uint4*>(base + OFF_U)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();

    for (int bp = cluster_id; bp < n_pairs; bp += n_clusters) {
      // ---- block metadata ----
      const int blk = 2 * bp + static_cast<int>(rank);
      int s0 = 0, nseq = 0, row0 = 0, ntok = 0;
      if (blk < a.n_blk) {
        s0 = __ldg(a.blk_seq + blk);
        nseq = __ldg(a.blk_seq + blk + 1) - s0;
        const int t0 = __ldg(a.cu + s0);
        row0 = t0 - a.tok0;
        ntok = __ldg(a.cu + s0 + nseq) - t0;
      }
      named_bar_sync(1, NSIMT * 32);   // everyone is done with the previous block's sequence table
      for (int i = threadIdx.x; i <= nseq; i += NSIMT * 32) seq_row[i] = __ldg(a.cu + s0 + i) - a.tok0 - row0;
      named_bar_sync(1, NSIMT * 32);
      if (threadIdx.x == 0) {
        // attention groups: runs of consecutive sequences with at most 16 tokens together share one 16 x 16 tile; a
        // longer sequence is alone in its group. Read after the barriers that open the first attention block.
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
        *n_grp_s = a.no_groups ? -1 : ng;
      }
      FZ_STAMP();
      // ---- load the block's residual rows: global fp32 -> TMEM x + XB ----
      {
        const float* src = a.x + static_cast<size_t>(row0 + r) * H;
#pragma unroll 1
        for (int k = part; k < H / 32; k += NW) {
          float f[32];
          if (r < ntok) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(src + k * 32) + j);
              f[4 * j] = t.x; f[4 * j + 1] = t.y; f[4 * j + 2] = t.z; f[4 * j + 3] = t.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = 0.f;
          }
          put_x_chunk(k, f);
        }
        publish_x();
      }
      FZ_STAMP();
      for (int l = 0; l < a.layers; ++l) {
        if constexpr (DBG) pa.tracing = (bp == cluster_id && l == 1);
        const size_t lo = static_cast<size_t>(l) * static_cast<size_t>(a.layer_stride);
        // =========== attention block ===========
        stage_ln_params(a.bo + lo, a.ln1g + lo, a.ln1b + lo, a.bqkv + lo);
        named_bar_sync(1, NSIMT * 32);   // staged QKV bias visible to both groups
        // The 16 SIMT warps work as two groups of 8 on alternating heads (group g: heads g, g+2, ...): while one
        // group runs the attention of head h out of its own QKV buffer, the other drains and converts head h+1,
        // so a head costs the tensor pipe's Q tile, not the drain + attention chain. Each group covers all four
        // TMEM lane quadrants with two warps (48 of the 96 columns each).
        {
          const int grp = warp_e >> 3;                 // SIMT group of this warp
          const int gw = warp_e & 7;                   // warp inside the group
          const int gpart = (warp_e >> 2) & 1;         // which 48 columns of T1 this warp drains
          uint8_t* qkv_g = reinterpret_cast<uint8_t*>(QKV) + grp * (128 * QROW);
          for (int h = grp; h < HEADS; h += 2) {
            // ---- drain T1 = [q|k|v] of head h: + bias -> bf16 rows in this group's QKV buffer ----
            constexpr int CP = 48;      // columns per warp
            const uint32_t tile = n_t1 + static_cast<uint32_t>(h);     // global index of this T1 tile
            pa.mark(15);
            mbar_wait(&t1_full[tile & 1], (tile >> 1) & 1);
            pa.mark(0);
            pa.event(0);   // t1_full observed (attention)
            tc_fence_after();
            uint32_t v[CP];
            if (!(ablate & 8)) {
              tmem_ld_cols<CP>(tq + TM_T1 + static_cast<uint32_t>(gpart * CP), v);
            } else {
#pragma unroll
              for (int i = 0; i < CP; ++i) v[i] = 0;
            }
            tmem_ld_wait_regs(v);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_leader(t1_free_a);
            pa.mark(1);
            pa.event(1);   // T1 drained, t1_free sent
#pragma unroll
            for (int i = 0; i < CP / 8; ++i) {
              const int col = gpart * CP + i * 8;                       // 8 columns never straddle q|k|v
              const float4* bp8 = reinterpret_cast<const float4*>(bqkv_s + (col >> 5) * H + h * D + (col & 31));
              const float4 b0 = bp8[0], b1 = bp8[1];
              uint4 w;
              w.x = pack_bf16x2(__uint_as_float(v[8 * i + 0]) + b0.x, __uint_as_float(v[8 * i + 1]) + b0.y);
              w.y = pack_bf16x2(__uint_as_float(v[8 * i + 2]) + b0.z, __uint_as_float(v[8 * i + 3]) + b0.w);
              w.z = pack_bf16x2(__uint_as_float(v[8 * i + 4]) + b1.x, __uint_as_float(v[8 * i + 5]) + b1.y);
              w.w = pack_bf16x2(__uint_as_float(v[8 * i + 6]) + b1.z, __uint_as_float(v[8 * i + 7]) + b1.w);
              *reinterpret_cast<uint4*>(const_cast<uint8_t*>(qkv_chunk(qkv_g, r, gpart * 6 + i))) = w;
            }
            pa.mark(2);
            named_bar_sync(6 + grp, (NSIMT / 2) * 32);   // every row of the head is in this group's QKV buffer
            pa.mark(3);
            pa.event(2);   // QKV rows complete (after barrier)
            // ---- attention of this head for the block's sequences (sequence i -> warp i % 8 of the group) ----
            const int cb = (h & 1) * 4;
            bool ao_ok = false;
            auto pre_store = [&]() {   // the O-projection of the previous head pair has read the AO tile
              if (!ao_ok) { mbar_wait(ao_free, (n_pair & 1) ^ 1); ao_ok = true; }
            };
            const int n_grp = *n_grp_s;
            if (n_grp >= 0) {
              for (int u = gw; u < ((ablate & 1) ? 0 : n_grp); u += NSIMT / 2) {
                const int s_first = grp_first[u], s_last = grp_first[u + 1];
                const int beg = seq_row[s_first];
                const int T = seq_row[s_last] - beg;
                if (T <= 0) continue;
                if (T <= 16) attn_group_smem(qkv_g, AO, beg, T, seq_row, s_first, s_last, lane, cb, a.scale_log2, pre_store);
                else if (T <= 32) attn_head_smem<32>(qkv_g, AO, beg, T, lane, cb, a.scale_log2, pre_store);
                else attn_head_smem<64>(qkv_g, AO, beg, T, lane, cb, a.scale_log2, pre_store);
              }
            } else {   // OWC_ATTN_GRP=0: one sequence per tile (A/B)
              for (int i = gw; i < ((ablate & 1) ? 0 : nseq); i += NSIMT / 2) {
                const int beg = seq_row[i];
                const int T = seq_row[i + 1] - beg;
                if (T <= 0) continue;
                if (T <= 16) attn_head_smem<16>(qkv_g, AO, beg, T, lane, cb, a.scale_log2, pre_store);
                else if (T <= 32) attn_head_smem<32>(qkv_g, AO, beg, T, lane, cb, a.scale_log2, pre_store);
                else attn_head_smem<64>(qkv_g, AO, beg, T, lane, cb, a.scale_log2, pre_store);
              }
            }
            pa.mark(4);
            pa.event(3);   // own attention done
            // A warp without sequences must not run ahead either: its arrival for THIS pair may only be counted once
            // the previous pair's phase of ao_full has completed (ao_free is committed behind that O-projection)
            pre_store();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) arrive_leader(ao_full);   // 2 * NSIMT arrivals per head pair: each group's warps for its head
            ++n_pair;
            named_bar_sync(6 + grp, (NSIMT / 2) * 32);   // this group's QKV buffer is rewritten by its next head
            pa.mark(5);
            pa.event(4);   // head done (after barrier)
          }
          n_t1 += HEADS;
        }
        FZ_STAMP();
        layer_norm(false, nullptr);
        publish_x();
        FZ_STAMP();
        // =========== feed-forward block ===========
        stage_ln_params(a.b2 + lo, a.ln2g + lo, a.ln2b + lo, nullptr);
        const float* b1p = a.b1 + lo;
        for (int c = 0; c < a.n_chunks; ++c) {
          constexpr int CP = 128 / NW;   // columns per warp
          float4 bf1[CP / 4];            // bias slice, fetched while the MMAs of the tile still run
#pragma unroll
          for (int i = 0; i < CP / 4; ++i) bf1[i] = __ldg(reinterpret_cast<const float4*>(b1p + c * 128 + part * CP) + i);
          pa.mark(15);
          mbar_wait(&t1_full[n_t1 & 1], (n_t1 >> 1) & 1); ++n_t1;
          pa.mark(10);
          pa.event(10);   // t1_full observed (FFN)
          tc_fence_after();
          uint32_t v[CP];
          if (!(ablate & 8)) {
            tmem_ld_cols<CP>(tq + TM_T1 + static_cast<uint32_t>(part * CP), v);
          } else {
#pragma unroll
            for (int i = 0; i < CP; ++i) v[i] = 0;
          }
          tmem_ld_wait_regs(v);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_leader(t1_free_f);
          pa.mark(11);
          pa.event(11);   // drained
          const int fb = c & 1;
          if (!(ablate & 2)) {
#pragma unroll
          for (int i = 0; i < CP / 4; ++i) {
            const float4 b = bf1[i];
            float x0 = __uint_as_float(v[4 * i + 0]) + b.x, x1 = __uint_as_float(v[4 * i + 1]) + b.y;
            float x2 = __uint_as_float(v[4 * i + 2]) + b.z, x3 = __uint_as_float(v[4 * i + 3]) + b.w;
            gelu_tanh2(x0, x1);   // tanh form: half the issue slots, error an order below the bf16 rounding (common.cuh)
            gelu_tanh2(x2, x3);
            v[2 * i + 0] = pack_bf16x2(x0, x1);   // packed results reuse the front of v[]
            v[2 * i + 1] = pack_bf16x2(x2, x3);
          }
          }
          pa.mark(12);
          pa.event(12);   // GELU done
          mbar_wait(&f_free[fb], (f_use[fb] & 1) ^ 1);   // FFN2 of the buffer's previous use has read it
          pa.mark(13);
          {
            // CP columns = CP/8 16-byte chunks, starting at chunk (part*CP/8) of the 128-column slice
            uint8_t* fbuf = FC + fb * 32768;
#pragma unroll
            for (int i = 0; i < CP / 8; ++i) {
              const int ch = part * (CP / 8) + i;          // 0..15 over the 128 columns
              uint8_t* rowp = fbuf + (ch >> 3) * 16384 + r * 128;
              *reinterpret_cast<uint4*>(rowp + (((ch & 7) ^ (r & 7)) << 4)) = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) arrive_leader(&f_full[fb]);
          ++f_use[fb];
          pa.mark(14);
          pa.event(13);   // FC stored, f_full sent
        }
        FZ_STAMP();
        const bool last = l + 1 == a.layers;
        float* gout = (last && r < ntok) ? a.x + static_cast<size_t>(row0 + r) * H : nullptr;
        layer_norm(last, gout);   // rows past ntok write nowhere
        if (!last) publish_x();
        FZ_STAMP();
      }
    }
#undef FZ_STAMP
    if constexpr (DBG) {
      if (stamp)
        for (int i = 0; i < 16; ++i) a.dbg[32 + i] = pa.a[i];
      // phase sums of four more observers: leader warp 15, peer warps 0 and 15, leader warp 7
      if (a.dbg != nullptr && blockIdx.x < 2 && lane == 0 && (warp == 0 || warp == 15 || warp == 7)) {
        const int slot = static_cast<int>(blockIdx.x) * 3 + (warp == 0 ? 0 : (warp == 15 ? 1 : 2));
        for (int i = 0; i < 16; ++i) a.dbg[1024 + slot * 16 + i] = pa.a[i];
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();   // the peer's shared memory and TMEM must outlive the leader's last MMA
  if (warp == W_TMA) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

}  // namespace owc
