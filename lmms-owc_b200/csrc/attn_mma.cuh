// attn_mma.cuh - short-sequence (<= 64 token) self-attention for the production (bf16) mode.
//
// One WARP per (sequence, head), no block-level synchronisation: each warp stages its head's Q/K/V
// slices (T x D bf16 each, 64-128 B rows) into a private shared-memory region with 16-byte loads,
// runs the two small products QK^T and PV on the warp-level tensor-core path (mma.sync m16n8k16
// bf16 -> fp32; a sequence is 3..64 tokens, far below the 128-row tcgen05 tile), keeps the softmax in
// registers (quad shuffles), rounds probabilities to bf16 only as the PV operand, and stages the
// output back through its region for 16-byte coalesced stores. Warps are persistent (grid-stride over
// the pairs) and pick the smallest tile padding (16/32/64) per sequence, so a 7-token label costs one
// m-tile and one key block regardless of the longest sequence in the chunk.
//   Row pitch D+8 elements: every fragment load (Q/K as 32-bit pairs, V via ldmatrix.trans) and every
//   16-byte staging store is bank-conflict free.
// Packed sequences have no padding, so the reference's additive padding mask (modeling_bert.py:114-140)
// is a no-op and only the tile padding (keys >= T) is masked. MPNet's relative-position bias
// (modeling_mpnet.py:163-173, 324-362) comes from the per-head LUT indexed by (key - query).
#pragma once

#include "common.cuh"
#include "kernels.cuh"

namespace owc {

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// Two transposed 8x8 b16 matrices: the B fragment (k = key, n = dim) of a row-major [key][dim] tile.
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t (&b)[2], const void* row_ptr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];"
               : "=r"(b[0]), "=r"(b[1])
               : "r"(smem_u32(row_ptr)));
}

// 16-byte asynchronous copy global -> shared (LDGSTS): a warp fires every chunk of a unit's Q/K/V rows back to back and
// waits once, instead of one load -> store round trip at a time (the staging loop ran at ~1 KB in flight per warp, which
// held the kernel at 2.8 TB/s against 6.5 TB/s of HBM).
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int D>
struct AttnWarpCfg {
  static constexpr int PITCH = D + 8;                       // bf16 elements per staged row
  __host__ __device__ static constexpr int region_elems(int tpmax) { return 3 * tpmax * PITCH; }
  __host__ __device__ static constexpr int smem_bytes(int tpmax) { return 4 * region_elems(tpmax) * 2; }   // 4 warps per CTA
};

// One (sequence, head) pair with tile padding TP, entirely inside the calling warp.
template <int D, int TP>
__device__ __forceinline__ void attn_pair(const AttnArgs& a, __nv_bfloat16* reg, int tpmax, int lane, int beg, int T,
                                          int head, float scale_log2) {
  constexpr int P = AttnWarpCfg<D>::PITCH;
  constexpr int NT = TP / 8;     // key n-tiles of S
  constexpr int MT = TP / 16;    // query m-tiles
  constexpr int DN = D / 8;      // output n-tiles
  constexpr int CPR = D / 8;     // 16-byte chunks per staged row
  const int g = lane >> 2, t4 = lane & 3;
  const int H = a.H;
  __nv_bfloat16* Qh = reg;
  __nv_bfloat16* Kh = reg + tpmax * P;
  __nv_bfloat16* Vh = reg + 2 * tpmax * P;
  const __nv_bfloat16* qkv = reinterpret_cast<const __nv_bfloat16*>(a.qkv) + static_cast<size_t>(beg) * 3 * H + head * D;

  // ---- stage Q, K, V rows of this head (row-major) ----
  for (int idx = lane; idx < T * 3 * CPR; idx += 32) {
    const int c = idx % CPR;
    const int p = (idx / CPR) % 3;
    const int t = idx / (3 * CPR);
    cp_async_16(reg + (p * tpmax + t) * P + c * 8, qkv + static_cast<size_t>(t) * 3 * H + p * H + c * 8);
  }
  cp_async_wait_all();
  __syncwarp();

  const float* lut = a.rel_lut ? a.rel_lut + static_cast<size_t>(head) * (2 * a.max_pos - 1) + a.max_pos - 1 : nullptr;
#pragma unroll 1
  for (int mt = 0; mt < MT; ++mt) {
    if (mt * 16 >= T) break;
    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) { s[nt][0] = 0.f; s[nt][1] = 0.f; s[nt][2] = 0.f; s[nt][3] = 0.f; }
    const int r0 = mt * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks) {
      uint32_t af[4];
      af[0] = *reinterpret_cast<const uint32_t*>(Qh + r0 * P + ks * 16 + t4 * 2);
      af[1] = *reinterpret_cast<const uint32_t*>(Qh + r1 * P + ks * 16 + t4 * 2);
      af[2] = *reinterpret_cast<const uint32_t*>(Qh + r0 * P + ks * 16 + 8 + t4 * 2);
      af[3] = *reinterpret_cast<const uint32_t*>(Qh + r1 * P + ks * 16 + 8 + t4 * 2);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        if (nt * 8 < T) {   // warp-uniform
          uint32_t bf[2];
          bf[0] = *reinterpret_cast<const uint32_t*>(Kh + (nt * 8 + g) * P + ks * 16 + t4 * 2);
          bf[1] = *reinterpret_cast<const uint32_t*>(Kh + (nt * 8 + g) * P + ks * 16 + 8 + t4 * 2);
          mma_bf16_16816(s[nt], af, bf);
        }
      }
    }
    // scale (log2 domain), relative bias, mask padded keys; row maxima
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (nt * 8 < T) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = nt * 8 + t4 * 2 + (e & 1);
          const int i = (e < 2) ? r0 : r1;
          float v = s[nt][e] * scale_log2;
          if (lut != nullptr && j < T && i < T) v += __ldg(lut + (j - i)) * 1.4426950408889634f;
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
    // probabilities rounded to bf16 (the PV operand); row sums of the ROUNDED values
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
    // O = P V  (V row-major; ldmatrix.trans yields the k=key, n=dim fragment)
    float o[DN][4];
#pragma unroll
    for (int dn = 0; dn < DN; ++dn) { o[dn][0] = 0.f; o[dn][1] = 0.f; o[dn][2] = 0.f; o[dn][3] = 0.f; }
#pragma unroll
    for (int kt = 0; kt < MT; ++kt) {
      if (kt * 16 < T) {   // warp-uniform
        const uint32_t af[4] = {pk[2 * kt][0], pk[2 * kt][1], pk[2 * kt + 1][0], pk[2 * kt + 1][1]};
        const __nv_bfloat16* vrow = Vh + (kt * 16 + (lane & 15)) * P;
#pragma unroll
        for (int dn = 0; dn < DN; ++dn) {
          uint32_t bf[2];
          ldmatrix_x2_trans(bf, vrow + dn * 8);
          mma_bf16_16816(o[dn], af, bf);
        }
      }
    }
    // normalise; stage this m-tile's output over its own (now dead) Q rows
    const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
    __syncwarp();
#pragma unroll
    for (int dn = 0; dn < DN; ++dn) {
      if (r0 < T) *reinterpret_cast<uint32_t*>(Qh + r0 * P + dn * 8 + t4 * 2) = pack_bf16x2(o[dn][0] * inv0, o[dn][1] * inv0);
      if (r1 < T) *reinterpret_cast<uint32_t*>(Qh + r1 * P + dn * 8 + t4 * 2) = pack_bf16x2(o[dn][2] * inv1, o[dn][3] * inv1);
    }
  }
  __syncwarp();
  // ---- coalesced write-back (16 B per lane) ----
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.out) + static_cast<size_t>(beg) * H + head * D;
  for (int idx = lane; idx < T * CPR; idx += 32) {
    const int c = idx % CPR;
    const int t = idx / CPR;
    *reinterpret_cast<uint4*>(out + static_cast<size_t>(t) * H + c * 8) = *reinterpret_cast<const uint4*>(Qh + t * P + c * 8);
  }
  __syncwarp();   // region is reused by the next pair
}

// Several SHORT sequences in one 16 x 16 tile. Label strings are 3-8 tokens, so a (sequence, head) pair fills a fraction
// of the smallest MMA tile and the kernel was issue-bound on the fixed per-pair cost (staging loops, 16 MMAs, softmax
// shuffles, write-back: ~700 warp instructions for 5 tokens). A GROUP is a run of consecutive sequences with at most 16
// tokens in total (the host builds the table, api.cu): their rows are contiguous in the packed activations, so one
// staging pass, one 16 x 16 score tile with a block-diagonal mask and one P.V product serve all of them - the same
// instruction count as ONE pair before. lo/hi: for tile row i, the key range [lo_i, hi_i) of its own sequence.
template <int D>
__device__ __forceinline__ void attn_group16(const AttnArgs& a, __nv_bfloat16* reg, int tpmax, int lane, int gbeg, int Tg,
                                             int s_first, int s_last, int head, float scale_log2) {
  constexpr int P = AttnWarpCfg<D>::PITCH;
  constexpr int NT = 2, DN = D / 8, CPR = D / 8;
  const int g = lane >> 2, t4 = lane & 3;
  const int H = a.H;
  __nv_bfloat16* Qh = reg;
  __nv_bfloat16* Kh = reg + tpmax * P;
  __nv_bfloat16* Vh = reg + 2 * tpmax * P;
  const __nv_bfloat16* qkv = reinterpret_cast<const __nv_bfloat16*>(a.qkv) + static_cast<size_t>(gbeg) * 3 * H + head * D;
  for (int idx = lane; idx < Tg * 3 * CPR; idx += 32) {
    const int c = idx % CPR;
    const int p = (idx / CPR) % 3;
    const int t = idx / (3 * CPR);
    cp_async_16(reg + (p * tpmax + t) * P + c * 8, qkv + static_cast<size_t>(t) * 3 * H + p * H + c * 8);
  }
  // key range of every tile row: lane r (< 16) finds the sequence that owns row r
  int lo = 0, hi = 0;
  for (int s = s_first; s < s_last; ++s) {
    const int b = a.cu[s] - a.tok0 - gbeg, e = a.cu[s + 1] - a.tok0 - gbeg;
    if (lane >= b && lane < e) { lo = b; hi = e; }
  }
  const int r0 = g, r1 = g + 8;
  const int lo0 = __shfl_sync(0xffffffffu, lo, r0), hi0 = __shfl_sync(0xffffffffu, hi, r0);
  const int lo1 = __shfl_sync(0xffffffffu, lo, r1), hi1 = __shfl_sync(0xffffffffu, hi, r1);
  cp_async_wait_all();   // the sequence lookup above ran under the copies
  __syncwarp();

  const float* lut = a.rel_lut ? a.rel_lut + static_cast<size_t>(head) * (2 * a.max_pos - 1) + a.max_pos - 1 : nullptr;
  float s[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) { s[nt][0] = 0.f; s[nt][1] = 0.f; s[nt][2] = 0.f; s[nt][3] = 0.f; }
#pragma unroll
  for (int ks = 0; ks < D / 16; ++ks) {
    uint32_t af[4];
    af[0] = *reinterpret_cast<const uint32_t*>(Qh + r0 * P + ks * 16 + t4 * 2);
    af[1] = *reinterpret_cast<const uint32_t*>(Qh + r1 * P + ks * 16 + t4 * 2);
    af[2] = *reinterpret_cast<const uint32_t*>(Qh + r0 * P + ks * 16 + 8 + t4 * 2);
    af[3] = *reinterpret_cast<const uint32_t*>(Qh + r1 * P + ks * 16 + 8 + t4 * 2);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (nt * 8 < Tg) {   // warp-uniform
        uint32_t bf[2];
        bf[0] = *reinterpret_cast<const uint32_t*>(Kh + (nt * 8 + g) * P + ks * 16 + t4 * 2);
        bf[1] = *reinterpret_cast<const uint32_t*>(Kh + (nt * 8 + g) * P + ks * 16 + 8 + t4 * 2);
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
      const int i = (e < 2) ? r0 : r1;
      const bool ok = (e < 2) ? (j >= lo0 && j < hi0) : (j >= lo1 && j < hi1);
      float v = s[nt][e] * scale_log2;
      if (lut != nullptr && ok) v += __ldg(lut + (j - i)) * 1.4426950408889634f;
      v = ok ? v : -INFINITY;
      s[nt][e] = v;
      if (e < 2) m0 = fmaxf(m0, v); else m1 = fmaxf(m1, v);
    }
  }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  // rows past the group (lo == hi) have no valid key: keep exp2(-inf - -inf) out of the arithmetic
  m0 = (hi0 > lo0) ? m0 : 0.f;
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
    const __nv_bfloat16* vrow = Vh + (lane & 15) * P;
#pragma unroll
    for (int dn = 0; dn < DN; ++dn) {
      uint32_t bf[2];
      ldmatrix_x2_trans(bf, vrow + dn * 8);
      mma_bf16_16816(o[dn], af, bf);
    }
  }
  const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;   // unused for rows past the group (sum 0)
  __syncwarp();
#pragma unroll
  for (int dn = 0; dn < DN; ++dn) {
    if (r0 < Tg) *reinterpret_cast<uint32_t*>(Qh + r0 * P + dn * 8 + t4 * 2) = pack_bf16x2(o[dn][0] * inv0, o[dn][1] * inv0);
    if (r1 < Tg) *reinterpret_cast<uint32_t*>(Qh + r1 * P + dn * 8 + t4 * 2) = pack_bf16x2(o[dn][2] * inv1, o[dn][3] * inv1);
  }
  __syncwarp();
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.out) + static_cast<size_t>(gbeg) * H + head * D;
  for (int idx = lane; idx < Tg * CPR; idx += 32) {
    const int c = idx % CPR;
    const int t = idx / CPR;
    *reinterpret_cast<uint4*>(out + static_cast<size_t>(t) * H + c * 8) = *reinterpret_cast<const uint4*>(Qh + t * P + c * 8);
  }
  __syncwarp();   // region is reused by the next unit
}

// Persistent warps over (group, head) units; a group is either a run of short sequences (<= 16 tokens together) or one
// longer sequence (17..64 tokens), which takes the single-sequence body above.
template <int D, int TPMAX>
__global__ void __launch_bounds__(128) k_attn_grp(const AttnArgs a) {
  using Cfg = AttnWarpCfg<D>;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __nv_bfloat16* reg = reinterpret_cast<__nv_bfloat16*>(smem_attn) + warp * Cfg::region_elems(TPMAX);
  for (int i = lane; i < Cfg::region_elems(TPMAX) / 8; i += 32) reinterpret_cast<uint4*>(reg)[i] = make_uint4(0, 0, 0, 0);
  __syncwarp();

  pdl_launch_dependents();
  pdl_wait();   // the QKV projections of this chunk are complete and visible

  const float scale_log2 = a.scale * 1.4426950408889634f;
  const long long units = static_cast<long long>(a.n_grp) * a.heads;
  const long long stride = static_cast<long long>(gridDim.x) * 4;
  for (long long p = static_cast<long long>(blockIdx.x) * 4 + warp; p < units; p += stride) {
    const int grp = static_cast<int>(p / a.heads);
    const int head = static_cast<int>(p % a.heads);
    const int s_first = a.grp[grp], s_last = a.grp[grp + 1];           // absolute sequence indices
    const int gbeg = a.cu[s_first] - a.tok0;
    const int Tg = a.cu[s_last] - a.tok0 - gbeg;
    if (Tg <= 0) continue;
    if (Tg <= 16) attn_group16<D>(a, reg, TPMAX, lane, gbeg, Tg, s_first, s_last, head, scale_log2);
    else if (TPMAX == 32 || Tg <= 32) attn_pair<D, (TPMAX >= 32 ? 32 : 16)>(a, reg, TPMAX, lane, gbeg, Tg, head, scale_log2);
    else attn_pair<D, (TPMAX >= 64 ? 64 : 16)>(a, reg, TPMAX, lane, gbeg, Tg, head, scale_log2);
  }
}

template <int D, int TPMAX>
__global__ void __launch_bounds__(128) k_attn_mma(const AttnArgs a) {
  using Cfg = AttnWarpCfg<D>;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __nv_bfloat16* reg = reinterpret_cast<__nv_bfloat16*>(smem_attn) + warp * Cfg::region_elems(TPMAX);

  // zero once: padded V rows must hold finite values (0 * garbage could be NaN); staging only ever
  // writes finite activations afterwards
  for (int i = lane; i < Cfg::region_elems(TPMAX) / 8; i += 32) reinterpret_cast<uint4*>(reg)[i] = make_uint4(0, 0, 0, 0);
  __syncwarp();

  pdl_launch_dependents();
  pdl_wait();   // the QKV projections of this chunk are complete and visible

  const float scale_log2 = a.scale * 1.4426950408889634f;
  const long long pairs = static_cast<long long>(a.n_seq) * a.heads;
  const long long stride = static_cast<long long>(gridDim.x) * 4;
  for (long long p = static_cast<long long>(blockIdx.x) * 4 + warp; p < pairs; p += stride) {
    const int seq = static_cast<int>(p / a.heads);
    const int head = static_cast<int>(p % a.heads);
    const int beg = a.cu[a.seq0 + seq] - a.tok0;
    const int T = a.cu[a.seq0 + seq + 1] - a.tok0 - beg;
    if (T <= 0) continue;
    if (TPMAX == 16 || T <= 16) attn_pair<D, 16>(a, reg, TPMAX, lane, beg, T, head, scale_log2);
    else if (TPMAX == 32 || T <= 32) attn_pair<D, (TPMAX >= 32 ? 32 : 16)>(a, reg, TPMAX, lane, beg, T, head, scale_log2);
    else attn_pair<D, (TPMAX >= 64 ? 64 : 16)>(a, reg, TPMAX, lane, beg, T, head, scale_log2);
  }
}

}  // namespace owc
