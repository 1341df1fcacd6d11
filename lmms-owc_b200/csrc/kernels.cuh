// kernels.cuh - the HBM-/latency-bound kernels of the scoring path (everything that is not a dense
// contraction) plus the fp32 "check mode" SIMT contraction.
//
//   k_embed_ln   K1  token+position(+type) embedding gather fused with LayerNorm
//   k_ln         K4/K6 tail: LayerNorm of the pre-LN sum written by the GEMM epilogue
//   k_attn       K3  short-sequence attention, K in shared memory, warp-shuffle softmax
//   k_pool_l2    K7  masked mean pooling fused with L2 normalisation
//   k_pair_dot   K8  paired score  s_i = <a_i, b_i>
//   k_rescore        exact fp32 re-scoring of the bf16 kernel's candidates + final top-k + paired
//   k_topk_rows      top-k of materialised score rows (check mode only)
//   k_gemm_f32       fp32 SIMT contraction with the same fused epilogues (check mode only)
#pragma once

#include "common.cuh"

namespace owc {

// ---- activation element I/O: bf16 in production mode, fp32 in check mode ----
__device__ __forceinline__ float ld_act(const float* p) { return *p; }
__device__ __forceinline__ float ld_act(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st_act(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_act(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// LayerNorm of one row held as NV float4 per lane (lane-strided), nn.LayerNorm semantics
// (biased variance, eps inside the sqrt); writes fp32 and optionally bf16.
template <int NV>
__device__ __forceinline__ void ln_row_store(float4 (&x)[NV], int lane, int H, const float* __restrict__ gamma,
                                             const float* __restrict__ beta, float eps, float* __restrict__ out_f32,
                                             __nv_bfloat16* __restrict__ out_bf16) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
  const float mean = warp_sum(s) / static_cast<float>(H);
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    x[i].x -= mean; x[i].y -= mean; x[i].z -= mean; x[i].w -= mean;
    v += (x[i].x * x[i].x + x[i].y * x[i].y) + (x[i].z * x[i].z + x[i].w * x[i].w);
  }
  const float rstd = rsqrtf(warp_sum(v) / static_cast<float>(H) + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c4 = lane + 32 * i;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
    float4 y;
    y.x = x[i].x * rstd * g.x + b.x;
    y.y = x[i].y * rstd * g.y + b.y;
    y.z = x[i].z * rstd * g.z + b.z;
    y.w = x[i].w * rstd * g.w + b.w;
    reinterpret_cast<float4*>(out_f32)[c4] = y;
    if (out_bf16 != nullptr) {
      uint2 p;
      p.x = pack_bf16x2(y.x, y.y);
      p.y = pack_bf16x2(y.z, y.w);
      reinterpret_cast<uint2*>(out_bf16)[c4] = p;
    }
  }
}

// The same with gamma / beta already in registers (the embedding kernel keeps them for the warp's lifetime: re-reading
// them per token through the L1 made that kernel L1-throughput-bound at 72 % l1tex against 36 % DRAM).
template <int NV>
__device__ __forceinline__ void ln_row_store_reg(float4 (&x)[NV], int lane, int H, const float4 (&g)[NV], const float4 (&b)[NV],
                                                 float eps, float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
  const float mean = warp_sum(s) / static_cast<float>(H);
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    x[i].x -= mean; x[i].y -= mean; x[i].z -= mean; x[i].w -= mean;
    v += (x[i].x * x[i].x + x[i].y * x[i].y) + (x[i].z * x[i].z + x[i].w * x[i].w);
  }
  const float rstd = rsqrtf(warp_sum(v) / static_cast<float>(H) + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c4 = lane + 32 * i;
    float4 y;
    y.x = x[i].x * rstd * g[i].x + b[i].x;
    y.y = x[i].y * rstd * g[i].y + b[i].y;
    y.z = x[i].z * rstd * g[i].z + b[i].z;
    y.w = x[i].w * rstd * g[i].w + b[i].w;
    reinterpret_cast<float4*>(out_f32)[c4] = y;
    if (out_bf16 != nullptr) {
      uint2 p;
      p.x = pack_bf16x2(y.x, y.y);
      p.y = pack_bf16x2(y.z, y.w);
      reinterpret_cast<uint2*>(out_bf16)[c4] = p;
    }
  }
}

struct EmbedArgs {
  const int32_t* ids;        // packed ids of the whole call
  const int32_t* cu;         // cu_seqlens of the whole call (device copy)
  int seq0, n_seq;           // sequences [seq0, seq0 + n_seq) form this chunk
  int tok0;                  // cu[seq0]: first token of the chunk (rows are chunk-relative)
  const float* word;         // [vocab, H]
  const float* pos;          // [max_pos, H]
  const float* type0;        // BERT: row 0 of the token-type table; MPNet: nullptr
  const float* gamma;
  const float* beta;
  float eps;
  int H, vocab, max_pos;
  int mpnet, pad_id;
  float* x;                  // [chunk tokens, H] fp32 residual stream
  __nv_bfloat16* xb;         // bf16 operand copy, or nullptr in check mode
  int* status;               // bit 0: an id was out of range (clamped)
  float* stats;              // wide path: [chunk tokens][2] row (sum, sumsq); the stream then is the UN-normalised sum
  __nv_bfloat16* xl;         // wide path: the stream as a bf16 PAIR, xb = bf16(v) and xl = bf16(v - xb) (16 mantissa bits), no fp32 copy
};

// One warp per sequence; tokens visited in order so MPNet's position id (running count of
// non-pad ids, modeling_mpnet.py:889-897) is a register counter. BERT: position = index
// (modeling_bert.py:53-112, token_type 0 as the tokenizer emits for single sentences).
// Two tokens are in flight per iteration (ids and positions first, then both gathers, both LayerNorms):
// a token is a dependent chain gather -> reduce -> reduce -> store, and one chain per warp left the
// kernel at 45 % of the HBM write roofline.
// Warps are persistent over the sequences (grid-stride): the LayerNorm parameters and the token-type row are loaded
// once per warp and live in registers.
// STATS (a.stats != nullptr): the deferred-LayerNorm form of the wide path - no normalisation here, so no parameters.
template <int NV, bool STATS>
__global__ void __launch_bounds__(128) k_embed_ln(const EmbedArgs a) {
  const int lane = threadIdx.x & 31;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  constexpr int NP = STATS ? 1 : NV;
  float4 gam[NP], bet[NP], typ[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c4 = lane + 32 * i;
    typ[i] = a.type0 != nullptr ? __ldg(reinterpret_cast<const float4*>(a.type0) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (!STATS) {
      gam[i] = __ldg(reinterpret_cast<const float4*>(a.gamma) + c4);
      bet[i] = __ldg(reinterpret_cast<const float4*>(a.beta) + c4);
    }
  }
  for (int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < a.n_seq; s += warps_total) {
  const int beg = a.cu[a.seq0 + s], end = a.cu[a.seq0 + s + 1];
  int running = 0;
  auto clamp_id = [&](int id) {
    if (id < 0 || id >= a.vocab) {
      if (lane == 0) atomicOr(a.status, 1);
      id = min(max(id, 0), a.vocab - 1);
    }
    return id;
  };
  auto position = [&](int id, int t) {
    int p;
    if (a.mpnet) {
      const int np = (id != a.pad_id) ? 1 : 0;
      running += np;
      p = running * np + a.pad_id;
    } else {
      p = t - beg;
    }
    return min(p, a.max_pos - 1);
  };
  auto gather = [&](int id, int p, float4 (&x)[NV]) {
    const float4* w4 = reinterpret_cast<const float4*>(a.word + static_cast<size_t>(id) * a.H);
    const float4* p4 = reinterpret_cast<const float4*>(a.pos + static_cast<size_t>(p) * a.H);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = lane + 32 * i;
      const float4 w = __ldg(w4 + c4);
      const float4 q = __ldg(p4 + c4);
      if (a.type0 != nullptr) {  // reference order: (word + type) + position
        const float4 ty = typ[i];
        x[i] = make_float4((w.x + ty.x) + q.x, (w.y + ty.y) + q.y, (w.z + ty.z) + q.z, (w.w + ty.w) + q.w);
      } else {
        x[i] = make_float4(w.x + q.x, w.y + q.y, w.z + q.z, w.w + q.w);
      }
    }
  };
  auto store = [&](int t, float4 (&x)[NV]) {
    const size_t row = static_cast<size_t>(t - a.tok0);
    if constexpr (STATS) {
      // deferred LayerNorm (wide path, gemm_wide.cuh): the un-normalised sum, its bf16 copy and the row statistics
      float s = 0.f, q = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c4 = lane + 32 * i;
        s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
        q += (x[i].x * x[i].x + x[i].y * x[i].y) + (x[i].z * x[i].z + x[i].w * x[i].w);
        uint2 p, l;
        split_bf16x2(x[i].x, x[i].y, p.x, l.x);
        split_bf16x2(x[i].z, x[i].w, p.y, l.y);
        reinterpret_cast<uint2*>(a.xb + row * a.H)[c4] = p;
        reinterpret_cast<uint2*>(a.xl + row * a.H)[c4] = l;
      }
      s = warp_sum(s);
      q = warp_sum(q);
      if (lane == 0) reinterpret_cast<float2*>(a.stats)[row] = make_float2(s, q);
    } else {
      ln_row_store_reg<NV>(x, lane, a.H, gam, bet, a.eps, a.x + row * a.H, a.xb ? a.xb + row * a.H : nullptr);
    }
  };
  int t = beg;
  for (; t + 1 < end; t += 2) {
    const int id0 = clamp_id(a.ids[t]), id1 = clamp_id(a.ids[t + 1]);
    const int p0 = position(id0, t), p1 = position(id1, t + 1);
    float4 x0[NV], x1[NV];
    gather(id0, p0, x0);
    gather(id1, p1, x1);
    store(t, x0);
    store(t + 1, x1);
  }
  if (t < end) {
    const int id0 = clamp_id(a.ids[t]);
    const int p0 = position(id0, t);
    float4 x0[NV];
    gather(id0, p0, x0);
    store(t, x0);
  }
  }
}

// y [M, H] fp32 (= dense + bias + residual) -> LayerNorm -> x fp32 (+ bf16 copy). Warp per row.
template <int NV>
__global__ void __launch_bounds__(256) k_ln(const float* __restrict__ y, int M, int H, const float* __restrict__ gamma,
                                            const float* __restrict__ beta, float eps, float* __restrict__ x,
                                            __nv_bfloat16* __restrict__ xb) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const float4* y4 = reinterpret_cast<const float4*>(y + static_cast<size_t>(row) * H);
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = y4[lane + 32 * i];
  ln_row_store<NV>(v, lane, H, gamma, beta, eps, x + static_cast<size_t>(row) * H,
                   xb ? xb + static_cast<size_t>(row) * H : nullptr);
}

struct AttnArgs {
  const void* qkv;           // [M, 3H]: q | k | v per token row (bias already added)
  void* out;                 // [M, H]
  const int32_t* cu;
  int seq0, n_seq, tok0;
  int H, heads;
  int max_t;                 // longest sequence in the chunk (sizes the shared buffers)
  float scale;               // head_dim^-0.5
  const float* rel_lut;      // MPNet: [heads, 2*max_pos-1] bias by (k - q); nullptr for BERT
  int max_pos;
  const int32_t* grp;        // k_attn_grp: [n_grp + 1] absolute index of the first sequence of each group (device)
  int n_grp;
};

// softmax(q k^T * scale (+ rel bias)) v for one (sequence, head) per group of WPP warps.
// Packed sequences carry no padding, so the reference's additive padding mask (finfo.min on pad
// keys, modeling_bert.py:114-140) contributes exactly-zero probabilities and is omitted.
// K of the head sits in shared memory as fp32 [T][D+1] (bank-conflict-free for lane = key);
// phase 1: lane = key, phase 2: lane = output dim with V streamed from global (coalesced).
template <typename T, int D, int WPP, int MAXKPL>
__global__ void __launch_bounds__(128) k_attn(const AttnArgs a) {
  constexpr int WARPS = 4;
  constexpr int GROUPS = WARPS / WPP;
  extern __shared__ float smem_f[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = warp / WPP, wq = warp % WPP;
  float* ks = smem_f + static_cast<size_t>(grp) * a.max_t * (D + 1);
  float* qs = smem_f + static_cast<size_t>(GROUPS) * a.max_t * (D + 1) + warp * D;
  float* ps = smem_f + static_cast<size_t>(GROUPS) * a.max_t * (D + 1) + WARPS * D + static_cast<size_t>(warp) * a.max_t;

  const long long pair = static_cast<long long>(blockIdx.x) * GROUPS + grp;
  const bool active = pair < static_cast<long long>(a.n_seq) * a.heads;
  int seq = 0, head = 0, beg = 0, Tn = 0;
  if (active) {
    seq = static_cast<int>(pair / a.heads);
    head = static_cast<int>(pair % a.heads);
    beg = a.cu[a.seq0 + seq] - a.tok0;
    Tn = a.cu[a.seq0 + seq + 1] - a.tok0 - beg;
  }
  const T* base = reinterpret_cast<const T*>(a.qkv);
  const int ld = 3 * a.H;
  // K rows of this head -> shared (each warp of the group takes rows wq, wq+WPP, ...)
  for (int j = wq; j < Tn; j += WPP) {
    const T* kr = base + static_cast<size_t>(beg + j) * ld + a.H + head * D;
#pragma unroll
    for (int d = lane; d < D; d += 32) ks[j * (D + 1) + d] = ld_act(kr + d);
  }
  if constexpr (WPP == 1) __syncwarp(); else __syncthreads();
  if (!active) return;

  for (int i = wq; i < Tn; i += WPP) {
    const T* qr = base + static_cast<size_t>(beg + i) * ld + head * D;
#pragma unroll
    for (int d = lane; d < D; d += 32) qs[d] = ld_act(qr + d);
    __syncwarp();
    float sc[MAXKPL];
    float mx = -INFINITY;
#pragma unroll
    for (int kk = 0; kk < MAXKPL; ++kk) {
      const int j = kk * 32 + lane;
      float s = -INFINITY;
      if (j < Tn) {
        const float* kr = ks + j * (D + 1);
        float acc = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) acc = fmaf(qs[d], kr[d], acc);
        s = acc * a.scale;
        if (a.rel_lut != nullptr) s += __ldg(a.rel_lut + static_cast<size_t>(head) * (2 * a.max_pos - 1) + (j - i) + a.max_pos - 1);
      }
      sc[kk] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int kk = 0; kk < MAXKPL; ++kk) {
      const int j = kk * 32 + lane;
      const float e = (j < Tn) ? expf(sc[kk] - mx) : 0.f;
      sc[kk] = e;
      sum += e;
    }
    const float inv = 1.0f / warp_sum(sum);
#pragma unroll
    for (int kk = 0; kk < MAXKPL; ++kk) {
      const int j = kk * 32 + lane;
      if (j < Tn) ps[j] = sc[kk] * inv;
    }
    __syncwarp();
    const T* vbase = base + static_cast<size_t>(beg) * ld + 2 * a.H + head * D;
    T* orow = reinterpret_cast<T*>(a.out) + static_cast<size_t>(beg + i) * a.H + head * D;
#pragma unroll
    for (int d = lane; d < D; d += 32) {
      float acc = 0.f;
      for (int j = 0; j < Tn; ++j) acc = fmaf(ps[j], ld_act(vbase + static_cast<size_t>(j) * ld + d), acc);
      st_act(orow + d, acc);
    }
    __syncwarp();
  }
}

// Masked mean pooling (_text.py:175-189: sum over real tokens / clamp(count, 1e-9)) fused with the
// L2 normalisation of _text.py:202 (no epsilon: an all-zero row yields NaN, as in the reference).
template <int NV>
__global__ void __launch_bounds__(128) k_pool_l2(const float* __restrict__ x, const int32_t* __restrict__ cu, int seq0,
                                                 int n_seq, int tok0, int H, float* __restrict__ z,
                                                 __nv_bfloat16* __restrict__ zb) {
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= n_seq) return;
  const int beg = cu[seq0 + s] - tok0, end = cu[seq0 + s + 1] - tok0;
  float4 acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = beg; t < end; ++t) {
    const float4* r4 = reinterpret_cast<const float4*>(x + static_cast<size_t>(t) * H);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 v = r4[lane + 32 * i];
      acc[i].x += v.x; acc[i].y += v.y; acc[i].z += v.z; acc[i].w += v.w;
    }
  }
  const float cnt = fmaxf(static_cast<float>(end - beg), 1e-9f);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    acc[i].x /= cnt; acc[i].y /= cnt; acc[i].z /= cnt; acc[i].w /= cnt;
    ss += (acc[i].x * acc[i].x + acc[i].y * acc[i].y) + (acc[i].z * acc[i].z + acc[i].w * acc[i].w);
  }
  const float nrm = sqrtf(warp_sum(ss));
  const size_t row = static_cast<size_t>(seq0 + s) * H;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c4 = lane + 32 * i;
    const float4 y = make_float4(acc[i].x / nrm, acc[i].y / nrm, acc[i].z / nrm, acc[i].w / nrm);
    reinterpret_cast<float4*>(z + row)[c4] = y;
    if (zb != nullptr) {
      uint2 p;
      p.x = pack_bf16x2(y.x, y.y);
      p.y = pack_bf16x2(y.z, y.w);
      reinterpret_cast<uint2*>(zb + row)[c4] = p;
    }
  }
}

// Wide path (gemm_wide.cuh): the last layer leaves the PRE-LayerNorm sum y with its row statistics; the mean over a
// sequence's tokens of LN(y_t) is gamma * mean_t((y_t - mu_t) * rstd_t) + beta, so the LayerNorm costs one FMA per
// element on the way into the pooling sum. Then the same masked mean + L2 as k_pool_l2.
template <int NV>
__global__ void __launch_bounds__(128) k_pool_ln_l2(const __nv_bfloat16* __restrict__ yh, const __nv_bfloat16* __restrict__ yl,
                                                    const float* __restrict__ stats, int slots,
                                                    long long stat_stride, const float* __restrict__ gamma,
                                                    const float* __restrict__ beta,
                                                    float eps, const int32_t* __restrict__ cu, int seq0, int n_seq, int tok0,
                                                    int H, float* __restrict__ z, __nv_bfloat16* __restrict__ zb) {
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= n_seq) return;
  const int beg = cu[seq0 + s] - tok0, end = cu[seq0 + s + 1] - tok0;
  const float inv_h = 1.0f / static_cast<float>(H);
  float4 acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = beg; t < end; ++t) {
    float2 st = reinterpret_cast<const float2*>(stats)[t];
    for (int sl = 1; sl < slots; ++sl) {   // partial statistics, summed in slot order (gemm_wide.cuh)
      const float2 p = reinterpret_cast<const float2*>(stats)[sl * stat_stride + t];
      st.x += p.x;
      st.y += p.y;
    }
    const float mu = st.x * inv_h;
    const float ra = rsqrtf(fmaxf(st.y * inv_h - mu * mu, 0.f) + eps);
    const float rb = -mu * ra;
    const uint2* h2 = reinterpret_cast<const uint2*>(yh + static_cast<size_t>(t) * H);
    const uint2* l2 = reinterpret_cast<const uint2*>(yl + static_cast<size_t>(t) * H);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const uint2 hh = h2[lane + 32 * i], ll = l2[lane + 32 * i];
      float4 v;
      join_bf16x2(hh.x, ll.x, v.x, v.y);
      join_bf16x2(hh.y, ll.y, v.z, v.w);
      acc[i].x += fmaf(v.x, ra, rb); acc[i].y += fmaf(v.y, ra, rb); acc[i].z += fmaf(v.z, ra, rb); acc[i].w += fmaf(v.w, ra, rb);
    }
  }
  const float n_tok = static_cast<float>(end - beg);
  const float cnt = fmaxf(n_tok, 1e-9f);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c4 = lane + 32 * i;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
    // sum_t LN(y_t) = gamma * sum_t yhat_t + n_tok * beta  (an empty sequence sums to zero -> 0/0 = NaN below, as the reference)
    acc[i].x = fmaf(acc[i].x, g.x, n_tok * b.x) / cnt; acc[i].y = fmaf(acc[i].y, g.y, n_tok * b.y) / cnt;
    acc[i].z = fmaf(acc[i].z, g.z, n_tok * b.z) / cnt; acc[i].w = fmaf(acc[i].w, g.w, n_tok * b.w) / cnt;
    ss += (acc[i].x * acc[i].x + acc[i].y * acc[i].y) + (acc[i].z * acc[i].z + acc[i].w * acc[i].w);
  }
  const float nrm = sqrtf(warp_sum(ss));
  const size_t row = static_cast<size_t>(seq0 + s) * H;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c4 = lane + 32 * i;
    const float4 o = make_float4(acc[i].x / nrm, acc[i].y / nrm, acc[i].z / nrm, acc[i].w / nrm);
    reinterpret_cast<float4*>(z + row)[c4] = o;
    if (zb != nullptr) {
      uint2 p;
      p.x = pack_bf16x2(o.x, o.y);
      p.y = pack_bf16x2(o.z, o.w);
      reinterpret_cast<uint2*>(zb + row)[c4] = p;
    }
  }
}

__device__ __forceinline__ float warp_dot(const float* __restrict__ a, const float* __restrict__ b, int dim, int lane) {
  float acc = 0.f;
  const int n4 = dim >> 2;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  for (int i = lane; i < n4; i += 32) {
    const float4 x = a4[i], y = b4[i];
    acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc); acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
  }
  for (int i = (n4 << 2) + lane; i < dim; i += 32) acc = fmaf(a[i], b[i], acc);
  return warp_sum(acc);
}

// out[i] = <a[a_index ? a_index[i] : i], b[b_index ? b_index[i] : i]>
// (_group.py:541/544 torch.bmm on (N,1,D)x(N,D,1))
// An index outside its table is clamped and reported through `status` bit 1 (the call then fails at its next sync)
// instead of reading out of bounds.
__global__ void __launch_bounds__(256) k_pair_dot(const float* __restrict__ a, const int32_t* __restrict__ a_index, long long na,
                                                  const float* __restrict__ b, const int32_t* __restrict__ b_index, long long nb,
                                                  long long n, int dim, float* __restrict__ out, int* __restrict__ status) {
  const int lane = threadIdx.x & 31;
  const long long i = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  long long ia = a_index ? a_index[i] : i;
  long long ib = b_index ? b_index[i] : i;
  if (ia < 0 || ia >= na || ib < 0 || ib >= nb) {
    if (lane == 0) atomicOr(status, 2);
    ia = min(max(ia, 0LL), na - 1);
    ib = min(max(ib, 0LL), nb - 1);
  }
  const float s = warp_dot(a + ia * dim, b + ib * dim, dim, lane);
  if (lane == 0) out[i] = s;
}

__global__ void __launch_bounds__(256) k_f32_to_bf16(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                     long long n) {
  const long long i4 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i4 + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(in + i4);
    uint2 p;
    p.x = pack_bf16x2(v.x, v.y);
    p.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(out + i4) = p;
  } else {
    for (long long i = i4; i < n; ++i) out[i] = __float2bfloat16_rn(in[i]);
  }
}

// (score desc, index asc) ordering used everywhere a top-k is finalised.
__device__ __forceinline__ bool better(float s1, int i1, float s2, int i2) {
  return (s1 > s2) || (s1 == s2 && i1 < i2);
}

// Exact fp32 re-scoring of the candidates produced by the bf16 top-k epilogue, final top-k, and the
// paired target score. Warp per prediction row. cand_idx: [splits, n, KC]; entries < 0 are empty.
__global__ void __launch_bounds__(128) k_rescore(const float* __restrict__ zp, const float* __restrict__ zc, long long n,
                                                 int dim, const int* __restrict__ cand_idx, int splits, int KC, int k,
                                                 const int32_t* __restrict__ targets, int n_classes, float* __restrict__ out_scores,
                                                 int32_t* __restrict__ out_idx, float* __restrict__ out_paired,
                                                 int* __restrict__ status) {
  extern __shared__ float smem_f[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncand = splits * KC;
  float* cs = smem_f + static_cast<size_t>(warp) * 2 * ncand;
  int* ci = reinterpret_cast<int*>(cs + ncand);
  const long long r = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + warp;
  if (r >= n) return;
  // Four candidates at a time: 8 lanes per candidate, each lane a strided eighth of the row, three shuffles per dot
  // (a whole warp per candidate spent 5 shuffles and a serialised store on every one of the 16 dots of a row).
  const float4* zr4 = reinterpret_cast<const float4*>(zp + r * dim);
  const int n4 = dim >> 2;                  // dim % 4 == 0 is checked by the callers
  const int sub = lane & 7, grp = lane >> 3;
  for (int c0 = 0; c0 < ncand; c0 += 4) {
    const int c = c0 + grp;
    int id = -1;
    if (c < ncand) id = cand_idx[(static_cast<long long>(c / KC) * n + r) * KC + (c % KC)];
    float acc = 0.f;
    if (id >= 0) {
      const float4* zc4 = reinterpret_cast<const float4*>(zc + static_cast<long long>(id) * dim);
      for (int i = sub; i < n4; i += 8) {
        const float4 x = zr4[i], y = __ldg(zc4 + i);
        acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc); acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (sub == 0 && c < ncand) { cs[c] = id >= 0 ? acc : -INFINITY; ci[c] = id; }
  }
  if (targets != nullptr && out_paired != nullptr) {
    int tg = targets[r];
    if (tg < 0 || tg >= n_classes) {   // clamped and reported (status bit 1): never an out-of-bounds read
      if (lane == 0) atomicOr(status, 2);
      tg = min(max(tg, 0), n_classes - 1);
    }
    const float s = warp_dot(zp + r * dim, zc + static_cast<long long>(tg) * dim, dim, lane);
    if (lane == 0) out_paired[r] = s;
  }
  __syncwarp();
  // rank of every candidate among the valid ones under (score desc, index asc); ranks < k are the result rows
  int n_valid = 0;
  for (int c = lane; c < ncand; c += 32) {
    const int id = ci[c];
    if (id < 0) continue;
    const float sc = cs[c];
    int rank = 0;
    for (int o = 0; o < ncand; ++o) {
      const int oid = ci[o];
      if (oid >= 0 && o != c && better(cs[o], oid, sc, id)) ++rank;
    }
    if (rank < k) { out_scores[r * k + rank] = sc; out_idx[r * k + rank] = id; }
  }
  for (int c = lane; c < ncand; c += 32) n_valid += ci[c] >= 0 ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_valid += __shfl_xor_sync(0xffffffffu, n_valid, o);
  for (int o = n_valid + lane; o < k; o += 32) { out_scores[r * k + o] = -INFINITY; out_idx[r * k + o] = -1; }
}

// Check mode: top-k of materialised fp32 score rows s[rows, C]. Warp per row; each lane keeps the
// best KL of its strided columns, lane 0 merges. (score desc, index asc).
template <int KL>
__global__ void __launch_bounds__(128) k_topk_rows(const float* __restrict__ s, long long rows, int C, int k,
                                                   float* __restrict__ out_scores, int32_t* __restrict__ out_idx) {
  extern __shared__ float smem_f[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* ms = smem_f + static_cast<size_t>(warp) * 2 * 32 * KL;
  int* mi = reinterpret_cast<int*>(ms + 32 * KL);
  const long long r = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + warp;
  if (r >= rows) return;
  float ls[KL];
  int li[KL];
#pragma unroll
  for (int i = 0; i < KL; ++i) { ls[i] = -INFINITY; li[i] = -1; }
  const float* sr = s + r * C;
  for (int c = lane; c < C; c += 32) {
    const float v = sr[c];
    if (v > ls[KL - 1]) {
      float cv = v;
      int cc = c;
#pragma unroll
      for (int i = 0; i < KL; ++i) {
        if (cv > ls[i]) { const float ts = ls[i]; const int ti = li[i]; ls[i] = cv; li[i] = cc; cv = ts; cc = ti; }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < KL; ++i) { ms[lane * KL + i] = ls[i]; mi[lane * KL + i] = li[i]; }
  __syncwarp();
  if (lane == 0) {
    for (int o = 0; o < k; ++o) {
      int best = -1;
      for (int c = 0; c < 32 * KL; ++c) {
        if (mi[c] < 0) continue;
        if (best < 0 || better(ms[c], mi[c], ms[best], mi[best])) best = c;
      }
      if (best >= 0) {
        out_scores[r * k + o] = ms[best];
        out_idx[r * k + o] = mi[best];
        mi[best] = -1;
      } else {
        out_scores[r * k + o] = -INFINITY;
        out_idx[r * k + o] = -1;
      }
    }
  }
}

// ---- fp32 SIMT contraction (check mode): C = A[M,K] . W[N,K]^T with the same fused epilogues ----
enum : int { EPF_BIAS = 0, EPF_BIAS_GELU = 1, EPF_BIAS_RES = 2 };

template <int EPF>
__global__ void __launch_bounds__(256) k_gemm_f32(const float* __restrict__ A, const float* __restrict__ W,
                                                  const float* __restrict__ bias, const float* __restrict__ residual,
                                                  float* __restrict__ out, int M, int N, int K, int ldo) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int lr = threadIdx.x >> 2;        // 0..63: tile row loaded by this thread
  const int lk = (threadIdx.x & 3) * 4;   // 0,4,8,12
  for (int k0 = 0; k0 < K; k0 += BK) {
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f), wv = av;
    if (m0 + lr < M) av = *reinterpret_cast<const float4*>(A + static_cast<size_t>(m0 + lr) * K + k0 + lk);
    if (n0 + lr < N) wv = *reinterpret_cast<const float4*>(W + static_cast<size_t>(n0 + lr) * K + k0 + lk);
    As[lk + 0][lr] = av.x; As[lk + 1][lr] = av.y; As[lk + 2][lr] = av.z; As[lk + 3][lr] = av.w;
    Ws[lk + 0][lr] = wv.x; Ws[lk + 1][lr] = wv.y; Ws[lk + 2][lr] = wv.z; Ws[lk + 3][lr] = wv.w;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 w4 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      const float ar[4] = {a4.x, a4.y, a4.z, a4.w};
      const float wr[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if constexpr (EPF == EPF_BIAS_GELU) v = gelu_erf(v);
      if constexpr (EPF == EPF_BIAS_RES) v += residual[static_cast<size_t>(m) * ldo + n];
      out[static_cast<size_t>(m) * ldo + n] = v;
    }
  }
}

}  // namespace owc
