// api.cu - the C ABI declared in include/owc_b200.h: context, encoder, and the orchestration of
// the kernels in kernels.cuh / gemm_tc.cuh on one CUDA stream. Host code only decides shapes,
// chunking and launch order; all arithmetic happens in the kernels. No CPU fallback exists.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <vector>

#include "../../include/owc_b200.h"
#include "../../include/owc_b200_debug.h"
#include "attn_mma.cuh"
#include "enc_fused.cuh"
#include "gemm_tc.cuh"
#include "gemm_wide.cuh"
#include "qkv_attn.cuh"
#include "kernels.cuh"

using namespace owc;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

#define CU_OK(expr)                                                                              \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define OWC_TRY(expr)           \
  do {                          \
    int _r = (expr);            \
    if (_r != 0) return _r;     \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// ------------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------------
struct Buf {
  void* p = nullptr;
  size_t cap = 0;
};

struct owc_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int chunk_tokens = 0;
  bool chunk_default = true;   // chunk_tokens was not given: encoders pick their own multiple of one block per SM
  // tcgen05 kernel variant: 0 = one CTA per 128 rows, 1 = CTA pairs (cta_group::2), 2 = CTA pairs + resident A.
  // Measured on B200: the encoder's short sweeps (2-8 column tiles per row block) run fastest on the 1-CTA
  // kernel, the similarity sweep (hundreds of tiles) on pairs with the A block resident.
  int gemm_mode = -1;      // encoder contractions (env OWC_GEMM_MODE); -1 = by width: 1-CTA for H < 768, CTA pairs from H = 768
  int sim_mode = 2;        // similarity + top-k     (env OWC_SIM_MODE)
  int fuse_ln = 1;         // fuse residual+LayerNorm into the O-proj / FFN2 epilogue (env OWC_FUSE_LN)
  int ares = 1;            // 1-CTA kernel keeps the A block resident in shared memory when K <= 384 (env OWC_ARES)
  int pdl = 1;             // programmatic dependent launch between the kernels of a layer (env OWC_PDL)
  int fused = 1;           // block-resident fused encoder for MiniLM-width BERT, sequences <= 64 tokens (env OWC_FUSED)
  int enc_wide = 1;        // 256-column tiles for wide encoders (env OWC_ENC_WIDE=0: kEncBN tiles everywhere)
  int fused_waves = 48;    // row-block pairs per cluster and launch of the fused encoder (env OWC_FUSED_WAVES): one launch per 100 k-row step
  int fused_dbg = 0;       // record the fused kernel's timeline stamps (env OWC_FUSED_DBG)
  int wide = 1;            // persistent CTA-pair GEMMs with deferred LayerNorm for H >= 768 (gemm_wide.cuh; env OWC_WIDE)
  int attn_grp = 1;        // short sequences share 16 x 16 attention tiles on the wide path (k_attn_grp; env OWC_ATTN_GRP)
  int wide_deep = 1;       // FFN2 on the 5-stage / single-residual-buffer variant of the residual epilogue (env OWC_WIDE_DEEP)
  int wide_rot = 0;        // clusters start their column-tile sweeps at different tiles (env OWC_WIDE_ROT=1; measured: QKV +5 %,
                           // the whole encode -1.5 % - the step is power-capped, not bound by any one kernel)
  int wide_ffn2_split = 1; // FFN2: the three column tiles of a row-block pair on three neighbouring clusters (env OWC_WIDE_FFN2_SPLIT)
  int qkv_attn = 1;        // wide path: attention in the QKV projection's epilogue for sequences <= 64 tokens (env OWC_QKV_ATTN)
  int wide_waves = 8;      // row-block pairs per cluster and chunk on that path (env OWC_WIDE_WAVES)
  std::map<const void*, size_t> smem_attr;   // kernels whose dynamic shared-memory opt-in was set on THIS device
  long long* wide_dbg = nullptr;             // [48] cycle counters of the instrumented wide GEMM (env OWC_WIDE_DBG=1), else nullptr
  int sm_count = 0;
  EncodeTiledFn encode_tiled = nullptr;
  owc_stats stats{};
  int* d_status = nullptr;
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;   // pairs: start, stop
  std::vector<int> prof_cats;
  size_t prof_used = 0;
  // per-call scratch (grown on demand, reused across calls)
  Buf ids, cu, targets, zp, zpb, zcb, zc32, cand_s, cand_i, out_s, out_i, out_p, tmp_a, tmp_b, tmp_c, tmp_d, blk, grp;
};

struct LayerW {
  const float *bqkv, *bo, *ln1g, *ln1b, *b1, *b2, *ln2g, *ln2b;   // fp32 (device)
  const float *wqkv32, *wo32, *w132, *w232;                        // fp32 masters (check mode)
  const __nv_bfloat16 *wqkv, *wo, *w1, *w2;                        // bf16 operand copies
  CUtensorMap tm_wqkv, tm_wo, tm_w1, tm_w2;          // W maps, box = kEncBN rows (1-CTA kernel)
  CUtensorMap tm2_wqkv, tm2_wo, tm2_w1, tm2_w2;      // W maps, box = kEncBN/2 rows (CTA-pair kernel)
  CUtensorMap tmw_wqkv, tmw_wo, tmw_w1, tmw_w2;      // W maps, box = 128 rows (CTA-pair kernel, 256-column tiles; wide encoders)
  // deferred-LayerNorm path (gemm_wide.cuh): gamma-folded bf16 copies of Wqkv / W1 and their epilogue vectors
  const __nv_bfloat16 *fqkv = nullptr, *f1 = nullptr;
  const float *c1q = nullptr, *c2q = nullptr, *c1f = nullptr, *c2f = nullptr;
  CUtensorMap tmf_wqkv, tmf_w1;                      // box = 128 rows
  // QKV projection with the attention in its epilogue (qkv_attn.cuh): the folded Wqkv with its rows regrouped per head
  const __nv_bfloat16* fqkv_h = nullptr;
  const float *c1qh = nullptr, *c2qh = nullptr;
  CUtensorMap tmh_wqkv;                              // box = 96 rows
};

struct owc_encoder {
  owc_ctx* ctx = nullptr;
  owc_arch arch{};
  float* w32 = nullptr;            // the whole blob
  __nv_bfloat16* w16 = nullptr;    // bf16 copies of the four GEMM weights of every layer
  const float *word = nullptr, *pos = nullptr, *type0 = nullptr, *emb_g = nullptr, *emb_b = nullptr, *rel_lut = nullptr;
  std::vector<LayerW> layers;
  int cap = 0;                     // workspace rows (tokens per chunk)
  float *x = nullptr, *y = nullptr;
  __nv_bfloat16 *xb = nullptr, *qkv = nullptr, *ao = nullptr, *f = nullptr;
  __nv_bfloat16* xl = nullptr;               // wide path: low half of the residual stream's bf16 pair (xb is the high half)
  float *qkv32 = nullptr, *ao32 = nullptr, *f32 = nullptr;   // check mode, allocated on first use
  CUtensorMap tm_xb, tm_ao, tm_f;            // A-operand maps (box 128 rows x 64)
  CUtensorMap tmo_qkv, tmo_f, tmo_y;         // epilogue TMA-store maps (box 32 x 32)
  CUtensorMap tmo_x, tmo_xb;                 // fused-LN epilogue: residual stream (fp32, load + store) and its bf16 copy
  CUtensorMap tmo_xl;                        // wide path: xl (load + store, as tmo_xb)
  // block-resident fused encoder (enc_fused.cuh): layer-stacked bf16 weights, QKV rows regrouped per head
  bool fused_ok = false;
  __nv_bfloat16 *fw_qkv = nullptr, *fw_o = nullptr, *fw_1 = nullptr, *fw_2 = nullptr;
  CUtensorMap tmf_qkv, tmf_o, tmf_1, tmf_2;
  bool wide = false;               // every GEMM width is a multiple of 256 (mpnet-base: 2304 / 768 / 3072): 256-column tiles
  bool wide_ln = false;            // the deferred-LayerNorm path is set up (wide && ctx->wide)
  bool qkv_attn = false;           // ... and the fused QKV + attention kernel (head_dim 64)
  __nv_bfloat16* wfold16h = nullptr;  // per-head regrouped folded Wqkv of every layer
  float* wfold32h = nullptr;          // c1qh | c2qh of every layer
  __nv_bfloat16* wfold16 = nullptr;   // gamma-folded Wqkv | W1 of every layer
  float* wfold32 = nullptr;           // c1q | c2q | c1f | c2f of every layer
  float *st1 = nullptr, *st2 = nullptr;   // [slots][cap][2] partial row (sum, sumsq) after the attention block / the feed-forward block
  int fcap = 0;                    // tokens per fused launch
  float* fx = nullptr;             // [fcap + 128, H] fp32 residual stream of a fused chunk
  long long* fdbg = nullptr;       // [1024] timeline stamps + event traces
};

constexpr int kWideStatSlots = 2 * 4;   // partial row statistics: (column split, chunk parity) slots; H <= 1024 = 4 tiles of 256
constexpr int kEncBN = 192;   // divides 3H, H and I of both architectures (1152/384/1536, 2304/768/3072)

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
static int ensure(owc_ctx* c, Buf& b, size_t bytes) {
  if (bytes <= b.cap) return 0;
  CU_OK(cudaStreamSynchronize(c->stream));  // earlier launches may still read the old buffer
  if (b.p) CU_OK(cudaFree(b.p));
  b.p = nullptr;
  b.cap = 0;
  size_t want = bytes + bytes / 4 + 256;
  CU_OK(cudaMalloc(&b.p, want));
  b.cap = want;
  return 0;
}

static int to_device(owc_ctx* c, Buf& b, const void* src, size_t bytes, int space, const void** dptr) {
  if (space == OWC_MEM_DEVICE) { *dptr = src; return 0; }
  OWC_TRY(ensure(c, b, bytes));
  CU_OK(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, c->stream));
  c->stats.h2d_bytes += static_cast<int64_t>(bytes);
  *dptr = b.p;
  return 0;
}

static int from_device(owc_ctx* c, void* dst, const void* src, size_t bytes, int space) {
  if (space == OWC_MEM_DEVICE) {
    if (dst != src) CU_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, c->stream));
    return 0;
  }
  CU_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  c->stats.d2h_bytes += static_cast<int64_t>(bytes);
  return 0;
}

// Brackets one kernel launch with CUDA events while profiling is on (owc_ctx_profile_begin/end).
struct ProfScope {
  owc_ctx* c;
  cudaEvent_t stop = nullptr;
  ProfScope(owc_ctx* ctx, int cat, int cat2 = -1) : c(ctx) {
    if (!c->profiling) return;
    if (c->prof_used + 2 > c->prof_events.size()) {
      cudaEvent_t a, b;
      if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
      c->prof_events.push_back(a);
      c->prof_events.push_back(b);
    }
    cudaEventRecord(c->prof_events[c->prof_used], c->stream);
    stop = c->prof_events[c->prof_used + 1];
    c->prof_used += 2;
    c->prof_cats.push_back(cat | ((cat2 < 0 ? 0xff : cat2) << 8));
  }
  ~ProfScope() {
    if (stop) cudaEventRecord(stop, c->stream);
  }
};

static int check_launch(owc_ctx* c, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("launch of %s failed: %s", what, cudaGetErrorString(e));
  c->stats.kernel_launches += 1;
  return 0;
}

// K-major bf16 matrix [rows, cols] -> 2-D tiled tensor map, box = 64 cols x box_rows, SWIZZLE_128B.
static int make_tmap(owc_ctx* c, CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail("tensor map base not 16-byte aligned");
  if (cols % 64 != 0) return fail("tensor map inner dimension %llu is not a multiple of 64", (unsigned long long)cols);
  if (rows == 0) rows = 1;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = c->encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return 0;
}

// Output matrix [rows, cols] (bf16 or fp32, row-major) -> tensor map for the epilogue's TMA stores:
// box = 32 x 32 elements; bf16 rows are 64 B (SWIZZLE_64B), fp32 rows 128 B (SWIZZLE_128B), matching
// the staging layout written by stage_chunk() in gemm_tc.cuh.
static int make_out_tmap(owc_ctx* c, CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, bool is_bf16) {
  const uint64_t esz = is_bf16 ? 2 : 4;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail("output tensor map base not 16-byte aligned");
  if ((cols * esz) % 16 != 0) return fail("output row pitch (%llu B) is not a multiple of 16", (unsigned long long)(cols * esz));
  if (rows == 0) rows = 1;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * esz};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = c->encode_tiled(m, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                               const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               is_bf16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (output) failed with CUresult %d", static_cast<int>(r));
  return 0;
}

// The dynamic shared-memory opt-in is a per-device (per-context) function attribute, and one process may hold contexts
// on several GPUs: track it per owc_ctx, not in a process-wide static.
template <typename F>
static int ensure_smem_attr(owc_ctx* c, F* func, size_t bytes) {
  if (bytes <= 48 * 1024) return 0;
  const void* key = reinterpret_cast<const void*>(func);
  auto it = c->smem_attr.find(key);
  if (it != c->smem_attr.end() && it->second >= bytes) return 0;
  CU_OK(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
  c->smem_attr[key] = bytes;
  return 0;
}

static int check_gemm_shape(const GemmArgs& g, int epi) {
  if (g.K % 64 != 0 || g.K <= 0) return fail("tcgen05 GEMM needs K %% 64 == 0 (K=%d)", g.K);
  if (epi == EPI_BIAS_RES_F32 && g.N % 32 != 0) return fail("residual epilogue needs N %% 32 == 0 (N=%d)", g.N);
  return 0;
}

template <int BN, int EPI, int KC, bool ARES>
static int launch_gemm_tc_impl(owc_ctx* c, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, GemmArgs g,
                               int splits, const CUtensorMap* td) {
  using Cfg = GemmCfg<BN, EPI, ARES>;
  OWC_TRY(ensure_smem_attr(c, k_gemm_tc<BN, EPI, KC, ARES>, Cfg::SMEM_BYTES));
  OWC_TRY(check_gemm_shape(g, EPI));
  const int tiles = (g.N + BN - 1) / BN;
  splits = std::max(1, std::min(splits, tiles));
  g.tiles_per_cta = (tiles + splits - 1) / splits;
  const int gy = (tiles + g.tiles_per_cta - 1) / g.tiles_per_cta;
  dim3 grid((g.M + 127) / 128, gy);
  ProfScope prof(c, EPI == EPI_TOPK ? OWC_KCAT_SIM : OWC_KCAT_GEMM);
  if (EPI == EPI_RES_LN && (g.N != 2 * BN || gy != 1 || !td || !g.gamma || !g.beta))
    return fail("fused LayerNorm epilogue needs N == %d and whole rows per CTA", 2 * BN);
  // Programmatic dependent launch: this kernel's CTAs may be scheduled (and run their prologue: TMEM
  // allocation, barrier init, descriptor prefetch) while the previous kernel in the stream drains;
  // griddepcontrol.wait inside the kernel orders every dependent global access.
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = c->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = c->pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CU_OK(cudaLaunchKernelEx(&cfg, k_gemm_tc<BN, EPI, KC, ARES>, ta, tb, tc, td ? *td : tc, g));
  return check_launch(c, "k_gemm_tc");
}

// 1-CTA kernel; the A block stays resident in shared memory when it fits (K <= 384) unless OWC_ARES=0.
template <int BN, int EPI, int KC>
static int launch_gemm_tc(owc_ctx* c, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, GemmArgs g,
                          int splits, const CUtensorMap* td = nullptr) {
  // only where at least three W stages still fit next to the 96 KB resident block
  constexpr int ares_stages = (kSmemLimit - epi_fixed_bytes(BN, EPI) - 6 * 128 * 64 * 2) / (BN * 64 * 2);
  if constexpr (ares_stages >= 3) {
    if (c->ares && g.K <= kAresMaxK) return launch_gemm_tc_impl<BN, EPI, KC, true>(c, ta, tb, tc, g, splits, td);
  }
  return launch_gemm_tc_impl<BN, EPI, KC, false>(c, ta, tb, tc, g, splits, td);
}

// CTA-pair kernel: `tb_half` is the W tensor map with a BN/2-row box (each CTA stages half a tile).
template <int BN, int EPI, int KC, bool ARES>
static int launch_gemm_tc2(owc_ctx* c, const CUtensorMap& ta, const CUtensorMap& tb_half, const CUtensorMap& tc, GemmArgs g,
                           int splits) {
  using Cfg = Gemm2Cfg<BN, EPI, ARES>;
  OWC_TRY(ensure_smem_attr(c, k_gemm_tc2<BN, EPI, KC, ARES>, Cfg::SMEM_BYTES));
  OWC_TRY(check_gemm_shape(g, EPI));
  if (ARES && g.K > kAresMaxK) return fail("resident-A GEMM needs K <= %d (K=%d)", kAresMaxK, g.K);
  const int tiles = (g.N + BN - 1) / BN;
  splits = std::max(1, std::min(splits, tiles));
  g.tiles_per_cta = (tiles + splits - 1) / splits;
  const int gy = (tiles + g.tiles_per_cta - 1) / g.tiles_per_cta;
  const int pairs = (g.M + 255) / 256;
  dim3 grid(2 * pairs, gy);   // the odd tail CTA of a pair only sees out-of-range rows (TMA zero-fills / clips)
  ProfScope prof(c, EPI == EPI_TOPK ? OWC_KCAT_SIM : OWC_KCAT_GEMM);
  k_gemm_tc2<BN, EPI, KC, ARES><<<grid, kGemmThreads, Cfg::SMEM_BYTES, c->stream>>>(ta, tb_half, tc, g);
  return check_launch(c, "k_gemm_tc2");
}

// column splits so that (row blocks x splits) covers the SMs when there are few row blocks
static int pick_splits(const owc_ctx* c, int m, int n, int bn) {
  const int mb = (m + 127) / 128, tiles = (n + bn - 1) / bn;
  if (mb >= c->sm_count) return 1;
  return std::max(1, std::min(tiles, (c->sm_count + mb - 1) / mb));
}

// Mode-dispatched contraction. gemm_mode: 0 = one CTA per 128 rows; 1 = CTA pairs (cta_group::2);
// 2 = CTA pairs with the A block resident in shared memory whenever K <= 384 (default).
template <int BN, int EPI, int KC>
static int launch_gemm(owc_ctx* c, const CUtensorMap& ta, const CUtensorMap& tb_full, const CUtensorMap& tb_half,
                       const CUtensorMap& tc, const GemmArgs& g) {
  const int splits = pick_splits(c, g.M, g.N, BN);
  // measured on B200: the short sweeps of a 384-wide encoder (2-8 column tiles per row block) run fastest on the 1-CTA
  // kernel; from H = 768 (12-16 tile sweeps, K up to 3072) the halved weight traffic of CTA pairs wins (+7 % GEMM rate)
  const int mode = c->gemm_mode >= 0 ? c->gemm_mode : ((g.K >= 768 || g.N >= 2304) ? 1 : 0);
  if (mode == 0) return launch_gemm_tc<BN, EPI, KC>(c, ta, tb_full, tc, g, splits);
  if (mode == 2 && g.K <= kAresMaxK) return launch_gemm_tc2<BN, EPI, KC, true>(c, ta, tb_half, tc, g, splits);
  return launch_gemm_tc2<BN, EPI, KC, false>(c, ta, tb_half, tc, g, splits);
}

// Encoder contraction: 256-column tiles on CTA pairs where the encoder is wide enough for pairs to win and every width
// divides (UMMA N = 256 is the shape whose operand fetch keeps up with the tensor pipe; measured on the mpnet-base
// shapes against N = 192: +2.4 % QKV, +1 % O, +2.6 % FFN1, +11.6 % FFN2), the kEncBN kernels otherwise.
template <int EPI>
static int launch_enc_gemm(owc_ctx* c, bool wide, const CUtensorMap& ta, const CUtensorMap& tb_full, const CUtensorMap& tb_half,
                           const CUtensorMap& tb_wide, const CUtensorMap& tc, const GemmArgs& g) {
  const int mode = c->gemm_mode >= 0 ? c->gemm_mode : ((g.K >= 768 || g.N >= 2304) ? 1 : 0);
  if (wide && c->enc_wide && mode != 0)
    return launch_gemm_tc2<256, EPI, 1, false>(c, ta, tb_wide, tc, g, pick_splits(c, g.M, g.N, 256));
  return launch_gemm<kEncBN, EPI, 1>(c, ta, tb_full, tb_half, tc, g);
}

// Persistent CTA-pair contraction of the wide encoder path (gemm_wide.cuh). `tb` is the W map with a 128-row box.
// column splits of the persistent kernel for an M-row problem with `tiles` 256-column tiles (1 once every cluster has a row-block pair)
static int wide_splits(const owc_ctx* c, int M, int tiles) {
  const int clusters_max = std::max(1, c->sm_count / 2);
  const int n_pairs = (M + 255) / 256;
  const int splits = n_pairs >= clusters_max ? 1 : std::min(tiles, (clusters_max + n_pairs - 1) / n_pairs);
  const int tpi = (tiles + splits - 1) / splits;
  return (tiles + tpi - 1) / tpi;
}

// FFN2 (K = 3072, three column tiles): with one row-block pair per cluster the 74 clusters hold 74 x 1.5 MB of A blocks in
// flight - more than the L2 keeps next to everything else - and every column tile re-reads its A block from DRAM. Splitting
// the three tiles of a pair over three NEIGHBOURING clusters (item = pair x tile) puts a third as many A blocks in flight,
// read from DRAM once and from L2 twice.
static int ffn2_splits(const owc_ctx* c, int M, int tiles) {
  const int s = wide_splits(c, M, tiles);
  return c->wide_ffn2_split ? tiles : s;
}

template <int EPI>
static int launch_gemm_wide(owc_ctx* c, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& td,
                            WideArgs g, int shape_cat, int force_splits = 0) {
  OWC_TRY(ensure_smem_attr(c, k_gemm_wide<EPI, false>, WideCfg<EPI>::SMEM_BYTES));
  if (c->wide_dbg) OWC_TRY(ensure_smem_attr(c, k_gemm_wide<EPI, true>, WideCfg<EPI>::SMEM_BYTES));
  if (g.K % 64 != 0 || g.K <= 0 || g.N % 256 != 0 || g.N <= 0 || g.M <= 0) return fail("wide GEMM: bad shape %d x %d x %d", g.M, g.N, g.K);
  const int clusters_max = std::max(1, c->sm_count / 2);
  const int tiles = g.N / 256;
  g.n_pairs = (g.M + 255) / 256;
  g.splits = force_splits > 0 ? std::min(force_splits, tiles) : wide_splits(c, g.M, tiles);
  g.tiles_per_item = (tiles + g.splits - 1) / g.splits;
  const int clusters = std::min(g.n_pairs * g.splits, clusters_max);
  g.dbg = c->wide_dbg;
  g.rotate = c->wide_rot;
  if (c->wide_dbg) {
    if (const char* ex = getenv("OWC_WIDE_EXP")) g.exp = atoi(ex);   // instrumented build only (wrong results by construction)
  }
  ProfScope prof(c, OWC_KCAT_GEMM, shape_cat);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = WideCfg<EPI>::SMEM_BYTES;
  cfg.stream = c->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = c->pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (c->wide_dbg) CU_OK(cudaLaunchKernelEx(&cfg, k_gemm_wide<EPI, true>, ta, tb, tc, td, g));
  else CU_OK(cudaLaunchKernelEx(&cfg, k_gemm_wide<EPI, false>, ta, tb, tc, td, g));
  return check_launch(c, "k_gemm_wide");
}

// QKV projection + attention of one layer on a chunk (qkv_attn.cuh). d_blk: the chunk's row-block table on the device.
static int launch_qkv_attn(owc_encoder* e, const LayerW& L, QkvAttnArgs g, bool long_units) {
  owc_ctx* c = e->ctx;
  const bool dbg = c->wide_dbg != nullptr;
  void (*kern)(const CUtensorMap, const CUtensorMap, const QkvAttnArgs) =
      dbg ? (long_units ? k_qkv_attn<true, true> : k_qkv_attn<true, false>) : (long_units ? k_qkv_attn<false, true> : k_qkv_attn<false, false>);
  OWC_TRY(ensure_smem_attr(c, kern, qa::SMEM_BYTES));
  g.dbg = dbg ? c->wide_dbg + 32 : nullptr;
  if (dbg) {
    if (const char* ex = getenv("OWC_WIDE_EXP")) g.exp = atoi(ex);   // instrumented build only (wrong results by construction)
  }
  g.burst = -3;  // at most three k-blocks in the tensor pipe; same box, ms per 300 k rows: 0 (no limit) 90.6, pairs (2) 94.2, -2 85.1, -3 84.4
  if (const char* bu = getenv("OWC_QKV_BURST")) g.burst = atoi(bu);
  const int n_items = (g.n_blk + 1) / 2;
  const int clusters = std::max(1, std::min(n_items, c->sm_count / 2));
  ProfScope prof(c, OWC_KCAT_GEMM, OWC_KCAT_GEMM_QKV);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(qa::THREADS);
  cfg.dynamicSmemBytes = qa::SMEM_BYTES;
  cfg.stream = c->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = c->pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CU_OK(cudaLaunchKernelEx(&cfg, kern, e->tm_xb, L.tmh_wqkv, g));
  return check_launch(c, "k_qkv_attn");
}

template <int EPF>
static int launch_gemm_f32(owc_ctx* c, const float* A, const float* W, const float* bias, const float* res, float* out,
                           int M, int N, int K, int ldo) {
  if (K % 16 != 0) return fail("fp32 GEMM needs K %% 16 == 0 (K=%d)", K);
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  ProfScope prof(c, OWC_KCAT_GEMM);
  k_gemm_f32<EPF><<<grid, 256, 0, c->stream>>>(A, W, bias, res, out, M, N, K, ldo);
  return check_launch(c, "k_gemm_f32");
}

static int launch_convert(owc_ctx* c, const float* in, __nv_bfloat16* out, long long n) {
  if (n <= 0) return 0;
  const long long threads = (n + 3) / 4;
  ProfScope prof(c, OWC_KCAT_MISC);
  k_f32_to_bf16<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, c->stream>>>(in, out, n);
  return check_launch(c, "k_f32_to_bf16");
}

#define DISPATCH_NV(H, ...)                                           \
  switch ((H) / 128) {                                                \
    case 2: { constexpr int NV = 2; __VA_ARGS__; } break;             \
    case 3: { constexpr int NV = 3; __VA_ARGS__; } break;             \
    case 4: { constexpr int NV = 4; __VA_ARGS__; } break;             \
    case 6: { constexpr int NV = 6; __VA_ARGS__; } break;             \
    case 8: { constexpr int NV = 8; __VA_ARGS__; } break;             \
    default: return fail("hidden size %d unsupported (need 256/384/512/768/1024)", (H)); \
  }

// ------------------------------------------------------------------------------------------------
// ABI: misc
// ------------------------------------------------------------------------------------------------
extern "C" int32_t owc_abi_version(void) { return OWC_ABI_VERSION; }
extern "C" const char* owc_last_error(void) { return g_err; }

extern "C" int32_t owc_ctx_create(int32_t device, void* stream, int32_t chunk_tokens, owc_ctx** out) {
  if (!out) return fail("owc_ctx_create: out is NULL");
  *out = nullptr;
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0)
    return fail("no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= n_dev) return fail("device %d out of range (have %d)", device, n_dev);
  CU_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
  owc_ctx* c = new (std::nothrow) owc_ctx();
  if (!c) return fail("out of host memory");
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  if (const char* gm = getenv("OWC_GEMM_MODE")) c->gemm_mode = std::max(0, std::min(2, atoi(gm)));
  if (const char* sm = getenv("OWC_SIM_MODE")) c->sim_mode = std::max(0, std::min(2, atoi(sm)));
  if (const char* fl = getenv("OWC_FUSE_LN")) c->fuse_ln = atoi(fl) != 0;
  if (const char* ar = getenv("OWC_ARES")) c->ares = atoi(ar) != 0;
  if (const char* pd = getenv("OWC_PDL")) c->pdl = atoi(pd) != 0;
  if (const char* fu = getenv("OWC_FUSED")) c->fused = atoi(fu) != 0;
  if (const char* ew = getenv("OWC_ENC_WIDE")) c->enc_wide = atoi(ew) != 0;
  if (const char* fw = getenv("OWC_FUSED_WAVES")) c->fused_waves = std::max(1, std::min(64, atoi(fw)));
  if (const char* fd = getenv("OWC_FUSED_DBG")) c->fused_dbg = atoi(fd) != 0;
  if (const char* wd = getenv("OWC_WIDE")) c->wide = atoi(wd) != 0;
  if (const char* ww = getenv("OWC_WIDE_WAVES")) c->wide_waves = std::max(1, std::min(64, atoi(ww)));
  if (const char* ag = getenv("OWC_ATTN_GRP")) c->attn_grp = atoi(ag) != 0;
  if (const char* wd2 = getenv("OWC_WIDE_DEEP")) c->wide_deep = atoi(wd2) != 0;
  if (const char* wr = getenv("OWC_WIDE_ROT")) c->wide_rot = atoi(wr) != 0;
  if (const char* fs = getenv("OWC_WIDE_FFN2_SPLIT")) c->wide_ffn2_split = atoi(fs) != 0;
  if (const char* qa_ = getenv("OWC_QKV_ATTN")) c->qkv_attn = atoi(qa_) != 0;
  const bool want_wide_dbg = getenv("OWC_WIDE_DBG") && atoi(getenv("OWC_WIDE_DBG")) != 0;
  c->chunk_default = chunk_tokens <= 0;
  c->chunk_tokens = chunk_tokens > 0 ? std::max(1024, (chunk_tokens + 127) / 128 * 128) : c->sm_count * 128;
  if (stream) {
    c->stream = static_cast<cudaStream_t>(stream);
  } else {
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return fail("cudaStreamCreate failed"); }
    c->own_stream = true;
  }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
      qres != cudaDriverEntryPointSuccess) {
    delete c;
    return fail("cuTensorMapEncodeTiled not available from the driver");
  }
  c->encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
  if (cudaMalloc(&c->d_status, sizeof(int)) != cudaSuccess) { delete c; return fail("cudaMalloc failed"); }
  cudaMemsetAsync(c->d_status, 0, sizeof(int), c->stream);
  if (want_wide_dbg) {
    if (cudaMalloc(&c->wide_dbg, 48 * sizeof(long long)) != cudaSuccess) { delete c; return fail("cudaMalloc failed"); }
    cudaMemsetAsync(c->wide_dbg, 0, 48 * sizeof(long long), c->stream);
  }
  *out = c;
  return 0;
}

extern "C" int32_t owc_ctx_destroy(owc_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  Buf* bufs[] = {&c->ids, &c->cu, &c->targets, &c->zp, &c->zpb, &c->zcb, &c->zc32, &c->cand_s, &c->cand_i,
                 &c->out_s, &c->out_i, &c->out_p, &c->tmp_a, &c->tmp_b, &c->tmp_c, &c->tmp_d, &c->blk, &c->grp};
  for (Buf* b : bufs)
    if (b->p) cudaFree(b->p);
  if (c->d_status) cudaFree(c->d_status);
  if (c->wide_dbg) cudaFree(c->wide_dbg);
  for (cudaEvent_t ev : c->prof_events) cudaEventDestroy(ev);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

static int sync_and_check(owc_ctx* c) {
  int status = 0;
  CU_OK(cudaMemcpyAsync(&status, c->d_status, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU_OK(cudaStreamSynchronize(c->stream));
  if (status != 0) {
    cudaMemsetAsync(c->d_status, 0, sizeof(int), c->stream);
    if (status & 2) return fail("target / row index out of range for its table (status %d)", status);
    return fail("token id out of range for the embedding table (status %d)", status);
  }
  return 0;
}

extern "C" int32_t owc_ctx_sync(owc_ctx* c) {
  if (!c) return fail("ctx is NULL");
  CU_OK(cudaSetDevice(c->device));
  return sync_and_check(c);
}

extern "C" int32_t owc_ctx_stats(owc_ctx* c, owc_stats* out) {
  if (!c || !out) return fail("NULL argument");
  *out = c->stats;
  return 0;
}

extern "C" int32_t owc_ctx_profile_begin(owc_ctx* c) {
  if (!c) return fail("ctx is NULL");
  CU_OK(cudaSetDevice(c->device));
  CU_OK(cudaStreamSynchronize(c->stream));
  c->prof_used = 0;
  c->prof_cats.clear();
  c->profiling = true;
  return 0;
}

extern "C" int32_t owc_ctx_profile_end(owc_ctx* c, owc_profile* out) {
  if (!c || !out) return fail("NULL argument");
  CU_OK(cudaSetDevice(c->device));
  c->profiling = false;
  CU_OK(cudaStreamSynchronize(c->stream));
  memset(out, 0, sizeof(*out));
  for (size_t i = 0; i < c->prof_cats.size(); ++i) {
    float ms = 0.f;
    CU_OK(cudaEventElapsedTime(&ms, c->prof_events[2 * i], c->prof_events[2 * i + 1]));
    const int cat = c->prof_cats[i] & 0xff, cat2 = (c->prof_cats[i] >> 8) & 0xff;
    out->ms[cat] += ms;
    out->launches[cat] += 1;
    if (cat2 != 0xff) {   // the per-shape breakdown of the wide encoder's contractions
      out->ms[cat2] += ms;
      out->launches[cat2] += 1;
    }
  }
  c->prof_used = 0;
  c->prof_cats.clear();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// ABI: encoder load / destroy
// ------------------------------------------------------------------------------------------------
// Wqkv [3H, H] (rows q | k | v, each head-major) -> rows grouped per head: [head][q 32 | k 32 | v 32][H],
// so one 96-row tile of the fused encoder is one head's q|k|v projection (H = 384, 12 heads of 32).
__global__ void __launch_bounds__(256) k_regroup_qkv(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst) {
  constexpr int H = fz::H, D = fz::D;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // one 16-byte piece (8 bf16) of a destination row
  if (i >= 3 * H * H / 8) return;
  const int drow = i / (H / 8), piece = i % (H / 8);
  const int head = drow / (3 * D), p = (drow / D) % 3, d = drow % D;
  const int srow = p * H + head * D + d;
  reinterpret_cast<uint4*>(dst)[static_cast<size_t>(drow) * (H / 8) + piece] =
      reinterpret_cast<const uint4*>(src)[static_cast<size_t>(srow) * (H / 8) + piece];
}

static size_t blob_floats(const owc_arch& a) {
  const size_t h = a.hidden, i = a.inter;
  size_t n = (size_t)a.vocab * h + (size_t)a.max_pos * h + (size_t)a.type_vocab * h + 2 * h;
  if (a.kind == OWC_ARCH_MPNET) n += (size_t)a.heads * (2 * (size_t)a.max_pos - 1);
  const size_t per_layer = 3 * h * h + 3 * h + h * h + h + 2 * h + i * h + i + h * i + h + 2 * h;
  return n + (size_t)a.layers * per_layer;
}

extern "C" int32_t owc_encoder_destroy(owc_encoder* e) {
  if (!e) return 0;
  cudaSetDevice(e->ctx->device);
  cudaStreamSynchronize(e->ctx->stream);
  void* ptrs[] = {e->w32, e->w16, e->x, e->y, e->xb, e->xl, e->qkv, e->ao, e->f, e->qkv32, e->ao32, e->f32,
                  e->fw_qkv, e->fw_o, e->fw_1, e->fw_2, e->fx, e->fdbg, e->wfold16, e->wfold32, e->st1, e->st2, e->wfold16h, e->wfold32h};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  delete e;
  return 0;
}

extern "C" int32_t owc_encoder_load(owc_ctx* c, const owc_arch* arch, const float* weights, size_t n_floats,
                                    owc_encoder** out) {
  if (!c || !arch || !weights || !out) return fail("owc_encoder_load: NULL argument");
  *out = nullptr;
  const owc_arch& a = *arch;
  if (a.kind != OWC_ARCH_BERT && a.kind != OWC_ARCH_MPNET) return fail("unknown arch kind %d", a.kind);
  if (a.hidden % 128 != 0 || a.hidden < 256 || a.hidden > 1024) return fail("hidden %d unsupported", a.hidden);
  if (a.heads <= 0 || a.hidden % a.heads != 0) return fail("heads %d does not divide hidden %d", a.heads, a.hidden);
  const int hd = a.hidden / a.heads;
  if (hd != 32 && hd != 64) return fail("head_dim %d unsupported (need 32 or 64)", hd);
  if (a.inter % 64 != 0 || a.layers <= 0 || a.vocab <= 0 || a.max_pos <= 0) return fail("bad architecture constants");
  if (a.kind == OWC_ARCH_BERT && a.type_vocab < 1) return fail("BERT needs a token-type table");
  if (a.kind == OWC_ARCH_MPNET && (a.type_vocab != 0 || a.heads % 4 != 0)) return fail("bad MPNet constants");
  if (n_floats != blob_floats(a)) return fail("weight blob has %zu floats, architecture needs %zu", n_floats, blob_floats(a));
  CU_OK(cudaSetDevice(c->device));

  owc_encoder* e = new (std::nothrow) owc_encoder();
  if (!e) return fail("out of host memory");
  e->ctx = c;
  e->arch = a;
  // one 128-row block per SM; wide encoders (CTA-pair GEMMs) run two waves per launch (measured on mpnet-base:
  // 129.1 ms per c3 step against 132.5 ms; more than two changes nothing)
  const size_t h = a.hidden, I = a.inter;
  e->wide = h >= 768 && h % 256 == 0 && I % 256 == 0;
  e->wide_ln = e->wide && c->wide;
  e->cap = (c->chunk_default && a.hidden >= 768) ? 2 * c->chunk_tokens : c->chunk_tokens;
  // the persistent kernels of the deferred-LayerNorm path amortise their launch over several waves of row-block pairs
  if (c->chunk_default && e->wide_ln) e->cap = (c->sm_count / 2) * 256 * c->wide_waves;
#define LOAD_FAIL(...) do { int _r = fail(__VA_ARGS__); owc_encoder_destroy(e); return _r; } while (0)
#define LOAD_CU(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) LOAD_FAIL("%s failed: %s", #expr, cudaGetErrorString(_e)); } while (0)
  LOAD_CU(cudaMalloc(&e->w32, n_floats * sizeof(float)));
  LOAD_CU(cudaMemcpyAsync(e->w32, weights, n_floats * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  c->stats.h2d_bytes += static_cast<int64_t>(n_floats * sizeof(float));
  const size_t per_layer16 = 3 * h * h + h * h + I * h + h * I;
  LOAD_CU(cudaMalloc(&e->w16, a.layers * per_layer16 * sizeof(__nv_bfloat16)));

  const float* p = e->w32;
  e->word = p; p += (size_t)a.vocab * h;
  e->pos = p; p += (size_t)a.max_pos * h;
  if (a.kind == OWC_ARCH_BERT) { e->type0 = p; p += (size_t)a.type_vocab * h; }
  e->emb_g = p; p += h;
  e->emb_b = p; p += h;
  if (a.kind == OWC_ARCH_MPNET) { e->rel_lut = p; p += (size_t)a.heads * (2 * (size_t)a.max_pos - 1); }
  e->layers.resize(a.layers);
  __nv_bfloat16* q = e->w16;
  for (int l = 0; l < a.layers; ++l) {
    LayerW& L = e->layers[l];
    L.wqkv32 = p; p += 3 * h * h;
    L.bqkv = p; p += 3 * h;
    L.wo32 = p; p += h * h;
    L.bo = p; p += h;
    L.ln1g = p; p += h;
    L.ln1b = p; p += h;
    L.w132 = p; p += I * h;
    L.b1 = p; p += I;
    L.w232 = p; p += h * I;
    L.b2 = p; p += h;
    L.ln2g = p; p += h;
    L.ln2b = p; p += h;
    L.wqkv = q; q += 3 * h * h;
    L.wo = q; q += h * h;
    L.w1 = q; q += I * h;
    L.w2 = q; q += h * I;
    if (launch_convert(c, L.wqkv32, const_cast<__nv_bfloat16*>(L.wqkv), 3 * h * h) ||
        launch_convert(c, L.wo32, const_cast<__nv_bfloat16*>(L.wo), h * h) ||
        launch_convert(c, L.w132, const_cast<__nv_bfloat16*>(L.w1), I * h) ||
        launch_convert(c, L.w232, const_cast<__nv_bfloat16*>(L.w2), h * I)) { owc_encoder_destroy(e); return 1; }
    if (make_tmap(c, &L.tm_wqkv, L.wqkv, 3 * h, h, kEncBN) || make_tmap(c, &L.tm_wo, L.wo, h, h, kEncBN) ||
        make_tmap(c, &L.tm_w1, L.w1, I, h, kEncBN) || make_tmap(c, &L.tm_w2, L.w2, h, I, kEncBN)) { owc_encoder_destroy(e); return 1; }
    if (make_tmap(c, &L.tm2_wqkv, L.wqkv, 3 * h, h, kEncBN / 2) || make_tmap(c, &L.tm2_wo, L.wo, h, h, kEncBN / 2) ||
        make_tmap(c, &L.tm2_w1, L.w1, I, h, kEncBN / 2) || make_tmap(c, &L.tm2_w2, L.w2, h, I, kEncBN / 2)) { owc_encoder_destroy(e); return 1; }
    if (e->wide &&
        (make_tmap(c, &L.tmw_wqkv, L.wqkv, 3 * h, h, 128) || make_tmap(c, &L.tmw_wo, L.wo, h, h, 128) ||
         make_tmap(c, &L.tmw_w1, L.w1, I, h, 128) || make_tmap(c, &L.tmw_w2, L.w2, h, I, 128))) { owc_encoder_destroy(e); return 1; }
  }
  if (static_cast<size_t>(p - e->w32) != n_floats) LOAD_FAIL("internal: blob walk mismatch");

  const size_t cap = e->cap;
  LOAD_CU(cudaMalloc(&e->x, cap * h * sizeof(float)));
  LOAD_CU(cudaMalloc(&e->y, cap * h * sizeof(float)));
  LOAD_CU(cudaMalloc(&e->xb, cap * h * sizeof(__nv_bfloat16)));
  LOAD_CU(cudaMalloc(&e->qkv, cap * 3 * h * sizeof(__nv_bfloat16)));
  LOAD_CU(cudaMalloc(&e->ao, cap * h * sizeof(__nv_bfloat16)));
  LOAD_CU(cudaMalloc(&e->f, cap * I * sizeof(__nv_bfloat16)));
  // Rows past a chunk's token count are read by TMA (never stored): keep them finite.
  LOAD_CU(cudaMemsetAsync(e->xb, 0, cap * h * sizeof(__nv_bfloat16), c->stream));
  LOAD_CU(cudaMemsetAsync(e->ao, 0, cap * h * sizeof(__nv_bfloat16), c->stream));
  LOAD_CU(cudaMemsetAsync(e->f, 0, cap * I * sizeof(__nv_bfloat16), c->stream));
  if (make_tmap(c, &e->tm_xb, e->xb, cap, h, 128) || make_tmap(c, &e->tm_ao, e->ao, cap, h, 128) ||
      make_tmap(c, &e->tm_f, e->f, cap, I, 128)) { owc_encoder_destroy(e); return 1; }
  if (make_out_tmap(c, &e->tmo_qkv, e->qkv, cap, 3 * h, true) || make_out_tmap(c, &e->tmo_f, e->f, cap, I, true) ||
      make_out_tmap(c, &e->tmo_y, e->y, cap, h, false) || make_out_tmap(c, &e->tmo_x, e->x, cap, h, false) ||
      make_out_tmap(c, &e->tmo_xb, e->xb, cap, h, true)) { owc_encoder_destroy(e); return 1; }
  // ---- deferred-LayerNorm path of wide encoders (gemm_wide.cuh): gamma-folded Wqkv / W1 and their epilogue vectors ----
  if (e->wide_ln) {
    const size_t L = a.layers;
    LOAD_CU(cudaMalloc(&e->xl, cap * h * sizeof(__nv_bfloat16)));
    LOAD_CU(cudaMemsetAsync(e->xl, 0, cap * h * sizeof(__nv_bfloat16), c->stream));
    if (make_out_tmap(c, &e->tmo_xl, e->xl, cap, h, true)) { owc_encoder_destroy(e); return 1; }
    LOAD_CU(cudaMalloc(&e->wfold16, L * (3 * h * h + I * h) * sizeof(__nv_bfloat16)));
    LOAD_CU(cudaMalloc(&e->wfold32, L * 2 * (3 * h + I) * sizeof(float)));
    const size_t stat_bytes = kWideStatSlots * cap * 2 * sizeof(float);
    LOAD_CU(cudaMalloc(&e->st1, stat_bytes));
    LOAD_CU(cudaMalloc(&e->st2, stat_bytes));
    LOAD_CU(cudaMemsetAsync(e->st1, 0, stat_bytes, c->stream));
    LOAD_CU(cudaMemsetAsync(e->st2, 0, stat_bytes, c->stream));
    __nv_bfloat16* w16 = e->wfold16;
    float* v32 = e->wfold32;
    for (size_t l = 0; l < L; ++l) {
      LayerW& Lw = e->layers[l];
      // the LayerNorm in front of the QKV projection: the embedding LayerNorm for layer 0, the previous layer's output LayerNorm after
      const float* g_in = l == 0 ? e->emb_g : e->layers[l - 1].ln2g;
      const float* b_in = l == 0 ? e->emb_b : e->layers[l - 1].ln2b;
      Lw.fqkv = w16; w16 += 3 * h * h;
      Lw.f1 = w16; w16 += I * h;
      Lw.c1q = v32; v32 += 3 * h;
      Lw.c2q = v32; v32 += 3 * h;
      Lw.c1f = v32; v32 += I;
      Lw.c2f = v32; v32 += I;
      k_fold_weight<<<static_cast<unsigned>((3 * h + 7) / 8), 256, 0, c->stream>>>(
          Lw.wqkv32, g_in, b_in, Lw.bqkv, static_cast<int>(3 * h), static_cast<int>(h), const_cast<__nv_bfloat16*>(Lw.fqkv),
          const_cast<float*>(Lw.c1q), const_cast<float*>(Lw.c2q));
      if (check_launch(c, "k_fold_weight")) { owc_encoder_destroy(e); return 1; }
      k_fold_weight<<<static_cast<unsigned>((I + 7) / 8), 256, 0, c->stream>>>(
          Lw.w132, Lw.ln1g, Lw.ln1b, Lw.b1, static_cast<int>(I), static_cast<int>(h), const_cast<__nv_bfloat16*>(Lw.f1),
          const_cast<float*>(Lw.c1f), const_cast<float*>(Lw.c2f));
      if (check_launch(c, "k_fold_weight")) { owc_encoder_destroy(e); return 1; }
      if (make_tmap(c, &Lw.tmf_wqkv, Lw.fqkv, 3 * h, h, 128) || make_tmap(c, &Lw.tmf_w1, Lw.f1, I, h, 128)) { owc_encoder_destroy(e); return 1; }
    }
    e->qkv_attn = hd == qa::D && a.heads <= qa::MAX_HEADS;
    if (e->qkv_attn) {
      LOAD_CU(cudaMalloc(&e->wfold16h, L * 3 * h * h * sizeof(__nv_bfloat16)));
      LOAD_CU(cudaMalloc(&e->wfold32h, L * 2 * 3 * h * sizeof(float)));
      for (size_t l = 0; l < L; ++l) {
        LayerW& Lw = e->layers[l];
        const float* g_in = l == 0 ? e->emb_g : e->layers[l - 1].ln2g;
        const float* b_in = l == 0 ? e->emb_b : e->layers[l - 1].ln2b;
        Lw.fqkv_h = e->wfold16h + l * 3 * h * h;
        Lw.c1qh = e->wfold32h + l * 6 * h;
        Lw.c2qh = Lw.c1qh + 3 * h;
        k_fold_weight_heads<<<static_cast<unsigned>((3 * h + 7) / 8), 256, 0, c->stream>>>(
            Lw.wqkv32, g_in, b_in, Lw.bqkv, static_cast<int>(h), a.heads, const_cast<__nv_bfloat16*>(Lw.fqkv_h),
            const_cast<float*>(Lw.c1qh), const_cast<float*>(Lw.c2qh));
        if (check_launch(c, "k_fold_weight_heads")) { owc_encoder_destroy(e); return 1; }
        if (make_tmap(c, &Lw.tmh_wqkv, Lw.fqkv_h, 3 * h, h, qa::BN / 2)) { owc_encoder_destroy(e); return 1; }
      }
    }
  }
  // ---- block-resident fused encoder: MiniLM-width BERT only (enc_fused.cuh) ----
  e->fused_ok = a.kind == OWC_ARCH_BERT && a.hidden == fz::H && a.heads == fz::HEADS && a.inter % 128 == 0 && a.inter >= 128;
  if (e->fused_ok) {
    const size_t L = a.layers;
    LOAD_CU(cudaMalloc(&e->fw_qkv, L * 3 * h * h * 2));
    LOAD_CU(cudaMalloc(&e->fw_o, L * h * h * 2));
    LOAD_CU(cudaMalloc(&e->fw_1, L * I * h * 2));
    LOAD_CU(cudaMalloc(&e->fw_2, L * h * I * 2));
    for (size_t l = 0; l < L; ++l) {
      const LayerW& Lw = e->layers[l];
      k_regroup_qkv<<<static_cast<unsigned>((3 * h * h / 8 + 255) / 256), 256, 0, c->stream>>>(Lw.wqkv, e->fw_qkv + l * 3 * h * h);
      if (check_launch(c, "k_regroup_qkv")) { owc_encoder_destroy(e); return 1; }
      LOAD_CU(cudaMemcpyAsync(e->fw_o + l * h * h, Lw.wo, h * h * 2, cudaMemcpyDeviceToDevice, c->stream));
      LOAD_CU(cudaMemcpyAsync(e->fw_1 + l * I * h, Lw.w1, I * h * 2, cudaMemcpyDeviceToDevice, c->stream));
      LOAD_CU(cudaMemcpyAsync(e->fw_2 + l * h * I, Lw.w2, h * I * 2, cudaMemcpyDeviceToDevice, c->stream));
    }
    if (make_tmap(c, &e->tmf_qkv, e->fw_qkv, L * 3 * h, h, 48) || make_tmap(c, &e->tmf_o, e->fw_o, L * h, h, 96) ||
        make_tmap(c, &e->tmf_1, e->fw_1, L * I, h, 64) || make_tmap(c, &e->tmf_2, e->fw_2, L * h, I, 96)) { owc_encoder_destroy(e); return 1; }
    e->fcap = (c->sm_count / 2) * 256 * c->fused_waves;   // rows of (SM pairs x waves) block pairs
    LOAD_CU(cudaMalloc(&e->fx, (static_cast<size_t>(e->fcap) + 128) * h * sizeof(float)));
    LOAD_CU(cudaMemsetAsync(e->fx, 0, (static_cast<size_t>(e->fcap) + 128) * h * sizeof(float), c->stream));
    LOAD_CU(cudaMalloc(&e->fdbg, 2048 * sizeof(long long)));
    LOAD_CU(cudaMemsetAsync(e->fdbg, 0, 2048 * sizeof(long long), c->stream));
    LOAD_CU(cudaFuncSetAttribute(k_encoder_fused<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, fz::SMEM_BYTES));
    LOAD_CU(cudaFuncSetAttribute(k_encoder_fused<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, fz::SMEM_BYTES));
    LOAD_CU(cudaFuncSetAttribute(k_encoder_fused<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, fz::SMEM_BYTES));
  }
  LOAD_CU(cudaStreamSynchronize(c->stream));  // `weights` may be freed by the caller after return
#undef LOAD_CU
#undef LOAD_FAIL
  *out = e;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// encoder forward over one chunk of whole sequences
// ------------------------------------------------------------------------------------------------
// k_embed_ln: persistent warps (LayerNorm parameters in registers for the warp's lifetime), a few CTAs per SM
static unsigned embed_grid(const owc_ctx* c, int n_seq) {
  return static_cast<unsigned>(std::max(1, std::min((n_seq + 3) / 4, c->sm_count * 8)));
}

template <typename T>
static int launch_attn(owc_encoder* e, const AttnArgs& aa) {
  owc_ctx* c = e->ctx;
  const int D = e->arch.hidden / e->arch.heads;
  const long long pairs = static_cast<long long>(aa.n_seq) * aa.heads;
  const bool is_short = aa.max_t <= 64;
  const int groups = is_short ? 4 : 1;
  const size_t smem = (static_cast<size_t>(groups) * aa.max_t * (D + 1) + 4 * D + 4 * static_cast<size_t>(aa.max_t)) * sizeof(float);
  const unsigned grid = static_cast<unsigned>((pairs + groups - 1) / groups);
  ProfScope prof(c, OWC_KCAT_ATTN);
#define ATTN_LAUNCH(DD, WPP, KPL)                                                                         \
  do {                                                                                                    \
    OWC_TRY(ensure_smem_attr(c, k_attn<T, DD, WPP, KPL>, smem));                                          \
    k_attn<T, DD, WPP, KPL><<<grid, 128, smem, c->stream>>>(aa);                                          \
  } while (0)
  if (D == 32) {
    if (aa.max_t <= 32) ATTN_LAUNCH(32, 1, 1);
    else if (is_short) ATTN_LAUNCH(32, 1, 2);
    else ATTN_LAUNCH(32, 4, 16);
  } else {
    if (aa.max_t <= 32) ATTN_LAUNCH(64, 1, 1);
    else if (is_short) ATTN_LAUNCH(64, 1, 2);
    else ATTN_LAUNCH(64, 4, 16);
  }
#undef ATTN_LAUNCH
  return check_launch(c, "k_attn");
}

// bf16 production path for sequences of at most 64 tokens: tensor-core attention, one persistent warp
// per (sequence, head) pair.
static int launch_attn_mma(owc_encoder* e, const AttnArgs& aa) {
  owc_ctx* c = e->ctx;
  const int D = e->arch.hidden / e->arch.heads;
  const long long pairs = static_cast<long long>(aa.n_seq) * aa.heads;
  ProfScope prof(c, OWC_KCAT_ATTN);
#define ATTN_MMA_LAUNCH(DD, TP)                                                                             \
  do {                                                                                                      \
    constexpr int smem = AttnWarpCfg<DD>::smem_bytes(TP);                                                   \
    OWC_TRY(ensure_smem_attr(c, k_attn_mma<DD, TP>, smem));                                                 \
    const int per_sm = std::max(1, std::min(16, (227 * 1024) / (smem + 1024)));                             \
    const long long want = (pairs + 3) / 4;                                                                 \
    const unsigned grid = static_cast<unsigned>(std::min<long long>(want, static_cast<long long>(c->sm_count) * per_sm)); \
    cudaLaunchConfig_t cfg{};                                                                               \
    cfg.gridDim = dim3(grid);                                                                               \
    cfg.blockDim = dim3(128);                                                                               \
    cfg.dynamicSmemBytes = smem;                                                                            \
    cfg.stream = c->stream;                                                                                 \
    cudaLaunchAttribute attr[1];                                                                            \
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                        \
    attr[0].val.programmaticStreamSerializationAllowed = c->pdl ? 1 : 0;                                    \
    cfg.attrs = attr;                                                                                       \
    cfg.numAttrs = 1;                                                                                       \
    CU_OK(cudaLaunchKernelEx(&cfg, k_attn_mma<DD, TP>, aa));                                                \
  } while (0)
  if (D == 32) {
    if (aa.max_t <= 16) ATTN_MMA_LAUNCH(32, 16);
    else if (aa.max_t <= 32) ATTN_MMA_LAUNCH(32, 32);
    else ATTN_MMA_LAUNCH(32, 64);
  } else {
    if (aa.max_t <= 16) ATTN_MMA_LAUNCH(64, 16);
    else if (aa.max_t <= 32) ATTN_MMA_LAUNCH(64, 32);
    else ATTN_MMA_LAUNCH(64, 64);
  }
#undef ATTN_MMA_LAUNCH
  return check_launch(c, "k_attn_mma");
}

// Short sequences grouped into shared 16 x 16 tiles (attn_mma.cuh: k_attn_grp). aa.grp / aa.n_grp describe the groups.
static int launch_attn_grp(owc_encoder* e, const AttnArgs& aa) {
  owc_ctx* c = e->ctx;
  const int D = e->arch.hidden / e->arch.heads;
  const long long units = static_cast<long long>(aa.n_grp) * aa.heads;
  ProfScope prof(c, OWC_KCAT_ATTN);
#define ATTN_GRP_LAUNCH(DD, TP)                                                                             \
  do {                                                                                                      \
    constexpr int smem = AttnWarpCfg<DD>::smem_bytes(TP);                                                   \
    OWC_TRY(ensure_smem_attr(c, k_attn_grp<DD, TP>, smem));                                                 \
    const int per_sm = std::max(1, std::min(16, (227 * 1024) / (smem + 1024)));                             \
    const long long want = (units + 3) / 4;                                                                 \
    const unsigned grid = static_cast<unsigned>(std::max<long long>(1, std::min<long long>(want, static_cast<long long>(c->sm_count) * per_sm))); \
    cudaLaunchConfig_t cfg{};                                                                               \
    cfg.gridDim = dim3(grid);                                                                               \
    cfg.blockDim = dim3(128);                                                                               \
    cfg.dynamicSmemBytes = smem;                                                                            \
    cfg.stream = c->stream;                                                                                 \
    cudaLaunchAttribute attr[1];                                                                            \
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                        \
    attr[0].val.programmaticStreamSerializationAllowed = c->pdl ? 1 : 0;                                    \
    cfg.attrs = attr;                                                                                       \
    cfg.numAttrs = 1;                                                                                       \
    CU_OK(cudaLaunchKernelEx(&cfg, k_attn_grp<DD, TP>, aa));                                                \
  } while (0)
  if (D == 32) {
    if (aa.max_t <= 16) ATTN_GRP_LAUNCH(32, 16);
    else if (aa.max_t <= 32) ATTN_GRP_LAUNCH(32, 32);
    else ATTN_GRP_LAUNCH(32, 64);
  } else {
    if (aa.max_t <= 16) ATTN_GRP_LAUNCH(64, 16);
    else if (aa.max_t <= 32) ATTN_GRP_LAUNCH(64, 32);
    else ATTN_GRP_LAUNCH(64, 64);
  }
#undef ATTN_GRP_LAUNCH
  return check_launch(c, "k_attn_grp");
}

static int ensure_f32_workspace(owc_encoder* e) {
  if (e->qkv32) return 0;
  const size_t cap = e->cap, h = e->arch.hidden, I = e->arch.inter;
  CU_OK(cudaMalloc(&e->qkv32, cap * 3 * h * sizeof(float)));
  CU_OK(cudaMalloc(&e->ao32, cap * h * sizeof(float)));
  CU_OK(cudaMalloc(&e->f32, cap * I * sizeof(float)));
  return 0;
}

// ids/cu: device copies for the whole call. Sequences [seq0, seq0+n_seq), tokens [tok0, tok0+M).
static int encode_chunk(owc_encoder* e, const int32_t* d_ids, const int32_t* d_cu, int seq0, int n_seq, int tok0, int M,
                        int max_t, int precision, float* z, __nv_bfloat16* zb) {
  owc_ctx* c = e->ctx;
  const owc_arch& a = e->arch;
  const int H = a.hidden, I = a.inter;
  const bool bf = precision == OWC_PREC_BF16;
  if (!bf) OWC_TRY(ensure_f32_workspace(e));

  EmbedArgs ea{};
  ea.ids = d_ids; ea.cu = d_cu; ea.seq0 = seq0; ea.n_seq = n_seq; ea.tok0 = tok0;
  ea.word = e->word; ea.pos = e->pos; ea.type0 = e->type0; ea.gamma = e->emb_g; ea.beta = e->emb_b;
  ea.eps = a.ln_eps; ea.H = H; ea.vocab = a.vocab; ea.max_pos = a.max_pos;
  ea.mpnet = a.kind == OWC_ARCH_MPNET; ea.pad_id = a.pad_id;
  ea.x = e->x; ea.xb = bf ? e->xb : nullptr; ea.status = c->d_status;
  {
    ProfScope prof(c, OWC_KCAT_EMBED);
    DISPATCH_NV(H, k_embed_ln<NV, false><<<embed_grid(c, n_seq), 128, 0, c->stream>>>(ea));
  }
  OWC_TRY(check_launch(c, "k_embed_ln"));

  AttnArgs aa{};
  aa.cu = d_cu; aa.seq0 = seq0; aa.n_seq = n_seq; aa.tok0 = tok0; aa.H = H; aa.heads = a.heads; aa.max_t = max_t;
  aa.scale = 1.0f / sqrtf(static_cast<float>(H / a.heads));
  aa.rel_lut = e->rel_lut; aa.max_pos = a.max_pos;

  const int ln_grid = (M + 7) / 8;
  // The residual+LayerNorm tail is fused into the O-projection / FFN2 epilogue when a whole row fits
  // the two TMEM accumulator buffers (H == 2 * kEncBN, i.e. the MiniLM width) and the 1-CTA kernel runs.
  const bool fuse_ln = bf && H == 2 * kEncBN && c->gemm_mode <= 0 && c->fuse_ln;
  for (int l = 0; l < a.layers; ++l) {
    const LayerW& L = e->layers[l];
    const bool last = l + 1 == a.layers;
    // ---- attention block ----
    if (bf) {
      GemmArgs g{};
      g.M = M; g.N = 3 * H; g.K = H; g.bias = L.bqkv;
      OWC_TRY((launch_enc_gemm<EPI_BIAS_BF16>(c, e->wide, e->tm_xb, L.tm_wqkv, L.tm2_wqkv, L.tmw_wqkv, e->tmo_qkv, g)));
      aa.qkv = e->qkv; aa.out = e->ao;
      if (aa.max_t <= 64) OWC_TRY(launch_attn_mma(e, aa));
      else OWC_TRY(launch_attn<__nv_bfloat16>(e, aa));
      g = GemmArgs{};
      g.M = M; g.N = H; g.K = H; g.bias = L.bo;
      if (fuse_ln) {
        g.gamma = L.ln1g; g.beta = L.ln1b; g.ln_eps = a.ln_eps; g.residual = e->x; g.ldo = H;
        OWC_TRY((launch_gemm_tc<kEncBN, EPI_RES_LN, 1>(c, e->tm_ao, L.tm_wo, e->tmo_x, g, 1, &e->tmo_xb)));
      } else {
        g.residual = e->x; g.ldo = H;
        OWC_TRY((launch_enc_gemm<EPI_BIAS_RES_F32>(c, e->wide, e->tm_ao, L.tm_wo, L.tm2_wo, L.tmw_wo, e->tmo_y, g)));
      }
    } else {
      OWC_TRY(launch_gemm_f32<EPF_BIAS>(c, e->x, L.wqkv32, L.bqkv, nullptr, e->qkv32, M, 3 * H, H, 3 * H));
      aa.qkv = e->qkv32; aa.out = e->ao32;
      OWC_TRY(launch_attn<float>(e, aa));
      OWC_TRY(launch_gemm_f32<EPF_BIAS_RES>(c, e->ao32, L.wo32, L.bo, e->x, e->y, M, H, H, H));
    }
    if (!fuse_ln) {
      {
        ProfScope prof(c, OWC_KCAT_LN);
        DISPATCH_NV(H, k_ln<NV><<<ln_grid, 256, 0, c->stream>>>(e->y, M, H, L.ln1g, L.ln1b, a.ln_eps, e->x, bf ? e->xb : nullptr));
      }
      OWC_TRY(check_launch(c, "k_ln"));
    }
    // ---- feed-forward block ----
    if (bf) {
      GemmArgs g{};
      g.M = M; g.N = I; g.K = H; g.bias = L.b1;
      OWC_TRY((launch_enc_gemm<EPI_BIAS_GELU_BF16>(c, e->wide, e->tm_xb, L.tm_w1, L.tm2_w1, L.tmw_w1, e->tmo_f, g)));
      g = GemmArgs{};
      g.M = M; g.N = H; g.K = I; g.bias = L.b2;
      if (fuse_ln) {
        g.gamma = L.ln2g; g.beta = L.ln2b; g.ln_eps = a.ln_eps; g.residual = e->x; g.ldo = H;
        OWC_TRY((launch_gemm_tc<kEncBN, EPI_RES_LN, 1>(c, e->tm_f, L.tm_w2, e->tmo_x, g, 1, &e->tmo_xb)));
      } else {
        g.residual = e->x; g.ldo = H;
        OWC_TRY((launch_enc_gemm<EPI_BIAS_RES_F32>(c, e->wide, e->tm_f, L.tm_w2, L.tm2_w2, L.tmw_w2, e->tmo_y, g)));
      }
    } else {
      OWC_TRY(launch_gemm_f32<EPF_BIAS_GELU>(c, e->x, L.w132, L.b1, nullptr, e->f32, M, I, H, I));
      OWC_TRY(launch_gemm_f32<EPF_BIAS_RES>(c, e->f32, L.w232, L.b2, e->x, e->y, M, H, I, H));
    }
    if (!fuse_ln) {
      {
        ProfScope prof(c, OWC_KCAT_LN);
        DISPATCH_NV(H, k_ln<NV><<<ln_grid, 256, 0, c->stream>>>(e->y, M, H, L.ln2g, L.ln2b, a.ln_eps, e->x, (bf && !last) ? e->xb : nullptr));
      }
      OWC_TRY(check_launch(c, "k_ln"));
    }
  }
  {
    ProfScope prof(c, OWC_KCAT_POOL);
    DISPATCH_NV(H, k_pool_l2<NV><<<(n_seq + 3) / 4, 128, 0, c->stream>>>(e->x, d_cu, seq0, n_seq, tok0, H, z, zb));
  }
  OWC_TRY(check_launch(c, "k_pool_l2"));
  c->stats.tokens_encoded += M;
  c->stats.sequences_encoded += n_seq;
  return 0;
}

// Wide encoders (H = 768) in bf16: persistent CTA-pair GEMMs with the LayerNorms deferred into their consumers
// (gemm_wide.cuh). Per layer: QKV (fold) -> attention -> O-projection (+residual, statistics) -> FFN1 (fold + GELU)
// -> FFN2 (+residual, statistics); the PRE-LayerNorm sum y lives as the bf16 pair e->xb (high half = the GEMMs' A operand) +
// e->xl (low half), st1 / st2 are the row statistics.
static int encode_chunk_wide(owc_encoder* e, const int32_t* d_ids, const int32_t* d_cu, const int32_t* cu, int seq0, int n_seq,
                             int tok0, int M, int max_t, float* z, __nv_bfloat16* zb) {
  owc_ctx* c = e->ctx;
  const owc_arch& a = e->arch;
  const int H = a.hidden, I = a.inter;

  EmbedArgs ea{};
  ea.ids = d_ids; ea.cu = d_cu; ea.seq0 = seq0; ea.n_seq = n_seq; ea.tok0 = tok0;
  ea.word = e->word; ea.pos = e->pos; ea.type0 = e->type0; ea.gamma = e->emb_g; ea.beta = e->emb_b;
  ea.eps = a.ln_eps; ea.H = H; ea.vocab = a.vocab; ea.max_pos = a.max_pos;
  ea.mpnet = a.kind == OWC_ARCH_MPNET; ea.pad_id = a.pad_id;
  ea.x = nullptr; ea.xb = e->xb; ea.xl = e->xl; ea.status = c->d_status; ea.stats = e->st2;
  {
    ProfScope prof(c, OWC_KCAT_EMBED);
    DISPATCH_NV(H, k_embed_ln<NV, true><<<embed_grid(c, n_seq), 128, 0, c->stream>>>(ea));
  }
  OWC_TRY(check_launch(c, "k_embed_ln"));

  AttnArgs aa{};
  aa.cu = d_cu; aa.seq0 = seq0; aa.n_seq = n_seq; aa.tok0 = tok0; aa.H = H; aa.heads = a.heads; aa.max_t = max_t;
  aa.scale = 1.0f / sqrtf(static_cast<float>(H / a.heads));
  aa.rel_lut = e->rel_lut; aa.max_pos = a.max_pos;
  aa.qkv = e->qkv; aa.out = e->ao;
  const bool fuse_attn = e->qkv_attn && c->qkv_attn && max_t <= 64;
  const void* d_blk = nullptr;
  int n_blk = 0;
  if (fuse_attn) {
    // row blocks of whole sequences (<= 128 tokens, <= 128 sequences), greedy in sequence order, as for the block-resident encoder
    std::vector<int32_t> blk;
    blk.reserve(static_cast<size_t>(M) / 96 + 2);
    blk.push_back(seq0);
    int rows = 0, cnt = 0;
    for (int s = seq0; s < seq0 + n_seq; ++s) {
      const int len = cu[s + 1] - cu[s];
      if (rows + len > 128 || cnt == 128) {
        blk.push_back(s);
        rows = 0;
        cnt = 0;
      }
      rows += len;
      ++cnt;
    }
    blk.push_back(seq0 + n_seq);
    n_blk = static_cast<int>(blk.size()) - 1;
    OWC_TRY(to_device(c, c->blk, blk.data(), blk.size() * sizeof(int32_t), OWC_MEM_HOST, &d_blk));
  }
  if (!fuse_attn && max_t <= 64 && c->attn_grp) {
    // attention groups: runs of consecutive sequences with at most 16 tokens together share one 16 x 16 tile; a longer
    // sequence is a group of its own
    std::vector<int32_t> grp;
    grp.reserve(static_cast<size_t>(M) / 8 + 2);
    grp.push_back(seq0);
    int rows = 0;
    for (int s = seq0; s < seq0 + n_seq; ++s) {
      const int len = cu[s + 1] - cu[s];
      if (rows > 0 && rows + len > 16) {
        grp.push_back(s);
        rows = 0;
      }
      rows += len;
    }
    grp.push_back(seq0 + n_seq);
    const void* d_grp = nullptr;
    OWC_TRY(to_device(c, c->grp, grp.data(), grp.size() * sizeof(int32_t), OWC_MEM_HOST, &d_grp));
    aa.grp = static_cast<const int32_t*>(d_grp);
    aa.n_grp = static_cast<int>(grp.size()) - 1;
  }

  const int o_slots = 2 * wide_splits(c, M, H / 256);     // partial-statistics slots the O-projection of this chunk writes (st1)
  const int f2_splits = ffn2_splits(c, M, H / 256);
  const int res_slots = 2 * f2_splits;                     // ... and FFN2 (st2)
  const long long stride = e->cap;
  for (int l = 0; l < a.layers; ++l) {
    const LayerW& L = e->layers[l];
    const float* g_in = l == 0 ? e->emb_g : e->layers[l - 1].ln2g;
    const float* b_in = l == 0 ? e->emb_b : e->layers[l - 1].ln2b;
    const int slots2 = l == 0 ? 1 : res_slots;   // the embedding kernel writes whole-row statistics into slot 0
    WideArgs g{};
    g.M = M; g.N = 3 * H; g.K = H; g.inv_h = 1.0f / static_cast<float>(H); g.ln_eps = a.ln_eps; g.stat_stride = stride;
    g.c1 = L.c1q; g.c2 = L.c2q; g.stats_in = e->st2; g.slots_in = slots2;
    if (fuse_attn) {
      QkvAttnArgs qa_args{};
      qa_args.n_blk = n_blk; qa_args.blk_seq = static_cast<const int32_t*>(d_blk); qa_args.cu = d_cu; qa_args.tok0 = tok0;
      qa_args.M = M; qa_args.K = H; qa_args.heads = a.heads;
      qa_args.c1 = L.c1qh; qa_args.c2 = L.c2qh; qa_args.stats_in = e->st2; qa_args.slots_in = slots2; qa_args.stat_stride = stride;
      qa_args.inv_h = 1.0f / static_cast<float>(H); qa_args.ln_eps = a.ln_eps;
      qa_args.ao = e->ao; qa_args.ldo = H;
      qa_args.scale_log2 = aa.scale * 1.4426950408889634f; qa_args.rel_lut = e->rel_lut; qa_args.max_pos = a.max_pos;
      OWC_TRY(launch_qkv_attn(e, L, qa_args, max_t > 16));
    } else {
      OWC_TRY(launch_gemm_wide<WEPI_FOLD_BF16>(c, e->tm_xb, L.tmf_wqkv, e->tmo_qkv, e->tmo_qkv, g, OWC_KCAT_GEMM_QKV));
      if (aa.grp != nullptr) OWC_TRY(launch_attn_grp(e, aa));
      else if (aa.max_t <= 64) OWC_TRY(launch_attn_mma(e, aa));
      else OWC_TRY(launch_attn<__nv_bfloat16>(e, aa));
    }
    g = WideArgs{};
    g.M = M; g.N = H; g.K = H; g.inv_h = 1.0f / static_cast<float>(H); g.ln_eps = a.ln_eps; g.stat_stride = stride;
    g.bias = L.bo; g.stats_prev = e->st2; g.slots_in = slots2; g.gamma_prev = g_in; g.beta_prev = b_in; g.stats_out = e->st1;
    OWC_TRY(launch_gemm_wide<WEPI_RES_STATS>(c, e->tm_ao, L.tmw_wo, e->tmo_xl, e->tmo_xb, g, OWC_KCAT_GEMM_O));
    g = WideArgs{};
    g.M = M; g.N = I; g.K = H; g.inv_h = 1.0f / static_cast<float>(H); g.ln_eps = a.ln_eps; g.stat_stride = stride;
    g.c1 = L.c1f; g.c2 = L.c2f; g.stats_in = e->st1; g.slots_in = o_slots;
    OWC_TRY(launch_gemm_wide<WEPI_FOLD_GELU_BF16>(c, e->tm_xb, L.tmf_w1, e->tmo_f, e->tmo_f, g, OWC_KCAT_GEMM_FFN1));
    g = WideArgs{};
    g.M = M; g.N = H; g.K = I; g.inv_h = 1.0f / static_cast<float>(H); g.ln_eps = a.ln_eps; g.stat_stride = stride;
    g.bias = L.b2; g.stats_prev = e->st1; g.slots_in = o_slots; g.gamma_prev = L.ln1g; g.beta_prev = L.ln1b; g.stats_out = e->st2;
    if (c->wide_deep) OWC_TRY(launch_gemm_wide<WEPI_RES_STATS_DEEP>(c, e->tm_f, L.tmw_w2, e->tmo_xl, e->tmo_xb, g, OWC_KCAT_GEMM_FFN2, f2_splits));
    else OWC_TRY(launch_gemm_wide<WEPI_RES_STATS>(c, e->tm_f, L.tmw_w2, e->tmo_xl, e->tmo_xb, g, OWC_KCAT_GEMM_FFN2, f2_splits));
  }
  const LayerW& LL = e->layers[a.layers - 1];
  {
    ProfScope prof(c, OWC_KCAT_POOL);
    DISPATCH_NV(H, k_pool_ln_l2<NV><<<(n_seq + 3) / 4, 128, 0, c->stream>>>(e->xb, e->xl, e->st2, res_slots, stride, LL.ln2g, LL.ln2b, a.ln_eps, d_cu, seq0, n_seq,
                                                                           tok0, H, z, zb));
  }
  OWC_TRY(check_launch(c, "k_pool_ln_l2"));
  c->stats.tokens_encoded += M;
  c->stats.sequences_encoded += n_seq;
  return 0;
}

// Block-resident fused path (enc_fused.cuh): sequences are grouped greedily into 128-token row blocks of
// whole sequences; each launch carries up to e->fcap tokens worth of blocks through every layer.
// Per launch: k_embed_ln -> fx (fp32), k_encoder_fused (fx in place), k_pool_l2.
static int encode_all_fused(owc_encoder* e, const int32_t* d_ids, const int32_t* d_cu, const int32_t* cu, int n_seq, float* z,
                            __nv_bfloat16* zb) {
  owc_ctx* c = e->ctx;
  const owc_arch& a = e->arch;
  const int H = a.hidden;
  // ---- row blocks over the whole call ----
  std::vector<int32_t> blk;
  blk.reserve(static_cast<size_t>(cu[n_seq]) / 96 + 2);
  blk.push_back(0);
  int rows = 0, cnt = 0;
  for (int s = 0; s < n_seq; ++s) {
    const int len = cu[s + 1] - cu[s];
    if (rows + len > 128 || cnt == 128) {
      blk.push_back(s);
      rows = 0;
      cnt = 0;
    }
    rows += len;
    ++cnt;
  }
  blk.push_back(n_seq);
  const int n_blk = static_cast<int>(blk.size()) - 1;
  const void* d_blk = nullptr;
  OWC_TRY(to_device(c, c->blk, blk.data(), blk.size() * sizeof(int32_t), OWC_MEM_HOST, &d_blk));

  FusedArgs fa{};
  fa.x = e->fx;
  fa.cu = d_cu;
  fa.layers = a.layers;
  fa.n_chunks = a.inter / 128;
  const LayerW& L0 = e->layers[0];
  fa.layer_stride = a.layers > 1 ? static_cast<long long>(e->layers[1].bqkv - L0.bqkv) : 0;
  fa.bqkv = L0.bqkv; fa.bo = L0.bo; fa.ln1g = L0.ln1g; fa.ln1b = L0.ln1b;
  fa.b1 = L0.b1; fa.b2 = L0.b2; fa.ln2g = L0.ln2g; fa.ln2b = L0.ln2b;
  fa.ln_eps = a.ln_eps;
  fa.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(H / a.heads));
  fa.dbg = c->fused_dbg ? e->fdbg : nullptr;
  fa.no_groups = c->attn_grp ? 0 : 1;
  // role ablation (wrong results by construction) exists for tools/fused_ablate.py only: it needs OWC_DEBUG_HOOKS=1 as well,
  // so a stray OWC_FUSED_ABLATE in a production environment cannot corrupt embeddings silently
  if (const char* ab = getenv("OWC_FUSED_ABLATE")) {
    const char* hooks = getenv("OWC_DEBUG_HOOKS");
    if (atoi(ab) != 0 && !(hooks && atoi(hooks) != 0))
      return fail("OWC_FUSED_ABLATE is set without OWC_DEBUG_HOOKS=1: refusing to run the ablated (wrong-result) kernel");
    fa.ablate = atoi(ab);
  }

  int b0 = 0;
  while (b0 < n_blk) {
    // whole waves: every cluster gets the same number of block pairs (a launch ends when its slowest cluster does)
    const int b1 = std::min(n_blk, b0 + e->fcap / 128);
    const int seq0 = blk[b0], nseq = blk[b1] - blk[b0];
    const int tok0 = cu[seq0], M = cu[blk[b1]] - tok0;
    if (M > 0) {
      EmbedArgs ea{};
      ea.ids = d_ids; ea.cu = d_cu; ea.seq0 = seq0; ea.n_seq = nseq; ea.tok0 = tok0;
      ea.word = e->word; ea.pos = e->pos; ea.type0 = e->type0; ea.gamma = e->emb_g; ea.beta = e->emb_b;
      ea.eps = a.ln_eps; ea.H = H; ea.vocab = a.vocab; ea.max_pos = a.max_pos;
      ea.mpnet = 0; ea.pad_id = a.pad_id;
      ea.x = e->fx; ea.xb = nullptr; ea.status = c->d_status;
      {
        ProfScope prof(c, OWC_KCAT_EMBED);
        k_embed_ln<3, false><<<embed_grid(c, nseq), 128, 0, c->stream>>>(ea);
      }
      OWC_TRY(check_launch(c, "k_embed_ln"));
      fa.blk_seq = static_cast<const int32_t*>(d_blk) + b0;
      fa.n_blk = b1 - b0;
      fa.tok0 = tok0;
      const int pairs = (fa.n_blk + 1) / 2;
      const int clusters = std::max(1, std::min(pairs, c->sm_count / 2));
      {
        ProfScope prof(c, OWC_KCAT_GEMM);
        if (c->fused_dbg) k_encoder_fused<1><<<2 * clusters, fz::THREADS, fz::SMEM_BYTES, c->stream>>>(e->tmf_qkv, e->tmf_o, e->tmf_1, e->tmf_2, fa);
        else if (fa.ablate) k_encoder_fused<2><<<2 * clusters, fz::THREADS, fz::SMEM_BYTES, c->stream>>>(e->tmf_qkv, e->tmf_o, e->tmf_1, e->tmf_2, fa);
        else k_encoder_fused<0><<<2 * clusters, fz::THREADS, fz::SMEM_BYTES, c->stream>>>(e->tmf_qkv, e->tmf_o, e->tmf_1, e->tmf_2, fa);
      }
      OWC_TRY(check_launch(c, "k_encoder_fused"));
    }
    {
      ProfScope prof(c, OWC_KCAT_POOL);
      k_pool_l2<3><<<(nseq + 3) / 4, 128, 0, c->stream>>>(e->fx, d_cu, seq0, nseq, tok0, H, z, zb);
    }
    OWC_TRY(check_launch(c, "k_pool_l2"));
    c->stats.tokens_encoded += M;
    c->stats.sequences_encoded += nseq;
    b0 = b1;
  }
  return 0;
}

// Validates cu_seqlens, stages ids/cu on the device and runs every chunk. z: device fp32 [n_seq,H];
// zb: device bf16 [n_seq,H] or nullptr.
static int encode_all(owc_encoder* e, const int32_t* ids, int ids_space, const int32_t* cu, int n_seq, int precision,
                      float* z, __nv_bfloat16* zb) {
  owc_ctx* c = e->ctx;
  if (precision != OWC_PREC_BF16 && precision != OWC_PREC_FP32) return fail("unknown precision %d", precision);
  if (n_seq < 0 || !cu) return fail("bad cu_seqlens");
  if (n_seq == 0) return 0;
  if (cu[0] != 0) return fail("cu_seqlens[0] must be 0");
  const int max_len = e->arch.kind == OWC_ARCH_MPNET ? e->arch.max_pos - 2 : e->arch.max_pos;
  int longest = 0;
  for (int s = 0; s < n_seq; ++s) {
    const int len = cu[s + 1] - cu[s];
    longest = std::max(longest, len);
    if (len < 0) return fail("cu_seqlens decreases at %d", s);
    if (len > max_len) return fail("sequence %d has %d tokens; the position table allows %d (truncate on the host)", s, len, max_len);
  }
  const long long total = cu[n_seq];
  if (total > 0 && !ids) return fail("ids is NULL");
  const void* d_ids = nullptr;
  const void* d_cu = nullptr;
  OWC_TRY(to_device(c, c->ids, ids, static_cast<size_t>(std::max<long long>(total, 1)) * sizeof(int32_t), ids_space, &d_ids));
  OWC_TRY(to_device(c, c->cu, cu, (static_cast<size_t>(n_seq) + 1) * sizeof(int32_t), OWC_MEM_HOST, &d_cu));
  if (precision == OWC_PREC_BF16 && c->fused && e->fused_ok && longest <= 64)
    return encode_all_fused(e, static_cast<const int32_t*>(d_ids), static_cast<const int32_t*>(d_cu), cu, n_seq, z, zb);
  // k_qkv_attn works on row blocks of whole sequences (<= 128 rows, packed greedily: ~126 on labels), two per cluster and
  // item. A chunk cut by tokens alone holds ~1.6 % more blocks than whole waves of clusters (602 items over 74 clusters: a
  // ninth, nearly empty wave = 10 % of that kernel's time), so on that path a chunk also ends after `wide_waves` waves of blocks.
  const bool by_blocks = e->wide_ln && precision == OWC_PREC_BF16 && e->qkv_attn && c->qkv_attn && longest <= 64;
  const int blk_cap = by_blocks ? (c->sm_count / 2) * 2 * c->wide_waves : 0;
  int s0 = 0;
  while (s0 < n_seq) {
    int s1 = s0, max_t = 1;
    int blocks = 1, rows = 0, cnt = 0;     // the packing rule of encode_chunk_wide
    while (s1 < n_seq && cu[s1 + 1] - cu[s0] <= e->cap) {
      const int len = cu[s1 + 1] - cu[s1];
      if (by_blocks) {
        if (rows + len > 128 || cnt == 128) {
          if (blocks == blk_cap) break;
          ++blocks;
          rows = 0;
          cnt = 0;
        }
        rows += len;
        ++cnt;
      }
      max_t = std::max(max_t, len);
      ++s1;
    }
    if (s1 == s0) return fail("internal: chunk capacity %d below one sequence", e->cap);
    if (e->wide_ln && precision == OWC_PREC_BF16) {
      OWC_TRY(encode_chunk_wide(e, static_cast<const int32_t*>(d_ids), static_cast<const int32_t*>(d_cu), cu, s0, s1 - s0, cu[s0],
                                cu[s1] - cu[s0], max_t, z, zb));
      s0 = s1;
      continue;
    }
    OWC_TRY(encode_chunk(e, static_cast<const int32_t*>(d_ids), static_cast<const int32_t*>(d_cu), s0, s1 - s0, cu[s0],
                         cu[s1] - cu[s0], max_t, precision, z, zb));
    s0 = s1;
  }
  return 0;
}

extern "C" int32_t owc_encode(owc_encoder* e, const int32_t* ids, int32_t ids_space, const int32_t* cu, int32_t n_seq,
                              int32_t precision, float* out, int32_t out_space) {
  if (!e || !out) return fail("owc_encode: NULL argument");
  owc_ctx* c = e->ctx;
  CU_OK(cudaSetDevice(c->device));
  if (n_seq == 0) return 0;
  const size_t bytes = static_cast<size_t>(n_seq) * e->arch.hidden * sizeof(float);
  float* z = out;
  if (out_space == OWC_MEM_HOST) {
    OWC_TRY(ensure(c, c->zp, bytes));
    z = static_cast<float*>(c->zp.p);
  }
  OWC_TRY(encode_all(e, ids, ids_space, cu, n_seq, precision, z, nullptr));
  if (out_space == OWC_MEM_HOST) {
    OWC_TRY(from_device(c, out, z, bytes, OWC_MEM_HOST));
    return sync_and_check(c);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// ABI: paired dot
// ------------------------------------------------------------------------------------------------
extern "C" int32_t owc_pair_dot(owc_ctx* c, const float* a, int64_t na, const int32_t* a_index, const float* b, int64_t nb,
                                const int32_t* b_index, int64_t n, int32_t dim, int32_t space, float* out) {
  if (!c || !a || !b || !out) return fail("owc_pair_dot: NULL argument");
  if (n < 0 || na < 0 || nb < 0 || dim <= 0 || dim % 4 != 0) return fail("owc_pair_dot: bad n/dim");
  if ((!a_index && na != n) || (!b_index && nb != n)) return fail("owc_pair_dot: without an index the table must have n rows");
  if (n == 0) return 0;
  CU_OK(cudaSetDevice(c->device));
  const void *da, *db, *dai = nullptr, *dbi = nullptr;
  OWC_TRY(to_device(c, c->tmp_a, a, static_cast<size_t>(na) * dim * sizeof(float), space, &da));
  OWC_TRY(to_device(c, c->tmp_b, b, static_cast<size_t>(nb) * dim * sizeof(float), space, &db));
  if (a_index) OWC_TRY(to_device(c, c->tmp_c, a_index, static_cast<size_t>(n) * sizeof(int32_t), space, &dai));
  if (b_index) OWC_TRY(to_device(c, c->tmp_d, b_index, static_cast<size_t>(n) * sizeof(int32_t), space, &dbi));
  float* dout = out;
  if (space == OWC_MEM_HOST) {
    OWC_TRY(ensure(c, c->out_p, static_cast<size_t>(n) * sizeof(float)));
    dout = static_cast<float*>(c->out_p.p);
  }
  ProfScope prof(c, OWC_KCAT_RESCORE);
  k_pair_dot<<<static_cast<unsigned>((n + 7) / 8), 256, 0, c->stream>>>(
      static_cast<const float*>(da), static_cast<const int32_t*>(dai), na, static_cast<const float*>(db),
      static_cast<const int32_t*>(dbi), nb, n, dim, dout, c->d_status);
  OWC_TRY(check_launch(c, "k_pair_dot"));
  if (space == OWC_MEM_HOST) {
    OWC_TRY(from_device(c, out, dout, static_cast<size_t>(n) * sizeof(float), OWC_MEM_HOST));
    return sync_and_check(c);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// similarity + top-k on device-resident embeddings
// ------------------------------------------------------------------------------------------------
template <int BN, int KC>
static int sim_candidates(owc_ctx* c, const __nv_bfloat16* zpb, long long n, const __nv_bfloat16* zcb, int C, int dim,
                          int* splits_out) {
  const bool pair = c->sim_mode != 0;
  const bool ares = c->sim_mode == 2 && dim <= kAresMaxK;
  CUtensorMap ta, tb;
  OWC_TRY(make_tmap(c, &ta, zpb, static_cast<uint64_t>(n), dim, 128));
  OWC_TRY(make_tmap(c, &tb, zcb, static_cast<uint64_t>(C), dim, pair ? BN / 2 : BN));
  int splits = pick_splits(c, static_cast<int>(n), C, BN);
  splits = std::min(splits, 64 / KC);   // 2 lists per split (chunk parity); re-score handles <= 128 candidates/row
  const int tiles = (C + BN - 1) / BN;
  splits = std::max(1, std::min(splits, tiles));
  const int tpc = (tiles + splits - 1) / splits;
  splits = (tiles + tpc - 1) / tpc;
  const size_t lists = 2 * static_cast<size_t>(splits);
  OWC_TRY(ensure(c, c->cand_s, lists * n * KC * sizeof(float)));
  OWC_TRY(ensure(c, c->cand_i, lists * n * KC * sizeof(int)));
  GemmArgs g{};
  g.M = static_cast<int>(n); g.N = C; g.K = dim;
  g.cand_score = static_cast<float*>(c->cand_s.p);
  g.cand_idx = static_cast<int*>(c->cand_i.p);
  if (!pair) OWC_TRY((launch_gemm_tc<BN, EPI_TOPK, KC>(c, ta, tb, ta, g, splits)));   // no output map: the epilogue keeps candidates
  else if (ares) OWC_TRY((launch_gemm_tc2<BN, EPI_TOPK, KC, true>(c, ta, tb, ta, g, splits)));
  else OWC_TRY((launch_gemm_tc2<BN, EPI_TOPK, KC, false>(c, ta, tb, ta, g, splits)));
  *splits_out = static_cast<int>(lists);   // candidate lists per row
  return 0;
}

// zp/zc fp32 device; zpb/zcb bf16 device copies (bf16 mode). Outputs device.
static int sim_topk_device(owc_ctx* c, const float* zp, const __nv_bfloat16* zpb, long long n, const float* zc,
                           const __nv_bfloat16* zcb, int C, int dim, int k, int precision, const int32_t* targets,
                           float* d_scores, int32_t* d_idx, float* d_paired) {
  if (k < 1 || k > 16 || k > C) return fail("top-k needs 1 <= k <= min(16, c) (k=%d, c=%d)", k, C);
  if (n >= (1LL << 31)) return fail("too many rows (%lld)", n);
  if (n == 0) return 0;
  if (precision == OWC_PREC_BF16) {
    if (dim % 64 != 0) return fail("bf16 similarity needs dim %% 64 == 0 (dim=%d)", dim);
    const int KC = (k <= 5) ? 8 : 16;   // candidates kept per row before the exact fp32 re-score
    int splits = 1;
    if (C <= 128) {
      if (KC == 8) OWC_TRY((sim_candidates<128, 8>(c, zpb, n, zcb, C, dim, &splits)));
      else OWC_TRY((sim_candidates<128, 16>(c, zpb, n, zcb, C, dim, &splits)));
    } else {
      if (KC == 8) OWC_TRY((sim_candidates<256, 8>(c, zpb, n, zcb, C, dim, &splits)));
      else OWC_TRY((sim_candidates<256, 16>(c, zpb, n, zcb, C, dim, &splits)));
    }
    const size_t smem = 4 * 2 * static_cast<size_t>(splits) * KC * sizeof(float);
    ProfScope prof(c, OWC_KCAT_RESCORE);
    k_rescore<<<static_cast<unsigned>((n + 3) / 4), 128, smem, c->stream>>>(zp, zc, n, dim, static_cast<const int*>(c->cand_i.p),
                                                                            splits, KC, k, targets, C, d_scores, d_idx, d_paired,
                                                                            c->d_status);
    return check_launch(c, "k_rescore");
  }
  // check mode: materialise fp32 score rows in slabs, then top-k per row
  const long long slab_rows = std::max<long long>(64, std::min<long long>(n, (256LL << 20) / (static_cast<long long>(C) * 4)));
  OWC_TRY(ensure(c, c->tmp_c, static_cast<size_t>(slab_rows) * C * sizeof(float)));
  float* S = static_cast<float*>(c->tmp_c.p);
  for (long long r0 = 0; r0 < n; r0 += slab_rows) {
    const int rows = static_cast<int>(std::min(slab_rows, n - r0));
    OWC_TRY(launch_gemm_f32<EPF_BIAS>(c, zp + r0 * dim, zc, nullptr, nullptr, S, rows, C, dim, C));
    const size_t smem = 4 * 2 * 32 * 16 * sizeof(float);
    ProfScope prof(c, OWC_KCAT_RESCORE);
    k_topk_rows<16><<<(rows + 3) / 4, 128, smem, c->stream>>>(S, rows, C, k, d_scores + r0 * k, d_idx + r0 * k);
    OWC_TRY(check_launch(c, "k_topk_rows"));
  }
  if (targets && d_paired) {
    ProfScope prof(c, OWC_KCAT_RESCORE);
    k_pair_dot<<<static_cast<unsigned>((n + 7) / 8), 256, 0, c->stream>>>(zc, targets, C, zp, nullptr, n, n, dim, d_paired, c->d_status);
    OWC_TRY(check_launch(c, "k_pair_dot"));
  }
  return 0;
}

extern "C" int32_t owc_sim_topk(owc_ctx* c, const float* zp, int64_t n, const float* zc, int32_t C, int32_t dim, int32_t k,
                                int32_t precision, int32_t space, float* scores, int32_t* idx) {
  if (!c || !zp || !zc || !scores || !idx) return fail("owc_sim_topk: NULL argument");
  if (n < 0 || C <= 0 || dim <= 0 || dim % 4 != 0) return fail("owc_sim_topk: bad shape");
  if (n == 0) return 0;
  CU_OK(cudaSetDevice(c->device));
  const void *dzp, *dzc;
  OWC_TRY(to_device(c, c->tmp_a, zp, static_cast<size_t>(n) * dim * sizeof(float), space, &dzp));
  OWC_TRY(to_device(c, c->zc32, zc, static_cast<size_t>(C) * dim * sizeof(float), space, &dzc));
  const __nv_bfloat16 *zpb = nullptr, *zcb = nullptr;
  if (precision == OWC_PREC_BF16) {
    OWC_TRY(ensure(c, c->zpb, static_cast<size_t>(n) * dim * 2));
    OWC_TRY(ensure(c, c->zcb, static_cast<size_t>(C) * dim * 2));
    OWC_TRY(launch_convert(c, static_cast<const float*>(dzp), static_cast<__nv_bfloat16*>(c->zpb.p), n * dim));
    OWC_TRY(launch_convert(c, static_cast<const float*>(dzc), static_cast<__nv_bfloat16*>(c->zcb.p), static_cast<long long>(C) * dim));
    zpb = static_cast<const __nv_bfloat16*>(c->zpb.p);
    zcb = static_cast<const __nv_bfloat16*>(c->zcb.p);
  }
  float* ds = scores;
  int32_t* di = idx;
  if (space == OWC_MEM_HOST) {
    OWC_TRY(ensure(c, c->out_s, static_cast<size_t>(n) * k * sizeof(float)));
    OWC_TRY(ensure(c, c->out_i, static_cast<size_t>(n) * k * sizeof(int32_t)));
    ds = static_cast<float*>(c->out_s.p);
    di = static_cast<int32_t*>(c->out_i.p);
  }
  OWC_TRY(sim_topk_device(c, static_cast<const float*>(dzp), zpb, n, static_cast<const float*>(dzc), zcb, C, dim, k,
                          precision, nullptr, ds, di, nullptr));
  if (space == OWC_MEM_HOST) {
    OWC_TRY(from_device(c, scores, ds, static_cast<size_t>(n) * k * sizeof(float), OWC_MEM_HOST));
    OWC_TRY(from_device(c, idx, di, static_cast<size_t>(n) * k * sizeof(int32_t), OWC_MEM_HOST));
    return sync_and_check(c);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// ABI: the whole path for one shard
// ------------------------------------------------------------------------------------------------
extern "C" int32_t owc_score(owc_encoder* e, const int32_t* ids, int32_t ids_space, const int32_t* cu, int32_t n_seq,
                             const float* zc, int32_t C, const int32_t* targets, int32_t k, int32_t precision,
                             float* out_scores, int32_t* out_idx, float* out_paired, float* out_embeds, int32_t out_space) {
  if (!e || !zc || !out_scores || !out_idx) return fail("owc_score: NULL argument");
  owc_ctx* c = e->ctx;
  CU_OK(cudaSetDevice(c->device));
  if (n_seq == 0) return 0;
  if (C <= 0) return fail("owc_score: no classes");
  const int H = e->arch.hidden;
  const size_t n = static_cast<size_t>(n_seq);
  const bool host_out = out_space == OWC_MEM_HOST;
  const bool bf = precision == OWC_PREC_BF16;

  float* z = (out_embeds && !host_out) ? out_embeds : nullptr;
  if (!z) {
    OWC_TRY(ensure(c, c->zp, n * H * sizeof(float)));
    z = static_cast<float*>(c->zp.p);
  }
  __nv_bfloat16* zb = nullptr;
  if (bf) {
    OWC_TRY(ensure(c, c->zpb, n * H * 2));
    OWC_TRY(ensure(c, c->zcb, static_cast<size_t>(C) * H * 2));
    zb = static_cast<__nv_bfloat16*>(c->zpb.p);
  }
  const void* d_targets = nullptr;
  if (targets) OWC_TRY(to_device(c, c->targets, targets, n * sizeof(int32_t), ids_space, &d_targets));
  float *ds = out_scores, *dp = out_paired;
  int32_t* di = out_idx;
  if (host_out) {
    OWC_TRY(ensure(c, c->out_s, n * k * sizeof(float)));
    OWC_TRY(ensure(c, c->out_i, n * k * sizeof(int32_t)));
    ds = static_cast<float*>(c->out_s.p);
    di = static_cast<int32_t*>(c->out_i.p);
    if (out_paired) {
      OWC_TRY(ensure(c, c->out_p, n * sizeof(float)));
      dp = static_cast<float*>(c->out_p.p);
    }
  }
  OWC_TRY(encode_all(e, ids, ids_space, cu, n_seq, precision, z, zb));
  if (bf) OWC_TRY(launch_convert(c, zc, static_cast<__nv_bfloat16*>(c->zcb.p), static_cast<long long>(C) * H));
  OWC_TRY(sim_topk_device(c, z, zb, n_seq, zc, static_cast<const __nv_bfloat16*>(c->zcb.p), C, H, k, precision,
                          static_cast<const int32_t*>(d_targets), ds, di, (targets && out_paired) ? dp : nullptr));
  if (host_out) {
    OWC_TRY(from_device(c, out_scores, ds, n * k * sizeof(float), OWC_MEM_HOST));
    OWC_TRY(from_device(c, out_idx, di, n * k * sizeof(int32_t), OWC_MEM_HOST));
    if (targets && out_paired) OWC_TRY(from_device(c, out_paired, dp, n * sizeof(float), OWC_MEM_HOST));
    if (out_embeds) OWC_TRY(from_device(c, out_embeds, z, n * H * sizeof(float), OWC_MEM_HOST));
    return sync_and_check(c);
  }
  return 0;
}

// The matching half of owc_score on embeddings that already sit in device memory.
extern "C" int32_t owc_match(owc_ctx* c, const float* zp, int64_t n, const float* zc, int32_t C, int32_t dim,
                             const int32_t* targets, int32_t targets_space, int32_t k, int32_t precision, float* out_scores,
                             int32_t* out_idx, float* out_paired, int32_t out_space) {
  if (!c || !zp || !zc || !out_scores || !out_idx) return fail("owc_match: NULL argument");
  if (n < 0 || C <= 0 || dim <= 0 || dim % 4 != 0) return fail("owc_match: bad shape");
  if (n == 0) return 0;
  CU_OK(cudaSetDevice(c->device));
  const size_t nn = static_cast<size_t>(n);
  const bool host_out = out_space == OWC_MEM_HOST;
  const __nv_bfloat16 *zpb = nullptr, *zcb = nullptr;
  if (precision == OWC_PREC_BF16) {
    OWC_TRY(ensure(c, c->zpb, nn * dim * 2));
    OWC_TRY(ensure(c, c->zcb, static_cast<size_t>(C) * dim * 2));
    OWC_TRY(launch_convert(c, zp, static_cast<__nv_bfloat16*>(c->zpb.p), n * dim));
    OWC_TRY(launch_convert(c, zc, static_cast<__nv_bfloat16*>(c->zcb.p), static_cast<long long>(C) * dim));
    zpb = static_cast<const __nv_bfloat16*>(c->zpb.p);
    zcb = static_cast<const __nv_bfloat16*>(c->zcb.p);
  }
  const void* d_targets = nullptr;
  if (targets) OWC_TRY(to_device(c, c->targets, targets, nn * sizeof(int32_t), targets_space, &d_targets));
  float *ds = out_scores, *dp = out_paired;
  int32_t* di = out_idx;
  if (host_out) {
    OWC_TRY(ensure(c, c->out_s, nn * k * sizeof(float)));
    OWC_TRY(ensure(c, c->out_i, nn * k * sizeof(int32_t)));
    ds = static_cast<float*>(c->out_s.p);
    di = static_cast<int32_t*>(c->out_i.p);
    if (out_paired) {
      OWC_TRY(ensure(c, c->out_p, nn * sizeof(float)));
      dp = static_cast<float*>(c->out_p.p);
    }
  }
  OWC_TRY(sim_topk_device(c, zp, zpb, n, zc, zcb, C, dim, k, precision, static_cast<const int32_t*>(d_targets), ds, di,
                          (targets && out_paired) ? dp : nullptr));
  if (host_out) {
    OWC_TRY(from_device(c, out_scores, ds, nn * k * sizeof(float), OWC_MEM_HOST));
    OWC_TRY(from_device(c, out_idx, di, nn * k * sizeof(int32_t), OWC_MEM_HOST));
    if (targets && out_paired) OWC_TRY(from_device(c, out_paired, dp, nn * sizeof(float), OWC_MEM_HOST));
    return sync_and_check(c);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// ABI: kernel-level test / bench hooks
// ------------------------------------------------------------------------------------------------
extern "C" int32_t owc_debug_gemm(owc_ctx* c, const float* a, const float* w, const float* bias, int32_t m, int32_t n,
                                  int32_t kdim, int32_t precision, int32_t tile_n, float* out) {
  if (!c || !a || !w || !out) return fail("owc_debug_gemm: NULL argument");
  if (m <= 0 || n <= 0 || kdim <= 0) return fail("owc_debug_gemm: bad shape");
  CU_OK(cudaSetDevice(c->device));
  if (precision == OWC_PREC_FP32) return launch_gemm_f32<EPF_BIAS>(c, a, w, bias, nullptr, out, m, n, kdim, n);
  OWC_TRY(ensure(c, c->tmp_a, static_cast<size_t>(m) * kdim * 2));
  OWC_TRY(ensure(c, c->tmp_b, static_cast<size_t>(n) * kdim * 2));
  OWC_TRY(launch_convert(c, a, static_cast<__nv_bfloat16*>(c->tmp_a.p), static_cast<long long>(m) * kdim));
  OWC_TRY(launch_convert(c, w, static_cast<__nv_bfloat16*>(c->tmp_b.p), static_cast<long long>(n) * kdim));
  if (tile_n == 0) tile_n = kEncBN + 1000 * std::max(c->gemm_mode, 0);
  const int variant = tile_n / 1000;   // 0: 1-CTA, 1: CTA pair, 2: CTA pair + resident A
  const int bn = tile_n % 1000;
  if (variant == 2 && kdim > 384) return fail("resident-A variant needs kdim <= 384");
  CUtensorMap ta, tb;
  OWC_TRY(make_tmap(c, &ta, c->tmp_a.p, m, kdim, 128));
  OWC_TRY(make_tmap(c, &tb, c->tmp_b.p, n, kdim, variant == 0 ? bn : bn / 2));
  GemmArgs g{};
  g.M = m; g.N = n; g.K = kdim; g.bias = bias;
  CUtensorMap tc;
  OWC_TRY(make_out_tmap(c, &tc, out, m, n, false));
#define DBG_CASE(BNV)                                                                                     \
  case BNV:                                                                                               \
    if (variant == 0) return launch_gemm_tc<BNV, EPI_BIAS_F32, 1>(c, ta, tb, tc, g, pick_splits(c, m, n, BNV));          \
    if (variant == 1) return launch_gemm_tc2<BNV, EPI_BIAS_F32, 1, false>(c, ta, tb, tc, g, pick_splits(c, m, n, BNV));  \
    return launch_gemm_tc2<BNV, EPI_BIAS_F32, 1, true>(c, ta, tb, tc, g, pick_splits(c, m, n, BNV));
  switch (bn) {
    DBG_CASE(128)
    DBG_CASE(192)
    DBG_CASE(256)
    default: return fail("tile_n must be 128, 192 or 256 (+1000: CTA pair, +2000: CTA pair with resident A)");
  }
#undef DBG_CASE
}

extern "C" int32_t owc_debug_gemm_timeline(owc_ctx* c, int32_t m, int32_t n, int32_t kdim, int32_t gelu, int64_t* stamps,
                                           int32_t max_ctas) {
  if (!c || !stamps) return fail("owc_debug_gemm_timeline: NULL argument");
  if (m <= 0 || n <= 0 || kdim <= 0 || n % kEncBN != 0) return fail("owc_debug_gemm_timeline: bad shape (n must be a multiple of %d)", kEncBN);
  CU_OK(cudaSetDevice(c->device));
  const int ctas = (m + 127) / 128;
  if (ctas > max_ctas) return fail("stamps buffer too small: need %d CTAs", ctas);
  OWC_TRY(ensure(c, c->tmp_a, static_cast<size_t>(m) * kdim * 2));
  OWC_TRY(ensure(c, c->tmp_b, static_cast<size_t>(n) * kdim * 2));
  OWC_TRY(ensure(c, c->tmp_c, static_cast<size_t>(m) * n * 2));
  OWC_TRY(ensure(c, c->tmp_d, static_cast<size_t>(ctas) * 8 * sizeof(long long)));
  CU_OK(cudaMemsetAsync(c->tmp_a.p, 0, static_cast<size_t>(m) * kdim * 2, c->stream));
  CU_OK(cudaMemsetAsync(c->tmp_b.p, 0, static_cast<size_t>(n) * kdim * 2, c->stream));
  CU_OK(cudaMemsetAsync(c->tmp_d.p, 0, static_cast<size_t>(ctas) * 8 * sizeof(long long), c->stream));
  CUtensorMap ta, tb, tc;
  OWC_TRY(make_tmap(c, &ta, c->tmp_a.p, m, kdim, 128));
  OWC_TRY(make_tmap(c, &tb, c->tmp_b.p, n, kdim, kEncBN));
  OWC_TRY(make_out_tmap(c, &tc, c->tmp_c.p, m, n, true));
  GemmArgs g{};
  g.M = m; g.N = n; g.K = kdim;
  g.dbg = static_cast<long long*>(c->tmp_d.p);
  CUtensorMap tx;
  if (gelu == 2) {   // fused residual + LayerNorm epilogue: needs n == 2 * kEncBN, an fp32 x and gamma/beta
    if (n != 2 * kEncBN) return fail("LayerNorm timeline needs n == %d", 2 * kEncBN);
    OWC_TRY(ensure(c, c->zc32, static_cast<size_t>(m) * n * sizeof(float)));
    OWC_TRY(ensure(c, c->zp, 3 * static_cast<size_t>(n) * sizeof(float)));
    CU_OK(cudaMemsetAsync(c->zc32.p, 0, static_cast<size_t>(m) * n * sizeof(float), c->stream));
    CU_OK(cudaMemsetAsync(c->zp.p, 0, 3 * static_cast<size_t>(n) * sizeof(float), c->stream));
    OWC_TRY(make_out_tmap(c, &tx, c->zc32.p, m, n, false));
    g.residual = static_cast<const float*>(c->zc32.p); g.ldo = n;
    g.gamma = static_cast<const float*>(c->zp.p); g.beta = g.gamma + n; g.bias = g.gamma + 2 * n; g.ln_eps = 1e-12f;
  }
  for (int rep = 0; rep < 3; ++rep) {   // the last (warm) launch's stamps are returned
    if (gelu == 2) OWC_TRY((launch_gemm_tc<kEncBN, EPI_RES_LN, 1>(c, ta, tb, tx, g, 1, &tc)));
    else if (gelu) OWC_TRY((launch_gemm_tc<kEncBN, EPI_BIAS_GELU_BF16, 1>(c, ta, tb, tc, g, 1)));
    else OWC_TRY((launch_gemm_tc<kEncBN, EPI_BIAS_BF16, 1>(c, ta, tb, tc, g, 1)));
  }
  CU_OK(cudaMemcpyAsync(stamps, c->tmp_d.p, static_cast<size_t>(ctas) * 8 * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
  CU_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int32_t owc_debug_gemm_loop(owc_ctx* c, int32_t m, int32_t n, int32_t kdim, int32_t gelu, int32_t iters,
                                       float* ms_per_launch) {
  if (!c || !ms_per_launch || iters <= 0) return fail("owc_debug_gemm_loop: bad argument");
  if (m <= 0 || n <= 0 || kdim <= 0 || n % kEncBN != 0) return fail("owc_debug_gemm_loop: bad shape");
  CU_OK(cudaSetDevice(c->device));
  OWC_TRY(ensure(c, c->tmp_a, static_cast<size_t>(m) * kdim * 2));
  OWC_TRY(ensure(c, c->tmp_b, static_cast<size_t>(n) * kdim * 2));
  OWC_TRY(ensure(c, c->tmp_c, static_cast<size_t>(m) * n * 2));
  CU_OK(cudaMemsetAsync(c->tmp_a.p, 0, static_cast<size_t>(m) * kdim * 2, c->stream));
  CU_OK(cudaMemsetAsync(c->tmp_b.p, 0, static_cast<size_t>(n) * kdim * 2, c->stream));
  CUtensorMap ta, tb, tc;
  OWC_TRY(make_tmap(c, &ta, c->tmp_a.p, m, kdim, 128));
  OWC_TRY(make_tmap(c, &tb, c->tmp_b.p, n, kdim, kEncBN));
  OWC_TRY(make_out_tmap(c, &tc, c->tmp_c.p, m, n, true));
  GemmArgs g{};
  g.M = m; g.N = n; g.K = kdim;
  CUtensorMap tx;
  if (gelu == 2) {
    if (n != 2 * kEncBN) return fail("LayerNorm loop needs n == %d", 2 * kEncBN);
    OWC_TRY(ensure(c, c->zc32, static_cast<size_t>(m) * n * sizeof(float)));
    OWC_TRY(ensure(c, c->zp, 3 * static_cast<size_t>(n) * sizeof(float)));
    CU_OK(cudaMemsetAsync(c->zc32.p, 0, static_cast<size_t>(m) * n * sizeof(float), c->stream));
    CU_OK(cudaMemsetAsync(c->zp.p, 0, 3 * static_cast<size_t>(n) * sizeof(float), c->stream));
    OWC_TRY(make_out_tmap(c, &tx, c->zc32.p, m, n, false));
    g.residual = static_cast<const float*>(c->zc32.p); g.ldo = n;
    g.gamma = static_cast<const float*>(c->zp.p); g.beta = g.gamma + n; g.bias = g.gamma + 2 * n; g.ln_eps = 1e-12f;
  }
  auto run = [&]() -> int {
    if (gelu == 2) return launch_gemm_tc<kEncBN, EPI_RES_LN, 1>(c, ta, tb, tx, g, 1, &tc);
    if (gelu) return launch_gemm_tc<kEncBN, EPI_BIAS_GELU_BF16, 1>(c, ta, tb, tc, g, 1);
    return launch_gemm_tc<kEncBN, EPI_BIAS_BF16, 1>(c, ta, tb, tc, g, 1);
  };
  for (int i = 0; i < 3; ++i) OWC_TRY(run());
  cudaEvent_t e0, e1;
  CU_OK(cudaEventCreate(&e0));
  CU_OK(cudaEventCreate(&e1));
  CU_OK(cudaEventRecord(e0, c->stream));
  for (int i = 0; i < iters; ++i) OWC_TRY(run());
  CU_OK(cudaEventRecord(e1, c->stream));
  CU_OK(cudaEventSynchronize(e1));
  float ms = 0.f;
  CU_OK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_per_launch = ms / iters;
  return 0;
}

extern "C" int32_t owc_debug_fused_stamps(owc_encoder* e, int64_t* out64) {
  if (!e || !out64) return fail("owc_debug_fused_stamps: NULL argument");
  if (!e->fdbg) return fail("fused encoder not available for this architecture");
  CU_OK(cudaSetDevice(e->ctx->device));
  CU_OK(cudaMemcpyAsync(out64, e->fdbg, 2048 * sizeof(long long), cudaMemcpyDeviceToHost, e->ctx->stream));
  CU_OK(cudaStreamSynchronize(e->ctx->stream));
  return 0;
}

extern "C" int32_t owc_debug_wide_counters(owc_ctx* c, int64_t* out48, int32_t reset) {
  if (!c || !out48) return fail("owc_debug_wide_counters: NULL argument");
  if (!c->wide_dbg) return fail("the ctx was not created with OWC_WIDE_DBG=1");
  CU_OK(cudaSetDevice(c->device));
  CU_OK(cudaMemcpyAsync(out48, c->wide_dbg, 48 * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
  if (reset) CU_OK(cudaMemsetAsync(c->wide_dbg, 0, 48 * sizeof(long long), c->stream));
  CU_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

__global__ void k_fill_unit_bf16(__nv_bfloat16* p, long long rows, int dim, unsigned seed) {
  // pseudo-random rows with norm ~1 (bench input only)
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * dim) return;
  unsigned x = static_cast<unsigned>(i) * 2654435761u + seed;
  x ^= x >> 15; x *= 2246822519u; x ^= x >> 13; x *= 3266489917u; x ^= x >> 16;
  const float u = (static_cast<float>(x & 0xFFFFFF) / 8388608.0f - 1.0f);   // [-1, 1)
  p[i] = __float2bfloat16_rn(u * 1.7320508f * rsqrtf(static_cast<float>(dim)));
}

extern "C" int32_t owc_bench_sim_kernel(owc_ctx* c, int64_t n, int32_t C, int32_t dim, int32_t k, int32_t iters,
                                        float* ms_per_launch) {
  if (!c || !ms_per_launch) return fail("owc_bench_sim_kernel: NULL argument");
  if (n <= 0 || C <= 0 || dim % 64 != 0 || iters <= 0 || k < 1 || k > 16) return fail("owc_bench_sim_kernel: bad arguments");
  CU_OK(cudaSetDevice(c->device));
  OWC_TRY(ensure(c, c->zpb, static_cast<size_t>(n) * dim * 2));
  OWC_TRY(ensure(c, c->zcb, static_cast<size_t>(C) * dim * 2));
  __nv_bfloat16* zpb = static_cast<__nv_bfloat16*>(c->zpb.p);
  __nv_bfloat16* zcb = static_cast<__nv_bfloat16*>(c->zcb.p);
  k_fill_unit_bf16<<<static_cast<unsigned>((n * dim + 255) / 256), 256, 0, c->stream>>>(zpb, n, dim, 17u);
  OWC_TRY(check_launch(c, "k_fill_unit_bf16"));
  k_fill_unit_bf16<<<static_cast<unsigned>((static_cast<long long>(C) * dim + 255) / 256), 256, 0, c->stream>>>(zcb, C, dim, 91u);
  OWC_TRY(check_launch(c, "k_fill_unit_bf16"));
  const int KC = (k <= 5) ? 8 : 16;
  int splits = 1;
  auto run = [&]() -> int {
    if (C <= 128) return KC == 8 ? sim_candidates<128, 8>(c, zpb, n, zcb, C, dim, &splits) : sim_candidates<128, 16>(c, zpb, n, zcb, C, dim, &splits);
    return KC == 8 ? sim_candidates<256, 8>(c, zpb, n, zcb, C, dim, &splits) : sim_candidates<256, 16>(c, zpb, n, zcb, C, dim, &splits);
  };
  OWC_TRY(run());  // warm-up (also sizes the candidate buffers)
  OWC_TRY(run());
  cudaEvent_t e0, e1;
  CU_OK(cudaEventCreate(&e0));
  CU_OK(cudaEventCreate(&e1));
  CU_OK(cudaEventRecord(e0, c->stream));
  for (int i = 0; i < iters; ++i) OWC_TRY(run());
  CU_OK(cudaEventRecord(e1, c->stream));
  CU_OK(cudaEventSynchronize(e1));
  float ms = 0.f;
  CU_OK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_per_launch = ms / iters;
  return 0;
}
