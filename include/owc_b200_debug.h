/*
 * owc_b200_debug.h - kernel-level test / measurement hooks of libowc_b200.so. NOT part of the product ABI
 * (include/owc_b200.h): nothing on the scoring path calls them; they exist for tests/ (mainloop parity against
 * torch), tools/ (in-kernel timelines) and bench.py's similarity-kernel roofline leg.
 */
#ifndef OWC_B200_DEBUG_H
#define OWC_B200_DEBUG_H

#include "owc_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Kernel-level entry points used by the parity tests and bench.py's roofline leg.
 * owc_debug_gemm: C[m,n] = A[m,kdim] . W[n,kdim]^T (+ bias[n]) in the given precision; all
 * pointers DEVICE fp32 (bf16 mode rounds operands to bf16 first, as the encoder does). tile_n
 * selects the tcgen05 column-tile width (128, 192 or 256; 0 = the encoder's default).            */
int32_t owc_debug_gemm(owc_ctx* ctx, const float* a, const float* w, const float* bias, int32_t m,
                       int32_t n, int32_t kdim, int32_t precision, int32_t tile_n, float* c);
/* Debug: run the 1-CTA tcgen05 kernel once on an m x n x kdim problem (bf16 output + bias epilogue;
 * gelu != 0 adds the GELU) on zero-filled operands and return per-CTA clock64 stamps, 8 int64 per CTA
 * (grid = ceil(m/128) CTAs): [0] entry, [1] setup done, [2] first operands landed, [3] last MMA issued,
 * [4] first accumulator ready, [5] epilogue done, [6] all warps done, [7] globaltimer (ns) at entry.  */
int32_t owc_debug_gemm_timeline(owc_ctx* ctx, int32_t m, int32_t n, int32_t kdim, int32_t gelu, int64_t* stamps,
                                int32_t max_ctas);
/* Debug: CUDA-event time of `iters` back-to-back launches of the same 1-CTA kernel (ms per launch). */
int32_t owc_debug_gemm_loop(owc_ctx* ctx, int32_t m, int32_t n, int32_t kdim, int32_t gelu, int32_t iters,
                            float* ms_per_launch);
/* Time `iters` back-to-back launches of the fused similarity+top-k kernel alone on bf16 operands
 * already resident in device memory (CUDA events on the ctx stream); ms_per_launch out.          */
int32_t owc_bench_sim_kernel(owc_ctx* ctx, int64_t n, int32_t c, int32_t dim, int32_t k, int32_t iters,
                             float* ms_per_launch);
/* Debug: timeline of the block-resident fused encoder kernel (enc_fused.cuh). When the ctx was created with
 * OWC_FUSED_DBG=1 in the environment, the first SIMT thread of CTA 0 records clock64 at: block start,
 * residual rows loaded, then per layer: attention block done, LayerNorm 1 done, feed-forward done,
 * LayerNorm 2 done. Copies the 2048 int64 debug words (stamps, phase sums, event traces) of the most recent launch to `stamps` (HOST).      */
int32_t owc_debug_fused_stamps(owc_encoder* enc, int64_t* stamps);

/* Debug: cycle counters of the instrumented persistent kernels of the wide path, available when the ctx was created with
 * OWC_WIDE_DBG=1: 48 int64, summed over launches since the last reset.
 *   [epilogue kind 0..3][8] of k_gemm_wide (gemm_wide.cuh): {0: MMA warp waiting for operands, 1: waiting for a free accumulator,
 *     2: MMA warp total, 3: epilogue warp waiting for an accumulator, 4: epilogue warp total, 5: leader CTAs counted};
 *   [32..47] of k_qkv_attn (qkv_attn.cuh): {32: MMA warp waiting for operands, 33: for a free accumulator, 34: MMA warp total,
 *     35: drain warp 0 waiting for an accumulator, 36: for the q|k|v buffer, 37: drain warp 0 total, 38: leader CTAs counted,
 *     39: attention units of attention warp 0, 40: that warp waiting for the rows, 41: building the block's tables, 42: total,
 *     43..47: its time in the phases of a unit (fragment loads, Q K^T, softmax, P V, output stores)}.                      */
int32_t owc_debug_wide_counters(owc_ctx* ctx, int64_t* out48, int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* OWC_B200_DEBUG_H */
