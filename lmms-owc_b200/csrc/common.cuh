// common.cuh - shared device helpers: sm_100a PTX wrappers (mbarrier, TMA, tcgen05/TMEM), warp
// reductions, vector loads. Everything here targets sm_100a only (no fallback paths).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace owc {

constexpr int kWarp = 32;

// ------------------------------------------------------------------------------------------
// warp-level reductions
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}

// The residual stream of the wide path as a bf16 pair: hi = bf16(v) (which is also the next GEMM's A operand) and
// lo = bf16(v - hi). hi + lo carries 16 mantissa bits (relative error 2^-17, against 2^-9 for one bf16), so no fp32 copy of the
// stream is written next to the operand copy: 4 instead of 6 bytes per element on every residual update.
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - __low2float(h), b - __high2float(h));
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void join_bf16x2(uint32_t hi, uint32_t lo, float& a, float& b) {
  a = __uint_as_float(hi << 16) + __uint_as_float(lo << 16);
  b = __uint_as_float(hi & 0xffff0000u) + __uint_as_float(lo & 0xffff0000u);
}

// exact erf GELU (hidden_act="gelu"; transformers modeling_bert.py:330-342)
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

// erf-GELU for the bf16 epilogue: erf(z) = 1 - 2^(-z*q(z)) with a degree-4 q fitted on [0, 4]
// (max |error| 7e-7 on erf, 1.2e-6 on gelu - three orders below the bf16 rounding of the result),
// z clamped at 4 where erfc < 2e-8. x*erf(x/sqrt2) is even, so only |x| enters the polynomial.
// ~12 issue slots per element against ~30 for erff: the FFN1 epilogue is issue-bound (4 warps/SM).
__device__ __forceinline__ float gelu_fast(float x) {
  const float ax = fabsf(x);
  const float z = fminf(ax * 0.70710678118654752440f, 4.0f);
  float p = 0.00294416095f;
  p = fmaf(p, z, -0.0295900535f);
  p = fmaf(p, z, 0.148665637f);
  p = fmaf(p, z, 0.918509364f);
  p = fmaf(p, z, 1.62788904f);
  p *= z;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-p));
  const float h = 0.5f * ax;
  return fmaf(-h, e, fmaf(0.5f, x, h));
}

// Packed fp32x2 arithmetic (sm_100 FFMA2/FMUL2): two elements per issue slot.
__device__ __forceinline__ uint64_t f2_pack(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// gelu_fast on a pair, polynomial in packed arithmetic (cubic q, |erf error| 1.5e-5: still two orders
// below the bf16 rounding of the result): gelu(x) = relu(x) - 0.5|x| * 2^(-z q(z)), z = min(|x|/sqrt2, 4).
__device__ __forceinline__ void gelu_fast2(float& x0, float& x1) {
  const float z0 = fminf(fabsf(x0) * 0.70710678118654752440f, 4.0f);
  const float z1 = fminf(fabsf(x1) * 0.70710678118654752440f, 4.0f);
  const uint64_t z = f2_pack(z0, z1);
  uint64_t p = f2_fma(f2_pack(-0.0191262871f, -0.0191262871f), z, f2_pack(0.136623386f, 0.136623386f));
  p = f2_fma(p, z, f2_pack(0.923601230f, 0.923601230f));
  p = f2_fma(p, z, f2_pack(1.62727115f, 1.62727115f));
  p = f2_mul(p, z);
  float p0, p1, e0, e1;
  f2_unpack(p, p0, p1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(-p0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(-p1));
  const uint64_t r = f2_fma(f2_pack(fabsf(x0) * -0.5f, fabsf(x1) * -0.5f), f2_pack(e0, e1),
                            f2_pack(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
  f2_unpack(r, x0, x1);
}

// GELU on a pair in its tanh form, 0.5 x (1 + tanh(0.79788456 (x + 0.044715 x^3))), five packed fp32x2 operations and
// two MUFU.TANH - half the issue slots of gelu_fast2. |tanh form - erf form| <= 4.8e-4 (rms 1.5e-4 on N(0,1.2) inputs)
// against 1.3e-3 rms for the bf16 rounding of the stored activation: the total error of the stored value grows by
// 0.7 %. Used where the GELU sits on the critical path between two tensor-core phases (fused encoder, enc_fused.cuh).
__device__ __forceinline__ void gelu_tanh2(float& x0, float& x1) {
  const uint64_t x = f2_pack(x0, x1);
  const uint64_t x2 = f2_mul(x, x);
  uint64_t u = f2_fma(x2, f2_pack(0.0356774081f, 0.0356774081f), f2_pack(0.7978845608f, 0.7978845608f));
  u = f2_mul(u, x);
  float u0, u1, t0, t1;
  f2_unpack(u, u0, u1);
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(u0));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(u1));
  const uint64_t h = f2_mul(x, f2_pack(0.5f, 0.5f));
  const uint64_t r = f2_fma(h, f2_pack(t0, t1), h);
  f2_unpack(r, x0, x1);
}

// Programmatic dependent launch (PDL). launch_dependents: the next kernel in the stream may start being
// scheduled; wait: block until the previous kernel has completed and its writes are visible. Both are
// no-ops for a kernel launched without the programmatic-serialization attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Non-blocking probe of a phase (try_wait may suspend for its hardware time window; this never does).
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap (sticky error, loud failure) instead of hanging the GPU.
// The bound is counted in failed try_waits (each suspends for the hardware's try_wait window), so
// the hot loop carries no clock reads and no printf/ERRBAR sequence.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > 200000000u) __trap();
  }
}

// ------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), 2-D tiled load global -> shared, completion on an mbarrier
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0,
                                            int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// 2-D tiled store shared -> global (bulk async group); out-of-range parts of the box are clipped.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() {   // <= N groups still reading their shared source
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_group() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp as alloc
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]; issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives TMEM lane (base_lane + i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// Store 32 consecutive fp32 columns of this thread's TMEM lane (inverse of tmem_ld_32x32).
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// Wait for outstanding tcgen05.ld and tie the destination registers to the wait, so the compiler
// cannot schedule a use of v[] above it.
__device__ __forceinline__ void tmem_ld_wait_dep(uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.wait::ld.sync.aligned;"
      : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
        "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
        "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
        "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
      :
      : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------------------------------
// CTA pairs (cluster of 2, cta_group::2): one MMA spans both SMs' tensor cores and shared memories
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {   // every thread of every CTA in the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// Arrive on a barrier of another CTA of the cluster. Default semantics (release at CTA scope), as CUTLASS's
// ClusterBarrier::arrive does: a `.release.cluster` arrive costs ~1.2k clk per call on B200 (measured in the
// fused encoder: it fences the whole cluster), and every payload it orders here is already fenced explicitly
// (tcgen05.fence / fence.proxy.async) before the arrive.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into this CTA's shared memory whose completion is signalled on a barrier that may live in
// the peer CTA (cluster address), as the leader's MMA thread is the only consumer.
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr,
                                                int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {  // one warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[256 rows: 128 from each CTA's smem] . B[N: N/2 rows from each CTA's smem]
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive (once all prior MMAs of this thread completed) on the barrier at this offset in BOTH CTAs.
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(mask)
      : "memory");
}

// Shared-memory matrix descriptor for a K-major bf16 tile stored by TMA with SWIZZLE_128B:
// rows of 64 elements (128 B), 8-row groups 1024 B apart (SBO), tile base 1024-B aligned.
// Field layout: cute/arch/mma_sm100_desc.hpp (SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type [61,64) with SWIZZLE_128B = 2.
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;               // LBO (unused for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;       // SBO = 8 rows * 128 B
  d |= static_cast<uint64_t>(1) << 46;               // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;               // SWIZZLE_128B
  return d;
}
// One lane of a fully active warp (always the same one), so a whole warp can run the issue loop with
// warp-uniform control flow and operands (kept in uniform registers) and only the issue is predicated.
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// Low word of make_sw128_kmajor_desc (start address >> 4, LBO = 1); the high word is the constant below.
__device__ __forceinline__ uint32_t sw128_desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
constexpr uint32_t kSw128DescHi = 0x40004040u;   // SBO = 1024 B, descriptor version 1, SWIZZLE_128B
// The four K = 16 steps of one 64-wide k-block (A and B advance 32 B per step: start field + 2), cta_group::2.
__device__ __forceinline__ void umma_kblock_2sm(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc_first) {
  asm volatile(
      "{\n\t"
      ".reg .pred p0, pt;\n\t"
      ".reg .b64 da, db;\n\t"
      ".reg .b32 ta, tb;\n\t"
      "setp.ne.b32 p0, %4, 0;\n\t"
      "setp.eq.b32 pt, %4, %4;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p0;\n\t"
      "add.u32 ta, %1, 2;\n\t"
      "add.u32 tb, %2, 2;\n\t"
      "mov.b64 da, {ta, %5};\n\t"
      "mov.b64 db, {tb, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, pt;\n\t"
      "add.u32 ta, %1, 4;\n\t"
      "add.u32 tb, %2, 4;\n\t"
      "mov.b64 da, {ta, %5};\n\t"
      "mov.b64 db, {tb, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, pt;\n\t"
      "add.u32 ta, %1, 6;\n\t"
      "add.u32 tb, %2, 6;\n\t"
      "mov.b64 da, {ta, %5};\n\t"
      "mov.b64 db, {tb, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, pt;\n\t"
      "}"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc_first), "r"(kSw128DescHi)
      : "memory");
}
// The four K = 16 steps of one 64-wide k-block (A and B advance 32 B per step: start field + 2), cta_group::1.
__device__ __forceinline__ void umma_kblock_1sm(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc_first) {
  asm volatile(
      "{\n\t"
      ".reg .pred p0, pt;\n\t"
      ".reg .b64 da, db;\n\t"
      ".reg .b32 ta, tb;\n\t"
      "setp.ne.b32 p0, %4, 0;\n\t"
      "setp.eq.b32 pt, %4, %4;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p0;\n\t"
      "add.u32 ta, %1, 2;\n\t"
      "add.u32 tb, %2, 2;\n\t"
      "mov.b64 da, {ta, %5};\n\t"
      "mov.b64 db, {tb, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n\t"
      "add.u32 ta, %1, 4;\n\t"
      "add.u32 tb, %2, 4;\n\t"
      "mov.b64 da, {ta, %5};\n\t"
      "mov.b64 db, {tb, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n\t"
      "add.u32 ta, %1, 6;\n\t"
      "add.u32 tb, %2, 6;\n\t"
      "mov.b64 da, {ta, %5};\n\t"
      "mov.b64 db, {tb, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n\t"
      "}"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc_first), "r"(kSw128DescHi)
      : "memory");
}
// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32, both operands K-major.
// (InstrDescriptor: c_format [4,6)=1 F32, a_format [7,10)=1 BF16, b_format [10,13)=1 BF16,
//  a_major bit 15 = 0, b_major bit 16 = 0, n>>3 [17,23), m>>4 [24,29).)
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

}  // namespace owc
