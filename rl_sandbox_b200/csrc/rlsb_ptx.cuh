// rlsb_ptx.cuh — thin inline-PTX wrappers for sm_100a (tcgen05 / TMEM / bulk-TMA / mbarrier).
//
// Everything here is hand-written PTX; there is no CUTLASS/CuTe dependency.  Encodings
// (shared-memory matrix descriptor, instruction descriptor) follow the PTX ISA tables for
// tcgen05.mma kind::f16.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace rlsb {

// ----------------------------------------------------------------------------------------
// shared-memory address helpers
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// explicit shared-space accesses (a pointer derived through integer arithmetic from the dynamic
// shared buffer loses its address space and compiles to slow generic LD/ST)
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint4 lds128u(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
// 16-byte asynchronous copy global -> shared (LDGSTS, L2 only), completion tracked per thread by cp.async groups
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ float2 lds64(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float lds32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts64(uint32_t a, float x, float y) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, float x) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(x) : "memory");
}

// ----------------------------------------------------------------------------------------
// packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2 on sm_100): two IEEE fp32 lanes per instruction.
// The GEMM epilogues are instruction-issue bound; pairing the LayerNorm / bias / affine math halves
// their FP instruction count without changing a single result bit.
// ----------------------------------------------------------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f32x2 pk2u(uint32_t lo, uint32_t hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void unpk2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fadd2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fmul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// 4 consecutive floats from shared memory as two fp32x2 pairs
__device__ __forceinline__ void lds_2x2(uint32_t a, f32x2& p0, f32x2& p1) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(p0), "=l"(p1) : "r"(a) : "memory");
}

// ----------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ----------------------------------------------------------------------------------------
// bulk async copy (TMA engine, non-tensor form): global -> shared, completes on an mbarrier
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// multicast variant: the bytes land at the same CTA-relative offset in every CTA of `cta_mask`
// and complete_tx is signalled on each destination CTA's mbarrier at the same offset.
__device__ __forceinline__ void bulk_g2s_multicast(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                                   uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
// shared -> global bulk store (bulk-group completion)
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ----------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA, commit, load
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// D[tmem] (+)= A[smem] * B[smem]; kind::f16 (bf16 inputs, fp32 accumulate), issued by one thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// same, arriving on the barrier at this offset in every CTA of `cta_mask` (cluster multicast)
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
// ---- CTA pair (cta_group::2): one MMA over two SMs; the even CTA of the pair (cluster rank 0) issues it ----------
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {   // one warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D (256 x N: rows 0-127 in this CTA's TMEM, 128-255 in the peer's) (+)= A * B with A = 128 rows per CTA and
// B = N/2 rows per CTA, each read from the SAME shared-memory offset in both CTAs.
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// arrive (once the pair's MMAs issued so far have completed) on the barrier at this offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit2_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
// shared::cluster address of `smem_addr` (a shared::cta address) in CTA `cta` of the cluster
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t smem_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// A local barrier whose arrivals come from the peer CTA is waited on with the plain (CTA-scope) mbar_wait: what the
// waiter goes on to touch is the peer's shared memory through the tensor core (async proxy) or TMEM, never its own L1,
// and a cluster-scope acquire per k-step in the MMA issue loop costs more than the four MMAs it guards.

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// TMEM -> registers: 32 lanes x 32 consecutive 32-bit columns (thread i gets lane base+i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
// registers -> TMEM (the mirror image of tmem_ld8): an epilogue parks an intermediate in the accumulator's own columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// the same wait, with the destination registers of the loads it completes tied to it: nothing that reads them can be
// scheduled above the wait while OTHER loads (issued after it, into other registers) are in flight
__device__ __forceinline__ void tmem_ld_wait2(uint32_t (&a)[8], uint32_t (&b)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]),
                 "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7])
               :
               : "memory");
}
// software-pipelined sweep over the 8-column chunks cq, cq + 4, ... of a thread's TMEM row: two chunks are processed
// while the next two are in flight (the wait then finds its loads complete).  f(r, i) handles chunk index cq + 4 i.
template <class F>
__device__ __forceinline__ void tmem_sweep(uint32_t tmem_d, int cq, int n_chunks, F&& f) {
  uint32_t a0[8] = {}, a1[8] = {}, b0[8] = {}, b1[8] = {};
  auto issue = [&](int i, uint32_t (&x)[8], uint32_t (&y)[8]) {
    if (i < n_chunks) tmem_ld8(tmem_d + static_cast<uint32_t>((cq + 4 * i) * 8), x);
    if (i + 1 < n_chunks) tmem_ld8(tmem_d + static_cast<uint32_t>((cq + 4 * (i + 1)) * 8), y);
  };
  issue(0, a0, a1);
  for (int i = 0; i < n_chunks; i += 4) {
    tmem_ld_wait2(a0, a1);
    issue(i + 2, b0, b1);
    f(a0, i);
    if (i + 1 < n_chunks) f(a1, i + 1);
    if (i + 2 >= n_chunks) break;
    tmem_ld_wait2(b0, b1);
    issue(i + 4, a0, a1);
    f(b0, i + 2);
    if (i + 3 < n_chunks) f(b1, i + 3);
  }
}

// the same sweep for a run of chunks that needs no bounds checks: groups of four chunks with their position inside the
// group known at compile time (f(r, std::integral_constant<int, J>) handles chunk 4 g + J of this warp, next_group() moves
// the caller's pointers on), so that all per-chunk addresses are a base register plus an immediate.  t0 = TMEM address of
// the warp's first chunk; consecutive chunks of a warp are 32 columns apart.  Returns the number of chunks handled
// (a multiple of four); the caller finishes the remaining n % 4 chunks.
template <int J>
struct ChunkPos { static constexpr int value = J; };
template <class F, class G>
__device__ __forceinline__ int tmem_sweep_groups(uint32_t t0, int n, F&& f, G&& next_group) {
  uint32_t a0[8] = {}, a1[8] = {}, b0[8] = {}, b1[8] = {};
  const int groups = n >> 2;
  if (groups > 0) {
    tmem_ld8(t0, a0);
    tmem_ld8(t0 + 32u, a1);
  }
  uint32_t t = t0;
  for (int g = 0; g < groups; ++g) {
    tmem_ld_wait2(a0, a1);
    tmem_ld8(t + 64u, b0);
    tmem_ld8(t + 96u, b1);
    f(a0, ChunkPos<0>{});
    f(a1, ChunkPos<1>{});
    tmem_ld_wait2(b0, b1);
    if (g + 1 < groups) {
      tmem_ld8(t + 128u, a0);
      tmem_ld8(t + 160u, a1);
    }
    f(b0, ChunkPos<2>{});
    f(b1, ChunkPos<3>{});
    t += 128u;
    next_group();
  }
  return groups << 2;
}

// tmem_sweep_groups with a second, global-memory operand per chunk (16 bytes per thread): gload(j) fetches the operand of
// chunk 4 g + j of the current group (j = 4, 5: first pair of the next group) two chunks ahead of its use, next to the TMEM
// load of the same chunk, so that neither latency is exposed per chunk.  f(r, bits, ChunkPos<J>, t): t = TMEM address of
// the group's first chunk.
template <class P, class F, class G>
__device__ __forceinline__ int tmem_sweep_groups_g(uint32_t t0, int n, P&& gload, F&& f, G&& next_group) {
  uint32_t a0[8] = {}, a1[8] = {}, b0[8] = {}, b1[8] = {};
  uint4 pa0 = make_uint4(0u, 0u, 0u, 0u), pa1 = pa0, pb0 = pa0, pb1 = pa0;
  const int groups = n >> 2;
  if (groups > 0) {
    tmem_ld8(t0, a0);
    tmem_ld8(t0 + 32u, a1);
    pa0 = gload(0);
    pa1 = gload(1);
  }
  uint32_t t = t0;
  for (int g = 0; g < groups; ++g) {
    tmem_ld_wait2(a0, a1);
    tmem_ld8(t + 64u, b0);
    tmem_ld8(t + 96u, b1);
    pb0 = gload(2);
    pb1 = gload(3);
    f(a0, pa0, ChunkPos<0>{}, t);
    f(a1, pa1, ChunkPos<1>{}, t);
    tmem_ld_wait2(b0, b1);
    if (g + 1 < groups) {
      tmem_ld8(t + 128u, a0);
      tmem_ld8(t + 160u, a1);
      pa0 = gload(4);
      pa1 = gload(5);
    }
    f(b0, pb0, ChunkPos<2>{}, t);
    f(b1, pb1, ChunkPos<3>{}, t);
    t += 128u;
    next_group();
  }
  return groups << 2;
}

// ----------------------------------------------------------------------------------------
// descriptors
// ----------------------------------------------------------------------------------------
// Shared-memory matrix descriptor for a K-major operand stored as rows of 128 bytes (64 bf16)
// with the 128-byte swizzle: 8-row core groups are 1024 bytes apart (SBO), version = 1 (sm_100),
// layout_type = 2 (SWIZZLE_128B).  The tile base must be 1024-byte aligned; stepping K by 16
// elements inside the 128-byte row is done by adding 32 bytes to the start address.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);        // start address   [0,14)
  d |= static_cast<uint64_t>(1) << 16;                        // LBO (ignored for swizzled K-major)
  d |= static_cast<uint64_t>(1024u >> 4) << 32;               // SBO = 1024 B    [32,46)
  d |= static_cast<uint64_t>(1) << 46;                        // version = 1     [46,48)
  d |= static_cast<uint64_t>(2) << 61;                        // SWIZZLE_128B    [61,64)
  return d;
}

// Instruction descriptor, kind::f16: D=f32, A=B=bf16, both K-major, M x N tile.
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t M, uint32_t N) {
  return (1u << 4)                 // c_format  = F32
         | (1u << 7)               // a_format  = BF16
         | (1u << 10)              // b_format  = BF16
         | (0u << 15) | (0u << 16) // a_major = b_major = K
         | ((N >> 3) << 17)        // n_dim
         | ((M >> 4) << 24);       // m_dim
}

// MN-major operand (the contraction index is the ROW index of the stored image): rows of 128 bytes
// hold 64 consecutive M/N elements, 8-row groups (8 contraction indices) are `SBO` = 1024 bytes
// apart, the next 64 M/N elements start `lbo_bytes` further on; SWIZZLE_128B as above.  Stepping
// the contraction by 16 rows = +2048 bytes on the start address.
__device__ __forceinline__ uint64_t make_smem_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;   // LBO: stride between 64-element M/N groups
  d |= static_cast<uint64_t>(1024u >> 4) << 32;                   // SBO: stride between 8-row K groups
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// kind::f16 instruction descriptor with both operands MN-major
__host__ __device__ constexpr uint32_t make_idesc_bf16_mn(uint32_t M, uint32_t N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ----------------------------------------------------------------------------------------
// packed ("UMMA-ready") global layout shared by every bf16 operand in this library.
//
// A logical [rows x K] bf16 matrix (K % 64 == 0, rows % RB == 0) is stored as tiles of
// RB rows x 64 columns; tiles are ordered (row_block, k_tile) with k_tile fastest; inside a
// tile row r occupies bytes [r*128, r*128+128) and its 16-byte chunk c (8 elements) is stored
// at chunk position c ^ (r & 7) — exactly the image the SWIZZLE_128B descriptor expects, so a
// tile is moved global->shared with ONE contiguous cp.async.bulk.
// ----------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ size_t packed_index(size_t row, size_t k, size_t K, int RB) {
  size_t rb = row / RB, r = row % RB;
  size_t kt = k >> 6, kk = k & 63;
  size_t chunk = (kk >> 3) ^ (r & 7);
  return ((rb * (K >> 6) + kt) * RB + r) * 64 + chunk * 8 + (kk & 7);
}

}  // namespace rlsb
