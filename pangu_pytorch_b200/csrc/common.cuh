// Thin inline-PTX layer for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA,
// TMEM alloc/ld, commit, fences) and a few packed-conversion helpers.  Everything here is
// hand-written against the PTX ISA; nothing is included from CUTLASS/CuTe.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace pg {

#ifndef PANGU_SPIN_LIMIT
// A mis-counted mbarrier would otherwise hang the GPU box; trap instead (≈ seconds).
#define PANGU_SPIN_LIMIT (1u << 26)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  // generic-proxy writes to smem -> visible to the async proxy (TMA / tcgen05.mma reads)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread for a system-dependent time: never use it to poll two barriers)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait with a suspend-time hint (ns): the warp is parked by the hardware until the phase completes or the time is up, and
// issues nothing meanwhile -- a "sleep that wakes up early", used where a warp polls several barriers
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    ++spins;
#ifdef PANGU_DEV_SWITCHES      // development builds: say who is stuck (the printf marshalling costs ~35 instructions per call site)
    if (spins == PANGU_SPIN_LIMIT && (blockIdx.x == 5 || threadIdx.x == 0))      // report, keep spinning so that every stuck role gets to report, then trap
      printf("pangu_b200: mbarrier timeout block %d thread %d smem 0x%x parity %u\n", (int)blockIdx.x,
             (int)threadIdx.x, smem_u32(bar), parity);
#endif
    if (spins > PANGU_SPIN_LIMIT + (PANGU_SPIN_LIMIT >> 1)) __trap();      // a mis-counted barrier must not hang the GPU
  }
}

// ------------------------------------------------------------------ TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* smem, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(const CUtensorMap* m, uint64_t* bar, void* smem, int c0, int c1,
                                                 uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
// bring one box of a tensor map into L2 (no shared-memory destination, no completion signal)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1) : "memory");
}
// ------------------------------------------------------------------ thread-block clusters
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// distributed shared memory: address of the same smem location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t raddr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t rbar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
}
// "slot free" style signal to a peer CTA: the reads it covers have already been consumed by this thread, so no
// release fence (the .release form costs a CCTL.IVALL + ERRBAR pair per arrive)
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t rbar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
}
// 16-byte store into a peer CTA's shared memory that signals complete_tx(16) on the peer's mbarrier when it lands:
// data and notification travel together, no fence on either side
__device__ __forceinline__ void st_async_cluster_f4(uint32_t raddr, float a, float b, float c, float d, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(raddr), "f"(a), "f"(b), "f"(c), "f"(d), "r"(rbar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0, ok = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) break;
    if (++spins > PANGU_SPIN_LIMIT) { printf("pangu_b200: cluster mbarrier timeout\n"); __trap(); }
  }
}
// TMA load multicast to every CTA in `mask`: the tile lands at the same smem offset in each
// destination CTA and complete_tx is signalled on the mbarrier at the same offset in each of them.
__device__ __forceinline__ void tma_load_2d_mcast(const CUtensorMap* m, uint64_t* bar, void* smem, int c0, int c1,
                                                  uint16_t mask, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5, %6;"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask),
        "l"(hint)
      : "memory");
}
// tcgen05.commit arriving on the mbarrier at this smem offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mcast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}

// ------------------------------------------------------------------ CTA pairs (tcgen05 cta_group::2)
// One tcgen05.mma.cta_group::2 spans the two SMs of a cluster pair: M = 256 (rows 0..127 from the A tile and into the TMEM of
// CTA 0, rows 128..255 from / into CTA 1), and each CTA holds HALF of the B tile (N/2 rows) at the same shared-memory offset.
// Only CTA 0 issues the MMA; both CTAs load their tiles with the cta_group::2 form of the TMA load, whose completion bytes
// are counted on CTA 0's mbarrier.
__device__ __forceinline__ void tma_load_2d_cta2(const CUtensorMap* m, uint32_t leader_bar, void* smem, int c0, int c1,
                                                 uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_alloc_cta2(uint32_t* dst_smem) {      // one warp (same warp index) in EACH CTA of the pair
  static_assert(kCols == 32 || kCols == 64 || kCols == 128 || kCols == 256 || kCols == 512, "pow2 >= 32");
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc_cta2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(kCols) : "memory");
}
__device__ __forceinline__ void umma_f16_ss_cta2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all prior tcgen05 ops of this thread -> one arrival on the mbarrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_cta2(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(uint16_t(3))
               : "memory");
}

// smem tile -> global (bulk async group); out-of-bounds rows/columns of the box are clipped
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// bring [p, p + bytes) into L2 without touching shared memory (bytes % 16 == 0)
__device__ __forceinline__ void bulk_prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// L2 eviction-priority descriptors (createpolicy.fractional encodings used by CUTLASS)
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

// ------------------------------------------------------------------ tcgen05 / TMEM
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  static_assert(kCols == 32 || kCols == 64 || kCols == 128 || kCols == 256 || kCols == 512, "pow2 >= 32");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16 (bf16 or fp16 operands, fp32 accumulate).
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Make the mbarrier track completion of all prior tcgen05 ops of this thread
// (implicitly performs tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// Instruction descriptor for kind::f16: fp32 accumulate, K-major A and B.
//  [4,6) c_format=1 (F32)  [7,10) a_format  [10,13) b_format (0=F16, 1=BF16)
//  [15] a_major=0 (K)  [16] b_major=0 (K)  [17,23) N>>3  [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, bool fp16) {
  return (1u << 4) | ((fp16 ? 0u : 1u) << 7) | ((fp16 ? 0u : 1u) << 10) | (uint32_t(N >> 3) << 17) |
         (uint32_t(M >> 4) << 24);
}

// Shared-memory matrix descriptor, K-major operand in the canonical SWIZZLE_128B layout
// produced by a TMA box of 64 16-bit elements x rows: row r at byte r*128 (XOR-swizzled in
// 16 B chunks within each 1024 B group of 8 rows).
//  [0,14) start>>4   [16,30) LBO>>4 (ignored for swizzled K-major; 1)   [32,46) SBO>>4 = 1024>>4
//  [46,48) version=1 (sm_100)   [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_sdesc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFF) >> 4);
  d |= uint64_t(1) << 16;
  d |= uint64_t(1024 >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}

// TMEM -> registers: 32 lanes x 32-bit, N consecutive columns; thread i of the warp reads
// lane (base_lane + i).  The warp may only touch lanes [32*(warp_id%4), +32).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------ programmatic dependent launch
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its predecessor in the
// stream is still running; pdl_wait() blocks until the predecessor grid has completed and its writes are visible,
// and must precede every access to global memory other than kernel parameters.  pdl_trigger() lets the successor's
// CTAs be scheduled as this grid's CTAs retire (only used by persistent kernels whose CTAs are all resident).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ------------------------------------------------------------------ misc
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// two fp32 -> packed 16-bit pair (lo = a, hi = b)
template <bool kFp16>
__device__ __forceinline__ uint32_t pack16(float a, float b) {
  uint32_t r;
  if constexpr (kFp16) {
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  } else {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  }
  return r;
}
template <bool kFp16>
__device__ __forceinline__ float unpack16_lo(uint32_t v) {
  if constexpr (kFp16) return __half2float(__ushort_as_half(uint16_t(v & 0xFFFF)));
  return __uint_as_float(v << 16);
}
template <bool kFp16>
__device__ __forceinline__ float unpack16_hi(uint32_t v) {
  if constexpr (kFp16) return __half2float(__ushort_as_half(uint16_t(v >> 16)));
  return __uint_as_float(v & 0xFFFF0000u);
}
template <bool kFp16>
__device__ __forceinline__ uint16_t cvt16(float a) {
  if constexpr (kFp16) return __half_as_ushort(__float2half_rn(a));
  return __bfloat16_as_ushort(__float2bfloat16_rn(a));
}

// ------------------------------------------------------------------ packed fp32x2 math (sm_100: FFMA2 / FADD2 / FMUL2)
typedef uint64_t f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Exact-erf GELU (nn.GELU default, reference models/layers.py:261), branch free:
//   gelu(x) = max(x, 0) - |x| * 0.5 erfc(|x| / sqrt(2)),     0.5 erfc(a / sqrt(2)) = 2^(a * P4(a) - 1),  a = |x|
// P4 is a degree-4 minimax fit (weighted by the error it causes in gelu) of log2(erfc(a / sqrt(2))) / a on [0, 6]; its
// leading coefficient is negative and P4 <= -1.15 everywhere, so the exponent only gets more negative for large |x| and
// no clamp is needed.  max |gelu - exact| = 5.3e-7 in exact arithmetic, 7e-7 with fp32 Horner + ex2.approx -- far below
// the 16-bit rounding of the stored activation.  On a pair the chain is 6 packed FFMA2 (4 Horner steps, the exponent, the
// final multiply-add) + 2 MUFU + 2 FMNMX: 5 issue slots per element (the degree-7 fit in |x| / sqrt(2) it replaces took 9).
constexpr float kGeluC0 = -1.151000543e+00f, kGeluC1 = -4.595958433e-01f, kGeluC2 = -5.214663275e-02f,
                kGeluC3 = 7.198718708e-03f, kGeluC4 = -4.881021609e-04f;
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
  const f32x2 ax = pack2(fabsf(x0), fabsf(x1));
  f32x2 p = fma2(ax, pack2(kGeluC4, kGeluC4), pack2(kGeluC3, kGeluC3));
  p = fma2(p, ax, pack2(kGeluC2, kGeluC2));
  p = fma2(p, ax, pack2(kGeluC1, kGeluC1));
  p = fma2(p, ax, pack2(kGeluC0, kGeluC0));
  float e0, e1;
  unpack2(fma2(p, ax, pack2(-1.0f, -1.0f)), e0, e1);
  const f32x2 u = pack2(-ex2_approx(e0), -ex2_approx(e1));
  const f32x2 r = fma2(ax, u, pack2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
  unpack2(r, x0, x1);
}

__device__ __forceinline__ float gelu_erf_libm(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// Eight elements at a time with the four pairs' chains interleaved step by step: one pair's chain is 8 dependent
// instructions (~90 clk), and left to itself ptxas overlaps only two of them (0.2 IPC in the GELU warps of the Mlp kernels).
__device__ __forceinline__ void gelu_erf8(float (&v)[8]) {
  f32x2 ax[4], p[4], m[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    ax[q] = pack2(fabsf(v[2 * q]), fabsf(v[2 * q + 1]));
    m[q] = pack2(fmaxf(v[2 * q], 0.f), fmaxf(v[2 * q + 1], 0.f));
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) p[q] = fma2(ax[q], pack2(kGeluC4, kGeluC4), pack2(kGeluC3, kGeluC3));
#pragma unroll
  for (int q = 0; q < 4; ++q) p[q] = fma2(p[q], ax[q], pack2(kGeluC2, kGeluC2));
#pragma unroll
  for (int q = 0; q < 4; ++q) p[q] = fma2(p[q], ax[q], pack2(kGeluC1, kGeluC1));
#pragma unroll
  for (int q = 0; q < 4; ++q) p[q] = fma2(p[q], ax[q], pack2(kGeluC0, kGeluC0));
#pragma unroll
  for (int q = 0; q < 4; ++q) p[q] = fma2(p[q], ax[q], pack2(-1.0f, -1.0f));
  float e[8];
#pragma unroll
  for (int q = 0; q < 4; ++q) unpack2(p[q], e[2 * q], e[2 * q + 1]);
#pragma unroll
  for (int i = 0; i < 8; ++i) e[i] = -ex2_approx(e[i]);
#pragma unroll
  for (int q = 0; q < 4; ++q) unpack2(fma2(ax[q], pack2(e[2 * q], e[2 * q + 1]), m[q]), v[2 * q], v[2 * q + 1]);
}

// scalar form of gelu_erf2 (same polynomial, same result)
__device__ __forceinline__ float gelu_erf(float x) {
  const float ax = fabsf(x);
  float p = fmaf(ax, kGeluC4, kGeluC3);
  p = fmaf(p, ax, kGeluC2);
  p = fmaf(p, ax, kGeluC1);
  p = fmaf(p, ax, kGeluC0);
  const float u = ex2_approx(fmaf(p, ax, -1.0f));
  return fmaf(-ax, u, fmaxf(x, 0.f));
}

// streaming 16-byte global accesses
__device__ __forceinline__ uint4 ldg_nc16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
// coherent 16-byte load with a memory clobber: may not be moved across barriers by the scheduler
__device__ __forceinline__ uint4 ldg16(const void* p) {
  uint4 r;
  asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void stg16(void* p, const uint4& v) {
  asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

}  // namespace pg
