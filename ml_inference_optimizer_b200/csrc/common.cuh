// Shared sm_100a device helpers: mbarrier, TMA, tcgen05 / TMEM, descriptors.
// Everything here is inline PTX; there is no CUTLASS/CuTe dependency.
//
// Compile ONLY with  -gencode arch=compute_100a,code=sm_100a  (SURVEY.md §7 step 2):
// a generic compute_100 pass rejects every tcgen05 instruction.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#if defined(__CUDA_ARCH__) && !defined(__CUDA_ARCH_FEAT_SM100_ALL)
#error "b200 kernels must be compiled for sm_100a (-gencode arch=compute_100a,code=sm_100a)"
#endif

namespace b200 {

// ---------------------------------------------------------------------------------------------
// misc
// ---------------------------------------------------------------------------------------------
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

template <int kRegs>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(kRegs));
}
template <int kRegs>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(kRegs));
}

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// Same with the barrier id as an immediate: with a run-time id ptxas reserves all 16 hardware barriers for the CTA, which
// can keep another kernel's CTAs (the co-resident tensor-parallel all-reduce) off the SM.
template <int ID>
__device__ __forceinline__ void named_bar_sync_imm(uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"n"(ID), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
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

// Bounded wait: a protocol bug turns into a trap (reported as a CUDA error by the C-ABI) instead of a
// GPU hang. try_wait suspends for a HW time slice per call, so the bound is minutes of wall time
// only in the deadlock case and costs nothing otherwise.
#ifndef B200_MBAR_SPIN_LIMIT
#define B200_MBAR_SPIN_LIMIT (1u << 22)
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t spins = 0;
  while (true) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) break;
    if (++spins > B200_MBAR_SPIN_LIMIT) {
      printf("b200: mbarrier wait timeout block=(%d,%d,%d) thread=%d bar=%u parity=%u\n", blockIdx.x, blockIdx.y,
             blockIdx.z, threadIdx.x, addr, parity);
      __trap();
    }
  }
}

// Variants taking a 32-bit shared-window address. A hot loop that is short of registers makes the compiler
// rematerialise `smem_u32(ptr)` (generic->shared conversion: two S2R plus the alignment arithmetic) in front of every
// barrier operation; an address kept in one opaque register (`smem_addr_opaque`) plus a constant offset avoids that.
__device__ __forceinline__ uint32_t smem_addr_opaque(const void* p) {
  uint32_t a;
  asm volatile("mov.u32 %0, %1;" : "=r"(a) : "r"(smem_u32(p)));
  return a;
}
__device__ __forceinline__ void mbar_arrive_addr(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t addr, uint32_t parity) {
  uint32_t spins = 0;
  while (true) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) break;
    if (++spins > B200_MBAR_SPIN_LIMIT) {
      printf("b200: mbarrier wait timeout block=(%d,%d,%d) thread=%d bar=%u parity=%u\n", blockIdx.x, blockIdx.y,
             blockIdx.z, threadIdx.x, addr, parity);
      __trap();
    }
  }
}

// Wait that may legitimately last as long as a whole work item of a persistent CTA (milliseconds): back off between
// polls; the bound (seconds) still turns a protocol bug into a trap.
__device__ __forceinline__ void mbar_wait_long(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(64);
    if (++spins > (1u << 24)) {
      printf("b200: mbarrier long-wait timeout block=(%d,%d,%d) thread=%d parity=%u\n", blockIdx.x, blockIdx.y, blockIdx.z,
             threadIdx.x, parity);
      __trap();
    }
  }
}

// ---- cluster launch control (sm_100): a running CTA takes over the block index of a not-yet-launched CTA of its own
// grid. The 16-byte response lands in shared memory through the async proxy and completes 16 tx bytes on `bar`. ----
__device__ __forceinline__ void clc_try_cancel(void* response16, uint64_t* bar) {
  asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.b128 [%0], [%1];" ::"r"(
                   smem_u32(response16)),
               "r"(smem_u32(bar))
               : "memory");
}
// Returns true and the stolen block index when the request succeeded; false = no blocks left (do not ask again).
__device__ __forceinline__ bool clc_query(const void* response16, int& x, int& y, int& z) {
  uint32_t ok, rx, ry, rz;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      ".reg .b128 R;\n\t"
      "mov.u32 %0, 0;\n\t"
      "mov.u32 %1, 0;\n\t"
      "mov.u32 %2, 0;\n\t"
      "ld.shared.b128 R, [%4];\n\t"
      "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 P, R;\n\t"
      "selp.u32 %3, 1, 0, P;\n\t"
      "@P clusterlaunchcontrol.query_cancel.get_first_ctaid.v4.b32.b128 {%0, %1, %2, _}, R;\n\t"
      "}\n"
      : "=&r"(rx), "=&r"(ry), "=&r"(rz), "=&r"(ok)
      : "r"(smem_u32(response16))
      : "memory");
  x = static_cast<int>(rx);
  y = static_cast<int>(ry);
  z = static_cast<int>(rz);
  return ok != 0;
}

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) — loads complete on an mbarrier, stores use bulk groups
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* desc, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* desc, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* desc, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* desc, const void* smem_src, int c0, int c1, int c2,
                                             int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------------------------------------
// TMEM allocation (one warp allocates and the same warp frees)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// tcgen05.commit: arrive once on the mbarrier when all previously issued tcgen05.mma of this thread retire.
// (implicitly performs tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------------------------------------
// CTA-pair (cta_group::2) variants: the two CTAs of a cluster issue one MMA over both SMs' shared memory / TMEM
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load issued by either CTA of the pair; the transaction bytes are counted on the LEADER CTA's mbarrier (peer bit 24 of
// the shared::cluster address cleared)
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* desc, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
// same, with an L2 eviction-priority hint (createpolicy encodings as used by CUTLASS' TMA::CacheHintSm90)
constexpr uint64_t kL2EvictNormal = 0x1000000000000000ull;
constexpr uint64_t kL2EvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kL2EvictLast = 0x14F0000000000000ull;
__device__ __forceinline__ void tma_load_2d_2sm_hint(void* smem_dst, const CUtensorMap* desc, uint64_t* bar, int c0, int c1,
                                                     uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap* desc, const void* smem_src, int c0, int c1, uint64_t hint) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
               :
               : "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "l"(hint)
               : "memory");
}
__device__ __forceinline__ void umma_ss_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit: arrive on the mbarrier at the same shared-memory offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
// TMA load multicast to the CTAs of `cta_mask`: the box lands at the same shared-memory offset in each of them and each
// CTA's mbarrier (same offset) receives the transaction bytes delivered to it
__device__ __forceinline__ void tma_load_4d_mcast(void* smem_dst, const CUtensorMap* desc, uint64_t* bar, int c0, int c1, int c2,
                                                  int c3, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5, %6}], "
      "[%2], %7;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
        "h"(cta_mask)
      : "memory");
}
// tcgen05.commit of a cta_group::1 MMA stream that arrives on the mbarrier at the same offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_mcast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
// arrive on the mbarrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

// ---------------------------------------------------------------------------------------------
// UMMA descriptors
// ---------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, 128-byte swizzle (layout_type 2), Blackwell version bit 46.
//   K-major operand (rows of 64 bf16 = 128 B, 8-row swizzle atoms stacked every 1024 B):
//       lbo ignored (1), sbo = 1024
//   MN-major operand (128-B rows hold 64 MN-elements of one K index; 8 K-rows per 1024-B atom):
//       lbo = byte distance between 64-element MN chunks, sbo = 1024 (next 8 K rows)
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version (sm_100)
  d |= 2ull << 61;  // SWIZZLE_128B
  return d;
}

// Instruction descriptor for kind::f16 (bf16 or fp16 inputs, fp32 accumulate).
//   bits [4,6) c_format=1 (f32); [7,10) a_format; [10,13) b_format (0=f16, 1=bf16);
//   bit 15 a_major, bit 16 b_major (0 = K-major, 1 = MN-major); [17,23) N>>3; [24,29) M>>4.
__host__ __device__ constexpr uint32_t make_idesc_f16_ab(int m, int n, bool a_is_bf16, bool b_is_bf16, bool a_mn_major,
                                                         bool b_mn_major) {
  return (1u << 4) | ((a_is_bf16 ? 1u : 0u) << 7) | ((b_is_bf16 ? 1u : 0u) << 10) | ((a_mn_major ? 1u : 0u) << 15) |
         ((b_mn_major ? 1u : 0u) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n, bool is_bf16, bool a_mn_major, bool b_mn_major) {
  return make_idesc_f16_ab(m, n, is_bf16, is_bf16, a_mn_major, b_mn_major);
}

// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// TMEM <-> registers. Shape 32x32b: thread i of the warp owns TMEM lane (32*(warp%4) + i) and receives
// N consecutive 32-bit columns.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
        "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]),
        "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// numeric helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x for a pair of values on the FMA pipe (no MUFU): round-to-nearest split x = n + f, |f| <= 0.5, cubic minimax
// polynomial for 2^f (max relative error 7.7e-5, far below the 16-bit rounding of the result's consumer), exponent
// inserted with an integer shift-add. x is clamped at -125, so -inf maps to 2^-125 (~2e-38), not to an exact zero.
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  const float kMagic = 12582912.0f;  // 1.5 * 2^23: adding it leaves round(x) in the low mantissa bits
  x.x = fmaxf(x.x, -125.0f);
  x.y = fmaxf(x.y, -125.0f);
  const float2 t = __fadd2_rn(x, make_float2(kMagic, kMagic));
  const float2 n = __fadd2_rn(t, make_float2(-kMagic, -kMagic));
  const float2 f = __ffma2_rn(n, make_float2(-1.0f, -1.0f), x);
  float2 p = __ffma2_rn(f, make_float2(0.05508868396282196f, 0.05508868396282196f),
                        make_float2(0.24260404706001282f, 0.24260404706001282f));
  p = __ffma2_rn(p, f, make_float2(0.6932762265205383f, 0.6932762265205383f));
  p = __ffma2_rn(p, f, make_float2(0.9999289512634277f, 0.9999289512634277f));
  float2 e;
  e.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
  e.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
  return e;
}
__device__ __forceinline__ float fast_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <typename T>
struct Pack2;
template <>
struct Pack2<__nv_bfloat16> {
  static constexpr bool kIsBf16 = true;
  // low 16 bits <- lo, high 16 bits <- hi
  __device__ __forceinline__ static uint32_t pack(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
  }
  __device__ __forceinline__ static float2 unpack(uint32_t v) {
    return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xFFFF0000u));
  }
};
template <>
struct Pack2<__half> {
  static constexpr bool kIsBf16 = false;
  __device__ __forceinline__ static uint32_t pack(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
  }
  __device__ __forceinline__ static float2 unpack(uint32_t v) {
    __half2 h = *reinterpret_cast<__half2*>(&v);
    return __half22float2(h);
  }
};

}  // namespace b200
