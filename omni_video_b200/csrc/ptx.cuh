// Thin inline-PTX wrappers for the sm_100a features the hot-path kernels use:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld / st / fences).
// Nothing here is generic: every wrapper is the exact form one of our kernels issues.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ovb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// Arrive from the lanes where `pred` holds and return (the low word of) the barrier's state token,
// 0 elsewhere.  The token is how a caller ORDERS later arithmetic after the arrival: ptxas moves
// register-only instructions (MUFU, FFMA) freely above an mbarrier.arrive, but not above the
// producer of a value they consume.
__device__ __forceinline__ uint32_t mbar_arrive_token(uint64_t* bar, bool pred) {
  uint32_t tok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 st;\n"
      "setp.ne.u32 p, %2, 0;\n"
      "mov.b64 st, 0;\n"
      "@p mbarrier.arrive.shared::cta.b64 st, [%1];\n"
      "cvt.u32.u64 %0, st;\n"
      "}\n"
      : "=r"(tok)
      : "r"(smem_u32(bar)), "r"(static_cast<uint32_t>(pred))
      : "memory");
  return tok;
}
// Suspend-time hint (ns): a waiting thread stays parked in hardware until the phase completes or
// this much time has passed, instead of returning early and re-issuing the poll loop on a
// scheduler it shares with working warps.
#ifndef OVB_MBAR_HINT_NS
#define OVB_MBAR_HINT_NS 20000
#endif
constexpr uint32_t kMbarSuspendHintNs = OVB_MBAR_HINT_NS;
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(kMbarSuspendHintNs)
      : "memory");
  return ok != 0;
}
// Non-blocking probe.  Issue it EARLY and consume the result late: the ~100-150 cycle round trip of
// the probe then overlaps whatever the thread issues in between (a tcgen05.mma issuer must never sit
// idle behind a barrier instruction: the MMA queue holds only a couple of instructions).
__device__ __forceinline__ uint32_t mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
// Bounded spin: a protocol bug must surface as a trap (-> CUDA error on the host), never as a
// hang of the GPU box.  2^28 polls of a HW-sleeping try_wait is far beyond any legitimate wait.
// Build with -DOVB_DEBUG_SYNC to get a printf naming the barrier before the trap.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 28)) {
#ifdef OVB_DEBUG_SYNC
      printf("ovb: mbarrier timeout block=(%d,%d,%d) thread=%d bar=%u parity=%u\n", blockIdx.x,
             blockIdx.y, blockIdx.z, threadIdx.x, smem_u32(bar), parity);
#endif
      __trap();
    }
  }
}

// Busy poll with the non-suspending probe: for the ONE thread whose reaction time is on the kernel's
// critical path (the tcgen05.mma issuer).  A suspended try_wait was measured to resume ~400 cycles
// after the phase completed; a probe round trip is ~150.
__device__ __forceinline__ void mbar_spin(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_test(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}

// ---------------------------------------------------------------- clusters / CTA pairs
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
// all threads of all CTAs of the cluster
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n"
               "barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// shared::cluster address of `saddr` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr)
               : "memory");
}

// ---------------------------------------------------------------- fences
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];\n" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
      "r"(c4)
      : "memory");
}

// Multicast form: the box lands at the same CTA-relative offset in the shared memory of every CTA
// of `cta_mask`, and each of those CTAs' mbarrier (same offset) receives the transaction bytes.
__device__ __forceinline__ void tma_load_3d_mcast(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                                  int c1, int c2, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;\n" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(cta_mask)
      : "memory");
}

// TMA store of a shared-memory box (bulk async group completion)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() {
  asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}
// at most kPending of this thread's committed bulk groups may still be READING shared memory
template <int kPending>
__device__ __forceinline__ void bulk_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(kPending) : "memory");
}
template <int kPending>
__device__ __forceinline__ void bulk_wait_group() {
  asm volatile("cp.async.bulk.wait_group %0;\n" ::"n"(kPending) : "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

// CTA-pair (cta_group::2) loads: the data lands in the executing CTA's shared memory, the
// transaction bytes are signalled on `bar_cluster_addr`, a shared::cluster mbarrier address that
// may belong to the peer CTA (the pair's MMA leader).
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr,
                                                 int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];\n" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr,
                                                 int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr,
                                                 int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ---------------------------------------------------------------- TMEM alloc
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                   smem_u32(dst_smem)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "n"(kCols)
               : "memory");
}

// CTA-pair forms: one warp of EACH CTA of the pair executes alloc / dealloc (same warp id, same
// destination offset); the allocation covers the same columns in both SMs.
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                   smem_u32(dst_smem)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "n"(kCols)
               : "memory");
}

// ---------------------------------------------------------------- UMMA descriptors
// Shared-memory matrix descriptor (sm_100 "version 1"), 128-byte swizzle.
//   bits [ 0,14) start address >> 4        bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4   bits [46,48) version = 1
//   bits [61,64) layout type (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes,
                                                    uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Same descriptor for the narrower swizzle modes (K-major tiles whose rows are 64 or 32 bytes):
// layout type 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B; SBO = 8 rows * row bytes.
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr, uint32_t row_bytes) {
  const uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(((8u * row_bytes) >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout) << 61;
  return d;
}

// Instruction descriptor for kind::f16 with bf16 A/B and fp32 accumulate.
//   c_format[4,6)=1(F32)  a_format[7,10)=1(BF16)  b_format[10,13)=1(BF16)
//   a_major bit15, b_major bit16 (0 = K-major, 1 = MN-major)
//   n_dim[17,23) = N>>3   m_dim[24,29) = M>>4
__host__ __device__ constexpr uint32_t umma_idesc_bf16(uint32_t M, uint32_t N, uint32_t a_mn_major,
                                                       uint32_t b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn_major << 15) | (b_mn_major << 16) |
         ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// CTA-pair MMA (issued by the leader CTA only): M = 256 rows split over the two SMs (128 TMEM
// lanes each), every SM contributes its own A rows and its half of the B rows (N/2) from the same
// shared-memory offsets.
__device__ __forceinline__ void umma_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier at the same shared-memory offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::
          "r"(smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed.
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::
                   "r"(smem_u32(bar))
               : "memory");
}

// 1-CTA MMAs, but the arrival is multicast: the mbarrier at the same shared-memory offset in every
// CTA of `cta_mask` receives one arrival (used where two CTAs share a multicast-loaded stage).
__device__ __forceinline__ void umma_commit_mcast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::
          "r"(smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

// ---------------------------------------------------------------- TMEM <-> registers
// 32x32b: lane i of the warp reads/writes TMEM lane (base_lane + i), N consecutive 32-bit columns.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
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
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(
          taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
      "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
      "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
      "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
// Compiler-level fence on 32 registers: an empty volatile asm that "rewrites" them, so that no use
// of these values can be scheduled above the preceding volatile asm (a tcgen05.wait::ld whose
// loads were issued earlier and are still landing in exactly these registers).  Emits nothing.
__device__ __forceinline__ void reg_fence32(float* r) {
  asm volatile("" : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]), "+f"(r[8]), "+f"(r[9]), "+f"(r[10]), "+f"(r[11]), "+f"(r[12]), "+f"(r[13]), "+f"(r[14]), "+f"(r[15]), "+f"(r[16]), "+f"(r[17]), "+f"(r[18]), "+f"(r[19]), "+f"(r[20]), "+f"(r[21]), "+f"(r[22]), "+f"(r[23]), "+f"(r[24]), "+f"(r[25]), "+f"(r[26]), "+f"(r[27]), "+f"(r[28]), "+f"(r[29]), "+f"(r[30]), "+f"(r[31]));
}
__device__ __forceinline__ void reg_fence32u(uint32_t* r) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31]));
}
__device__ __forceinline__ void tmem_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}

// ---------------------------------------------------------------- misc math
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// bf16x2 pack on the integer ALU pipe: add half an ulp and keep the high halves (round to nearest,
// ties away from zero: differs from F2FP's ties-to-even only on exact ties, 2^-16 of the inputs;
// valid for finite values).  F2FP shares the XU pipe (4 lanes/clk per scheduler) with MUFU.EX2,
// and the softmax loop is bound by exactly that pipe: 1 MUFU + 0.5 F2FP per element = 12 cycles
// per element per warp (measured 1536 cycles per 128-element row pass, scripts/attn_trace.py).
__device__ __forceinline__ uint32_t pack_bf16x2_alu(float lo, float hi) {
  return __byte_perm(__float_as_uint(lo) + 0x8000u, __float_as_uint(hi) + 0x8000u, 0x7632);
}
__device__ __forceinline__ float bf16_round(float x) {
  return __bfloat162float(__float2bfloat16_rn(x));
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}
// `volatile` forms: the compiler keeps volatile asm statements in source order, which is how the
// softmax pass software-pipelines its MUFU stream by hand (consumers one chunk behind the producers).
__device__ __forceinline__ float fast_exp2_v(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2_v(float lo, float hi) {
  uint32_t r;
  asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// Packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 / FMUL2 process two fp32 lanes per instruction).
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};\n" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;\n" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;\n" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.f32x2 %0, %1, %2;\n" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t fadd2_v(uint64_t a, uint64_t b) {
  uint64_t d;
  asm volatile("add.f32x2 %0, %1, %2;\n" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t fmax2(uint64_t a, uint64_t b) {
  float a0, a1, b0, b1;
  unpack_f32x2(a, a0, a1);
  unpack_f32x2(b, b0, b1);
  return pack_f32x2(fmaxf(a0, b0), fmaxf(a1, b1));
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.f32x2 %0, %1, %2;\n" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// 2^x on the FMA/ALU pipes instead of the MUFU (16 ex2 per clock per SM is the scarce resource of
// the softmax): round-to-nearest split x = n + f, f in [-0.5, 0.5], cubic minimax for 2^f
// (relative error < 7.5e-5, far below the bf16 rounding applied to the probabilities), exponent
// patched in with integer arithmetic.  Valid for x in [-126, 126].
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -126.0f);
  const float r = x + 12582912.0f;            // 1.5 * 2^23: low mantissa bits = round(x)
  const float f = x - (r - 12582912.0f);
  float p = fmaf(f, 0.0551716685295105f, 0.2426111251115799f);
  p = fmaf(p, f, 0.6932609677314758f);
  p = fmaf(p, f, 0.9999280571937561f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}

// Two lanes at once (FFMA2 / FADD2).  Inputs must be >= -120 (callers clamp).
__device__ __forceinline__ uint64_t poly_exp2_x2(uint64_t x) {
  const uint64_t r = fadd2(x, pack_f32x2(12582912.0f, 12582912.0f));
  const uint64_t n = fadd2(r, pack_f32x2(-12582912.0f, -12582912.0f));
  const uint64_t f = ffma2(n, pack_f32x2(-1.f, -1.f), x);
  uint64_t p = ffma2(f, pack_f32x2(0.0551716685295105f, 0.0551716685295105f),
                     pack_f32x2(0.2426111251115799f, 0.2426111251115799f));
  p = ffma2(p, f, pack_f32x2(0.6932609677314758f, 0.6932609677314758f));
  p = ffma2(p, f, pack_f32x2(0.9999280571937561f, 0.9999280571937561f));
  float p0, p1, r0, r1;
  unpack_f32x2(p, p0, p1);
  unpack_f32x2(r, r0, r1);
  return pack_f32x2(__int_as_float(__float_as_int(p0) + (__float_as_int(r0) << 23)),
                    __int_as_float(__float_as_int(p1) + (__float_as_int(r1) << 23)));
}

// GELU(tanh) exactly as torch.nn.GELU(approximate='tanh') evaluates it in fp32.
__device__ __forceinline__ float gelu_tanh(float x) {
  const float kBeta = 0.7978845608028654f;  // sqrt(2/pi)
  const float kKappa = 0.044715f;
  float inner = kBeta * (x + kKappa * x * x * x);
  // tanh(u) = 1 - 2 / (1 + e^(2u)); e^(2u) = 2^(2u*log2 e).  ex2.approx keeps ~2^-22 relative
  // accuracy, far below the bf16 rounding that follows; saturates correctly at +-1.
  const float e = fast_exp2(inner * 2.885390081777927f);
  const float th = 1.0f - __fdividef(2.0f, 1.0f + e);
  return 0.5f * x * (1.0f + th);
}

// The same GELU(tanh) for two values at once, in the form  0.5 x (1 + tanh(u)) = x * sigmoid(2u)
// = x / (1 + 2^(-2 u log2 e)),  u = sqrt(2/pi) (x + 0.044715 x^3)  — an identity, no approximation
// beyond ex2.approx / rcp.approx (~2^-22), far below the bf16 rounding that follows.  5 packed fp32x2
// instructions + 4 MUFU per pair instead of ~13 scalar instructions per value: the GEMM epilogue is
// executed by ONE warp per scheduler and must fit behind a K = 1536 main loop.
__device__ __forceinline__ uint64_t gelu_tanh_x2(uint64_t x) {
  const float c1 = -2.0f * 0.7978845608028654f * 1.4426950408889634f;            // -2 sqrt(2/pi) log2(e)
  const float c3 = c1 * 0.044715f;
  const uint64_t x2 = fmul2(x, x);
  const uint64_t t = ffma2(x2, pack_f32x2(c3, c3), pack_f32x2(c1, c1));          // -2 log2e beta (1 + kappa x^2)
  float a0, a1;
  unpack_f32x2(fmul2(t, x), a0, a1);                                              // -2 u log2(e)
  // x -> -inf: 2^(+big) = inf, 1/(1+inf) = 0, x * 0 = -0: the correct limit; x -> +inf: 2^(-big) = 0, x / 1
  const uint64_t d = fadd2(pack_f32x2(fast_exp2(a0), fast_exp2(a1)), pack_f32x2(1.0f, 1.0f));
  float d0, d1, r0, r1;
  unpack_f32x2(d, d0, d1);
  asm("rcp.approx.ftz.f32 %0, %1;\n" : "=f"(r0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;\n" : "=f"(r1) : "f"(d1));
  return fmul2(x, pack_f32x2(r0, r1));
}

}  // namespace ovb
