// Inline-PTX wrappers shared by the tensor-core kernels (mlp_tc.cu: K1; gemm_tc.cu: K5-TC) — sm_100a only:
// mbarriers, cluster addressing, TMA tile loads, tcgen05 fences / commit / TMEM loads.
#pragma once
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>

namespace cfn {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
// arrive on a barrier addressed in the shared::cluster window (own CTA or the pair's leader)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  // default .release.cta semantics, as cutlass::arch::ClusterBarrier::arrive: a cluster-scope release costs a
  // MEMBAR.ALL.GPU per arrival; the generic->async proxy fence issued before it is what orders the smem writes
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(bar_cluster_addr) : "memory");
}
// One probe of a phase.  Default .acquire.cta: a cluster-scope acquire makes ptxas emit CCTL.IVALL (L1 invalidate) after
// every wait, which round 1's first profile showed to be the single largest stall of the MMA-issuing thread.  The
// suspend-time hint lets the warp SLEEP in hardware until the phase completes (woken by the arrival) instead of
// re-issuing the probe every few hundred ns: spinning waiters steal issue slots from the MMA-issuing warp.
__device__ __forceinline__ bool mbar_try_wait_sleep(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
  return ok != 0;
}
// Short probe: mbarrier.try_wait WITHOUT a suspend-time hint.  It is "potentially blocking": when the phase is not
// complete it waits for a short implementation-defined time and wakes with low latency when the phase completes inside
// that window.  The MMA issuer uses it for the NEXT stage / chunk before issuing the current MMAs: measured on B200 it
// beats both a truly non-blocking mbarrier.test_wait followed by a sleeping wait (19.5 vs 17.5 ms per launch: the wake-up
// of the long-hint sleep costs ~300 cycles on the critical path) and a tight spin (steals issue slots).
__device__ __forceinline__ bool mbar_probe(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// truly non-blocking test (returns at once): for speculative look-ahead that usually fails
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t globaltimer_ns() { uint64_t t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// A phase that never completes (a bad tensor map, a protocol bug, a lost TMA transaction) must not hang the GPU: after
// kMbarTimeoutNs the waiter reports where it was stuck and traps; the host sees a launch failure instead of a dead device.
constexpr uint64_t kMbarTimeoutNs = 4000000000ull;   // 4 s: every legitimate wait in these kernels is far below 1 ms
static __device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity) {
  printf("cfnerf_b200: mbarrier wait timed out (block %d, thread %d, barrier smem 0x%x, parity %u) - aborting the kernel\n",
         (int)blockIdx.x, (int)threadIdx.x, bar, parity);
  __trap();
}
// The retry loop with its clock reads is OUT of line: inlined at every wait site it put ~20 cold instructions into the hot
// loops of single-warp roles whose instruction stream sets the pace of the kernel.
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  const uint64_t t0 = globaltimer_ns();
  while (!mbar_try_wait_sleep(bar, parity))
    if (globaltimer_ns() - t0 > kMbarTimeoutNs) mbar_timeout(bar, parity);
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait_sleep(bar, parity)) return;
  mbar_wait_slow(bar, parity);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int CG>
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  if (CG == 1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
  } else {
    // executed by both CTAs of the pair; the peer bit of the barrier address is cleared so the transaction bytes land
    // on the leader CTA's barrier
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar & 0xFEFFFFFFu) : "memory");
  }
}

// tcgen05.commit: the barrier receives one arrival when every MMA issued so far by this thread has completed.
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(bar), "h"((uint16_t)3) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
  uint32_t r; asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory"); return r;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace ptx
}  // namespace cfn
