// Minimal mbarrier + bulk-async-copy (TMA engine, SASS UBLKCP) helpers for sm_100a.
//
// The stencil kernels stage row segments of their input planes in shared memory with
// `cp.async.bulk.shared::cluster.global` (one instruction per row, issued by one elected
// thread, completion counted in bytes on an mbarrier) and run a multi-stage
// producer/consumer ring over them.  Tensor-map TMA cannot be used for these arrays: the
// reference's layouts have row pitches of (n+1)*8 bytes, which is not a multiple of 16
// for even n, and tensor maps need 16-byte global strides.  A bulk copy only needs a
// 16-byte aligned start and size, which the kernels arrange per row by starting the copy
// at the aligned-down address and remembering the one-element lead.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dzb {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make barrier inits (and prior generic-proxy shared writes) visible to the async proxy
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy; src and dst 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}


// ---- signals between GPUs (peer memory over NVLink) ---------------------------------------------------------------
// A 32-bit word in a peer's memory is written with system-scope release semantics and polled with system-scope
// acquire loads; fence_proxy_async_global orders the acquire before TMA (async proxy) reads of the data it guards.
__device__ __forceinline__ void signal_release_sys(unsigned *flag, unsigned value) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}
__device__ __forceinline__ unsigned load_acquire_sys(const unsigned *flag) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// waits until *flag >= value (wrap-around safe); gives up after ~2 s and reports it through *timed_out
__device__ __forceinline__ void wait_signal_sys(const unsigned *flag, unsigned value, unsigned *timed_out) {
  if ((int)(load_acquire_sys(flag) - value) >= 0) return;
  const unsigned long long t0 = global_timer_ns();
  while ((int)(load_acquire_sys(flag) - value) < 0) {
    __nanosleep(64);
    if (global_timer_ns() - t0 > 2000000000ULL) {
      atomicExch(timed_out, 1u);
      return;
    }
  }
}

// non-blocking probe: issue early, test the predicate after independent work, fall back to mbar_wait
__device__ __forceinline__ bool mbar_test_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// Exact staging of the global elements [lo, hi) of an array that occupies exactly [abeg, aend) (all four are
// absolute element addresses = byte address / 8; hi is clamped to aend here).  A bulk copy needs 16-byte aligned
// ends, so it covers the widest even-aligned range that stays INSIDE the array; the at most one element that
// range misses at either end (only when the array itself starts / ends on an odd element address) is flagged for
// a plain 8-byte load + shared store by the calling lane.  Nothing outside [abeg, aend) is ever read.
struct RowCopy {
  long long tlo;     // first element of the bulk copy (even)
  unsigned bytes;    // size of the bulk copy (multiple of 16, may be 0)
  bool head, tail;   // element lo / element hi-1 must be copied by hand
  long long hi;      // clamped end
};
__device__ __forceinline__ RowCopy plan_row_copy(long long lo, long long hi, long long abeg, long long aend) {
  RowCopy c;
  hi = hi < aend ? hi : aend;
  long long tlo = lo & ~1ll;
  if (tlo < abeg) tlo += 2;
  long long thi = (hi + 1) & ~1ll;
  if (thi > aend) thi -= 2;
  c.tlo = tlo;
  c.bytes = thi > tlo ? (unsigned)((thi - tlo) << 3) : 0u;
  c.head = (lo < tlo) && (lo < hi);
  c.tail = hi > (thi > tlo ? thi : tlo);
  c.hi = hi;
  return c;
}

}  // namespace dzb
