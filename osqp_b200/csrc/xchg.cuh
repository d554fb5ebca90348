// xchg.cuh -- device primitives of the peer-memory exchange (row-sharded solve): buffers, slots and the
// two-set protocol are described in common.cuh / dist.cu; users: pcg_graph.cu (CG loop), dist.cu
// (scalar blocks and vector heads outside the loop).
#pragma once

#include "common.cuh"

namespace b200 {

constexpr unsigned long long kXchgTimeoutNs = 4ull * 1000ull * 1000ull * 1000ull;   // a peer that never arrives

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// spin until the sequence word reaches `seq` (written by a peer over NVLink); bounded, so that a rank
// whose peer died reports an error instead of hanging the GPU; after one failure nothing waits again
__device__ __forceinline__ bool xchg_wait(XchgState* S, const unsigned long long* flag, unsigned long long seq) {
  if (*(volatile int*)&S->err) return false;
  const unsigned long long t0 = globaltimer_ns();
  unsigned spins = 0;
  while (ld_acquire_sys(flag) < seq) {
    if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > kXchgTimeoutNs) return false;
  }
  return true;
}

}  // namespace b200
