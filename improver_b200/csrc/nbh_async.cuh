// mbarrier / bulk asynchronous copy helpers (sm_100a) shared by the row-streaming kernels.
#pragma once
#include "nbh_common.cuh"

namespace nbh {

__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "NBH_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra NBH_DONE_%=;\n"
      "bra NBH_WAIT_%=;\n"
      "NBH_DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
// try_wait with a suspend-time hint: a waiting thread sleeps in hardware instead of spinning
// through the issue port
__device__ __forceinline__ void mbar_wait_hint(unsigned bar, unsigned parity, unsigned hint_ns) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "NBH_WAITH_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra NBH_DONEH_%=;\n"
      "bra NBH_WAITH_%=;\n"
      "NBH_DONEH_%=:\n"
      "}\n" ::"r"(bar), "r"(parity), "r"(hint_ns)
      : "memory");
}
// producer-side wait: the producer thread has stages of slack, so it sleeps between tries instead
// of taking issue slots from the math warps with a polling loop
__device__ __forceinline__ void mbar_wait_backoff(unsigned bar, unsigned parity, unsigned sleep_ns) {
  for (;;) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    __nanosleep(sleep_ns);
  }
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

}  // namespace nbh
