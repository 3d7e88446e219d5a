// Small PTX wrappers shared by the sm_100a kernels: mbarrier + TMA bulk copy (SASS: UBLKCP), proxy fence.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tb {

#ifdef TB_EMU  // tests/emu: the kernels stepped on the CPU (test infrastructure; see tests/emu/cuda_emu.h)
// An mbarrier is a count of completed phases; a bulk copy completes at once.
inline void mbar_init(uint64_t *bar, uint32_t) { *bar = 0; }
inline void mbar_expect_tx(uint64_t *, uint32_t) {}
inline void mbar_wait(uint64_t *bar, uint32_t parity) {
    while ((*bar & 1u) == parity) emu::yield();
}
inline void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    memcpy(dst, src, bytes);
    *bar += 1;
    emu::g_cta->progress++;
}
inline void fence_proxy_async() {}
#else

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier.
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif  // TB_EMU

}  // namespace tb
