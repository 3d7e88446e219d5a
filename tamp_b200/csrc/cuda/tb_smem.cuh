// Shared-memory accessors by .shared byte address (32-bit), for the loops in which the compiler would otherwise
// re-derive generic addresses: plain ld/st.shared on the device, the emulator's flat shared array in tests/emu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tb {
namespace smem {

#ifndef TB_EMU
__device__ __forceinline__ uint32_t ld32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t ld16(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t ld8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void st32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void st16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory"); }
__device__ __forceinline__ void st8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
#else  // tests/emu: the kernel stepped on the CPU (test infrastructure; see tests/emu/cuda_emu.h)
inline uint32_t ld32(uint32_t a) { return *reinterpret_cast<const uint32_t *>(emu_shared_ptr(a)); }
inline uint32_t ld16(uint32_t a) { return *reinterpret_cast<const uint16_t *>(emu_shared_ptr(a)); }
inline uint32_t ld8(uint32_t a) { return *reinterpret_cast<const uint8_t *>(emu_shared_ptr(a)); }
inline void st32(uint32_t a, uint32_t v) { *reinterpret_cast<uint32_t *>(emu_shared_ptr(a)) = v; }
inline void st16(uint32_t a, uint32_t v) { *reinterpret_cast<uint16_t *>(emu_shared_ptr(a)) = (uint16_t)v; }
inline void st8(uint32_t a, uint32_t v) { *reinterpret_cast<uint8_t *>(emu_shared_ptr(a)) = (uint8_t)v; }
#endif

// 16 bytes starting at shared byte address `sa` (any alignment; the arrays read this way keep 32 bytes of slack).
__device__ __forceinline__ void load16(uint32_t sa, uint32_t (&w)[4]) {
    const uint32_t q = sa & ~3u;
    const int sh = (int)(sa << 3);  // funnel shifts use the low 5 bits: (sa & 3) * 8
    const uint32_t a0 = ld32(q), a1 = ld32(q + 4), a2 = ld32(q + 8), a3 = ld32(q + 12), a4 = ld32(q + 16);
    w[0] = __funnelshift_r(a0, a1, sh);
    w[1] = __funnelshift_r(a1, a2, sh);
    w[2] = __funnelshift_r(a2, a3, sh);
    w[3] = __funnelshift_r(a3, a4, sh);
}

// Length of the common prefix of two 16-byte strings held as four little-endian words each (0..16).
__device__ __forceinline__ int common_prefix16(const uint32_t (&w)[4], const uint32_t (&la)[4]) {
    const uint32_t d0 = w[0] ^ la[0], d1 = w[1] ^ la[1], d2 = w[2] ^ la[2], d3 = w[3] ^ la[3];
    uint32_t d = d0;
    int nb = 0;
    if (!d) { d = d1; nb = 4; }
    if (!d) { d = d2; nb = 8; }
    if (!d) { d = d3; nb = 12; }
    return d ? nb + ((__ffs(d) - 1) >> 3) : 16;
}

}  // namespace smem
}  // namespace tb
