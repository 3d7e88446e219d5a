// The history ring shared by the history-walk compressors (hwalk_compress.cu, cwalk_compress.cu): v1 format, any
// window.  V = dictionary ++ input in "virtual time"; the last W + C (+ 32 lookahead) bytes live in shared memory at
// position (time mod R), 32 mirror bytes behind the end making reads across the wrap contiguous.
#pragma once
#include "tb_device_common.cuh"
#include "tb_smem.cuh"

namespace tb {
namespace {

constexpr uint32_t kFull = 0xffffffffu;
constexpr int kPad = 32;  // mirror / lookahead bytes

__host__ __device__ inline uint32_t up16(uint32_t x) { return (x + 15u) & ~15u; }
// ring capacity: the multiple of C that holds W + C + 32 bytes, so that a chunk never wraps
__host__ __device__ inline uint32_t ring_size(uint32_t W, uint32_t C) { return C * ((W + C + kPad + C - 1u) / C); }

__device__ __forceinline__ uint32_t bigram_hash(uint32_t key16, int hbits) { return (key16 * 2654435761u) >> (32 - hbits); }

// One candidate of the poll at q: time distance D (1..W), its bytes at shared address sHB + ca.
__device__ __forceinline__ uint32_t eval_candidate(int D, uint32_t ca, const uint32_t (&la)[4], int L, uint32_t qmr, uint32_t sHB,
                                                   int pq, int W, int R) {
    uint32_t w[4];
    smem::load16(sHB + ca, w);
    int n = smem::common_prefix16(w, la);
    const uint32_t ridx = (qmr - (uint32_t)D) & (uint32_t)(W - 1);
    const int room = W - (int)ridx;  // never past the end of the window buffer
    const int lim0 = D < room ? D : room;
    const int lim = lim0 < L ? lim0 : L;
    if (n >= lim) {
        n = lim;
        if (D < room && D < L) {  // ran into the write position: the window continues with the bytes one lap older
            int ao = pq - W;
            if (ao < 0) ao += R;
            const int lim2 = room < L ? room : L;
            while (n < lim2 && smem::ld8(sHB + (uint32_t)(ao + n - D)) == smem::ld8(sHB + (uint32_t)(pq + n))) n++;
        }
    }
    return ((uint32_t)n << 16) | (ridx ^ 0xFFFFu);
}

}  // namespace
}  // namespace tb
