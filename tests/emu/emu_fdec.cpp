// TEST INFRASTRUCTURE ONLY (see cuda_emu.h): k_fast_decompress compiled with g++ over the SIMT emulator.
#define TB_EMU 1
#include "cuda_emu.h"

#include "../../tamp_b200/csrc/cuda/fast_decompress.cu"

// One launch of k_fast_decompress<wmaxbits> over host buffers.  `seed` = 3 x 32 KiB seeded dictionaries.
static void run_fast_decompress(int wmaxbits, const uint8_t *seed_tables, const uint8_t *custom, int window_bits_max,
                                const uint8_t *in, const uint64_t *in_offsets, const uint32_t *in_sizes,
                                uint64_t in_stride, uint8_t *out, uint64_t out_stride, uint32_t *out_sizes,
                                int8_t *status, uint64_t n, unsigned grid, uint64_t seed, int only_deferred) {
    using namespace tb;
    FastDecArgs a;
    a.only_deferred = only_deferred;
    a.b.in = in;
    a.b.in_offsets = in_offsets;
    a.b.in_sizes = in_sizes;
    a.b.in_stride = in_stride;
    a.b.out = out;
    a.b.out_stride = out_stride;
    a.b.out_sizes = out_sizes;
    a.b.status = status;
    a.b.n_streams = n;
    a.b.seg_header = emu::g_seg_header;
    a.seed = seed_tables;
    a.custom = custom;
    a.window_bits_max = window_bits_max;
    a.aligned_io = ((out_stride & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) ? 1 : 0;
    a.lut = kHuff.lut;
    memset(emu::g_smem, 0xA5, sizeof emu::g_smem);
    if (wmaxbits == 8)
        emu::launch(grid, kWarpsPerCtaDec<8> * 32, seed, [&] { k_fast_decompress<8>(a); });
    else if (wmaxbits == 9)
        emu::launch(grid, kWarpsPerCtaDec<9> * 32, seed, [&] { k_fast_decompress<9>(a); });
    else
        emu::launch(grid, kWarpsPerCtaDec<10> * 32, seed, [&] { k_fast_decompress<10>(a); });
}

extern "C" void emu_fast_decompress(int wmaxbits, const uint8_t *seed_tables, const uint8_t *custom, int window_bits_max,
                                    const uint8_t *in, const uint64_t *in_offsets, const uint32_t *in_sizes,
                                    uint64_t in_stride, uint8_t *out, uint64_t out_stride, uint32_t *out_sizes,
                                    int8_t *status, uint64_t n, unsigned grid, uint64_t seed) {
    run_fast_decompress(wmaxbits, seed_tables, custom, window_bits_max, in, in_offsets, in_sizes, in_stride, out, out_stride,
                        out_sizes, status, n, grid, seed, 0);
}

// The pick-up pass behind k_split_decompress: only the streams whose out_sizes entry is kDeferred.
extern "C" void emu_fast_decompress_pickup(int wmaxbits, const uint8_t *seed_tables, const uint8_t *custom, int window_bits_max,
                                           const uint8_t *in, const uint64_t *in_offsets, const uint32_t *in_sizes,
                                           uint64_t in_stride, uint8_t *out, uint64_t out_stride, uint32_t *out_sizes,
                                           int8_t *status, uint64_t n, unsigned grid, uint64_t seed) {
    run_fast_decompress(wmaxbits, seed_tables, custom, window_bits_max, in, in_offsets, in_sizes, in_stride, out, out_stride,
                        out_sizes, status, n, grid, seed, 1);
}
