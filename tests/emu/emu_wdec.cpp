// TEST INFRASTRUCTURE ONLY (see cuda_emu.h): k_wide_decompress (one warp per stream) compiled with g++ over the emulator.
#define TB_EMU 1
#include "cuda_emu.h"

#include "../../tamp_b200/csrc/cuda/wide_decompress.cu"

extern "C" void emu_wide_decompress(const uint8_t *seed_tables, const uint8_t *custom, int window_bits_max, const uint8_t *in,
                                    const uint64_t *in_offsets, const uint32_t *in_sizes, uint64_t in_stride, uint8_t *out,
                                    uint64_t out_stride, uint32_t *out_sizes, int8_t *status, uint64_t n, unsigned grid,
                                    int wpc, uint64_t seed) {
    using namespace tb;
    WideDecArgs a;
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
    a.only_deferred = 0;
    memset(emu::g_smem, 0xA5, sizeof emu::g_smem);
    emu::launch(grid, (unsigned)wpc * 32, seed, [&] { k_wide_decompress(a); });
}
