// TEST INFRASTRUCTURE ONLY (see cuda_emu.h): k_lsplit_decompress (parse / copy split decompressor for any window and any
// row length) compiled with g++ over the SIMT emulator.
#define TB_EMU 1
#include "cuda_emu.h"

#include <vector>

#include "../../tamp_b200/csrc/cuda/lsplit_decompress.cu"

// One launch of k_lsplit_decompress over host buffers.  Returns the number of streams it marked as deferred.
extern "C" int emu_lsplit_decompress(const uint8_t *seed_tables, const uint8_t *custom, int window_bits_max, const uint8_t *in,
                                     const uint64_t *in_offsets, const uint32_t *in_sizes, uint64_t in_stride, uint8_t *out,
                                     uint64_t out_stride, uint32_t *out_sizes, int8_t *status, uint64_t n, unsigned grid,
                                     uint64_t seed) {
    using namespace tb;
    std::vector<uint32_t> scratch((size_t)grid * kLsWarps * kChunk * 32, 0xDEADBEEFu);
    LsplitArgs a;
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
    a.scratch = scratch.data();
    d_lsplit_deferred_total = 0;
    memset(emu::g_smem, 0xA5, sizeof emu::g_smem);
    emu::launch(grid, kLsWarps * 32, seed, [&] { k_lsplit_decompress(a); });
    return (int)d_lsplit_deferred_total;
}
