// TEST INFRASTRUCTURE ONLY (see cuda_emu.h): k_walk_compress (segment-walk v1 compressor) compiled with g++ over the
// SIMT emulator, exported through a tiny C interface for tests/test_emulated_kernels.py.
#define TB_EMU 1
#include "cuda_emu.h"

#include "../../tamp_b200/csrc/cuda/walk_compress.cu"

extern "C" int emu_walk_warps() { return tb::kWarps; }

// One launch of k_walk_compress over host buffers.  Returns the number of streams it marked as deferred.
extern "C" int emu_walk_compress(const uint8_t *dict, int window, int literal, int flags, int write_token, int max_pairs,
                                 const uint8_t *in, const uint32_t *in_sizes, uint64_t in_stride, uint8_t *out,
                                 uint64_t out_stride, uint32_t *out_sizes, int8_t *status, uint64_t n, unsigned grid,
                                 uint64_t seed) {
    using namespace tb;
    WalkArgs a;
    a.b.in = in;
    a.b.in_offsets = nullptr;
    a.b.in_sizes = in_sizes;
    a.b.in_stride = in_stride;
    a.b.out = out;
    a.b.out_stride = out_stride;
    a.b.out_sizes = out_sizes;
    a.b.status = status;
    a.b.n_streams = n;
    a.dict = dict;
    a.window_bits = window;
    a.literal = literal;
    a.flags = flags;
    a.write_token = write_token;
    a.max_pairs = max_pairs;
    d_walk_deferred_total = 0;
    memset(emu::g_smem, 0xA5, sizeof emu::g_smem);  // shared memory starts out as garbage
    if (flags & TB_F_EXTENDED)
        emu::launch(grid, kWarps * 32, seed, [&] { k_walk_compress<true>(a); });
    else
        emu::launch(grid, kWarps * 32, seed, [&] { k_walk_compress<false>(a); });
    return (int)d_walk_deferred_total;
}
