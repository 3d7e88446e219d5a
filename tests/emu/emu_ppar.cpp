// TEST INFRASTRUCTURE ONLY (see cuda_emu.h): the kernel sources compiled with g++ over the SIMT emulator, exported
// through a tiny C interface for tests/test_emulated_kernels.py.  Nothing here is reachable from the product.
#define TB_EMU 1
#include "cuda_emu.h"

#include "../../tamp_b200/csrc/cuda/ppar_compress.cu"

extern "C" int emu_ppar_warps(int mode) {
    using namespace tb;
    return mode == kModeLazy ? Lay<kModeLazy>::kWarps
         : mode == kModeExt  ? Lay<kModeExt>::kWarps
         : mode == kModeLaps ? Lay<kModeLaps>::kWarps
         : mode == kModeLazyLaps ? Lay<kModeLazyLaps>::kWarps
                             : Lay<kModeV1>::kWarps;
}

// One launch of k_ppar_compress<mode> over host buffers.  Returns the number of streams it marked as deferred.
extern "C" int emu_ppar_compress(int mode, const uint8_t *dict, int window, int literal, int flags, int write_token,
                                 int max_pairs, const uint8_t *in, const uint32_t *in_sizes, uint64_t in_stride,
                                 uint8_t *out, uint64_t out_stride, uint32_t *out_sizes, int8_t *status, uint64_t n,
                                 unsigned grid, uint64_t seed) {
    using namespace tb;
    PparArgs a;
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
    d_deferred_total = 0;
    memset(emu::g_smem, 0xA5, sizeof emu::g_smem);  // shared memory starts out as garbage
    if (mode == kModeLazy)
        emu::launch(grid, Lay<kModeLazy>::kWarps * 32, seed, [&] { k_ppar_compress<kModeLazy>(a); });
    else if (mode == kModeExt)
        emu::launch(grid, Lay<kModeExt>::kWarps * 32, seed, [&] { k_ppar_compress<kModeExt>(a); });
    else if (mode == kModeLaps)
        emu::launch(grid, Lay<kModeLaps>::kWarps * 32, seed, [&] { k_ppar_compress<kModeLaps>(a); });
    else if (mode == kModeLazyLaps)
        emu::launch(grid, Lay<kModeLazyLaps>::kWarps * 32, seed, [&] { k_ppar_compress<kModeLazyLaps>(a); });
    else
        emu::launch(grid, Lay<kModeV1>::kWarps * 32, seed, [&] { k_ppar_compress<kModeV1>(a); });
    return (int)d_deferred_total;
}
