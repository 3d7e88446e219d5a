// TEST INFRASTRUCTURE ONLY (see cuda_emu.h): k_build_dictrows + k_fast_compress compiled with g++ over the SIMT emulator.
#define TB_EMU 1
#include "cuda_emu.h"

#include "../../tamp_b200/csrc/cuda/fast_compress.cu"

namespace {
template <int WBITS, bool EXT>
void run(const tb::FastCompArgs &a, int wpc, unsigned grid, uint64_t seed) {
    using namespace tb;
    if (wpc == 1)
        emu::launch(grid, 32, seed, [&] { k_fast_compress<WBITS, EXT, 1>(a); });
    else
        emu::launch(grid, kWarpsPerCta * 32, seed, [&] { k_fast_compress<WBITS, EXT, kWarpsPerCta>(a); });
}
}  // namespace

// k_build_dictrows, then one launch of k_fast_compress<window, extended, wpc>.  only_deferred: pick-up pass (streams
// whose out_sizes entry is 0xFFFFFFFF).  Returns 0, or -1 for an unsupported window.
extern "C" int emu_fast_compress(const uint8_t *dict, int window, int literal, int flags, int write_token, int only_deferred,
                                 const uint8_t *in, const uint32_t *in_sizes, uint64_t in_stride, uint8_t *out,
                                 uint64_t out_stride, uint32_t *out_sizes, int8_t *status, uint64_t n, unsigned grid,
                                 uint64_t seed) {
    using namespace tb;
    const int W = 1 << window, rs = W / 32 + 1;
    const int total_words = ((32 * rs * 4 + 15) / 16 * 16) / 4;
    alignas(16) static uint32_t rows[2048];
    emu::launch(1, 256, seed, [&] { k_build_dictrows(dict, W, rows, rs, total_words); });
    FastCompArgs a;
    a.b.in = in;
    a.b.in_offsets = nullptr;
    a.b.in_sizes = in_sizes;
    a.b.in_stride = in_stride;
    a.b.out = out;
    a.b.out_stride = out_stride;
    a.b.out_sizes = out_sizes;
    a.b.status = status;
    a.b.n_streams = n;
    a.dictrows = rows;
    a.literal = literal;
    a.flags = flags;
    a.write_token = write_token;
    a.only_deferred = only_deferred;
    a.small_grid = 0;
    memset(emu::g_smem, 0xA5, sizeof emu::g_smem);
    const int wpc = only_deferred ? 1 : kWarpsPerCta;
    switch (window * 2 + ((flags & TB_F_EXTENDED) ? 1 : 0)) {
        case 16: run<8, false>(a, wpc, grid, seed); break;
        case 17: run<8, true>(a, wpc, grid, seed); break;
        case 18: run<9, false>(a, wpc, grid, seed); break;
        case 19: run<9, true>(a, wpc, grid, seed); break;
        case 20: run<10, false>(a, wpc, grid, seed); break;
        case 21: run<10, true>(a, wpc, grid, seed); break;
        default: return -1;
    }
    return 0;
}
