// TEST INFRASTRUCTURE ONLY (see cuda_emu.h): k_build_dictrows_wide + k_wide_compress (windows 11..15, one CTA per
// stream) compiled with g++ over the SIMT emulator.
#define TB_EMU 1
#include "cuda_emu.h"

#include "../../tamp_b200/csrc/cuda/wide_compress.cu"

namespace {
template <int WBITS, int NWARPS>
void run(const tb::WideCompArgs &a, bool ext, bool multi, unsigned grid, uint64_t seed) {
    using namespace tb;
    (void)multi;
    if (ext) {
        emu::launch(grid, NWARPS * 32, seed, [&] { k_wide_compress<WBITS, true, NWARPS>(a); });
    } else {
        emu::launch(grid, NWARPS * 32, seed, [&] { k_wide_compress<WBITS, false, NWARPS>(a); });
    }
}
}  // namespace

extern "C" int emu_wide_compress(const uint8_t *dict, int window, int literal, int flags, int write_token, const uint8_t *in,
                                 const uint32_t *in_sizes, uint64_t in_stride, uint8_t *out, uint64_t out_stride,
                                 uint32_t *out_sizes, int8_t *status, uint64_t n, unsigned grid, uint64_t seed, int multi) {
    using namespace tb;
    if (window < 11 || window > 15) return -1;
    const int W = 1 << window, rs = W / 32 + 4;
    alignas(16) static uint32_t rows[32 * (1024 + 4)];
    emu::launch(1, 1024, seed, [&] { k_build_dictrows_wide(dict, W, rows, rs, 32 * rs); });
    WideCompArgs a;
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
    a.only_deferred = 0;
    memset(emu::g_smem, 0xA5, sizeof emu::g_smem);
    const bool ext = (flags & TB_F_EXTENDED) != 0;
    switch (window) {
        case 11: run<11, 1>(a, ext, multi != 0, grid, seed); break;
        case 12: run<12, 1>(a, ext, multi != 0, grid, seed); break;
        case 13: run<13, 2>(a, ext, multi != 0, grid, seed); break;
        case 14: run<14, 4>(a, ext, multi != 0, grid, seed); break;
        default: run<15, 8>(a, ext, multi != 0, grid, seed); break;
    }
    return 0;
}
