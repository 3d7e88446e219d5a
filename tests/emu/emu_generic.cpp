// TEST INFRASTRUCTURE ONLY (see cuda_emu.h): the general kernels (any window 8..15, any option) and the synthetic
// input generator compiled with g++ over the SIMT emulator.
#define TB_EMU 1
#include "cuda_emu.h"

#include "../../tamp_b200/csrc/cuda/generic_kernels.cu"

static tb::BatchArgs batch_args(const uint8_t *in, const uint64_t *in_offsets, const uint32_t *in_sizes, uint64_t in_stride,
                                uint8_t *out, uint64_t out_stride, uint32_t *out_sizes, int8_t *status, uint64_t n) {
    tb::BatchArgs b;
    b.in = in;
    b.in_offsets = in_offsets;
    b.in_sizes = in_sizes;
    b.in_stride = in_stride;
    b.out = out;
    b.out_stride = out_stride;
    b.out_sizes = out_sizes;
    b.status = status;
    b.n_streams = n;
    b.seg_header = emu::g_seg_header;
    return b;
}

// k_generic_compress_batch: one warp per stream, `wpc` warps per CTA (window + ring of each warp in shared memory).
extern "C" void emu_generic_compress(const uint8_t *dict, int window, int literal, int flags, int write_token,
                                     const uint8_t *in, const uint32_t *in_sizes, uint64_t in_stride, uint8_t *out,
                                     uint64_t out_stride, uint32_t *out_sizes, int8_t *status, uint64_t n, int wpc,
                                     uint64_t seed) {
    using namespace tb;
    CompBatchConf cf{window, literal, flags, write_token};
    const BatchArgs b = batch_args(in, nullptr, in_sizes, in_stride, out, out_stride, out_sizes, status, n);
    memset(emu::g_smem, 0xA5, sizeof emu::g_smem);
    emu::launch((unsigned)((n + wpc - 1) / wpc), (unsigned)wpc * 32, seed, [&] { k_generic_compress_batch(cf, dict, b); });
}

// k_generic_decompress_batch: one thread per stream, windows in `scratch` (n_slots << window_bits_max bytes).
extern "C" void emu_generic_decompress(const uint8_t *seed_tables, const uint8_t *custom, int window_bits_max,
                                       uint8_t *scratch, uint64_t n_slots, const uint8_t *in, const uint64_t *in_offsets,
                                       const uint32_t *in_sizes, uint64_t in_stride, uint8_t *out, uint64_t out_stride,
                                       uint32_t *out_sizes, int8_t *status, uint64_t n, uint64_t seed) {
    using namespace tb;
    const BatchArgs b = batch_args(in, in_offsets, in_sizes, in_stride, out, out_stride, out_sizes, status, n);
    emu::launch((unsigned)(n_slots / 128), 128, seed,
                [&] { k_generic_decompress_batch(seed_tables, custom, window_bits_max, scratch, b); });
}

// BatchArgs::seg_header for the decompressor launches that follow (0 = frames carry their own headers)
extern "C" void emu_set_seg_header(uint32_t v) { emu::g_seg_header = v; }

// k_synth: the synthetic input generators as the bench uses them.
extern "C" void emu_synth(int kind, uint64_t first_k, uint64_t n_streams, uint64_t stream_len, uint8_t *out) {
    using namespace tb;
    synth_build_vocab(g_vocab);
    emu::launch((unsigned)((n_streams + 127) / 128), 128, 0, [&] { k_synth(kind, first_k, n_streams, stream_len, out); });
}
