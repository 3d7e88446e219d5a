// TEST INFRASTRUCTURE ONLY (see cuda_emu.h): k_cwalk_compress (cooperative history-walk v1 compressor, windows 11..15)
// compiled with g++ over the SIMT emulator, exported through a tiny C interface for tests/test_emulated_kernels.py.
#define TB_EMU 1
#include "cuda_emu.h"

#include "../../tamp_b200/csrc/cuda/cwalk_compress.cu"

// One launch over host buffers.  cbits / hbits / threads = 0: the plan the launcher would pick for the window.
// Returns the number of streams marked as deferred, or -1 if the layout does not fit shared memory.
extern "C" int emu_cwalk_compress(const uint8_t *dict, int window, int literal, int flags, int write_token, int cbits, int hbits,
                                  int threads, int gl, int budget, const uint8_t *in, const uint32_t *in_sizes, uint64_t in_stride,
                                  uint8_t *out, uint64_t out_stride, uint32_t *out_sizes, int8_t *status, uint64_t n, unsigned grid,
                                  uint64_t seed) {
    using namespace tb;
    const CwalkPlan plan = cwalk_plan(window < 11 ? 11 : window);
    if (!cbits) cbits = plan.cbits;
    if (!hbits) hbits = plan.hbits;
    if (!threads) threads = plan.threads;
    const CwalkLayout Lo = cwalk_layout(window, cbits, hbits);
    if (Lo.total > emu::kSmemBytes) return -1;
    CwalkArgs a;
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
    a.chunk_bits = cbits;
    a.hash_bits = hbits;
    a.budget = budget ? budget : 16 * ((1 << cbits) / (threads / (gl == 16 ? 16 : 32)));
    d_cwalk_deferred_total = 0;
    memset(emu::g_smem, 0xA5, sizeof emu::g_smem);  // shared memory starts out as garbage
    if (gl == 16)
        emu::launch(grid, threads, seed, [&] { k_cwalk_compress<80, 16>(a); });
    else
        emu::launch(grid, threads, seed, [&] { k_cwalk_compress<80, 32>(a); });
    return (int)d_cwalk_deferred_total;
}
