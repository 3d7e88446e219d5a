// TEST INFRASTRUCTURE ONLY (see cuda_emu.h): k_hwalk_dict + k_hwalk_compress (history-walk v1 compressor) compiled with
// g++ over the SIMT emulator, exported through a tiny C interface for tests/test_emulated_kernels.py.
#define TB_EMU 1
#include "cuda_emu.h"

#include "../../tamp_b200/csrc/cuda/hwalk_compress.cu"

#include <vector>

// One launch over host buffers.  cbits / hbits / seg / threads = 0: the plan the launcher would pick for the window.
// Returns the number of streams marked as deferred, or -1 if the layout does not fit shared memory.
extern "C" int emu_hwalk_compress(const uint8_t *dict, int window, int literal, int flags, int write_token, int cbits, int hbits,
                                  int seg, int threads, int budget, int max_pairs, const uint8_t *in, const uint32_t *in_sizes,
                                  uint64_t in_stride, uint8_t *out, uint64_t out_stride, uint32_t *out_sizes, int8_t *status,
                                  uint64_t n, unsigned grid, uint64_t seed) {
    using namespace tb;
    const HwalkPlan plan = hwalk_plan(window);
    if (!cbits) cbits = plan.cbits;
    if (!hbits) hbits = plan.hbits;
    if (!seg) seg = plan.seg;
    if (!threads) threads = plan.threads;
    const HwalkLayout Lo = hwalk_layout(window, cbits, hbits, seg);
    if (Lo.total > emu::kSmemBytes) return -1;
    const int W = 1 << window, HS = 1 << hbits;
    std::vector<uint16_t> links(W), heads(HS);
    memset(emu::g_smem, 0xA5, sizeof emu::g_smem);  // shared memory starts out as garbage
    emu::launch(1, W >= 4096 ? 512 : 128, seed, [&] { k_hwalk_dict(dict, window, hbits, Lo.R - Lo.W, links.data(), heads.data()); });
    HwalkArgs a;
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
    a.dict_links = links.data();
    a.dict_heads = heads.data();
    a.window_bits = window;
    a.literal = literal;
    a.flags = flags;
    a.write_token = write_token;
    a.chunk_bits = cbits;
    a.hash_bits = hbits;
    const int nseg = (int)Lo.nseg, segs_per_lane = (nseg + threads - 1) / threads;
    a.budget = budget ? budget : segs_per_lane * (1024 + W);
    a.max_pairs = max_pairs ? max_pairs : 16;
    d_hwalk_deferred_total = 0;
    memset(emu::g_smem, 0xA5, sizeof emu::g_smem);
    if (seg == 32)
        emu::launch(grid, threads, seed, [&] { k_hwalk_compress<32>(a); });
    else
        emu::launch(grid, threads, seed, [&] { k_hwalk_compress<16>(a); });
    return (int)d_hwalk_deferred_total;
}
