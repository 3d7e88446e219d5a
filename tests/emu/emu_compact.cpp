// TEST INFRASTRUCTURE ONLY (see cuda_emu.h): the output-compaction kernels compiled with g++ over the SIMT emulator.
#define TB_EMU 1
#include "cuda_emu.h"

#include "../../tamp_b200/csrc/cuda/compact.cu"

// block sums -> scan -> offsets -> copies, as launch_compact issues them.  offsets[n + 1]; returns the packed size.
extern "C" uint64_t emu_compact(const uint8_t *rows, uint64_t stride, const uint32_t *sizes, uint64_t n, uint8_t *packed,
                                uint64_t capacity, uint64_t *offsets, uint64_t seed) {
    using namespace tb;
    if (n == 0) {
        offsets[0] = 0;
        return 0;
    }
    const uint64_t n_blocks = (n + kPerBlock - 1) / kPerBlock;
    std::vector<uint64_t> sums(n_blocks + 1);
    emu::launch((unsigned)n_blocks, kThreads, seed, [&] { k_block_sums(sizes, n, sums.data()); });
    emu::launch(1, 1024, seed, [&] { k_scan_block_sums(sums.data(), n_blocks, sums.data() + n_blocks); });
    emu::launch((unsigned)n_blocks, kThreads, seed,
                [&] { k_pack_rows(rows, stride, sizes, n, sums.data(), packed, capacity, offsets); });
    const int cta_per_row = stride > kWarpRow ? 1 : 0;
    emu::launch(3, kThreads, seed, [&] { k_copy_rows(rows, stride, sizes, n, offsets, packed, capacity, cta_per_row); });
    return sums[n_blocks];
}
