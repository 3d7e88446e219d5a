// Internal launch interface between the engine (engine.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../tb_wire.h"

namespace tb {

// Device-side description of a batch (mirrors TampB200Batch with device pointers).
struct BatchArgs {
    const uint8_t *in;
    const uint64_t *in_offsets;
    const uint32_t *in_sizes;
    uint64_t in_stride;
    uint8_t *out;
    uint64_t out_stride;
    uint32_t *out_sizes;
    int8_t *status;
    uint64_t n_streams;
    // Decompress, the segments of ONE long stream (tamp_b200_decompress_segmented): 0 = every frame carries its own
    // header; otherwise 0x100 | the header byte of segment 0 (| 0x200: stream 0 of the batch is a later segment as well)
    // — the frames behind the first start with the 16-bit
    // append-mode marker (a FLUSH padded to 16 bits, compressor.c:227-234) where a dictionary_reset header would be.
    uint32_t seg_header = 0;
};

struct CompBatchConf {
    int window, literal, flags;  // flags: TB_F_*
    int write_token;
};

// generic_kernels.cu
void launch_comp_job(TbCompJob *d_job, uint8_t *d_window, const uint8_t *d_in, uint8_t *d_out, int window_bits,
                     cudaStream_t st);
void launch_dec_job(TbDecJob *d_job, uint8_t *d_window, const uint8_t *d_in, uint8_t *d_out, cudaStream_t st);
void launch_generic_compress_batch(const CompBatchConf &cf, const uint8_t *d_dict, const BatchArgs &b,
                                   cudaStream_t st);
// d_seed: three seeded tables of 32 KiB each (literal classes 5, 6, 7/8); d_custom may be NULL.
// d_scratch: n_slots windows of (1 << window_bits_max) bytes.
void launch_generic_decompress_batch(const uint8_t *d_seed, const uint8_t *d_custom, int window_bits_max,
                                     uint8_t *d_scratch, uint64_t n_slots, const BatchArgs &b, cudaStream_t st);
uint64_t generic_decompress_slots(uint64_t n_streams, int window_bits_max);
void launch_synth(int kind, uint64_t first_k, uint64_t n_streams, uint64_t stream_len, uint8_t *d_out,
                  cudaStream_t st);

// fast_compress.cu / fast_decompress.cu: return false when the configuration has no specialised kernel.
// only_deferred: process just the streams whose out_sizes entry is kDeferred (left behind by the position-parallel
// kernel); small_grid: few are expected, one warp per SM is enough to scan for them.
constexpr uint32_t kDeferred = 0xFFFFFFFFu;
bool launch_fast_compress_batch(const CompBatchConf &cf, const uint8_t *d_dict, const BatchArgs &b, cudaStream_t st,
                                bool only_deferred = false, bool small_grid = false);
// ppar_compress.cu: position-parallel v1 compressor for streams no longer than the window (<= 1024).
// allow_laps: also take v1 streams longer than the window (lap variant; off in kernel mode 4 = the round-1 dispatch).
bool launch_ppar_compress_batch(const CompBatchConf &cf, const uint8_t *d_dict, const BatchArgs &b, cudaStream_t st,
                                bool allow_laps = false);
// walk_compress.cu: segment-walk v1 compressor for streams no longer than the window (<= 1024): matches are evaluated
// only at the offsets the greedy parse reaches.  Same contract (incl. the pick-up pass for deferred streams).
bool launch_walk_compress_batch(const CompBatchConf &cf, const uint8_t *d_dict, const BatchArgs &b, cudaStream_t st);
// fast_compress.cu: nibble bitmaps of a dictionary (row stride rs words) built into a scratch slot on `st`.
const uint32_t *stage_dictrows(const uint8_t *d_dict, int W, int rs, cudaStream_t st);
// Dictionaries inside [lo, lo + bytes) never change (the engine's seeded tables): their bitmaps are cached.
void register_static_dictionaries(const uint8_t *lo, size_t bytes);
bool launch_wide_compress_batch(const CompBatchConf &cf, const uint8_t *d_dict, const BatchArgs &b, cudaStream_t st,
                                bool only_deferred = false);
// hwalk_compress.cu: history-walk v1 compressor for any window and any stream length (one CTA per stream; time-ordered
// hash chains over a ring of the last W + C bytes).  Same contract as the segment-walk kernel, incl. the pick-up pass.
bool launch_hwalk_compress_batch(const CompBatchConf &cf, const uint8_t *d_dict, const BatchArgs &b, cudaStream_t st);
// cwalk_compress.cu: cooperative history walk for windows 11..15 (v1, any stream length): candidates of a bigram as an
// array (counting sort per chunk), every poll evaluated by a whole warp, the parse walked by warps.
bool launch_cwalk_compress_batch(const CompBatchConf &cf, const uint8_t *d_dict, const BatchArgs &b, cudaStream_t st);
bool launch_fast_decompress_batch(const uint8_t *d_seed, const uint8_t *d_custom, int window_bits_max,
                                  const BatchArgs &b, cudaStream_t st, bool only_deferred = false, bool small_grid = false);
// split_decompress.cu: parse (lane per stream, no window) + copy (warp per stream) for frames whose output fits the
// window (<= 1024, no wrap); everything else is marked kDeferred and picked up by fast_decompress.cu behind it.
bool launch_split_decompress_batch(const uint8_t *d_seed, const uint8_t *d_custom, int window_bits_max, const BatchArgs &b,
                                   cudaStream_t st);

// wide_decompress.cu: one warp per stream, window in shared memory, windows 11..15.
bool launch_wide_decompress_batch(const uint8_t *d_seed, const uint8_t *d_custom, int window_bits_max, const BatchArgs &b,
                                  cudaStream_t st, bool only_deferred = false);
// lsplit_decompress.cu: parse (lane per stream) + copy (warp per stream) for any window and any row length: the frame's
// own output row in global memory is the history.  Everything it cannot decide is marked kDeferred and picked up by
// fast_decompress.cu / wide_decompress.cu behind it.
bool launch_lsplit_decompress_batch(const uint8_t *d_seed, const uint8_t *d_custom, int window_bits_max, const BatchArgs &b,
                                    cudaStream_t st);

// compact.cu: pack fixed-stride rows into contiguous frames; offsets[n_streams + 1] (exclusive prefix sum of sizes).
bool launch_compact(const uint8_t *rows, uint64_t stride, const uint32_t *sizes, uint64_t n, uint8_t *packed,
                    uint64_t capacity, uint64_t *offsets, cudaStream_t st);

// sizes[i] = offsets[i + 1] - offsets[i] (frame sizes of a packed layout)
void launch_offsets_to_sizes(const uint64_t *offsets, uint64_t n, uint32_t *sizes, cudaStream_t st);

void count_launch();

}  // namespace tb
