/* tamp-b200: batch extension of the Tamp C API (new, additive — the reference has no batch call).
 *
 * One call (de)compresses many independent streams on one B200.  For every stream the bytes produced
 * are exactly what the reference's per-stream call sequence produces on the same input:
 *
 *   compress   == tamp_compressor_init(conf, window)  +  tamp_compressor_compress_and_flush(..., write_token)
 *                 (compressor.h:84, :280; the sequence used by devices/common/tamp_bench.c:113-120,
 *                  fuzz/fuzz_round_trip.c:43-55 and the Python one-shot tamp/_c_compressor.pyx:189-199)
 *   decompress == tamp_decompressor_init(conf=NULL, window, window_bits_max)  +  ONE
 *                 tamp_decompressor_decompress(out, out_stride, ..., in, in_size)
 *                 (decompressor.h:83, :128; devices/common/tamp_bench.c:156-162)
 *
 * Plain C ABI: pointers and sizes only.  `*_device` variants take device pointers (data resident in
 * HBM, nothing copied) and a cudaStream_t passed as void*; the host variants stage through pinned
 * memory.  There is no CPU fallback: without a usable CUDA device every entry point returns TAMP_ERROR
 * and tamp_b200_last_error() says why.
 */
#ifndef TAMP_B200_H
#define TAMP_B200_H

#include "tamp/common.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct TampB200Batch {
    const unsigned char *in;    /* base of all stream inputs */
    const uint64_t *in_offsets; /* [n_streams] byte offset of stream i, or NULL => i * in_stride */
    const uint32_t *in_sizes;   /* [n_streams] byte length of stream i, or NULL => in_stride */
    uint64_t in_stride;
    unsigned char *out;   /* stream i's output starts at out + i * out_stride */
    uint64_t out_stride;  /* output capacity per stream */
    uint32_t *out_sizes;  /* [n_streams] bytes produced */
    int8_t *status;       /* [n_streams] tamp_res of the per-stream call sequence; may be NULL */
    uint64_t n_streams;
} TampB200Batch;

/* Worst-case compressed size of an n-byte stream (all literals + headers + flush slack). */
size_t tamp_b200_compress_bound(const TampConf *conf, size_t n);

/* Host-pointer entry points (H2D / D2H inside the call).  `dictionary` is NULL unless
 * conf->use_custom_dictionary (then 1 << conf->window bytes, shared by all streams). */
tamp_res tamp_b200_compress_batch(const TampConf *conf, const unsigned char *dictionary, const TampB200Batch *batch,
                                  bool write_token);
/* Host-pointer compress into CONTIGUOUS frames instead of fixed-stride rows: frame i is
 * packed[offsets[i] .. offsets[i] + batch->out_sizes[i]); `offsets` has n_streams + 1 entries (offsets[n] = total bytes).
 * batch->out / out_stride are not used (worst-case rows exist on the device only); the input is strided
 * (in_offsets == NULL).  Only payload bytes cross the bus, in one contiguous copy per chunk.  TAMP_OUTPUT_FULL if
 * packed_capacity is too small (sum of tamp_b200_compress_bound() always fits).  The frames are what
 * tamp_b200_decompress_batch takes back through in = packed, in_offsets = offsets, in_sizes = out_sizes. */
tamp_res tamp_b200_compress_batch_packed(const TampConf *conf, const unsigned char *dictionary, const TampB200Batch *batch,
                                         bool write_token, unsigned char *packed, uint64_t packed_capacity,
                                         uint64_t *offsets);

/* Threads: one host-pointer compress call and one host-pointer decompress call may be in flight together (two host
 * threads; separate staging slots per direction): the first is bound by host -> device traffic, the second by device ->
 * host, so a stream of batches keeps both PCIe directions busy.  Two calls of the same direction serialise. */

/* Decompress: `window_bits_max` is the size (log2) of the window buffer a per-stream caller would hand to
 * tamp_decompressor_init (decompressor.h:67-79): frames that ask for a larger window end with TAMP_INVALID_CONF.  A
 * non-NULL `dictionary` must hold exactly 1 << window_bits_max bytes: that many are read.  Batches whose
 * window_bits_max is <= 10 take the specialised kernels; pass the real bound, not 15, when it is known. */
tamp_res tamp_b200_decompress_batch(const unsigned char *dictionary, uint8_t window_bits_max,
                                    const TampB200Batch *batch);

/* Device-pointer entry points: every pointer in `batch` (and `dictionary`) is a device pointer;
 * work is enqueued on `cuda_stream` (cudaStream_t, NULL = default stream) and NOT synchronised.  Calls on different
 * streams may overlap: every scratch buffer of a call (dictionary copy, token records, windows of the general
 * decompressor, compaction sums) is allocated and freed in stream order on `cuda_stream`.  The engine is bound to the
 * CUDA device that was current at its first use (one process per GPU); calls with another device current fail. */
tamp_res tamp_b200_compress_batch_device(const TampConf *conf, const unsigned char *dictionary,
                                         const TampB200Batch *batch, bool write_token, void *cuda_stream);
tamp_res tamp_b200_decompress_batch_device(const unsigned char *dictionary, uint8_t window_bits_max,
                                           const TampB200Batch *batch, void *cuda_stream);

/* Output compaction (the reference's .tamp files / wire frames are contiguous; the batch kernels write fixed-stride
 * rows): packs rows [out + i * out_stride, + out_sizes[i]) of a finished batch into `packed` and writes the frame
 * offsets — offsets[i] = sum of out_sizes[0..i), offsets[n_streams] = total bytes (n_streams + 1 entries).  Frames
 * that would end beyond packed_capacity are not copied; offsets[n_streams] still tells the room needed.  The result
 * is the layout tamp_b200_decompress_batch_device takes through in = packed, in_offsets = offsets, in_sizes.
 * Device pointers; enqueued on `cuda_stream`, not synchronised. */
tamp_res tamp_b200_compact_batch_device(const TampB200Batch *batch, unsigned char *packed, uint64_t packed_capacity,
                                        uint64_t *offsets, void *cuda_stream);

/* ONE long stream, segment-parallel (the format's own mechanism: dictionary_reset + append mode, compressor.c:227-234,
 * :847-881; decompressor.c:501-514).  `in` is cut into segments of segment_size bytes (a multiple of 16, at most 2^30;
 * the last one may be shorter).  The bytes written are exactly those of ONE reference compressor driven as
 *
 *     tamp_compressor_init(conf with dictionary_reset = 1, window)
 *     for every segment:  tamp_compressor_compress(segment) ; tamp_compressor_reset_dictionary()   [not after the last]
 *     tamp_compressor_flush(write_token = true)
 *
 * i.e. segment 0 is a dictionary_reset stream, every later segment starts with a FLUSH padded to 16 bits (what
 * conf.append writes in place of a header), every segment ends with a FLUSH token: one valid Tamp stream that any
 * Tamp decompressor reads from front to back, and whose segments the decompress call below reads in parallel given
 * their offsets.  conf: NULL = the library default; use_custom_dictionary and append are rejected (a reset re-seeds the
 * dictionary).  seg_offsets: tamp_b200_segment_count() + 1 entries (the last one = total bytes), may be NULL for compress.
 * The calls return when the work is done (*out_size is a host variable); TAMP_OUTPUT_FULL if out_capacity is too small
 * (*out_size then tells the room needed; tamp_b200_segmented_bound() always fits).  A segment's error status
 * (e.g. TAMP_EXCESS_BITS) is returned as such.  `_device`: in / out / seg_offsets are device pointers.  The host-pointer
 * compress call pipelines chunks of segments over both PCIe directions like tamp_b200_compress_batch_packed (pinned
 * buffers reach the link rate: ~33 GB/s for a 1 GiB stream); after TAMP_OUTPUT_FULL from that path *out_size is
 * tamp_b200_segmented_bound(), a size that suffices. */
uint64_t tamp_b200_segment_count(uint64_t in_size, uint64_t segment_size);
uint64_t tamp_b200_segmented_bound(const TampConf *conf, uint64_t in_size, uint64_t segment_size);
tamp_res tamp_b200_compress_segmented(const TampConf *conf, const unsigned char *in, uint64_t in_size, uint64_t segment_size,
                                      unsigned char *out, uint64_t out_capacity, uint64_t *seg_offsets, uint64_t *out_size);
tamp_res tamp_b200_compress_segmented_device(const TampConf *conf, const unsigned char *in, uint64_t in_size,
                                             uint64_t segment_size, unsigned char *out, uint64_t out_capacity,
                                             uint64_t *seg_offsets, uint64_t *out_size, void *cuda_stream);
/* Segment-parallel decompress of such a stream: segment i is in[seg_offsets[i] .. seg_offsets[i + 1]) and decodes to
 * segment_size bytes (the last one to at most that).  *out_size = bytes written; TAMP_OUTPUT_FULL if out_capacity ends
 * before the data does.  window_bits_max as for tamp_b200_decompress_batch.  The host form checks that the offsets
 * ascend; the `_device` form trusts them like TampB200Batch::in_offsets (in[0 .. seg_offsets[n_segments]) must be
 * readable). */
tamp_res tamp_b200_decompress_segmented(const unsigned char *in, const uint64_t *seg_offsets, uint64_t n_segments,
                                        uint64_t segment_size, uint8_t window_bits_max, unsigned char *out,
                                        uint64_t out_capacity, uint64_t *out_size);
tamp_res tamp_b200_decompress_segmented_device(const unsigned char *in, const uint64_t *seg_offsets, uint64_t n_segments,
                                               uint64_t segment_size, uint8_t window_bits_max, unsigned char *out,
                                               uint64_t out_capacity, uint64_t *out_size, void *cuda_stream);

/* Kernel selection for the batch entry points: 0 = auto (specialised kernels when the configuration
 * has one, otherwise the general kernel), 1 = force the general kernel.  Both are CUDA. */
void tamp_b200_set_kernel_mode(int mode);

/* Deterministic synthetic streams: kind 0 text, 1 printable-random, 2 alpha16,
 * 3 periodic, 4 binary, 5 run-heavy.  Stream i is generated from index first_k + i. */
tamp_res tamp_b200_synth_device(int kind, uint64_t first_k, uint64_t n_streams, uint64_t stream_len,
                                unsigned char *d_out, void *cuda_stream);

/* Engine control / introspection. */
int tamp_b200_device_count(void);
tamp_res tamp_b200_set_device(int device);
const char *tamp_b200_last_error(void);
uint64_t tamp_b200_launch_count(void); /* kernels launched by this library since load */
/* Cumulative bytes the host-pointer batch entry points have copied host->device and device->host. */
void tamp_b200_copy_bytes(uint64_t *h2d, uint64_t *d2h);
const char *tamp_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif
