/* tamp-b200: ABI-compatible compressor API.
 *
 * Drop-in for the reference's tamp/_c_src/tamp/compressor.h (BrianPugh/tamp @ 48880ad):
 *   TampCompressor ........................ compressor.h:13-66  (48 bytes on LP64; offsets asserted below)
 *   tamp_compressor_init .................. compressor.h:84   (compressor.c:191-245)  host C
 *   tamp_compressor_sink .................. compressor.h:101  (compressor.c:665-679)  host C
 *   tamp_compressor_full .................. compressor.h:154  (compressor.c:77-79)    host C
 *   tamp_compressor_poll .................. compressor.h:142  (compressor.c:532-660)  -> CUDA
 *   tamp_compressor_flush ................. compressor.h:193  (compressor.c:728-810)  -> CUDA
 *   tamp_compressor_compress_cb ........... compressor.h:227  (compressor.c:681-722)  -> CUDA
 *   tamp_compressor_compress_and_flush_cb . compressor.h:259  (compressor.c:815-845)  -> CUDA
 *   tamp_compressor_reset_dictionary ...... compressor.h:207  (compressor.c:847-881)  -> CUDA
 *
 * The library never allocates on behalf of a compressor object and there is no destructor: the
 * caller-owned struct and window ARE the per-stream state; every CUDA-backed call ships
 * (state, window, input chunk) to the device, runs the codec kernel on a batch of one, and
 * ships (state', window', output chunk) back (SURVEY 8b "state-in / state-out").
 */
#ifndef TAMP_COMPRESSOR_H
#define TAMP_COMPRESSOR_H

#include "common.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct TampCompressor {
    unsigned char *window;   /* caller-owned, >= 1 << conf.window bytes */
    uint32_t bit_buffer;     /* pending output bits, MSb-aligned */
    uint16_t window_pos;     /* next write index into window */
    uint8_t bit_buffer_pos;  /* number of pending bits */
    uint8_t input_size;      /* bytes held in the input ring, 0..16 */
    uint8_t input_pos;       /* ring read index, 0..15 */
    unsigned char input[16]; /* 16-byte input ring */
    uint8_t min_pattern_size;
    TampConf conf;
#if TAMP_LAZY_MATCHING
    int16_t cached_match_index;
#endif
    uint16_t extended_match_position;
#if TAMP_LAZY_MATCHING
    uint8_t cached_match_size;
#endif
    uint8_t rle_count;
    uint8_t extended_match_count;
    uint8_t last_was_flush;
} TampCompressor;

tamp_res tamp_compressor_init(TampCompressor *compressor, const TampConf *conf, unsigned char *window);

void tamp_compressor_sink(TampCompressor *compressor, const unsigned char *input, size_t input_size,
                          size_t *consumed_size);

tamp_res tamp_compressor_poll(TampCompressor *compressor, unsigned char *output, size_t output_size,
                              size_t *output_written_size);
#define tamp_compressor_compress_poll tamp_compressor_poll

bool tamp_compressor_full(const TampCompressor *compressor);

tamp_res tamp_compressor_flush(TampCompressor *compressor, unsigned char *output, size_t output_size,
                               size_t *output_written_size, bool write_token);

tamp_res tamp_compressor_reset_dictionary(TampCompressor *compressor, unsigned char *output, size_t output_size,
                                          size_t *output_written_size);

tamp_res tamp_compressor_compress_cb(TampCompressor *compressor, unsigned char *output, size_t output_size,
                                     size_t *output_written_size, const unsigned char *input, size_t input_size,
                                     size_t *input_consumed_size, tamp_callback_t callback, void *user_data);

tamp_res tamp_compressor_compress_and_flush_cb(TampCompressor *compressor, unsigned char *output,
                                               size_t output_size, size_t *output_written_size,
                                               const unsigned char *input, size_t input_size,
                                               size_t *input_consumed_size, bool write_token,
                                               tamp_callback_t callback, void *user_data);

/* Reference compressor.h:338 (compressor.c:891-955): pull input through read_cb until it returns 0, push the
 * compressed bytes through write_cb, then flush (no FLUSH token).  callback(user_data, input bytes so far, 0) after
 * every chunk read; a non-zero return aborts with that value.  The compressor must be initialised. */
tamp_res tamp_compress_stream(TampCompressor *compressor, tamp_read_t read_cb, void *read_handle, tamp_write_t write_cb,
                              void *write_handle, size_t *input_consumed_size, size_t *output_written_size,
                              tamp_callback_t callback, void *user_data);

static inline tamp_res tamp_compressor_compress(TampCompressor *compressor, unsigned char *output,
                                                size_t output_size, size_t *output_written_size,
                                                const unsigned char *input, size_t input_size,
                                                size_t *input_consumed_size) {
    return tamp_compressor_compress_cb(compressor, output, output_size, output_written_size, input, input_size,
                                       input_consumed_size, NULL, NULL);
}

static inline tamp_res tamp_compressor_compress_and_flush(TampCompressor *compressor, unsigned char *output,
                                                          size_t output_size, size_t *output_written_size,
                                                          const unsigned char *input, size_t input_size,
                                                          size_t *input_consumed_size, bool write_token) {
    return tamp_compressor_compress_and_flush_cb(compressor, output, output_size, output_written_size, input,
                                                 input_size, input_consumed_size, write_token, NULL, NULL);
}

#ifdef __cplusplus
}
#endif
#endif
