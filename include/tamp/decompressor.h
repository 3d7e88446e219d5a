/* tamp-b200: ABI-compatible decompressor API.
 *
 * Drop-in for the reference's tamp/_c_src/tamp/decompressor.h (BrianPugh/tamp @ 48880ad):
 *   TampDecompressor ................... decompressor.h:13-57  (24 bytes on LP64)
 *   tamp_decompressor_read_header ...... decompressor.h:67    (decompressor.c:276-297)  host C
 *   tamp_decompressor_init ............. decompressor.h:83    (decompressor.c:331-347)  host C
 *   tamp_decompressor_decompress_cb .... decompressor.h:93    (decompressor.c:371-578)  -> CUDA
 * Same state-in / state-out contract as the compressor (see compressor.h).
 */
#ifndef TAMP_DECOMPRESSOR_H
#define TAMP_DECOMPRESSOR_H

#include "common.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    unsigned char *window; /* caller-owned, 1 << window_bits_max bytes */
    uint32_t bit_buffer;   /* unread bits, MSb-aligned */
    uint16_t window_pos;
    union {
        struct {
            uint8_t bit_buffer_pos; /* number of unread bits in bit_buffer */
            uint8_t token_state;    /* 0 none, 1 RLE pending, 2 ext-match fresh, 3 ext-match needs offset */
        };
        uint16_t pos_and_state;
    };
    uint16_t pending_window_offset; /* resume data for an interrupted extended token */
    uint16_t pending_match_size;
    uint8_t conf_window : 4;
    uint8_t conf_literal : 4;
    uint8_t min_pattern_size : 2;
    uint8_t conf_extended : 1;
    uint8_t conf_dictionary_reset : 1;
    union {
        uint8_t skip_bytes;          /* configured: bytes of the current token already delivered */
        uint8_t stashed_header_byte; /* not configured: first header byte awaiting the second */
    };
    uint8_t window_bits_max : 4;
    uint8_t configured : 1;
    uint8_t header_bytes_read : 2;
    uint8_t last_was_flush : 1;
} TampDecompressor;

tamp_res tamp_decompressor_read_header(TampConf *conf, const unsigned char *input, size_t input_size,
                                       size_t *input_consumed_size);

tamp_res tamp_decompressor_init(TampDecompressor *decompressor, const TampConf *conf, unsigned char *window,
                                uint8_t window_bits);

tamp_res tamp_decompressor_decompress_cb(TampDecompressor *decompressor, unsigned char *output, size_t output_size,
                                         size_t *output_written_size, const unsigned char *input, size_t input_size,
                                         size_t *input_consumed_size, tamp_callback_t callback, void *user_data);

/* Reference decompressor.h (decompressor.c:585-640): pull compressed bytes through read_cb, push the decoded bytes
 * through write_cb until the input is at its end and fully consumed.  callback(user_data, compressed bytes read so
 * far, 0) after every call into the decoder; non-zero aborts with that value. */
tamp_res tamp_decompress_stream(TampDecompressor *decompressor, tamp_read_t read_cb, void *read_handle,
                                tamp_write_t write_cb, void *write_handle, size_t *input_consumed_size,
                                size_t *output_written_size, tamp_callback_t callback, void *user_data);

static inline tamp_res tamp_decompressor_decompress(TampDecompressor *decompressor, unsigned char *output,
                                                    size_t output_size, size_t *output_written_size,
                                                    const unsigned char *input, size_t input_size,
                                                    size_t *input_consumed_size) {
    return tamp_decompressor_decompress_cb(decompressor, output, output_size, output_written_size, input, input_size,
                                           input_consumed_size, NULL, NULL);
}

#ifdef __cplusplus
}
#endif
#endif
