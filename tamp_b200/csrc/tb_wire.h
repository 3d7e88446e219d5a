/* Internal: plain (bit-field free) per-stream state exchanged between the host C API layer and the
 * CUDA kernels.  The public structs (include/tamp/compressor.h, decompressor.h) keep the
 * reference's bit-field layouts; host code converts to/from these before/after each launch. */
#ifndef TB_WIRE_H
#define TB_WIRE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { TB_OP_POLL = 0, TB_OP_COMPRESS = 1, TB_OP_FLUSH = 2, TB_OP_COMPRESS_AND_FLUSH = 3, TB_OP_RESET_DICT = 4 };

/* TB_F_APPEND (batch kernels; compressor.c:227-234): the frame starts with a FLUSH token padded to 16 bits instead of a
 * header (conf.append; needs TB_F_DICT_RESET).  With TB_F_APPEND_TAIL stream 0 of the batch keeps its header and only
 * the streams behind it append: the segments of ONE long stream (tamp_b200_compress_segmented). */
enum { TB_F_EXTENDED = 1, TB_F_DICT_RESET = 2, TB_F_LAZY = 4, TB_F_CUSTOM_DICT = 8, TB_F_APPEND = 16, TB_F_APPEND_TAIL = 32 };

typedef struct TbCompState {
    uint32_t bit_buffer;
    uint16_t window_pos;
    uint8_t bit_buffer_pos;
    uint8_t input_size;
    uint8_t input_pos;
    uint8_t min_pattern_size;
    uint8_t window_bits;
    uint8_t literal_bits;
    uint8_t flags; /* TB_F_* */
    uint8_t rle_count;
    uint8_t ext_count;
    uint8_t last_was_flush;
    uint16_t ext_pos;
    int16_t cached_index;
    uint8_t cached_size;
    uint8_t pad[3];
    uint8_t input[16];
} TbCompState;

typedef struct TbCompJob {
    TbCompState st;
    uint32_t op;
    uint32_t write_token;
    uint64_t in_size;
    uint64_t out_cap;
    /* results */
    int32_t res;
    uint32_t pad;
    uint64_t out_written;
    uint64_t in_consumed;
} TbCompJob;

typedef struct TbDecState {
    uint32_t bit_buffer;
    uint16_t window_pos;
    uint8_t bit_buffer_pos;
    uint8_t token_state;
    uint16_t pending_window_offset;
    uint16_t pending_match_size;
    uint8_t window_bits;
    uint8_t literal_bits;
    uint8_t min_pattern_size;
    uint8_t flags; /* TB_F_EXTENDED | TB_F_DICT_RESET */
    uint8_t skip_bytes; /* doubles as stashed_header_byte before configuration */
    uint8_t window_bits_max;
    uint8_t configured;
    uint8_t header_bytes_read;
    uint8_t last_was_flush;
    uint8_t pad[3];
} TbDecState;

typedef struct TbDecJob {
    TbDecState st;
    uint64_t in_size;
    uint64_t out_cap;
    int32_t res;
    uint32_t pad;
    uint64_t out_written;
    uint64_t in_consumed;
} TbDecJob;

/* Launch shims implemented in cuda/engine.cu (extern "C"); host pointers in, results copied back. */
int tb_engine_run_comp_job(TbCompJob *job, unsigned char *window, const unsigned char *in, unsigned char *out);
int tb_engine_run_dec_job(TbDecJob *job, unsigned char *window, const unsigned char *in, unsigned char *out);
void tb_set_error(const char *fmt, ...);

#ifdef __cplusplus
}
#endif
#endif
