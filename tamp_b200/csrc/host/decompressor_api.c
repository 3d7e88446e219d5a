/* Host C layer of the decompressor API (include/tamp/decompressor.h).
 *
 * read_header / init and the header bookkeeping at the top of decompress are host logic as in the
 * reference (decompressor.c:276-347, :386-421; dictionary seeding = common.c, kept).  The token
 * loop itself — bit reader, Huffman decode, window copy — runs in the CUDA codec kernel on a batch
 * of one via tb_engine_run_dec_job(); the caller's struct and window round-trip through it, so
 * partial-input / output-full resumption behaves call by call like the reference.
 */
#include <string.h>

#include "../tb_wire.h"
#include "tamp/decompressor.h"

tamp_res tamp_decompressor_read_header(TampConf *conf, const unsigned char *input, size_t input_size,
                                       size_t *input_consumed_size) {
    if (input_consumed_size) *input_consumed_size = 0;
    if (input_size == 0) return TAMP_INPUT_EXHAUSTED;
    const size_t need = 1u + (input[0] & 1u); /* bit 0: a second header byte follows */
    if (input_size < need) return TAMP_INPUT_EXHAUSTED;
    if (need == 2 && input[1] != 0) return TAMP_INVALID_CONF; /* byte 2 is reserved */
    conf->window = (uint16_t)(((input[0] >> 5) & 7u) + 8u);
    conf->literal = (uint16_t)(((input[0] >> 3) & 3u) + 5u);
    conf->use_custom_dictionary = (input[0] >> 2) & 1u;
    conf->extended = (input[0] >> 1) & 1u;
    conf->dictionary_reset = input[0] & 1u;
    if (input_consumed_size) *input_consumed_size = need;
    return TAMP_OK;
}

static tamp_res configure(TampDecompressor *d, unsigned window, unsigned literal, unsigned custom, unsigned extended,
                          unsigned dict_reset) {
    if (window < 8 || window > 15) return TAMP_INVALID_CONF;
    if (literal < 5 || literal > 8) return TAMP_INVALID_CONF;
    if (window > d->window_bits_max) return TAMP_INVALID_CONF;
    if (!custom) tamp_initialize_dictionary(d->window, (size_t)1 << window, (uint8_t)(extended ? literal : 8));
    d->conf_window = (uint8_t)window;
    d->conf_literal = (uint8_t)literal;
    d->min_pattern_size = (uint8_t)tamp_compute_min_pattern_size((uint8_t)window, (uint8_t)literal);
    d->configured = 1;
    d->conf_extended = (uint8_t)extended;
    d->conf_dictionary_reset = (uint8_t)dict_reset;
    return TAMP_OK;
}

tamp_res tamp_decompressor_init(TampDecompressor *decompressor, const TampConf *conf, unsigned char *window,
                                uint8_t window_bits) {
    if (window_bits < 8 || window_bits > 15) return TAMP_INVALID_CONF;
    TampConf kept;
    if (conf) kept = *conf;
    memset(decompressor, 0, sizeof *decompressor);
    decompressor->window = window;
    decompressor->window_bits_max = window_bits;
    if (!conf) return TAMP_OK; /* header will be read from the stream */
    return configure(decompressor, kept.window, kept.literal, kept.use_custom_dictionary, kept.extended,
                     kept.dictionary_reset);
}

tamp_res tamp_decompressor_decompress_cb(TampDecompressor *d, unsigned char *output, size_t output_size,
                                         size_t *output_written_size, const unsigned char *input, size_t input_size,
                                         size_t *input_consumed_size, tamp_callback_t callback, void *user_data) {
    size_t consumed = 0;
    if (output_written_size) *output_written_size = 0;
    if (input_consumed_size) *input_consumed_size = 0;

    if (!d->configured) {
        TampConf conf;
        size_t used = 0;
        tamp_res res;
        if (d->header_bytes_read) { /* second call of a split two-byte header */
            unsigned char both[2] = {d->stashed_header_byte, 0};
            size_t have = 1;
            if (input_size) {
                both[1] = input[0];
                have = 2;
            }
            res = tamp_decompressor_read_header(&conf, both, have, &used);
            if (res != TAMP_OK) return res;
            consumed = used - 1;
        } else {
            res = tamp_decompressor_read_header(&conf, input, input_size, &used);
            if (res == TAMP_INPUT_EXHAUSTED && input_size) {
                d->stashed_header_byte = input[0];
                d->header_bytes_read = 1;
                if (input_consumed_size) *input_consumed_size = 1;
                return TAMP_INPUT_EXHAUSTED;
            }
            if (res != TAMP_OK) return res;
            consumed = used;
        }
        if (input_consumed_size) *input_consumed_size = consumed;
        res = configure(d, conf.window, conf.literal, conf.use_custom_dictionary, conf.extended,
                        conf.dictionary_reset);
        if (res != TAMP_OK) return res;
        d->skip_bytes = 0;
    }

    TbDecJob job;
    memset(&job, 0, sizeof job);
    job.st.bit_buffer = d->bit_buffer;
    job.st.window_pos = d->window_pos;
    job.st.bit_buffer_pos = d->bit_buffer_pos;
    job.st.token_state = d->token_state;
    job.st.pending_window_offset = d->pending_window_offset;
    job.st.pending_match_size = d->pending_match_size;
    job.st.window_bits = d->conf_window;
    job.st.literal_bits = d->conf_literal;
    job.st.min_pattern_size = d->min_pattern_size;
    job.st.flags = (uint8_t)((d->conf_extended ? TB_F_EXTENDED : 0) | (d->conf_dictionary_reset ? TB_F_DICT_RESET : 0));
    job.st.skip_bytes = d->skip_bytes;
    job.st.window_bits_max = d->window_bits_max;
    job.st.configured = 1;
    job.st.last_was_flush = d->last_was_flush;
    job.in_size = input_size - consumed;
    job.out_cap = output_size;
    if (tb_engine_run_dec_job(&job, d->window, input + consumed, output) != 0) return TAMP_ERROR;
    d->bit_buffer = job.st.bit_buffer;
    d->window_pos = job.st.window_pos;
    d->bit_buffer_pos = job.st.bit_buffer_pos;
    d->token_state = job.st.token_state;
    d->pending_window_offset = job.st.pending_window_offset;
    d->pending_match_size = job.st.pending_match_size;
    d->skip_bytes = job.st.skip_bytes;
    d->last_was_flush = job.st.last_was_flush;
    consumed += (size_t)job.in_consumed;
    if (output_written_size) *output_written_size = (size_t)job.out_written;
    if (input_consumed_size) *input_consumed_size = consumed;
    tamp_res res = (tamp_res)job.res;
    /* One progress report per call (documented deviation: the reference reports once per token). */
    if (callback && res >= 0 && job.in_consumed) {
        int cb = callback(user_data, consumed, input_size);
        if (cb) return (tamp_res)cb;
    }
    return res;
}
