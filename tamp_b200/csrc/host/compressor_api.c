/* Host C layer of the compressor API (include/tamp/compressor.h).
 *
 * init / sink / full are pure host logic, exactly as in the reference (compressor.c:191-245,
 * :665-679, :77-79).  Every entry point that produces compressed bytes converts the caller's
 * TampCompressor into the wire state (tb_wire.h), runs the CUDA codec kernel on a batch of one
 * through tb_engine_run_comp_job(), and converts the state back (SURVEY 8b state-in / state-out).
 * There is no CPU implementation of poll/flush/compress in this library.
 */
#include <string.h>

#include "../tb_wire.h"
#include "tamp/compressor.h"

#define FLUSH_CODE 0xABu

static void put_bits(TampCompressor *c, uint32_t bits, unsigned n) {
    c->bit_buffer_pos = (uint8_t)(c->bit_buffer_pos + n);
    c->bit_buffer |= bits << (32 - c->bit_buffer_pos);
}

tamp_res tamp_compressor_init(TampCompressor *compressor, const TampConf *conf, unsigned char *window) {
    TampConf defaults;
    memset(&defaults, 0, sizeof defaults);
    defaults.window = 10;
    defaults.literal = 8;
    defaults.extended = 1; /* conf == NULL selects the v2 format (compressor.c:193-204) */
    if (!conf) conf = &defaults;

    const unsigned wbits = conf->window, lbits = conf->literal; /* bit-fields: copy out before the range checks */
    if (wbits < 8 || wbits > 15) return TAMP_INVALID_CONF;
    if (lbits < 5 || lbits > 8) return TAMP_INVALID_CONF;
    if (conf->append && (!conf->dictionary_reset || conf->use_custom_dictionary)) return TAMP_INVALID_CONF;

    const TampConf kept = *conf; /* conf may alias compressor->conf */
    memset(compressor, 0, sizeof *compressor);
    compressor->conf = kept;
    compressor->window = window;
    compressor->min_pattern_size = (uint8_t)tamp_compute_min_pattern_size(kept.window, kept.literal);
#if TAMP_LAZY_MATCHING
    compressor->cached_match_index = -1;
#endif
    if (!kept.use_custom_dictionary)
        tamp_initialize_dictionary(window, (size_t)1 << kept.window, kept.extended ? kept.literal : 8);

    if (kept.append) {
        /* FLUSH padded to two bytes: pairs with the previous segment's trailing FLUSH (compressor.c:227-234) */
        put_bits(compressor, FLUSH_CODE, 9);
        compressor->bit_buffer_pos = 16;
        compressor->last_was_flush = 1;
    } else {
        uint32_t header = ((uint32_t)(kept.window - 8) << 5) | ((uint32_t)(kept.literal - 5) << 3) |
                          ((uint32_t)kept.use_custom_dictionary << 2) | ((uint32_t)kept.extended << 1) |
                          (uint32_t)kept.dictionary_reset;
        put_bits(compressor, header, 8);
        if (kept.dictionary_reset) compressor->bit_buffer_pos = (uint8_t)(compressor->bit_buffer_pos + 8);
    }
    return TAMP_OK;
}

bool tamp_compressor_full(const TampCompressor *compressor) {
    return compressor->input_size == sizeof compressor->input;
}

void tamp_compressor_sink(TampCompressor *compressor, const unsigned char *input, size_t input_size,
                          size_t *consumed_size) {
    size_t room = sizeof compressor->input - compressor->input_size;
    size_t n = input_size < room ? input_size : room;
    for (size_t i = 0; i < n; i++)
        compressor->input[(compressor->input_pos + compressor->input_size + i) & 0xF] = input[i];
    compressor->input_size = (uint8_t)(compressor->input_size + n);
    if (consumed_size) *consumed_size = n;
}

static void to_wire(const TampCompressor *c, TbCompState *w) {
    memset(w, 0, sizeof *w);
    w->bit_buffer = c->bit_buffer;
    w->window_pos = c->window_pos;
    w->bit_buffer_pos = c->bit_buffer_pos;
    w->input_size = c->input_size;
    w->input_pos = c->input_pos;
    w->min_pattern_size = c->min_pattern_size;
    w->window_bits = (uint8_t)c->conf.window;
    w->literal_bits = (uint8_t)c->conf.literal;
    w->flags = (uint8_t)((c->conf.extended ? TB_F_EXTENDED : 0) | (c->conf.dictionary_reset ? TB_F_DICT_RESET : 0));
    w->rle_count = c->rle_count;
    w->ext_count = c->extended_match_count;
    w->ext_pos = c->extended_match_position;
    w->last_was_flush = c->last_was_flush;
    w->cached_index = -1;
#if TAMP_LAZY_MATCHING
    if (c->conf.lazy_matching) w->flags |= TB_F_LAZY;
    w->cached_index = c->cached_match_index;
    w->cached_size = c->cached_match_size;
#endif
    memcpy(w->input, c->input, 16);
}

static void from_wire(TampCompressor *c, const TbCompState *w) {
    c->bit_buffer = w->bit_buffer;
    c->window_pos = w->window_pos;
    c->bit_buffer_pos = w->bit_buffer_pos;
    c->input_size = w->input_size;
    c->input_pos = w->input_pos;
    c->rle_count = w->rle_count;
    c->extended_match_count = w->ext_count;
    c->extended_match_position = w->ext_pos;
    c->last_was_flush = w->last_was_flush;
#if TAMP_LAZY_MATCHING
    c->cached_match_index = w->cached_index;
    c->cached_match_size = w->cached_size;
#endif
    memcpy(c->input, w->input, 16);
}

static tamp_res run_job(TampCompressor *c, uint32_t op, unsigned char *output, size_t output_size,
                        size_t *output_written_size, const unsigned char *input, size_t input_size,
                        size_t *input_consumed_size, bool write_token) {
    TbCompJob job;
    memset(&job, 0, sizeof job);
    to_wire(c, &job.st);
    job.op = op;
    job.write_token = write_token ? 1u : 0u;
    job.in_size = input_size;
    job.out_cap = output_size;
    if (output_written_size) *output_written_size = 0;
    if (input_consumed_size) *input_consumed_size = 0;
    if (tb_engine_run_comp_job(&job, c->window, input, output) != 0) return TAMP_ERROR;
    from_wire(c, &job.st);
    if (output_written_size) *output_written_size = (size_t)job.out_written;
    if (input_consumed_size) *input_consumed_size = (size_t)job.in_consumed;
    return (tamp_res)job.res;
}

tamp_res tamp_compressor_poll(TampCompressor *compressor, unsigned char *output, size_t output_size,
                              size_t *output_written_size) {
    return run_job(compressor, TB_OP_POLL, output, output_size, output_written_size, NULL, 0, NULL, false);
}

tamp_res tamp_compressor_flush(TampCompressor *compressor, unsigned char *output, size_t output_size,
                               size_t *output_written_size, bool write_token) {
    return run_job(compressor, TB_OP_FLUSH, output, output_size, output_written_size, NULL, 0, NULL, write_token);
}

tamp_res tamp_compressor_compress_cb(TampCompressor *compressor, unsigned char *output, size_t output_size,
                                     size_t *output_written_size, const unsigned char *input, size_t input_size,
                                     size_t *input_consumed_size, tamp_callback_t callback, void *user_data) {
    size_t consumed = 0;
    tamp_res res = run_job(compressor, TB_OP_COMPRESS, output, output_size, output_written_size, input, input_size,
                           &consumed, false);
    if (input_consumed_size) *input_consumed_size = consumed;
    if (res != TAMP_OK) return res;
    /* One progress report per call (documented deviation: the reference reports once per token). */
    if (callback && input_size) {
        int cb = callback(user_data, consumed, input_size);
        if (cb) return (tamp_res)cb;
    }
    return TAMP_OK;
}

tamp_res tamp_compressor_compress_and_flush_cb(TampCompressor *compressor, unsigned char *output,
                                               size_t output_size, size_t *output_written_size,
                                               const unsigned char *input, size_t input_size,
                                               size_t *input_consumed_size, bool write_token,
                                               tamp_callback_t callback, void *user_data) {
    tamp_res res = run_job(compressor, TB_OP_COMPRESS_AND_FLUSH, output, output_size, output_written_size, input,
                           input_size, input_consumed_size, write_token);
    if (res != TAMP_OK) return res;
    if (callback) {
        int cb = callback(user_data, input_size, input_size); /* 100 % report, compressor.c:837-842 */
        if (cb) return (tamp_res)cb;
    }
    return TAMP_OK;
}

tamp_res tamp_compressor_reset_dictionary(TampCompressor *compressor, unsigned char *output, size_t output_size,
                                          size_t *output_written_size) {
    /* compressor.c:847-881: two FLUSH tokens (the double-FLUSH reset signal), then re-initialise the
     * window and discard the header that init queues. */
    if (!compressor->conf.dictionary_reset) return TAMP_INVALID_CONF;
    size_t total = 0;
    if (output_written_size) *output_written_size = 0;
    for (int i = 0; i < 2; i++) {
        size_t n = 0;
        compressor->last_was_flush = 0;
        tamp_res res = tamp_compressor_flush(compressor, output, output_size, &n, true);
        total += n;
        if (output_written_size) *output_written_size = total;
        if (res != TAMP_OK) return res;
        output += n;
        output_size -= n;
    }
    TampConf conf = compressor->conf;
    conf.use_custom_dictionary = 0;
    tamp_res res = tamp_compressor_init(compressor, &conf, compressor->window);
    compressor->bit_buffer = 0;
    compressor->bit_buffer_pos = 0;
    return res;
}
