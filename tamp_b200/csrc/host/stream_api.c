/* Stream API of the Tamp C library over the per-call API: callback readers / writers on the
 * host, the codec calls behind them are the same device round trips as everywhere else (no CPU codec here).
 *
 * Replaces tamp_compress_stream (reference compressor.c:891-955), tamp_decompress_stream (decompressor.c:585-640)
 * and the memory / stdio handlers declared in common.h:257-330.  Observable behaviour kept: the bytes handed to
 * write_cb, concatenated; the byte counters; TAMP_READ_ERROR / TAMP_WRITE_ERROR on a negative read, a negative or
 * short write; codec errors passed through; the progress callback's arguments and its abort rule.  What differs is
 * granularity only: the reference moves 16 bytes per callback, this moves TAMP_B200_STREAM_CHUNK. */
#include <stdio.h>
#include <string.h>

#include "tamp/compressor.h"
#include "tamp/decompressor.h"

typedef struct Outlet {
    tamp_write_t write;
    void *handle;
    size_t *total;
} Outlet;

/* All n bytes or TAMP_WRITE_ERROR. */
static tamp_res outlet_put(const Outlet *o, const unsigned char *bytes, size_t n) {
    if (n == 0) return TAMP_OK;
    const int took = o->write(o->handle, bytes, n);
    if (took < 0 || (size_t)took != n) return TAMP_WRITE_ERROR;
    *o->total += n;
    return TAMP_OK;
}

tamp_res tamp_compress_stream(TampCompressor *compressor, tamp_read_t read_cb, void *read_handle, tamp_write_t write_cb,
                              void *write_handle, size_t *input_consumed_size, size_t *output_written_size,
                              tamp_callback_t callback, void *user_data) {
    size_t consumed_local, written_local;
    size_t *consumed = input_consumed_size ? input_consumed_size : &consumed_local;
    size_t *written = output_written_size ? output_written_size : &written_local;
    *consumed = 0;
    *written = 0;
    const Outlet outlet = {write_cb, write_handle, written};
    unsigned char fresh[TAMP_B200_STREAM_CHUNK], packed[TAMP_B200_STREAM_CHUNK];

    for (;;) {
        const int got = read_cb(read_handle, fresh, sizeof fresh);
        if (got < 0) return TAMP_READ_ERROR;
        if (got == 0) break; /* end of input */
        *consumed += (size_t)got;
        for (size_t at = 0; at < (size_t)got;) {
            size_t used = 0, made = 0;
            tamp_res res = tamp_compressor_compress(compressor, packed, sizeof packed, &made, fresh + at,
                                                    (size_t)got - at, &used);
            if (res < TAMP_OK) return res;
            at += used;
            if ((res = outlet_put(&outlet, packed, made)) != TAMP_OK) return res;
        }
        if (callback) {
            const int stop = callback(user_data, *consumed, 0);
            if (stop) return (tamp_res)stop;
        }
    }
    /* drain the ring and the bit buffer; no FLUSH token (the stream simply ends) */
    for (;;) {
        size_t made = 0;
        const tamp_res res = tamp_compressor_flush(compressor, packed, sizeof packed, &made, false);
        if (res < TAMP_OK) return res;
        const tamp_res put = outlet_put(&outlet, packed, made);
        if (put != TAMP_OK) return put;
        if (res == TAMP_OK) return TAMP_OK; /* TAMP_OUTPUT_FULL: more to come */
    }
}

tamp_res tamp_decompress_stream(TampDecompressor *decompressor, tamp_read_t read_cb, void *read_handle,
                                tamp_write_t write_cb, void *write_handle, size_t *input_consumed_size,
                                size_t *output_written_size, tamp_callback_t callback, void *user_data) {
    size_t consumed_local, written_local;
    size_t *consumed = input_consumed_size ? input_consumed_size : &consumed_local;
    size_t *written = output_written_size ? output_written_size : &written_local;
    *consumed = 0;
    *written = 0;
    const Outlet outlet = {write_cb, write_handle, written};
    unsigned char packed[TAMP_B200_STREAM_CHUNK], plain[TAMP_B200_STREAM_CHUNK];
    size_t at = 0, left = 0; /* unread part of packed[] */
    bool at_end = false;

    for (;;) {
        if (left == 0 && !at_end) {
            const int got = read_cb(read_handle, packed, sizeof packed);
            if (got < 0) return TAMP_READ_ERROR;
            at_end = got == 0;
            at = 0;
            left = (size_t)got;
            *consumed += (size_t)got;
        }
        size_t used = 0, made = 0;
        tamp_res res = tamp_decompressor_decompress(decompressor, plain, sizeof plain, &made, packed + at, left, &used);
        if (res < TAMP_OK) return res;
        at += used;
        left -= used;
        const tamp_res put = outlet_put(&outlet, plain, made);
        if (put != TAMP_OK) return put;
        if (res == TAMP_INPUT_EXHAUSTED && at_end) return TAMP_OK;
        if (callback) {
            const int stop = callback(user_data, *consumed, 0);
            if (stop) return (tamp_res)stop;
        }
    }
}

/* ---- built-in handlers ---------------------------------------------------------------------------------------- */

int tamp_stream_mem_read(void *handle, unsigned char *buffer, size_t size) {
    TampMemReader *r = (TampMemReader *)handle;
    const size_t rest = r->size - r->pos;
    const size_t n = size < rest ? size : rest;
    memcpy(buffer, r->data + r->pos, n);
    r->pos += n;
    return (int)n;
}

int tamp_stream_mem_write(void *handle, const unsigned char *buffer, size_t size) {
    TampMemWriter *w = (TampMemWriter *)handle;
    if (size > w->capacity - w->pos) return -1; /* would overflow: nothing is written */
    memcpy(w->data + w->pos, buffer, size);
    w->pos += size;
    return (int)size;
}

int tamp_stream_stdio_read(void *handle, unsigned char *buffer, size_t size) {
    FILE *f = (FILE *)handle;
    const size_t n = fread(buffer, 1, size, f);
    return (n == 0 && ferror(f)) ? -1 : (int)n;
}

int tamp_stream_stdio_write(void *handle, const unsigned char *buffer, size_t size) {
    FILE *f = (FILE *)handle;
    const size_t n = fwrite(buffer, 1, size, f);
    return (n < size && ferror(f)) ? -1 : (int)n; /* a short count is the stream functions' TAMP_WRITE_ERROR */
}
