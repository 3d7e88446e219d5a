/* tamp-b200: ABI-compatible declarations for the shared part of the Tamp C API.
 *
 * Drop-in for the declarations in the reference's tamp/_c_src/tamp/common.h (BrianPugh/tamp @ 48880ad):
 *   - tamp_res status codes ............ common.h:145-168
 *   - TampConf bit-field ............... common.h:170-182
 *   - tamp_callback_t .................. common.h:210
 *   - tamp_initialize_dictionary ....... common.h:395  (common.c:37-52)
 *   - tamp_compute_min_pattern_size .... common.h:405  (common.c:54-56)
 *   - tamp_window_copy ................. common.h:424  (common.c:58-86)
 * Layouts, enum values and symbol names are identical so that callers compiled against the
 * reference headers (Cython ctamp.pxd:33-106, mpy_bindings, wasm, C programs) link unchanged.
 *
 * Build-macro coupling: this library is built with the reference's defaults —
 * TAMP_EXTENDED=1, TAMP_ESP32=0 — and honours TAMP_LAZY_MATCHING (default 0) because that macro
 * changes TampConf/TampCompressor field sets.  The stream API (tamp_compress_stream / tamp_decompress_stream,
 * common.h:225-330) is provided with the memory and stdio handlers; the LittleFS / FatFs
 * handlers (embedded file systems) are not.
 */
#ifndef TAMP_COMMON_H
#define TAMP_COMMON_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef TAMP_LAZY_MATCHING
#define TAMP_LAZY_MATCHING 0
#endif
#define TAMP_EXTENDED 1
#define TAMP_EXTENDED_COMPRESS 1
#define TAMP_EXTENDED_DECOMPRESS 1

/* Token symbols / limits of the extended (v2) format. */
#define TAMP_RLE_SYMBOL 12
#define TAMP_EXTENDED_MATCH_SYMBOL 13
#define TAMP_LEADING_EXTENDED_MATCH_BITS 3
#define TAMP_LEADING_RLE_BITS 4
#define TAMP_RLE_MAX_WINDOW 8

/* Status codes.  >= 0: recoverable / informational, < 0: error.  [100,127] and [-128,-100] are left to
 * user callbacks. */
enum {
    TAMP_OK = 0,
    TAMP_OUTPUT_FULL = 1,     /* call again with more output room */
    TAMP_INPUT_EXHAUSTED = 2, /* normal end-of-input result of decompress */
    TAMP_ERROR = -1,
    TAMP_EXCESS_BITS = -2,  /* a literal does not fit in conf.literal bits */
    TAMP_INVALID_CONF = -3,
    TAMP_OOB = -4,          /* compressed data references bytes outside the window */
    TAMP_IO_ERROR = -10,
    TAMP_READ_ERROR = -11,
    TAMP_WRITE_ERROR = -12
};
typedef int8_t tamp_res;

typedef struct TampConf {
    uint16_t window : 4;                /* window bits, 8..15 */
    uint16_t literal : 4;               /* literal bits, 5..8 */
    uint16_t use_custom_dictionary : 1; /* caller pre-filled the window */
    uint16_t extended : 1;              /* v2 format: RLE + extended match */
    uint16_t dictionary_reset : 1;      /* two-byte header; stream may carry double-FLUSH resets */
    uint16_t append : 1;                /* start with a FLUSH instead of a header */
#if TAMP_LAZY_MATCHING
    uint16_t lazy_matching : 1;
#endif
} TampConf;

/* Progress callback.  Return non-zero to abort; the value is truncated to tamp_res.
 * DEVIATION (documented in INTEGRATION.md): the reference fires it once per token on the host
 * (compressor.c:717, decompressor.c:574); here a call is one kernel launch, so it fires once per call. */
typedef int (*tamp_callback_t)(void *user_data, size_t bytes_processed, size_t total_bytes);

/* Stream callbacks (reference common.h:225, :242): fread / fwrite shaped.  read: bytes read, 0 at end of input,
 * negative on error (-> TAMP_READ_ERROR).  write: must take all `size` bytes; negative or short -> TAMP_WRITE_ERROR.
 * DEVIATION: the reference moves TAMP_STREAM_WORK_BUFFER_SIZE/2 (16) bytes per callback; here a call into the codec
 * is a device round trip, so the stream functions move TAMP_B200_STREAM_CHUNK bytes per callback.  The bytes that
 * flow through the callbacks, concatenated, are identical. */
typedef int (*tamp_read_t)(void *handle, unsigned char *buffer, size_t size);
typedef int (*tamp_write_t)(void *handle, const unsigned char *buffer, size_t size);
#ifndef TAMP_B200_STREAM_CHUNK
#define TAMP_B200_STREAM_CHUNK 16384
#endif

/* Built-in handlers: memory buffers (reference common.h:257-313) and stdio FILE* (:316-330). */
typedef struct TampMemReader {
    const unsigned char *data;
    size_t size;
    size_t pos; /* start at 0 */
} TampMemReader;
typedef struct TampMemWriter {
    unsigned char *data;
    size_t capacity;
    size_t pos; /* start at 0; bytes written so far */
} TampMemWriter;
int tamp_stream_mem_read(void *handle, unsigned char *buffer, size_t size);        /* handle: TampMemReader* */
int tamp_stream_mem_write(void *handle, const unsigned char *buffer, size_t size); /* handle: TampMemWriter*; -1 if it would overflow */
int tamp_stream_stdio_read(void *handle, unsigned char *buffer, size_t size);        /* handle: FILE* */
int tamp_stream_stdio_write(void *handle, const unsigned char *buffer, size_t size); /* handle: FILE* */

void tamp_initialize_dictionary(unsigned char *buffer, size_t size, uint8_t literal);
int8_t tamp_compute_min_pattern_size(uint8_t window, uint8_t literal);
void tamp_window_copy(unsigned char *window, uint16_t *window_pos, uint16_t window_offset, uint8_t match_size,
                      uint16_t window_mask);

#ifdef __cplusplus
}
#endif
#endif
