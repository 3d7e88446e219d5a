/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement ("oracle") of the Tamp LZSS hot path of BrianPugh/tamp @ 48880ad:
 * tamp/_c_src/tamp/compressor.c (sink/poll match search, bit writer, RLE + extended-match state
 * machine, flush) and tamp/_c_src/tamp/decompressor.c (bit reader, Huffman decode, window copy).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  The product (tamp_b200/) never includes, links or calls anything in oracle/.
 *
 * Parity pin: this restatement is checked (tests/test_oracle_*.py) against
 *   - every golden bitstream the reference's own tests hold for this path (tests/golden/reference_kats.json,
 *     transcribed from tests/test_compressor.py, tests/test_decompressor.py, ctests/test_compressor.c,
 *     ctests/test_decompressor.c, tests/test_pseudorandom.py), and
 *   - the unmodified reference C compiled here into oracle/_ref/libtamp_ref.so, differentially on
 *     thousands of seeded streams (fixtures in tests/golden/ref_fixtures.json made by
 *     tests/golden/make_fixtures.py).
 */
#ifndef TAMP_ORACLE_H
#define TAMP_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes: same numeric values as tamp_res (common.h:145-168). */
enum {
    ORC_OK = 0,
    ORC_OUTPUT_FULL = 1,
    ORC_INPUT_EXHAUSTED = 2,
    ORC_ERROR = -1,
    ORC_EXCESS_BITS = -2,
    ORC_INVALID_CONF = -3,
    ORC_OOB = -4,
};

typedef struct OracleConf {
    int32_t window;            /* 8..15 */
    int32_t literal;           /* 5..8 */
    int32_t use_custom_dictionary;
    int32_t extended;
    int32_t dictionary_reset;
    int32_t lazy_matching;     /* compile-time TAMP_LAZY_MATCHING + conf.lazy_matching in the reference */
} OracleConf;

typedef struct OracleEnc OracleEnc;

/* common.c */
void oracle_initialize_dictionary(uint8_t *buf, size_t size, int literal);
int oracle_min_pattern_size(int window, int literal);

/* Streaming encoder with unbounded output (mirrors init / compress / flush call sequences). */
OracleEnc *oracle_enc_new(const OracleConf *conf, const uint8_t *dictionary, int *status);
int oracle_enc_write(OracleEnc *e, const uint8_t *data, size_t n);   /* == tamp_compressor_compress */
int oracle_enc_flush(OracleEnc *e, int write_token);                  /* == tamp_compressor_flush */
size_t oracle_enc_size(const OracleEnc *e);
const uint8_t *oracle_enc_data(const OracleEnc *e);
const uint8_t *oracle_enc_window(const OracleEnc *e);
void oracle_enc_free(OracleEnc *e);

/* One-shot: init + compress_and_flush(write_token).  Returns compressed size, or a negative
 * status; -100 if `cap` is too small. */
long oracle_compress(const OracleConf *conf, const uint8_t *dictionary, const uint8_t *in, size_t n, uint8_t *out,
                     size_t cap, int write_token);

/* One-shot decoder: tamp_decompressor_init(conf=NULL, window_bits_max) + one decompress call over the
 * whole frame with `cap` bytes of output room.  `dictionary` (1<<window bytes) is used when the
 * header's custom-dictionary bit is set.  Returns bytes written; *status receives the tamp_res. */
long oracle_decompress(const uint8_t *dictionary, int window_bits_max, const uint8_t *in, size_t n, uint8_t *out,
                       size_t cap, int *status);

/* Executable spec of find_best_match (fuzz/esp32_host/differential.cpp:50-67 semantics): exhaustive,
 * lowest index among the longest, >=2 bytes.  Returns length (0 if none); *index set when >0. */
int oracle_find_best_match(const uint8_t *window, int window_size, const uint8_t *pattern, int max_len,
                           int *index);

#ifdef __cplusplus
}
#endif
#endif
