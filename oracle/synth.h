/* TEST INFRASTRUCTURE (oracle/): deterministic synthetic stream generators, CPU side.
 *
 * Definitions follow SURVEY.md 8(d).  The PRNG is xorshift32, the same generator the reference
 * uses for its seed dictionary (tamp/_c_src/tamp/common.c:28-35) and for its on-device stress
 * streams (devices/common/tamp_bench.c:33-40).  Kinds 2 and 3 restate the reference's stress
 * generators 1 and 2 (devices/common/tamp_bench.c:57-65) with per-stream seeding.
 *
 * The product library has its own device-side generator (tamp_b200/csrc/synth.cuh); tests check the
 * two agree byte for byte.  Nothing in the product path includes this file.
 */
#ifndef ORACLE_SYNTH_H
#define ORACLE_SYNTH_H
#include <stddef.h>
#include <stdint.h>
#include <string.h>

enum {
    SYNTH_TEXT = 0,     /* G_text: 256-word vocabulary, Zipf-ish word picks, ' ' / '\n' separators */
    SYNTH_RAND = 1,     /* G_rand: uniform printable ASCII 0x20..0x7e (no matches: worst case) */
    SYNTH_ALPHA16 = 2,  /* reference stress generator 1: 16-letter alphabet */
    SYNTH_PERIODIC = 3, /* reference stress generator 2: period-64 ramp with a random byte every 50 */
    SYNTH_BINARY = 4,   /* reference stress generator 0: uniform bytes (needs literal=8) */
    SYNTH_RUNS = 5,     /* run-heavy text: exercises RLE / extended-match paths */
};

static inline uint32_t synth_xs(uint32_t *s) {
    uint32_t x = *s;
    x ^= x << 13;
    x ^= x >> 17;
    x ^= x << 5;
    return *s = x;
}

static inline uint32_t synth_stream_seed(uint64_t k) {
    uint32_t s = (uint32_t)(0xC0FFEE01u + (uint32_t)k * 0x9E3779B9u);
    return s ? s : 1u;
}

/* vocab: 256 entries, each up to 9 chars; vocab_len[i] in 2..9 */
typedef struct {
    uint8_t len[256];
    uint8_t chars[256][12];
} SynthVocab;

static inline void synth_build_vocab(SynthVocab *v) {
    uint32_t s = 0x1234ABCDu;
    for (int i = 0; i < 256; i++) {
        int len = 2 + (int)(synth_xs(&s) % 8u);
        v->len[i] = (uint8_t)len;
        memset(v->chars[i], 0, sizeof v->chars[i]);
        for (int j = 0; j < len; j++) v->chars[i][j] = (uint8_t)('a' + synth_xs(&s) % 26u);
    }
}

static inline void synth_fill(int kind, uint64_t k, uint8_t *out, size_t n, const SynthVocab *vocab) {
    uint32_t s = synth_stream_seed(k);
    size_t i = 0;
    switch (kind) {
        case SYNTH_TEXT:
            while (i < n) {
                uint32_t r = synth_xs(&s);
                uint32_t w = r & 0xFFu;
                if (r & 0x100u) w &= 0x3Fu;
                if (r & 0x200u) w &= 0x0Fu;
                for (int j = 0; j < vocab->len[w] && i < n; j++) out[i++] = vocab->chars[w][j];
                if (i < n) out[i++] = (((r >> 12) & 15u) == 0) ? '\n' : ' ';
            }
            break;
        case SYNTH_RAND:
            for (; i < n; i++) out[i] = (uint8_t)(0x20u + synth_xs(&s) % 95u);
            break;
        case SYNTH_ALPHA16:
            for (; i < n; i++) out[i] = (uint8_t)('a' + (synth_xs(&s) & 0x0Fu));
            break;
        case SYNTH_PERIODIC:
            for (; i < n; i++) {
                if ((i % 50) == 0)
                    out[i] = (uint8_t)synth_xs(&s);
                else
                    out[i] = (uint8_t)(((i & 63) * 37 + 11) & 0xFF);
            }
            break;
        case SYNTH_BINARY:
            for (; i < n; i++) out[i] = (uint8_t)synth_xs(&s);
            break;
        default: /* SYNTH_RUNS: words interleaved with runs of one character, run length 1..40 */
            while (i < n) {
                uint32_t r = synth_xs(&s);
                if (r & 1u) {
                    uint32_t w = (r >> 8) & 0x1Fu;
                    for (int j = 0; j < vocab->len[w] && i < n; j++) out[i++] = vocab->chars[w][j];
                } else {
                    uint32_t run = 1u + ((r >> 4) % 40u);
                    uint8_t c = (uint8_t)("ab \n0x"[(r >> 16) % 6u]);
                    for (uint32_t j = 0; j < run && i < n; j++) out[i++] = c;
                }
            }
            break;
    }
}
#endif
