// Deterministic synthetic stream generators, device side: one thread per stream.
// Definitions: xorshift32 PRNG (same generator as the reference's seed dictionary, common.c:28-35, and
// its stress streams, devices/common/tamp_bench.c:33-40); kinds 2/3/4 restate the reference's stress
// generators 1/2/0 (tamp_bench.c:49-74) with per-stream seeding.
#pragma once
#include <stdint.h>

namespace tb {

struct SynthVocab {
    uint8_t len[256];
    uint8_t chars[256][12];
};

__host__ __device__ inline uint32_t synth_next(uint32_t &s) {
    s ^= s << 13;
    s ^= s >> 17;
    s ^= s << 5;
    return s;
}

inline void synth_build_vocab(SynthVocab &v) {
    uint32_t s = 0x1234ABCDu;
    for (int i = 0; i < 256; i++) {
        int len = 2 + (int)(synth_next(s) % 8u);
        v.len[i] = (uint8_t)len;
        for (int j = 0; j < 12; j++) v.chars[i][j] = 0;
        for (int j = 0; j < len; j++) v.chars[i][j] = (uint8_t)('a' + synth_next(s) % 26u);
    }
}

__device__ inline void synth_fill(int kind, uint64_t k, uint8_t *out, uint64_t n, const SynthVocab *vocab) {
    uint32_t s = 0xC0FFEE01u + (uint32_t)k * 0x9E3779B9u;
    if (s == 0) s = 1;
    uint64_t i = 0;
    if (kind == 0) {  // word text
        while (i < n) {
            uint32_t r = synth_next(s);
            uint32_t w = r & 0xFFu;
            if (r & 0x100u) w &= 0x3Fu;
            if (r & 0x200u) w &= 0x0Fu;
            int len = vocab->len[w];
            for (int j = 0; j < len && i < n; j++) out[i++] = vocab->chars[w][j];
            if (i < n) out[i++] = (((r >> 12) & 15u) == 0) ? '\n' : ' ';
        }
    } else if (kind == 1) {  // printable random
        for (; i < n; i++) out[i] = (uint8_t)(0x20u + synth_next(s) % 95u);
    } else if (kind == 2) {  // 16-letter alphabet
        for (; i < n; i++) out[i] = (uint8_t)('a' + (synth_next(s) & 0x0Fu));
    } else if (kind == 3) {  // periodic ramp with a random byte every 50
        for (; i < n; i++) out[i] = (i % 50) == 0 ? (uint8_t)synth_next(s) : (uint8_t)(((i & 63) * 37 + 11) & 0xFF);
    } else if (kind == 4) {  // uniform bytes
        for (; i < n; i++) out[i] = (uint8_t)synth_next(s);
    } else {  // run-heavy
        const char pick[7] = "ab \n0x";
        while (i < n) {
            uint32_t r = synth_next(s);
            if (r & 1u) {
                uint32_t w = (r >> 8) & 0x1Fu;
                int len = vocab->len[w];
                for (int j = 0; j < len && i < n; j++) out[i++] = vocab->chars[w][j];
            } else {
                uint32_t run = 1u + ((r >> 4) % 40u);
                uint8_t c = (uint8_t)pick[(r >> 16) % 6u];
                for (uint32_t j = 0; j < run && i < n; j++) out[i++] = c;
            }
        }
    }
}

}  // namespace tb
