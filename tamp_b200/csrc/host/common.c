/* Host-side shared helpers of the Tamp C API, kept on the CPU as in the reference
 * (BASELINE.json north_star: "common.c dictionary-init kept").
 *
 * Behaviour follows tamp/_c_src/tamp/common.c of BrianPugh/tamp @ 48880ad:
 *   tamp_initialize_dictionary ...... common.c:37-52 (seed tables :18-25, xorshift32 :28-35)
 *   tamp_compute_min_pattern_size ... common.c:54-56
 *   tamp_window_copy ................ common.c:58-86
 */
#include "tamp/common.h"

/* Seed alphabet for literal 7/8 (and every v1 stream); literal 5/6 use common English letters masked
 * to the literal width so that every seeded byte is representable. */
static const unsigned char k_seed_wide[16] = {0x20, 0x00, 0x30, 0x65, 0x69, 0x3e, 0x74, 0x6f,
                                              0x3c, 0x61, 0x6e, 0x73, 0x0a, 0x72, 0x2f, 0x2e};
static const char k_seed_english[] = " etaoinshrdlcumw";

void tamp_initialize_dictionary(unsigned char *buffer, size_t size, uint8_t literal) {
    unsigned char alphabet[16];
    const unsigned narrow_mask = literal <= 5 ? 0x1Fu : 0x3Fu;
    for (int i = 0; i < 16; i++)
        alphabet[i] = literal <= 6 ? (unsigned char)((unsigned)k_seed_english[i] & narrow_mask) : k_seed_wide[i];

    uint32_t state = 3758097560u; /* seed fixed by the format (specification.rst:90-157) */
    size_t i = 0;
    while (i < size) {
        state ^= state << 13;
        state ^= state >> 17;
        state ^= state << 5;
        uint32_t nibbles = state; /* eight 4-bit picks per draw, low nibble first */
        for (int j = 0; j < 8 && i < size; j++, i++, nibbles >>= 4) buffer[i] = alphabet[nibbles & 0xFu];
    }
}

int8_t tamp_compute_min_pattern_size(uint8_t window, uint8_t literal) {
    /* A 2-byte match only pays off while (huffman + window) bits < 2 literals. */
    return (int8_t)(window > 10 + 2 * (literal - 5) ? 3 : 2);
}

void tamp_window_copy(unsigned char *window, uint16_t *window_pos, uint16_t window_offset, uint8_t match_size,
                      uint16_t window_mask) {
    /* The destination wraps, the source never does (validated by the caller).  When the destination
     * starts inside the source range the bytes are moved back to front so that every source byte is
     * read before it can be overwritten; otherwise front to back.  Either way the result equals
     * "snapshot the source, then write". */
    const uint16_t pos = *window_pos;
    const uint16_t gap = (uint16_t)((pos - window_offset) & window_mask);
    if (gap != 0 && gap < match_size) {
        for (int i = (int)match_size - 1; i >= 0; i--)
            window[(pos + i) & window_mask] = window[window_offset + i];
    } else {
        for (int i = 0; i < (int)match_size; i++) window[(pos + i) & window_mask] = window[window_offset + i];
    }
    *window_pos = (uint16_t)((pos + match_size) & window_mask);
}
