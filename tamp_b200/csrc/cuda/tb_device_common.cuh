// Device-side constants and helpers shared by every tamp-b200 kernel.
//
// Format facts restated from the Tamp specification (docs/source/specification.rst:159-186) and the
// reference tables (tamp/_c_src/tamp/compressor.c:33-36, decompressor.c:52-57).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tb {

constexpr int kOk = 0, kOutputFull = 1, kInputExhausted = 2, kError = -1, kExcessBits = -2, kInvalidConf = -3,
              kOob = -4;
constexpr int kPollContinue = 127;

constexpr int kSymRle = 12, kSymExt = 13, kSymFlush = 14;
constexpr int kRleMax = 241;       // (14 << 4) + 15 + 2
constexpr int kRleWindowMax = 8;   // an RLE token writes at most 8 bytes into the window
constexpr int kExtExtraMax = 120;  // (14 << 3) + 7 + 1
constexpr int kExtMinOutput = 6;   // output bytes that must be free before an extended-match token

// Static Huffman code of (match_len - min_pattern) 0..13 and FLUSH (14).  Bit counts include the
// leading 0 "token" flag.
struct HuffTables {
    uint8_t code[16];
    uint8_t bits[16];
    uint8_t lut[128];  // decoder: index = 7 bits after the code's leading '1'; (extra_bits << 4) | symbol
};

constexpr HuffTables make_huff_tables() {
    HuffTables t{};
    constexpr uint8_t code[15] = {0x00, 0x03, 0x08, 0x0b, 0x14, 0x24, 0x26, 0x2b, 0x4b, 0x54, 0x94, 0x95, 0xaa, 0x27, 0xab};
    constexpr uint8_t bits[15] = {2, 3, 5, 5, 6, 7, 7, 7, 8, 8, 9, 9, 9, 7, 9};
    for (int i = 0; i < 15; i++) {
        t.code[i] = code[i];
        t.bits[i] = bits[i];
    }
    // Derive the decoder LUT from the code table: symbol s >= 1 is '1' followed by (bits-2) more bits.
    for (int idx = 0; idx < 128; idx++) {
        for (int s = 1; s < 15; s++) {
            int extra = bits[s] - 2;                              // bits after flag and leading '1'
            int tail = code[s] & ((1 << extra) - 1);              // those bits
            if ((idx >> (7 - extra)) == tail) t.lut[idx] = (uint8_t)((extra << 4) | s);
        }
    }
    return t;
}

static __device__ __constant__ HuffTables kHuff = make_huff_tables();

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// Tamp seed dictionary (common.c:37-52) restated for the device.  One xorshift32 draw feeds 8 bytes, so
// the fill is inherently serial; it is used only on the rare double-FLUSH reset path (the batch kernels
// stage a host-seeded table instead).
__device__ inline void seed_dictionary_serial(uint8_t *dst, int size, int literal_for_seed) {
    const uint8_t wide[16] = {0x20, 0x00, 0x30, 0x65, 0x69, 0x3e, 0x74, 0x6f,
                              0x3c, 0x61, 0x6e, 0x73, 0x0a, 0x72, 0x2f, 0x2e};
    const char english[17] = " etaoinshrdlcumw";
    uint32_t s = 3758097560u;
    for (int base = 0; base < size; base += 8) {
        s ^= s << 13;
        s ^= s >> 17;
        s ^= s << 5;
        uint32_t r = s;
        for (int j = 0; j < 8 && base + j < size; j++, r >>= 4) {
            uint32_t n = r & 15u;
            uint8_t c = literal_for_seed <= 5   ? (uint8_t)(english[n] & 0x1F)
                        : literal_for_seed <= 6 ? (uint8_t)(english[n] & 0x3F)
                                                : wide[n];
            dst[base + j] = c;
        }
    }
}

// How a batch stream's frame starts (compressor.c:227-243): the header byte (+ a zero byte with dictionary_reset), or —
// append mode — FLUSH (0xAB in 9 bits) padded to 16 bits: 55 80.  Both forms of a dictionary_reset stream are 16 bits.
constexpr uint32_t kAppendStart = 0x55800000u;
__device__ __forceinline__ bool stream_appends(int flags, uint64_t stream) {
    return (flags & TB_F_APPEND) && (stream > 0 || !(flags & TB_F_APPEND_TAIL));
}
// compressor.c:784-794: flush(write_token) ends the frame with a FLUSH token unless the frame is byte-aligned without
// dictionary_reset — or the FLUSH of an append-mode start is still the last thing written (empty input).
__device__ __forceinline__ bool ends_with_flush(int write_token, uint32_t nbits, int flags, uint64_t stream, uint64_t n_in) {
    return write_token && ((nbits & 7u) || (flags & TB_F_DICT_RESET)) && !(n_in == 0 && stream_appends(flags, stream));
}

// The first two bytes of a batch frame as a decoder reads them: header byte | second byte << 8 (0 if the frame has one
// byte).  Behind the first segment of a segmented stream (seg_header) the frame starts with the append-mode marker 55 80:
// it reads as segment 0's header followed by a zero byte; any other start reads as a bad second header byte
// (INVALID_CONF, decompressor.c:276-297).  `in` must hold n >= 1 bytes.
__device__ __forceinline__ uint32_t frame_start(uint32_t seg_header, uint64_t stream, const uint8_t *in, uint32_t n) {
    uint32_t h = in[0], b1 = n >= 2 ? in[1] : 0u;
    if (seg_header && (stream > 0 || (seg_header & 0x200u))) {  // 0x200: stream 0 of this batch is not segment 0 either
        const bool marker = n >= 2 && h == (kAppendStart >> 24) && b1 == ((kAppendStart >> 16) & 0xFFu);
        h = (seg_header & 0xFFu) | 1u;
        b1 = marker || n < 2 ? 0u : 0xFFu;
    }
    return h | (b1 << 8);
}

__device__ __forceinline__ int min_pattern_size(int window, int literal) {
    return window > 10 + 2 * (literal - 5) ? 3 : 2;
}

}  // namespace tb
