// General-purpose Tamp decompressor for ONE stream, executed by one thread.
//
// Fully resumable (partial input, output-full, split extended tokens), every window size and both
// formats; behaviour restated from tamp/_c_src/tamp/decompressor.c @ 48880ad (line refs inline).
// It backs the per-call C API (batch of one) and the batch path for configurations without a
// specialised kernel, where each thread of the grid decodes its own stream against a window kept in
// global memory (L2-resident scratch).
#pragma once
#include "../tb_wire.h"
#include "tb_device_common.cuh"

namespace tb {

struct DecIo {
    const uint8_t *in;
    size_t in_size, in_pos;
    uint8_t *out;
    size_t out_cap, out_pos;
};

// refill_bit_buffer, decompressor.c:357-365
__device__ __forceinline__ void dec_refill(TbDecState &s, DecIo &io) {
    while (io.in_pos != io.in_size && s.bit_buffer_pos <= 24) {
        s.bit_buffer_pos += 8;
        s.bit_buffer |= (uint32_t)io.in[io.in_pos++] << (32 - s.bit_buffer_pos);
    }
}

// decode_huffman, decompressor.c:71-104.  Works on copies; the caller commits on success.
__device__ __forceinline__ bool dec_huffman(uint32_t &buf, int &pos, int trailing, int &value) {
    if (pos < 1 + trailing) return false;
    int sym;
    pos--;
    if ((buf >> 31) == 0) {
        buf <<= 1;
        sym = 0;
    } else {
        buf <<= 1;
        uint32_t e = kHuff.lut[buf >> 25];
        int extra = (int)(e >> 4);
        if (pos < extra + trailing) return false;
        buf <<= extra;
        pos -= extra;
        sym = (int)(e & 15u);
    }
    if (trailing) {
        uint32_t t = buf >> (32 - trailing);
        buf <<= trailing;
        pos -= trailing;
        value = (sym << trailing) + (int)t;
    } else {
        value = sym;
    }
    return true;
}

// Copy n window bytes [src, src+n) to window_pos.. (destination wraps) with tamp_window_copy's
// snapshot semantics (common.c:58-86): back-to-front when the destination starts inside the source.
__device__ __forceinline__ void dec_window_copy(uint8_t *win, int &wpos, int src, int n, int mask) {
    int gap = (wpos - src) & mask;
    if (gap != 0 && gap < n) {
        for (int i = n - 1; i >= 0; i--) win[(wpos + i) & mask] = win[src + i];
    } else {
        for (int i = 0; i < n; i++) win[(wpos + i) & mask] = win[src + i];
    }
    wpos = (wpos + n) & mask;
}

// decode_rle, decompressor.c:114-174
__device__ inline int dec_rle(TbDecState &s, uint8_t *win, DecIo &io) {
    int count;
    int skip = s.skip_bytes;
    if (skip > 0) {
        count = s.pending_window_offset;
    } else {
        uint32_t buf = s.bit_buffer;
        int pos = s.bit_buffer_pos;
        int raw;
        if (!dec_huffman(buf, pos, 4, raw)) return kInputExhausted;
        s.bit_buffer = buf;
        s.bit_buffer_pos = (uint8_t)pos;
        count = raw + 2;
    }
    const int W = 1 << s.window_bits;
    uint8_t sym = win[(s.window_pos - 1) & (W - 1)];
    int remaining = count - skip;
    size_t space = io.out_cap - io.out_pos;
    int n;
    if ((size_t)remaining > space) {
        n = (int)space;
        s.skip_bytes = (uint8_t)(skip + n);
        s.token_state = 1;
        s.pending_window_offset = (uint16_t)count;
    } else {
        n = remaining;
        s.skip_bytes = 0;
        s.token_state = 0;
    }
    for (int i = 0; i < n; i++) io.out[io.out_pos + i] = sym;
    io.out_pos += (size_t)n;
    if (skip == 0) {  // the window is updated with the first delivered chunk only
        int room = W - s.window_pos;
        int nw = count < kRleWindowMax ? count : kRleWindowMax;
        nw = nw < room ? nw : room;
        for (int i = 0; i < nw; i++) win[s.window_pos + i] = sym;
        s.window_pos = (uint16_t)((s.window_pos + nw) & (W - 1));
    }
    return s.token_state == 0 ? kOk : kOutputFull;
}

// decode_extended_match, decompressor.c:187-273
__device__ inline int dec_ext_match(TbDecState &s, uint8_t *win, DecIo &io) {
    const int wbits = s.window_bits;
    const int W = 1 << wbits;
    int off, len;
    int skip = s.skip_bytes;
    if (skip > 0) {
        off = s.pending_window_offset;
        len = s.pending_match_size;
    } else if (s.token_state == 3) {  // size known, offset still to come
        len = s.pending_match_size;
        if (s.bit_buffer_pos < wbits) return kInputExhausted;
        off = (int)(s.bit_buffer >> (32 - wbits));
        s.bit_buffer <<= wbits;
        s.bit_buffer_pos = (uint8_t)(s.bit_buffer_pos - wbits);
    } else {
        uint32_t buf = s.bit_buffer;
        int pos = s.bit_buffer_pos;
        int raw;
        if (!dec_huffman(buf, pos, 3, raw)) return kInputExhausted;
        len = raw + s.min_pattern_size + 12;
        if (pos < wbits) {  // park the size; the offset arrives with more input
            s.bit_buffer = buf;
            s.bit_buffer_pos = (uint8_t)pos;
            s.token_state = 3;
            s.pending_match_size = (uint16_t)len;
            return kInputExhausted;
        }
        off = (int)(buf >> (32 - wbits));
        buf <<= wbits;
        pos -= wbits;
        s.bit_buffer = buf;
        s.bit_buffer_pos = (uint8_t)pos;
    }
    if (off >= W || off + len > W) return kOob;  // decompressor.c:231-236
    int remaining = len - skip;
    size_t space = io.out_cap - io.out_pos;
    int n;
    if ((size_t)remaining > space) {
        n = (int)space;
        s.skip_bytes = (uint8_t)(skip + n);
        s.token_state = 3;
        s.pending_window_offset = (uint16_t)off;
        s.pending_match_size = (uint16_t)len;
    } else {
        n = remaining;
        s.skip_bytes = 0;
        s.token_state = 0;
    }
    for (int i = 0; i < n; i++) io.out[io.out_pos + i] = win[off + skip + i];
    io.out_pos += (size_t)n;
    if (s.token_state == 0) {  // window is written once, when the token completes; no wrap
        int wp = s.window_pos;
        int room = W - wp;
        dec_window_copy(win, wp, off, len < room ? len : room, W - 1);
        s.window_pos = (uint16_t)wp;
    }
    return s.token_state == 0 ? kOk : kOutputFull;
}

// Main loop of tamp_decompressor_decompress_cb, decompressor.c:423-577 (header already handled).
__device__ inline int dec_run(TbDecState &s, uint8_t *win, DecIo &io) {
    const int wbits = s.window_bits, lbits = s.literal_bits, min_pat = s.min_pattern_size;
    const int W = 1 << wbits, mask = W - 1;
    const bool extended = (s.flags & TB_F_EXTENDED) != 0;
    while (io.in_pos != io.in_size || s.bit_buffer_pos || s.token_state) {
        if (io.out_pos == io.out_cap) return kOutputFull;
        dec_refill(s, io);
        if (s.token_state) {
        extended_dispatch:
            int r = (s.token_state == 1) ? dec_rle(s, win, io) : dec_ext_match(s, win, io);
            if (r == kInputExhausted) {
                int before = s.bit_buffer_pos;
                dec_refill(s, io);
                if (s.bit_buffer_pos == before && io.in_pos == io.in_size) return kInputExhausted;
                continue;
            }
            if (r != kOk) return r;
            continue;
        }
        if (s.bit_buffer_pos == 0) return kInputExhausted;
        if (s.bit_buffer >> 31) {  // literal: 1 | literal bits
            s.last_was_flush = 0;
            if (s.bit_buffer_pos < 1 + lbits) return kInputExhausted;
            uint32_t b = (s.bit_buffer << 1) >> (32 - lbits);
            s.bit_buffer <<= (1 + lbits);
            s.bit_buffer_pos = (uint8_t)(s.bit_buffer_pos - (1 + lbits));
            io.out[io.out_pos++] = (uint8_t)b;
            win[s.window_pos] = (uint8_t)b;
            s.window_pos = (uint16_t)((s.window_pos + 1) & mask);
            continue;
        }
        // token: 0 | huffman(len - min) | offset.  Decode on copies; commit when complete.
        uint32_t buf = s.bit_buffer << 1;
        int pos = s.bit_buffer_pos - 1;
        int sym;
        if (!dec_huffman(buf, pos, 0, sym)) return kInputExhausted;
        if (sym == kSymFlush) {  // drop to the next byte boundary; decompressor.c:501-514
            s.bit_buffer = buf << (pos & 7);
            s.bit_buffer_pos = (uint8_t)(pos & ~7);
            if ((s.flags & TB_F_DICT_RESET) && s.last_was_flush) {
                s.window_pos = 0;
                seed_dictionary_serial(win, W, extended ? lbits : 8);
            }
            s.last_was_flush = 1;
            continue;
        }
        s.last_was_flush = 0;
        if (extended && sym >= kSymRle) {
            s.bit_buffer = buf;
            s.bit_buffer_pos = (uint8_t)pos;
            s.token_state = (uint8_t)(sym - (kSymRle - 1));  // 12 -> 1 (RLE), 13 -> 2 (fresh ext match)
            goto extended_dispatch;
        }
        if (pos < wbits) return kInputExhausted;  // nothing committed: token is re-decoded next call
        int len = sym + min_pat;
        int off = (int)(buf >> (32 - wbits));
        if (off >= W || off + len > W) return kOob;  // decompressor.c:540-544
        int len_left = len - s.skip_bytes;
        int src = off + s.skip_bytes;
        size_t space = io.out_cap - io.out_pos;
        if ((size_t)len_left > space) {  // deliver what fits; the token is decoded again next call
            s.skip_bytes = (uint8_t)(s.skip_bytes + space);
            len_left = (int)space;
        } else {
            s.skip_bytes = 0;
            s.bit_buffer = buf << wbits;
            s.bit_buffer_pos = (uint8_t)(pos - wbits);
        }
        for (int i = 0; i < len_left; i++) io.out[io.out_pos + i] = win[src + i];
        io.out_pos += (size_t)len_left;
        if (s.skip_bytes == 0) {
            int wp = s.window_pos;
            dec_window_copy(win, wp, off, len, mask);
            s.window_pos = (uint16_t)wp;
        }
    }
    return kInputExhausted;
}

}  // namespace tb
