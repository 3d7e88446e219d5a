// General-purpose Tamp compressor state machine for ONE stream, executed by one warp.
//
// Coverage: every configuration the reference accepts (window 8..15, literal 5..8, v1 and extended
// format, custom dictionary, dictionary_reset header, lazy matching) and every resumable entry point
// (poll / compress / flush with arbitrary output room).  It is the engine behind the per-call C API
// and the batch path for configurations without a specialised kernel.
//
// Semantics restated from tamp/_c_src/tamp/compressor.c @ 48880ad (line refs on each function).
// Work split: all 32 lanes execute the (warp-uniform) token state machine redundantly; the two O(W)
// searches — find_best_match and find_extended_match — are spread over the lanes and reduced with
// __reduce_max_sync.  Window and input ring live in shared memory; only lane 0 stores to them.
#pragma once
#include "../tb_wire.h"
#include "tb_device_common.cuh"

namespace tb {

struct CompCtx {
    uint32_t bitbuf;
    int bitpos;
    int wpos;
    int qn, qpos;
    int min_pat, wbits, lbits, flags;
    int rle, ext_n, ext_pos, last_flush, cache_idx, cache_len;
    int W, mask;
    uint8_t *win;   // shared memory, W bytes, 4-byte aligned
    uint8_t *ring;  // shared memory, 16 bytes
    uint8_t *out;   // global
    size_t out_room;
    size_t written;
};

__device__ __forceinline__ uint32_t ring_at(const CompCtx &c, int i) { return c.ring[(c.qpos + i) & 15]; }

__device__ __forceinline__ void ring_consume(CompCtx &c, int n) {
    c.qpos = (c.qpos + n) & 15;
    c.qn -= n;
}

// write_to_bit_buffer, compressor.c:49-52
__device__ __forceinline__ void put_bits(CompCtx &c, uint32_t bits, int n) {
    c.bitpos += n;
    c.bitbuf |= bits << (32 - c.bitpos);
}

// partial_flush, compressor.c:65-75
__device__ __forceinline__ int drain_bytes(CompCtx &c) {
    while (c.bitpos >= 8 && c.out_room) {
        if (lane_id() == 0) *c.out = (uint8_t)(c.bitbuf >> 24);
        c.out++;
        c.out_room--;
        c.written++;
        c.bitpos -= 8;
        c.bitbuf <<= 8;
    }
    return c.bitpos >= 8 ? kOutputFull : kOk;
}

// write_extended_huffman, compressor.c:257-263
__device__ __forceinline__ void put_exthuff(CompCtx &c, int v, int t) {
    int i = v >> t;
    put_bits(c, ((uint32_t)kHuff.code[i] << t) | (uint32_t)(v & ((1 << t) - 1)), kHuff.bits[i] - 1 + t);
}

__device__ __forceinline__ uint32_t last_window_byte(const CompCtx &c) { return c.win[(c.wpos - 1) & c.mask]; }

// Append n bytes taken from the ring (starting at ring offset 0) to the window, wrapping.
__device__ __forceinline__ void window_push_from_ring(CompCtx &c, int n) {
    int l = lane_id();
    uint32_t b = ring_at(c, l & 15);
    __syncwarp();
    if (l < n) c.win[(c.wpos + l) & c.mask] = (uint8_t)b;
    __syncwarp();
    c.wpos = (c.wpos + n) & c.mask;
}

__device__ __forceinline__ void window_push_byte(CompCtx &c, uint32_t b) {
    if (lane_id() == 0) c.win[c.wpos] = (uint8_t)b;
    __syncwarp();
    c.wpos = (c.wpos + 1) & c.mask;
}

// find_best_match: compressor_find_match_desktop.c:82-167 (same contract as compressor.c:113-172 and
// fuzz/esp32_host/differential.cpp:50-67): longest common prefix of ring[o..o+L) with window[idx..],
// idx in [0, W-2], never reading past W-1, >= 2 bytes, lowest idx among the longest.
// Each lane filters 4 start positions per step with a SWAR test on the first two bytes, extends the
// survivors byte-wise, and the warp takes max over key = len << 16 | (0xFFFF - idx).
__device__ inline void find_best_match(const CompCtx &c, int o, int avail, int &idx, int &len) {
    len = 0;
    idx = 0;
    if (avail < c.min_pat) return;
    int cap = (c.flags & TB_F_EXTENDED) ? c.min_pat + 11 + kExtExtraMax : c.min_pat + 13;
    int L = avail < cap ? avail : cap;
    if (L < 2) return;
    const uint32_t p0 = ring_at(c, o) * 0x01010101u, p1 = ring_at(c, o + 1) * 0x01010101u;
    const uint32_t *w32 = reinterpret_cast<const uint32_t *>(c.win);
    const int nwords = c.W >> 2;
    uint32_t best = 0;
    for (int base = 0; base < nwords; base += 32) {
        int wi = base + lane_id();
        if (wi < nwords) {
            uint32_t w0 = w32[wi];
            uint32_t w1 = (wi + 1 < nwords) ? w32[wi + 1] : 0u;
            uint32_t z = (w0 ^ p0) | (__funnelshift_r(w0, w1, 8) ^ p1);  // byte b == 0 <=> bigram match at 4*wi+b
            if ((z - 0x01010101u) & ~z & 0x80808080u) {
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    int pos = wi * 4 + b;
                    if (((z >> (8 * b)) & 0xFFu) == 0 && pos + 1 < c.W) {
                        int l = 2;
                        while (l < L && pos + l < c.W && c.win[pos + l] == ring_at(c, o + l)) l++;
                        uint32_t key = ((uint32_t)l << 16) | (uint32_t)(0xFFFF - pos);
                        best = key > best ? key : best;
                    }
                }
            }
        }
    }
    best = __reduce_max_sync(0xffffffffu, best);
    len = (int)(best >> 16);
    idx = 0xFFFF - (int)(best & 0xFFFFu);
}

// find_extended_match, compressor.c:297-333: among cand >= cur_pos whose first cur_n bytes equal
// window[cur_pos..+cur_n) and whose next byte equals ring[0], the longest extension (<= cur_n + ring
// fill, <= min_pat + 131, never past W-1); lowest cand among the longest.
__device__ inline void find_extended_match(const CompCtx &c, int cur_pos, int cur_n, int &new_pos, int &new_n) {
    int maxp = cur_n + c.qn;
    if (maxp > c.min_pat + 11 + kExtExtraMax) maxp = c.min_pat + 11 + kExtExtraMax;
    const uint32_t next = ring_at(c, 0);
    uint32_t best = 0;
    for (int base = cur_pos; base + cur_n + 1 <= c.W; base += 32) {
        int cand = base + lane_id();
        if (cand + cur_n + 1 <= c.W && c.win[cand + cur_n] == next) {
            int i = 0;
            while (i < cur_n && c.win[cand + i] == c.win[cur_pos + i]) i++;
            if (i == cur_n) {
                int lim = maxp < c.W - cand ? maxp : c.W - cand;
                int l = cur_n + 1;
                while (l < lim && c.win[cand + l] == ring_at(c, l - cur_n)) l++;
                uint32_t key = ((uint32_t)l << 16) | (uint32_t)(0xFFFF - cand);
                best = key > best ? key : best;
            }
        }
    }
    best = __reduce_max_sync(0xffffffffu, best);
    new_n = (int)(best >> 16);
    new_pos = 0xFFFF - (int)(best & 0xFFFFu);
}

// write_rle_token, compressor.c:342-359
__device__ inline void emit_rle(CompCtx &c, int count) {
    uint32_t sym = last_window_byte(c);
    put_bits(c, kHuff.code[kSymRle], kHuff.bits[kSymRle]);
    put_exthuff(c, count - 2, 4);
    int room = c.W - c.wpos;
    int n = count < kRleWindowMax ? count : kRleWindowMax;
    n = n < room ? n : room;
    __syncwarp();
    if (lane_id() < n) c.win[c.wpos + lane_id()] = (uint8_t)sym;  // never wraps: n <= room
    __syncwarp();
    c.wpos = (c.wpos + n) & c.mask;
}

// write_extended_match_token, compressor.c:377-415
__device__ inline int emit_ext_match(CompCtx &c) {
    if (c.out_room < (size_t)kExtMinOutput) return kOutputFull;
    put_bits(c, kHuff.code[kSymExt], kHuff.bits[kSymExt]);
    put_exthuff(c, c.ext_n - c.min_pat - 12, 3);
    int r = drain_bytes(c);
    if (r != kOk) return r;
    put_bits(c, (uint32_t)c.ext_pos, c.wbits);
    r = drain_bytes(c);
    if (r != kOk) return r;
    // window[wpos..] <- window[ext_pos..ext_pos+n), n = min(count, bytes to the buffer end): no wrap.
    // Source and destination may overlap; tamp_window_copy's direction rule equals "read all, then
    // write" (common.c:58-86), done here in register-sized rounds of 32 bytes, highest chunk first
    // when the destination sits above the source.
    int room = c.W - c.wpos;
    int n = c.ext_n < room ? c.ext_n : room;
    bool backwards = c.wpos > c.ext_pos;
    int nchunks = (n + 31) >> 5;
    for (int k = 0; k < nchunks; k++) {
        int chunk = backwards ? nchunks - 1 - k : k;
        int i = chunk * 32 + lane_id();
        uint32_t b = 0;
        if (i < n) b = c.win[c.ext_pos + i];
        __syncwarp();
        if (i < n) c.win[c.wpos + i] = (uint8_t)b;
        __syncwarp();
    }
    c.wpos = (c.wpos + n) & c.mask;
    c.ext_n = 0;
    return kOk;
}

__device__ inline void emit_literal_of_last(CompCtx &c) {  // compressor.c:512-523 and :748-756
    uint32_t b = last_window_byte(c);
    put_bits(c, (1u << c.lbits) | b, c.lbits + 1);
    window_push_byte(c, b);
}

// poll_extended_handling, compressor.c:437-525
__device__ inline int extended_handling(CompCtx &c, int &m_idx, int &m_len) {
    if (c.ext_n) {
        const int cap = c.min_pat + 11 + kExtExtraMax;
        while (c.qn > 0) {
            if (c.ext_pos + c.ext_n >= c.W || c.ext_n >= cap) return emit_ext_match(c);
            int npos, nlen;
            find_extended_match(c, c.ext_pos, c.ext_n, npos, nlen);
            if (nlen > c.ext_n) {
                ring_consume(c, nlen - c.ext_n);
                c.ext_pos = npos;
                c.ext_n = nlen;
                continue;
            }
            return emit_ext_match(c);
        }
        return kOk;
    }
    const uint32_t last = last_window_byte(c);
    int avail = 0;
    while (avail < c.qn && c.rle + avail < kRleMax && ring_at(c, avail) == last) avail++;
    const int total = c.rle + avail;
    const bool ended = (avail < c.qn) || (total >= kRleMax);
    if (!ended && total > 0) {  // whole ring is run: keep counting, emit nothing yet
        c.rle = total;
        ring_consume(c, avail);
        return kOk;
    }
    if (total >= 2) {
        if (total == avail && total <= 6) {  // short run seen entirely in this poll may lose to a match
            int idx, len;
            find_best_match(c, 0, c.qn, idx, len);
            if (len > total) {
                c.rle = 0;
                m_idx = idx;
                m_len = len;
                return kPollContinue;
            }
        }
        ring_consume(c, avail);
        emit_rle(c, total);
        c.rle = 0;
        return kOk;
    }
    if (c.rle == 1) {  // lone byte swallowed by an earlier poll: re-emit as literal
        emit_literal_of_last(c);
        c.rle = 0;
        return kOk;
    }
    return kPollContinue;
}

// tamp_compressor_poll, compressor.c:532-660
__device__ inline int poll(CompCtx &c) {
    if (c.qn == 0) return kOk;
    c.last_flush = 0;
    if (drain_bytes(c) != kOk) return kOutputFull;
    if (c.out_room == 0) return kOutputFull;

    int idx = 0, len = 0;
    if (c.flags & TB_F_EXTENDED) {
        int r = extended_handling(c, idx, len);
        if (r != kPollContinue) {
            c.cache_idx = -1;
            return r;
        }
    }
    if (c.flags & TB_F_LAZY) {  // compressor.c:576-616
        if (c.cache_idx >= 0) {
            idx = c.cache_idx;
            len = c.cache_len;
            c.cache_idx = -1;
        } else if (len == 0) {
            find_best_match(c, 0, c.qn, idx, len);
        }
        if (len >= c.min_pat && len <= 8 && c.qn > len + 2) {
            int nidx, nlen;
            find_best_match(c, 1, c.qn - 1, nidx, nlen);
            bool clear = c.wpos < nidx || c.wpos >= nidx + nlen;  // the literal must not clobber the match
            if (nlen > len && clear) {
                c.cache_idx = nidx;
                c.cache_len = nlen;
                len = 0;
            } else {
                c.cache_idx = -1;
            }
        } else {
            c.cache_idx = -1;
        }
    } else if (len == 0) {
        find_best_match(c, 0, c.qn, idx, len);
    }

    int n;
    if (len < c.min_pat) {
        uint32_t ch = ring_at(c, 0);
        if (ch >> c.lbits) return kExcessBits;
        put_bits(c, (1u << c.lbits) | ch, c.lbits + 1);
        n = 1;
    } else if ((c.flags & TB_F_EXTENDED) && len > c.min_pat + 11) {
        c.ext_n = len;  // start an extended match; the token is written once it stops growing
        c.ext_pos = idx;
        ring_consume(c, len);
        return kOk;
    } else {
        int h = len - c.min_pat;
        put_bits(c, ((uint32_t)kHuff.code[h] << c.wbits) | (uint32_t)idx, kHuff.bits[h] + c.wbits);
        n = len;
    }
    window_push_from_ring(c, n);
    ring_consume(c, n);
    return kOk;
}

// tamp_compressor_compress_cb, compressor.c:681-722
__device__ inline int compress_loop(CompCtx &c, const uint8_t *in, size_t in_size, size_t &consumed) {
    consumed = 0;
    while (consumed < in_size && c.out_room > 0) {
        size_t room = (size_t)(16 - c.qn);
        size_t left = in_size - consumed;
        int n = (int)(left < room ? left : room);
        int l = lane_id();
        __syncwarp();  // ring reads of the previous poll precede the refill
        if (l < n) c.ring[(c.qpos + c.qn + l) & 15] = in[consumed + l];
        __syncwarp();
        c.qn += n;
        consumed += (size_t)n;
        if (c.qn == 16) {
            int r = poll(c);
            if (r != kOk) return r;
        }
    }
    return kOk;
}

// tamp_compressor_flush, compressor.c:728-810
__device__ inline int flush(CompCtx &c, bool write_token) {
    for (;;) {
        int r = drain_bytes(c);
        if (r != kOk) return r;
        if (c.qn) {
            r = poll(c);
        } else if ((c.flags & TB_F_EXTENDED) && c.rle >= 1) {
            if (c.rle == 1)
                emit_literal_of_last(c);
            else
                emit_rle(c, c.rle);
            c.rle = 0;
        } else if ((c.flags & TB_F_EXTENDED) && c.ext_n) {
            r = emit_ext_match(c);
        } else {
            break;
        }
        if (r != kOk) return r;
    }
    if (write_token && !c.last_flush && (c.bitpos || (c.flags & TB_F_DICT_RESET))) {
        if (c.out_room < 2) return kOutputFull;
        put_bits(c, kHuff.code[kSymFlush], kHuff.bits[kSymFlush]);
        c.last_flush = 1;
    }
    int r = drain_bytes(c);
    if (c.bitpos) {
        if (c.out_room == 0) return kOutputFull;
        if (lane_id() == 0) *c.out = (uint8_t)(c.bitbuf >> 24);
        c.written++;
        c.bitpos = 0;
        c.bitbuf = 0;
    }
    return r;
}

__device__ inline int run_op(CompCtx &c, int op, const uint8_t *in, size_t in_size, bool write_token,
                             size_t &consumed) {
    consumed = 0;
    switch (op) {
        case TB_OP_POLL:
            return poll(c);
        case TB_OP_COMPRESS:
            return compress_loop(c, in, in_size, consumed);
        case TB_OP_FLUSH:
            return flush(c, write_token);
        default: {  // TB_OP_COMPRESS_AND_FLUSH, compressor.c:815-845
            int r = compress_loop(c, in, in_size, consumed);
            if (r != kOk) return r;
            return flush(c, write_token);
        }
    }
}

__device__ inline void ctx_from_state(CompCtx &c, const TbCompState &s) {
    c.bitbuf = s.bit_buffer;
    c.bitpos = s.bit_buffer_pos;
    c.wpos = s.window_pos;
    c.qn = s.input_size;
    c.qpos = s.input_pos;
    c.min_pat = s.min_pattern_size;
    c.wbits = s.window_bits;
    c.lbits = s.literal_bits;
    c.flags = s.flags;
    c.rle = s.rle_count;
    c.ext_n = s.ext_count;
    c.ext_pos = s.ext_pos;
    c.last_flush = s.last_was_flush;
    c.cache_idx = s.cached_index;
    c.cache_len = s.cached_size;
    c.W = 1 << s.window_bits;
    c.mask = c.W - 1;
    c.written = 0;
}

__device__ inline void ctx_to_state(const CompCtx &c, TbCompState &s) {
    s.bit_buffer = c.bitbuf;
    s.bit_buffer_pos = (uint8_t)c.bitpos;
    s.window_pos = (uint16_t)c.wpos;
    s.input_size = (uint8_t)c.qn;
    s.input_pos = (uint8_t)c.qpos;
    s.rle_count = (uint8_t)c.rle;
    s.ext_count = (uint8_t)c.ext_n;
    s.ext_pos = (uint16_t)c.ext_pos;
    s.last_was_flush = (uint8_t)c.last_flush;
    s.cached_index = (int16_t)c.cache_idx;
    s.cached_size = (uint8_t)c.cache_len;
}

// Fresh-stream state, equal to what tamp_compressor_init leaves behind (compressor.c:191-245).
__device__ inline void ctx_init(CompCtx &c, int window, int literal, int flags) {
    c.wbits = window;
    c.lbits = literal;
    c.flags = flags;
    c.W = 1 << window;
    c.mask = c.W - 1;
    c.min_pat = min_pattern_size(window, literal);
    c.wpos = 0;
    c.qn = 0;
    c.qpos = 0;
    c.rle = 0;
    c.ext_n = 0;
    c.ext_pos = 0;
    c.last_flush = 0;
    c.cache_idx = -1;
    c.cache_len = 0;
    c.written = 0;
    uint32_t header = ((uint32_t)(window - 8) << 5) | ((uint32_t)(literal - 5) << 3) |
                      ((flags & TB_F_CUSTOM_DICT) ? 4u : 0u) | ((flags & TB_F_EXTENDED) ? 2u : 0u) |
                      ((flags & TB_F_DICT_RESET) ? 1u : 0u);
    c.bitbuf = header << 24;
    c.bitpos = (flags & TB_F_DICT_RESET) ? 16 : 8;
    if (flags & TB_F_APPEND) {  // compressor.c:227-234: a FLUSH padded to 16 bits instead of the header
        c.bitbuf = kAppendStart;
        c.bitpos = 16;
        c.last_flush = 1;
    }
}

}  // namespace tb
