// Specialised batch decompressor: one LANE per stream, 32 streams per warp, windows <= 1 KiB in shared memory.
//
// The reference decoder (tamp/_c_src/tamp/decompressor.c:371-578) is a strictly serial bit-stream walk:
// refill a 32-bit buffer byte by byte, decode one literal or Huffman(len)+offset token, copy len bytes
// window -> output and window -> window.  Nothing inside one stream is parallel, so the B200 mapping
// puts the parallelism ACROSS streams: every lane of a warp walks its own frame.
//
//  * Windows live in shared memory, word-interleaved across the 32 lanes of a warp
//    (word k of lane l at k*32 + l), so any per-lane access pattern is bank-conflict free.
//  * The bit reader is a 64-bit MSb-aligned register refilled with aligned 32-bit big-endian loads.
//  * In the v1 format (and between extended tokens) the bytes written to the window ARE the output, so
//    output is not written token by token: whenever 16 more output bytes exist, a lane re-reads them from
//    its window (4 conflict-free LDS) and issues one 128-bit store.  RLE / extended-match tokens
//    (decompressor.c:114-273), whose window writes are truncated, take a byte-wise side path.
//  * Per-stream header parsing, OOB / invalid-header statuses and the OUTPUT_FULL / INPUT_EXHAUSTED
//    rules restate tamp_decompressor_read_header (:276-297), populate_from_conf (:304-329) and the loop
//    conditions of decompress_cb (:433-463) for a single whole-frame call.
#include "../tb_wire.h"
#include "tb_cuda.h"
#include "tb_device_common.cuh"

namespace tb {

namespace {

// One CTA per SM holding as many warps (32 streams, 32 windows each) as 224 KiB of shared memory take: 7 for
// 1 KiB windows.  (One-warp CTAs would stop at 6: every CTA costs 1 KiB of reserved shared memory.)
template <int WMAXBITS>
constexpr int kWarpsPerCtaDec = (224 * 1024) / (32 << WMAXBITS);

struct FastDecArgs {
    BatchArgs b;
    const uint8_t *seed;    // 3 x 32 KiB seeded dictionaries (literal classes 5, 6, 7/8)
    const uint8_t *custom;  // caller dictionary or nullptr
    int window_bits_max;
    int aligned_io;  // out rows 16-byte aligned (128-bit stores allowed)
    const uint8_t *lut;  // 128-entry Huffman decode LUT in global memory (L1-resident)
    int only_deferred;   // process just the streams whose out_sizes entry is kDeferred (left by split_decompress.cu)
};

// Shared-memory accessors with explicit .shared addressing (keeps generic->shared conversions out of the loop).
#ifndef TB_EMU
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ uint32_t lds8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v)); }
#else  // tests/emu: the kernel stepped on the CPU (test infrastructure; see tests/emu/cuda_emu.h)
inline uint32_t lds32(uint32_t a) { return *reinterpret_cast<const uint32_t *>(emu_shared_ptr(a)); }
inline void sts32(uint32_t a, uint32_t v) { *reinterpret_cast<uint32_t *>(emu_shared_ptr(a)) = v; }
inline uint32_t lds8(uint32_t a) { return *reinterpret_cast<const uint8_t *>(emu_shared_ptr(a)); }
inline void sts8(uint32_t a, uint32_t v) { *reinterpret_cast<uint8_t *>(emu_shared_ptr(a)) = (uint8_t)v; }
#endif

// Word-interleaved window of one lane: word k of lane l lives at warp_base + (k*32 + l)*4, so every
// per-lane access pattern is bank-conflict free.
struct LaneWindow {
    uint32_t sbase;  // .shared address of this lane's word 0
    __device__ __forceinline__ uint32_t baddr(int off) const { return sbase + (uint32_t)(((off >> 2) << 7) | (off & 3)); }
    __device__ __forceinline__ uint32_t ld(int off) const { return lds8(baddr(off)); }
    __device__ __forceinline__ void st(int off, uint32_t b) const { sts8(baddr(off), b); }
    __device__ __forceinline__ uint32_t ld_word(int widx) const { return lds32(sbase + ((uint32_t)widx << 7)); }
    __device__ __forceinline__ void st_word(int widx, uint32_t v) const { sts32(sbase + ((uint32_t)widx << 7), v); }
};

// byte mask with the k lowest bytes set (k <= 0 -> 0, k >= 4 -> all)
__device__ __forceinline__ uint32_t low_bytes(int k) {
    return k <= 0 ? 0u : (k >= 4 ? 0xffffffffu : ((1u << (8 * k)) - 1u));
}

// Per-lane decoder state.
struct LaneDec {
    // frame
    const uint8_t *in;
    uint32_t n, ip;
    uint32_t next_word;  // prefetched aligned input word at in + ip (valid when has_next)
    bool has_next;
    // bit reader: MSb-aligned unread bits
    uint64_t bb;
    int nb;
    // output row
    uint8_t *out;
    uint32_t cap, opos, flushed;
    // window
    LaneWindow win;
    int wpos, mask, wmask;
    // configuration from the frame header
    int wbits, lbits, min_pat, max_plain_sym;
    bool extended, dict_reset, last_flush;
    // result
    int status;
    bool active;
    // the token decoded for the next copy phase
    int t_len;      // bytes to append (0: nothing)
    int t_src;      // >= 0: window offset to copy from; -1: literal bytes only
    uint32_t t_lit; // literal bytes, first one in the low byte
    int t_mlen;     // t_src >= 0: how many of the t_len bytes come from the window (the rest are t_lit)
};

__device__ __forceinline__ bool word_loadable(const LaneDec &d) {
    return d.ip + 4 <= d.n && ((reinterpret_cast<uintptr_t>(d.in) + d.ip) & 3) == 0;
}

// Top up the bit buffer (decompressor.c:357-365, done a word at a time).  The aligned word for the NEXT
// refill is requested now, so its global-memory latency is hidden behind a couple of tokens.
__device__ __forceinline__ void refill(LaneDec &d) {
    if (d.nb > 32) return;
    if (!d.has_next) {
        while (d.nb <= 32 && d.ip < d.n && !word_loadable(d)) {  // ragged head / tail of the frame
            d.bb |= (uint64_t)d.in[d.ip] << (56 - d.nb);
            d.nb += 8;
            d.ip += 1;
        }
        if (d.nb <= 32 && word_loadable(d)) {
            d.next_word = *reinterpret_cast<const uint32_t *>(d.in + d.ip);
            d.has_next = true;
        }
    }
    if (d.has_next && d.nb <= 32) {
        d.bb |= (uint64_t)__byte_perm(d.next_word, 0, 0x0123) << (32 - d.nb);
        d.nb += 32;
        d.ip += 4;
        d.has_next = word_loadable(d);
        if (d.has_next) d.next_word = *reinterpret_cast<const uint32_t *>(d.in + d.ip);
    }
}

__device__ __forceinline__ void deliver_pending_bytes(LaneDec &d) {
    for (; d.flushed < d.opos; d.flushed++)
        d.out[d.flushed] = (uint8_t)d.win.ld((d.wpos - (int)(d.opos - d.flushed)) & d.mask);
}

// Everything that is not a literal or a complete, in-bounds plain token that fits the output row:
// end of frame, errors, FLUSH, RLE and extended-match tokens, the partial token at the end of the row.
// Runs rarely; byte-wise.  Semantics: decompressor.c:433-577 (+ :114-273 for the extended tokens).
__device__ __forceinline__ void decode_slow(LaneDec &d, const uint8_t *lut, const uint8_t *seed, uint32_t top, int sym,
                                         int used) {
    if (d.nb == 0) {
        d.active = false;  // INPUT_EXHAUSTED: frame fully consumed
        return;
    }
    if (d.opos == d.cap) {
        d.status = kOutputFull;
        d.active = false;
        return;
    }
    if (top >> 31) {  // literal without enough bits
        d.last_flush = false;
        d.active = false;
        return;
    }
    if (d.nb < used) {  // Huffman code incomplete
        d.active = false;
        return;
    }
    const int W = d.mask + 1;
    if (sym <= d.max_plain_sym) {
        d.last_flush = false;
        if (d.nb < used + d.wbits) {  // offset not there: nothing is consumed, the frame ends here
            d.active = false;
            return;
        }
        const int tlen = sym + d.min_pat;
        const int off = (int)((d.bb << used) >> (64 - d.wbits));
        if (off + tlen > W) {  // also covers off >= W (decompressor.c:540-544)
            d.status = kOob;
            d.active = false;
            return;
        }
        // partial token at the end of the output row: bytes only, no window update (:547-562)
        deliver_pending_bytes(d);
        const uint32_t space = d.cap - d.opos;
        for (uint32_t i = 0; i < space; i++) d.out[d.opos + i] = (uint8_t)d.win.ld(off + (int)i);
        d.opos += space;
        d.flushed = d.opos;
        d.status = kOutputFull;
        d.active = false;
        return;
    }
    if (sym == kSymFlush) {  // drop to the next byte boundary of the frame (:501-514)
        const int drop = used + ((d.nb - used) & 7);
        d.bb <<= drop;
        d.nb -= drop;
        if (d.dict_reset && d.last_flush) {  // double FLUSH: deliver what the window still owes, re-seed it
            deliver_pending_bytes(d);
            const int seed_lit = d.extended ? d.lbits : 8;
            const uint32_t *sd =
                reinterpret_cast<const uint32_t *>(seed + (seed_lit <= 5 ? 0 : seed_lit <= 6 ? 1 : 2) * 32768);
            for (int k = 0; k <= d.wmask; k++) d.win.st_word(k, __ldg(sd + k));
            d.wpos = 0;
        }
        d.last_flush = true;
        return;
    }
    // ---- RLE / extended match: the symbol is consumed (:521-526), then the count / size code --------
    d.last_flush = false;
    uint64_t b2 = d.bb << used;
    int n2 = d.nb - used;
    const int trailing = sym == kSymRle ? 4 : 3;
    int value = -1;
    if (n2 >= 1 + trailing) {
        const uint32_t t2 = (uint32_t)(b2 >> 32);
        int s2 = -1, u2 = 0;
        if ((t2 >> 31) == 0) {
            s2 = 0;
            u2 = 1;
        } else {
            const uint32_t e2 = lut[(t2 << 1) >> 25];
            const int extra = (int)(e2 >> 4);
            if (n2 >= 1 + extra + trailing) {
                s2 = (int)(e2 & 15u);
                u2 = 1 + extra;
            }
        }
        if (s2 >= 0) {
            const uint32_t tr = (uint32_t)((b2 << u2) >> (64 - trailing));
            value = (s2 << trailing) + (int)tr;
            b2 <<= u2 + trailing;
            n2 -= u2 + trailing;
        }
    }
    if (value < 0) {
        d.active = false;
        return;
    }
    int xlen, off = 0, nwin = 0;
    if (sym == kSymRle) {
        xlen = value + 2;
        nwin = xlen < kRleWindowMax ? xlen : kRleWindowMax;
        if (nwin > W - d.wpos) nwin = W - d.wpos;
    } else {
        xlen = value + d.min_pat + 12;
        if (n2 < d.wbits) {  // one more top-up for the offset (the reference parks the size, :216-223)
            d.has_next = false;
            while (n2 <= 32 && d.ip < d.n) {
                b2 |= (uint64_t)d.in[d.ip] << (56 - n2);
                n2 += 8;
                d.ip += 1;
            }
        }
        if (n2 < d.wbits) {
            d.active = false;
            return;
        }
        off = (int)(b2 >> (64 - d.wbits));
        b2 <<= d.wbits;
        n2 -= d.wbits;
        if (off >= W || off + xlen > W) {  // :231-236
            d.status = kOob;
            d.active = false;
            return;
        }
        nwin = xlen < W - d.wpos ? xlen : W - d.wpos;
    }
    d.bb = b2;
    d.nb = n2;
    // deliver pending window-backed output, then this token's bytes directly
    deliver_pending_bytes(d);
    const uint32_t space = d.cap - d.opos;
    const uint32_t nout = (uint32_t)xlen < space ? (uint32_t)xlen : space;
    const uint32_t rsym = d.win.ld((d.wpos - 1) & d.mask);
    for (uint32_t i = 0; i < nout; i++) d.out[d.opos + i] = (uint8_t)(sym == kSymRle ? rsym : d.win.ld(off + (int)i));
    d.opos += nout;
    d.flushed = d.opos;
    if (nout < (uint32_t)xlen) {
        d.status = kOutputFull;
        d.active = false;
    } else if (sym == kSymRle) {
        for (int i = 0; i < nwin; i++) d.win.st(d.wpos + i, rsym);
        d.wpos = (d.wpos + nwin) & d.mask;
    } else {  // window[wpos..] <- window[off..off+nwin): no wrap, snapshot semantics
        if (d.wpos > off) {
            for (int i = nwin - 1; i >= 0; i--) d.win.st(d.wpos + i, d.win.ld(off + i));
        } else {
            for (int i = 0; i < nwin; i++) d.win.st(d.wpos + i, d.win.ld(off + i));
        }
        d.wpos = (d.wpos + nwin) & d.mask;
    }
}

// Extended format: the common small cases of its two extra tokens go through the normal copy phase — a run of up to 4
// bytes is just that many literal bytes (the byte before the write position, decompressor.c:114-174), an extended match
// of up to 16 bytes a plain window copy (:187-273) — as long as the window takes every byte (no clipping at its end),
// the source is in bounds and the row has room.  Everything else (longer ones, missing bits, errors) is decode_slow's.
__device__ __forceinline__ bool decode_small_extended(LaneDec &d, int sym, int used) {
    const uint64_t b2 = d.bb << used;  // the bits behind the symbol: count / size code, then its raw bits
    if (b2 >> 63) return false;        // only code 0 (one '0' bit) keeps the value this small
    const int W = d.mask + 1;
    if (sym == kSymRle) {
        const int need = used + 1 + 4;
        const int count = (int)((b2 << 1) >> 60) + 2;
        if (d.nb < need || count > 4 || d.wpos + count > W || (uint32_t)count > d.cap - d.opos) return false;
        d.t_len = count;
        d.t_lit = d.win.ld((d.wpos - 1) & d.mask) * 0x01010101u;
        d.bb <<= need;
        d.nb -= need;
    } else {
        const int need = used + 1 + 3 + d.wbits;
        const int xlen = (int)((b2 << 1) >> 61) + d.min_pat + 12;
        const int off = (int)((b2 << 4) >> (64 - d.wbits));
        if (d.nb < need || xlen > 16 || off + xlen > W || d.wpos + xlen > W || (uint32_t)xlen > d.cap - d.opos) return false;
        d.t_len = xlen;
        d.t_mlen = xlen;
        d.t_src = off;
        d.t_lit = 0;
        d.bb <<= need;
        d.nb -= need;
    }
    d.last_flush = false;
    return true;
}

// Decode the next item of the frame into (t_len, t_src, t_lit).  The common cases — a literal or a complete
// in-bounds plain token that fits the row — are straight-line, register-only code (no window access), so
// they can overlap with the previous token's copy; everything else goes to decode_slow.
__device__ __forceinline__ void decode_next(LaneDec &d, const uint8_t *lut, const uint8_t *seed) {
    d.t_len = 0;
    d.t_src = -1;
    d.t_mlen = 0;
    if (!d.active) return;
    refill(d);
    const uint32_t top = (uint32_t)(d.bb >> 32);
    const bool is_lit = (top >> 31) != 0;
    const uint32_t e = __ldg(lut + ((top << 2) >> 25));
    const bool long_code = ((top >> 30) & 1u) != 0;
    const int sym = long_code ? (int)(e & 15u) : 0;
    const int used = long_code ? 2 + (int)(e >> 4) : 2;
    const int need = is_lit ? 1 + d.lbits : used + d.wbits;
    const int tlen = is_lit ? 1 : sym + d.min_pat;
    const int off = (int)((d.bb << used) >> (64 - d.wbits));
    const bool fast = d.nb >= need && (is_lit || (sym <= d.max_plain_sym && off + tlen <= d.mask + 1)) &&
                      (uint32_t)tlen <= d.cap - d.opos;
    if (fast) {
        d.last_flush = false;
        d.t_len = tlen;
        d.t_src = is_lit ? -1 : off;
        d.t_lit = (top << 1) >> (32 - d.lbits);
        d.bb <<= need;
        d.nb -= need;
        // Literals right behind this item ride along: their bytes land in the same copy (the lanes of a warp
        // advance in lock-step, so fewer, fatter steps are what counts).  Up to three literals in a row, or up
        // to two behind a match as long as the step stays within 16 bytes.
        const int lneed = 1 + d.lbits;
        if (is_lit) {
#pragma unroll
            for (int extra = 1; extra <= 2; extra++) {
                const uint32_t t2 = (uint32_t)(d.bb >> 32);
                if ((t2 >> 31) && d.nb >= lneed && d.t_len == extra && (uint32_t)(extra + 1) <= d.cap - d.opos) {
                    d.t_lit |= ((t2 << 1) >> (32 - d.lbits)) << (8 * extra);
                    d.t_len = extra + 1;
                    d.bb <<= lneed;
                    d.nb -= lneed;
                }
            }
        } else {
            d.t_mlen = tlen;
            d.t_lit = 0;
#pragma unroll
            for (int extra = 0; extra < 2; extra++) {
                const uint32_t t2 = (uint32_t)(d.bb >> 32);
                if ((t2 >> 31) && d.nb >= lneed && d.t_len == tlen + extra && d.t_len < 16 &&
                    (uint32_t)(d.t_len + 1) <= d.cap - d.opos) {
                    d.t_lit |= ((t2 << 1) >> (32 - d.lbits)) << (8 * extra);
                    d.t_len += 1;
                    d.bb <<= lneed;
                    d.nb -= lneed;
                }
            }
        }
    } else if (!is_lit && d.extended && (sym == kSymRle || sym == kSymExt) && decode_small_extended(d, sym, used)) {
        // handled: the copy phase does the rest
    } else {
        decode_slow(d, lut, seed, top, sym, used);
    }
}

template <int WMAXBITS>
__global__ void __launch_bounds__(kWarpsPerCtaDec<WMAXBITS> * 32) k_fast_decompress(FastDecArgs a) {
    constexpr int WMAX = 1 << WMAXBITS;
#ifndef TB_EMU
    extern __shared__ __align__(128) uint8_t smem[];
#else
    uint8_t *smem = emu::g_smem;
#endif

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    LaneDec d;
    d.win.sbase = (uint32_t)__cvta_generic_to_shared(smem + (size_t)warp * (32 * WMAX) + lane * 4);
    const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t first = (uint64_t)blockIdx.x * blockDim.x + warp * 32;  // first stream of this warp's batch

    for (uint64_t batch = first; batch < a.b.n_streams; batch += nthreads) {
        const uint64_t stream = batch + lane;
        const bool mine = stream < a.b.n_streams && (!a.only_deferred || a.b.out_sizes[stream] == kDeferred);
        if (!__any_sync(0xffffffffu, mine)) continue;
        d.active = mine;
        d.in = nullptr;
        d.n = 0;
        d.ip = 0;
        d.out = nullptr;
        d.cap = 0;
        d.status = kInputExhausted;
        d.wbits = 10;
        d.lbits = 8;
        d.min_pat = 2;
        d.mask = WMAX - 1;
        d.extended = false;
        d.dict_reset = false;
        d.last_flush = false;
        const uint8_t *dict_src = a.seed + 2 * 32768;

        if (d.active) {
            d.in = a.b.in + (a.b.in_offsets ? a.b.in_offsets[stream] : stream * a.b.in_stride);
            d.n = a.b.in_sizes ? a.b.in_sizes[stream] : (uint32_t)a.b.in_stride;
            d.out = a.b.out + stream * a.b.out_stride;
            d.cap = (uint32_t)a.b.out_stride;
            // header (decompressor.c:276-329)
            if (d.n == 0) {
                d.active = false;
            } else {
                const uint32_t hs = frame_start(a.b.seg_header, stream, d.in, d.n), h = hs & 0xFFu;
                const uint32_t hdr = 1 + (h & 1u);
                if (d.n < hdr) {
                    d.active = false;
                } else if (hdr == 2 && (hs >> 8) != 0) {
                    d.status = kInvalidConf;
                    d.active = false;
                } else {
                    d.wbits = (int)((h >> 5) & 7u) + 8;
                    d.lbits = (int)((h >> 3) & 3u) + 5;
                    d.extended = (h & 2u) != 0;
                    d.dict_reset = (h & 1u) != 0;
                    const bool use_custom = (h & 4u) != 0;
                    if (d.wbits > a.window_bits_max || d.wbits > WMAXBITS || (use_custom && !a.custom)) {
                        d.status = kInvalidConf;
                        d.active = false;
                    } else {
                        d.min_pat = min_pattern_size(d.wbits, d.lbits);
                        d.mask = (1 << d.wbits) - 1;
                        const int seed_lit = d.extended ? d.lbits : 8;
                        dict_src = use_custom ? a.custom : a.seed + (seed_lit <= 5 ? 0 : seed_lit <= 6 ? 1 : 2) * 32768;
                        d.ip = hdr;
                    }
                }
            }
        }
        d.wmask = d.mask >> 2;
        // normal-token symbols: v1 uses 0..13, the extended format 0..11 (12 = RLE, 13 = extended match)
        d.max_plain_sym = d.extended ? kSymRle - 1 : kSymFlush - 1;
        // seed the windows: lanes usually share one dictionary, so every word is a broadcast load
        {
            const int words = (d.mask + 1) >> 2;
            const int maxw = __reduce_max_sync(0xffffffffu, d.active ? words : 0);
            const uint4 *src = reinterpret_cast<const uint4 *>(dict_src);
            for (int k = 0; k < maxw; k += 4) {
                if (d.active && k < words) {
                    const uint4 v = __ldg(src + (k >> 2));
                    d.win.st_word(k, v.x);
                    d.win.st_word(k + 1, v.y);
                    d.win.st_word(k + 2, v.z);
                    d.win.st_word(k + 3, v.w);
                }
            }
        }
        __syncwarp();

        d.bb = 0;
        d.nb = 0;
        d.has_next = false;
        d.next_word = 0;
        d.opos = 0;
        d.flushed = 0;
        d.wpos = 0;
        decode_next(d, a.lut, a.seed);

        while (__any_sync(0xffffffffu, d.active)) {
            // ---- copy phase: literals (1 byte) and plain tokens (<= 16 window bytes) -------------------------
            // All source words are read before any destination word is written, which is exactly the
            // "snapshot" behaviour of tamp_window_copy's direction rule (common.c:58-86).
            const int len = d.t_len;
            uint32_t s0 = d.t_lit, s1 = 0, s2 = 0, s3 = 0;
            if (d.t_src >= 0) {
                const int w0 = d.t_src >> 2, sh = (d.t_src & 3) * 8;
                const uint32_t a0 = d.win.ld_word(w0 & d.wmask), a1 = d.win.ld_word((w0 + 1) & d.wmask),
                               a2 = d.win.ld_word((w0 + 2) & d.wmask), a3 = d.win.ld_word((w0 + 3) & d.wmask),
                               a4 = d.win.ld_word((w0 + 4) & d.wmask);
                s0 = __funnelshift_r(a0, a1, sh);
                s1 = __funnelshift_r(a1, a2, sh);
                s2 = __funnelshift_r(a2, a3, sh);
                s3 = __funnelshift_r(a3, a4, sh);
                if (len > d.t_mlen) {  // literal bytes behind the match: spliced in at byte t_mlen (at most 2, len <= 16)
                    const int wi = d.t_mlen >> 2, bo = (d.t_mlen & 3) * 8;
                    const uint32_t lo = d.t_lit << bo, hi = bo ? d.t_lit >> (32 - bo) : 0u, keep = (1u << bo) - 1u;
                    if (wi == 0) { s0 = (s0 & keep) | lo; s1 = hi; }
                    if (wi == 1) { s1 = (s1 & keep) | lo; s2 = hi; }
                    if (wi == 2) { s2 = (s2 & keep) | lo; s3 = hi; }
                    if (wi == 3) { s3 = (s3 & keep) | lo; }
                }
            }
            {
                const int dd = d.wpos & 3, dw = d.wpos >> 2;
                const int sh = dd * 8;
                const int nwords = len ? (dd + len + 3) >> 2 : 0;
                // destination word j receives source bytes [4j - dd, 4j - dd + 4)
                uint32_t prev = 0;
#pragma unroll
                for (int j = 0; j < 5; j++) {
                    const uint32_t cur = j == 0 ? s0 : j == 1 ? s1 : j == 2 ? s2 : j == 3 ? s3 : 0u;
                    if (j < nwords) {
                        const uint32_t x = __funnelshift_l(prev, cur, sh);
                        const uint32_t m = low_bytes(len + dd - 4 * j) & ~low_bytes(dd - 4 * j);
                        const int wi = (dw + j) & d.wmask;
                        const uint32_t old = d.win.ld_word(wi);
                        d.win.st_word(wi, (x & m) | (old & ~m));
                    }
                    prev = cur;
                }
                d.wpos = (d.wpos + len) & d.mask;
                d.opos += (uint32_t)len;
            }
            // ---- deliver output 16 bytes at a time from the window ------------------------------------------
            if (d.opos - d.flushed >= 16) {
                if (a.aligned_io && (d.flushed & 15u) == 0) {
                    // 16 window bytes from any window alignment (extended tokens shift it): 5 words + funnel shifts
                    const int wsrc = (d.wpos - (int)(d.opos - d.flushed)) & d.mask;
                    const int w0 = wsrc >> 2, sh = (wsrc & 3) * 8;
                    const uint32_t a0 = d.win.ld_word(w0 & d.wmask), a1 = d.win.ld_word((w0 + 1) & d.wmask),
                                   a2 = d.win.ld_word((w0 + 2) & d.wmask), a3 = d.win.ld_word((w0 + 3) & d.wmask);
                    uint4 v;
                    if (sh == 0) {
                        v = make_uint4(a0, a1, a2, a3);
                    } else {
                        const uint32_t a4 = d.win.ld_word((w0 + 4) & d.wmask);
                        v.x = __funnelshift_r(a0, a1, sh);
                        v.y = __funnelshift_r(a1, a2, sh);
                        v.z = __funnelshift_r(a2, a3, sh);
                        v.w = __funnelshift_r(a3, a4, sh);
                    }
                    *reinterpret_cast<uint4 *>(d.out + d.flushed) = v;
                    d.flushed += 16;
                } else {
                    // output row position not 16-byte aligned (after an extended token, or unaligned rows):
                    // byte-wise until it is
                    do {
                        d.out[d.flushed] = (uint8_t)d.win.ld((d.wpos - (int)(d.opos - d.flushed)) & d.mask);
                        d.flushed += 1;
                    } while ((d.flushed & 15u) != 0 && d.flushed < d.opos);
                }
            }
            // ---- decode the next token (bit stream only; overlaps with the copy above) ------------------------
            decode_next(d, a.lut, a.seed);
        }
        // stream ends: deliver the tail, publish size and status
        if (mine) {
            deliver_pending_bytes(d);
            a.b.out_sizes[stream] = d.opos;
            if (a.b.status) a.b.status[stream] = (int8_t)d.status;
        }
        __syncwarp();
    }
}

#ifndef TB_EMU
__global__ void k_store_lut(uint8_t *dst) {
    if (threadIdx.x < 128) dst[threadIdx.x] = kHuff.lut[threadIdx.x];
}
uint8_t *g_lut = nullptr;

template <int WMAXBITS>
void launch_dec(const FastDecArgs &a, cudaStream_t st, bool small_grid) {
    static int blocks_per_sm = 0, sms = 0;
    constexpr int kWarps = kWarpsPerCtaDec<WMAXBITS>;
    const size_t smem = (size_t)kWarps * 32 * (1 << WMAXBITS);
    if (!blocks_per_sm) {
        cudaFuncSetAttribute(k_fast_decompress<WMAXBITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_fast_decompress<WMAXBITS>, kWarps * 32, smem);
        if (blocks_per_sm < 1) blocks_per_sm = 1;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    if (small_grid) {
        // pick-up pass with (probably) nothing to do: one-warp CTAs (32 KiB of windows) find room beside the resident
        // CTAs of the split decompressor instead of waiting for a whole SM's shared memory
        const uint64_t want1 = (a.b.n_streams + 31) / 32;
        const unsigned grid1 = (unsigned)(want1 < (uint64_t)sms ? want1 : (uint64_t)sms);
        k_fast_decompress<WMAXBITS><<<grid1, 32, (size_t)32 * (1 << WMAXBITS), st>>>(a);
        count_launch();
        return;
    }
    const uint64_t per_block = kWarps * 32;
    uint64_t want = (a.b.n_streams + per_block - 1) / per_block;
    uint64_t persistent = (uint64_t)sms * blocks_per_sm;
    unsigned grid = (unsigned)(want < persistent ? want : persistent);
    k_fast_decompress<WMAXBITS><<<grid, kWarps * 32, smem, st>>>(a);
    count_launch();
}

#endif  // TB_EMU

}  // namespace

#ifndef TB_EMU
bool launch_fast_decompress_batch(const uint8_t *d_seed, const uint8_t *d_custom, int window_bits_max,
                                  const BatchArgs &b, cudaStream_t st, bool only_deferred, bool small_grid) {
    if (window_bits_max > 10) return false;
    if (b.out_stride > 0xFFFFFFF0ull) return false;
    if (b.n_streams == 0) return true;
    if (!g_lut) {
        if (cudaMalloc(&g_lut, 128) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        k_store_lut<<<1, 128, 0, st>>>(g_lut);
        count_launch();
        cudaStreamSynchronize(st);  // the table is published to every stream from here on
    }
    FastDecArgs a;
    a.lut = g_lut;
    a.b = b;
    a.seed = d_seed;
    a.custom = d_custom;
    a.window_bits_max = window_bits_max;
    a.aligned_io = ((b.out_stride & 15) == 0 && (reinterpret_cast<uintptr_t>(b.out) & 15) == 0) ? 1 : 0;
    a.only_deferred = only_deferred ? 1 : 0;
    switch (window_bits_max) {
        case 8: launch_dec<8>(a, st, small_grid); break;
        case 9: launch_dec<9>(a, st, small_grid); break;
        default: launch_dec<10>(a, st, small_grid); break;
    }
    return true;
}

#endif  // TB_EMU

}  // namespace tb
