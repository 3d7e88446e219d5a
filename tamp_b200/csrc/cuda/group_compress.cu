// Batch compressor for windows <= 1024 bytes: several independent streams per warp.
//
// Same arithmetic as fast_compress.cu (the window is kept as 32 nibble bitmaps; one "level" ANDs in the
// positions that also match the next lookahead byte; the longest match is the last non-empty level and its
// index the lowest set bit — find_best_match of compressor_find_match_desktop.c:82-167), re-laid-out so
// that the per-token scalar work (lookahead fetch, token assembly, bit writer, window-update bookkeeping,
// loop control) is paid once per warp instruction for SPW = 32 / LPS streams instead of once per stream:
//
//   * a stream is owned by a GROUP of LPS lanes; lane gl of the group holds WPL = W / 32 / LPS consecutive
//     bitmap words of every row (W = 1024, LPS = 8: one 128-bit LDS per row and lane);
//   * the groups of a warp run the poll loop in lock-step.  Inside a poll the level chain runs until every
//     group's match has ended (one full-warp ballot per level decides both "which groups go on" and "is
//     anybody left"); finished groups keep their candidate set through `m &= shifted | keep`;
//   * every group-uniform quantity (p, window position, bit accumulator ...) is held redundantly by the
//     group's lanes, so the code below reads like scalar code per stream;
//   * window update: lane gl owns rows gl, gl + LPS, ...; the row words of the next 32 window bytes are
//     built by scattering the bytes' nibbles into a 32-word scratch with shared-memory atomics
//     (a 32-byte x 32-row transpose), then merged into the bitmap column as tokens consume them;
//   * bit writer: a 64-bit accumulator per group; completed 32-bit words are parked one per lane and
//     leave with one LPS-word coalesced store.
//
// Reference semantics restated here: tamp_compressor_poll (compressor.c:532-660), poll_extended_handling
// (:437-525), find_extended_match (:297-333), write_rle_token / write_extended_match_token (:342-415),
// tamp_compressor_flush (:728-810).  Ring fill at poll entry is min(16, N - p).
#include "../tb_wire.h"
#include "tb_cuda.h"
#include "tb_device_common.cuh"
#include "tb_ptx.cuh"

namespace tb {

namespace {

constexpr int kRing = 512;     // per-stream input ring (power of two)
constexpr int kMirror = 32;    // first ring bytes repeated behind it: unwrapped 20-byte lookahead reads
constexpr int kWarpsPerCta = 2;
constexpr uint32_t kFull = 0xffffffffu;

template <int N> struct Log2 { static constexpr int v = 1 + Log2<N / 2>::v; };
template <> struct Log2<1> { static constexpr int v = 0; };

template <int WBITS, int LPS>
struct Geo {
    static constexpr int W = 1 << WBITS;
    static constexpr int WW = W / 32;                 // words per bitmap row
    static constexpr int WPL = WW / LPS;              // words per lane
    static constexpr int SPW = 32 / LPS;              // streams per warp
    static constexpr int RPL = 32 / LPS;              // bitmap rows owned per lane (window update)
    static constexpr int BPL = 32 / LPS;              // bytes of a 32-byte block handled per lane
    static constexpr int PAD = WPL >= 4 ? 4 : WPL;    // row padding: column accesses of LPS rows hit LPS banks
    static constexpr int RS = WW + PAD;               // row stride in words
    static constexpr int ROW_BYTES = 32 * RS * 4;     // multiple of 16
    static constexpr int OFF_RING = ROW_BYTES;
    static constexpr int OFF_SCR = OFF_RING + kRing + kMirror;
    static constexpr int SCR_WORDS = WW > 32 ? WW : 32;
    static constexpr int OFF_MBAR = OFF_SCR + SCR_WORDS * 4;
    static constexpr int PER_STREAM = OFF_MBAR + 16;
    static constexpr int LUT_BYTES = 64;
    static constexpr int CTA_BYTES = LUT_BYTES + PER_STREAM * SPW * kWarpsPerCta;
    static_assert(WPL * LPS == WW && WPL >= 1 && WPL <= 8, "bad lanes-per-stream for this window");
    static_assert(LPS >= 4 && LPS <= 32, "4..32 lanes per stream");
    static_assert(PER_STREAM % 16 == 0, "per-stream regions stay 16-byte aligned");
};

struct GroupCompArgs {
    BatchArgs b;
    const uint32_t *dictrows;
    int literal, flags, write_token;
};

template <int WBITS, int LPS, bool EXT>
struct Stream {
    using G = Geo<WBITS, LPS>;
    static constexpr int W = G::W, WW = G::WW, WPL = G::WPL, RS = G::RS, RPL = G::RPL, BPL = G::BPL, MASK = W - 1;

    // ---- shared-memory views ----------------------------------------------------------------------
    uint32_t *rows;        // [32][RS] nibble bitmaps of the window
    const uint32_t *rowp;  // rows + gl * WPL          (this lane's words of row 0)
    uint32_t *myrow;       // rows + gl * RS           (row gl; row gl + LPS * i sits i * LPS * RS words on)
    uint8_t *ring;
    uint32_t *scr;         // 32-word scratch: block transposes, wide shifts
    const uint32_t *lut;   // Huffman (code | bits << 16) per symbol, CTA-shared
    int gl;
    uint32_t gmask;

    // ---- per-stream state (group-uniform unless noted) ----------------------------------------
    const uint8_t *src;
    int N, npad, loaded;
    int p, res;
    int trig_p;        // p at which this group needs a slow step: stream end / ring refill / shrinking lookahead
    bool active;
    uint32_t in[4];    // 16 lookahead bytes at p (fetched at the end of the previous step)

    int wpos, cb, blk_src;
    uint32_t old_r[RPL], next_r[RPL];  // per lane: its rows' words of the block being filled
    uint32_t last;                     // last byte written to the window (RLE reference byte)

    uint32_t acc_lo, acc_hi;  // bit accumulator: the low `cnt` bits are pending
    int cnt;
    uint32_t *out32;
    uint32_t ow;              // words emitted so far
    uint32_t myword;          // per lane: parked output word (ow % LPS == gl)

    int lbits, min_pat;
    int rle, ext_n, ext_pos, ext_start;
    uint32_t ext_set[WPL];    // per lane

    __device__ __forceinline__ bool group_any(bool pred) const { return __ballot_sync(gmask, pred) != 0u; }
    __device__ __forceinline__ void group_sync() const { __syncwarp(gmask); }
    __device__ __forceinline__ uint32_t T(int pos) const { return ring[pos & (kRing - 1)]; }

    // ---- bit writer (write_to_bit_buffer / partial_flush, compressor.c:49-75) ----------------------------
    __device__ __forceinline__ void put(uint32_t bits, int nb) {  // nb < 32
        acc_hi = __funnelshift_l(acc_lo, acc_hi, nb);
        acc_lo = (acc_lo << nb) | bits;
        cnt += nb;
        if (cnt >= 32) {
            cnt -= 32;
            const uint32_t word = __funnelshift_r(acc_lo, acc_hi, cnt);
            if (gl == (int)(ow & (LPS - 1))) myword = word;
            ow++;
            if ((ow & (LPS - 1)) == 0) out32[ow - LPS + gl] = __byte_perm(myword, 0, 0x0123);
        }
    }
    __device__ __forceinline__ uint32_t exthuff_bits(int v, int t, int &nb) const {  // write_extended_huffman, :257-263
        const uint32_t e = lut[v >> t];
        nb = (int)(e >> 16) - 1 + t;
        return ((e & 0xFFFFu) << t) | (uint32_t)(v & ((1 << t) - 1));
    }

    // ---- bitmap primitives -------------------------------------------------------------------------------
    __device__ __forceinline__ void load_row(uint32_t (&e)[WPL], const uint32_t *q) const {
        if constexpr (WPL == 8) {
            const uint4 a = *reinterpret_cast<const uint4 *>(q), b = *reinterpret_cast<const uint4 *>(q + 4);
            e[0] = a.x; e[1] = a.y; e[2] = a.z; e[3] = a.w; e[4] = b.x; e[5] = b.y; e[6] = b.z; e[7] = b.w;
        } else if constexpr (WPL == 4) {
            const uint4 a = *reinterpret_cast<const uint4 *>(q);
            e[0] = a.x; e[1] = a.y; e[2] = a.z; e[3] = a.w;
        } else if constexpr (WPL == 2) {
            const uint2 a = *reinterpret_cast<const uint2 *>(q);
            e[0] = a.x; e[1] = a.y;
        } else {
            e[0] = *q;
        }
    }
    // E(c): window positions holding byte c (this lane's words).
    __device__ __forceinline__ void row_of(uint32_t (&e)[WPL], uint32_t c) const {
        uint32_t a[WPL], b[WPL];
        load_row(a, rowp + (c >> 4) * RS);
        load_row(b, rowp + (16 + (c & 15)) * RS);
#pragma unroll
        for (int i = 0; i < WPL; i++) e[i] = a[i] & b[i];
    }
    // this lane's words of (E >> k), 0 < k < 32.  Lock-step code only: the shuffle names the whole warp (a
    // per-group member mask would make the compiler verify the mask with MATCH/REDUX on every call).
    __device__ __forceinline__ void shifted_small(uint32_t (&sh)[WPL], const uint32_t (&e)[WPL], int k) const {
        uint32_t nb = __shfl_down_sync(kFull, e[0], 1, LPS);
        if (gl == LPS - 1) nb = 0;
#pragma unroll
        for (int i = 0; i < WPL; i++) sh[i] = __funnelshift_r(e[i], i + 1 < WPL ? e[i + 1] : nb, k);
    }
    // this lane's words of (E >> k), any k (rare: extended-match continuation) — through the scratch line
    __device__ __forceinline__ void shifted_any(uint32_t (&sh)[WPL], const uint32_t (&e)[WPL], int k) const {
        group_sync();
#pragma unroll
        for (int i = 0; i < WPL; i++) scr[gl * WPL + i] = e[i];
        group_sync();
        const int s = k >> 5, sb = k & 31;
        uint32_t a[WPL + 1];
#pragma unroll
        for (int i = 0; i <= WPL; i++) {
            const int j = gl * WPL + i + s;
            a[i] = j < WW ? scr[j] : 0u;
        }
#pragma unroll
        for (int i = 0; i < WPL; i++) sh[i] = __funnelshift_r(a[i], a[i + 1], sb);
    }
    // Lowest set position of the group's candidate set.  CONVERGED: called from lock-step code (whole-warp
    // butterflies that stay inside the group); otherwise from a group-divergent branch.
    template <bool CONVERGED>
    __device__ __forceinline__ int lowest_pos(const uint32_t (&m)[WPL]) const {
        uint32_t w = m[WPL - 1];
        int wi = WPL - 1;
#pragma unroll
        for (int i = WPL - 2; i >= 0; i--)
            if (m[i]) {
                w = m[i];
                wi = i;
            }
        uint32_t mine = w ? (uint32_t)((gl * WPL + wi) * 32 + __ffs(w) - 1) : 0xFFFFu;
        if (CONVERGED) {
            // the group's first lane with a candidate holds the lowest position: one ballot, one shuffle
            const uint32_t have = __ballot_sync(kFull, w != 0u) & gmask;
            const int src = have ? __ffs(have) - 1 : 0;
            return (int)__shfl_sync(kFull, mine, src);
        }
        return (int)__reduce_min_sync(gmask, mine);
    }

    // ---- window update -----------------------------------------------------------------------------------
    // Row words (this lane's RPL rows) of input bytes [s, s+32); bytes at or past N never reach the window.
    __device__ __forceinline__ void block_rows(int s, uint32_t (&out)[RPL]) const {
        constexpr int LB = Log2<LPS>::v;
        group_sync();
#pragma unroll
        for (int i = 0; i < RPL; i++) scr[gl * RPL + i] = 0u;
        group_sync();
        const int q0 = s + gl * BPL;
#pragma unroll
        for (int j = 0; j < BPL; j++) {
            const int q = q0 + j;
            if (q < N) {
                const uint32_t c = T(q);
                const uint32_t bit = 1u << (gl * BPL + j);
                const uint32_t rh = c >> 4, rl = 16 + (c & 15);
                atomicOr(&scr[(rh & (LPS - 1)) * RPL + (rh >> LB)], bit);
                atomicOr(&scr[(rl & (LPS - 1)) * RPL + (rl >> LB)], bit);
            }
        }
        group_sync();
#pragma unroll
        for (int i = 0; i < RPL; i++) out[i] = scr[gl * RPL + i];
    }
    // Lock-step variant: every group of the (converged) warp takes part, `go` selects the groups that rebuild.
    __device__ __forceinline__ void block_rows_lockstep(bool go, int s, uint32_t (&out)[RPL]) const {
        constexpr int LB = Log2<LPS>::v;
        __syncwarp();
        if (go) {
#pragma unroll
            for (int i = 0; i < RPL; i++) scr[gl * RPL + i] = 0u;
        }
        __syncwarp();
        if (go) {
            const int q0 = s + gl * BPL;
#pragma unroll
            for (int j = 0; j < BPL; j++) {
                const int q = q0 + j;
                if (q < N) {
                    const uint32_t c = T(q);
                    const uint32_t bit = 1u << (gl * BPL + j);
                    const uint32_t rh = c >> 4, rl = 16 + (c & 15);
                    atomicOr(&scr[(rh & (LPS - 1)) * RPL + (rh >> LB)], bit);
                    atomicOr(&scr[(rl & (LPS - 1)) * RPL + (rl >> LB)], bit);
                }
            }
        }
        __syncwarp();
        if (go) {
#pragma unroll
            for (int i = 0; i < RPL; i++) out[i] = scr[gl * RPL + i];
        }
    }
    __device__ __forceinline__ void store_col(uint32_t lm2) {
#pragma unroll
        for (int i = 0; i < RPL; i++) myrow[i * LPS * RS + cb] = (next_r[i] & lm2) | (old_r[i] & ~lm2);
    }
    // Block boundary crossed or (extended format) the pending block was built from other input bytes.
    __device__ __forceinline__ void window_write_slow(int s, int m) {
        while (m > 0) {
            const int off = wpos & 31;
            if (EXT && blk_src + off != s) {
                const uint32_t lm = (1u << off) - 1u;  // off < 32
                blk_src = s - off;
                uint32_t nr[RPL];
                block_rows(blk_src, nr);
#pragma unroll
                for (int i = 0; i < RPL; i++) {
                    old_r[i] = (next_r[i] & lm) | (old_r[i] & ~lm);
                    next_r[i] = (nr[i] & ~lm) | (old_r[i] & lm);
                }
            }
            const int take = m < 32 - off ? m : 32 - off;
            const int off2 = off + take;
            store_col(off2 == 32 ? 0xffffffffu : ((1u << off2) - 1u));
            wpos = (wpos + take) & MASK;
            s += take;
            m -= take;
            if (off2 == 32) {
                cb = wpos >> 5;
                blk_src = s;
                group_sync();
#pragma unroll
                for (int i = 0; i < RPL; i++) old_r[i] = myrow[i * LPS * RS + cb];
                block_rows(s, next_r);
            }
        }
    }
    // Append m input bytes starting at input position s to the window (destination wraps).  Called by the
    // converged warp.  Three cases: the bytes stay inside the pending 32-byte block (one column store);
    // they cross into the next block once (lock-step: every group with such a crossing rebuilds its block
    // rows together); anything else (extended format only: long copies, re-aligned sources) per group.
    __device__ __forceinline__ void window_write(int s, int m) {
        const int off = wpos & 31, off2 = off + m;
        const bool aligned = !EXT || blk_src + off == s;
        const bool inside = m > 0 && off2 < 32 && aligned;
        const bool cross = m > 0 && off2 >= 32 && m <= 32 && aligned;
        if (inside) {
            store_col((1u << off2) - 1u);
            wpos += m;
        }
        if (EXT && m > 0 && !inside && !cross) window_write_slow(s, m);
        if (__any_sync(kFull, cross)) {
            if (cross) {
                store_col(0xffffffffu);
                const int take = 32 - off;
                wpos = (wpos + take) & MASK;
                s += take;
                m -= take;
                cb = wpos >> 5;
                blk_src = s;
#pragma unroll
                for (int i = 0; i < RPL; i++) old_r[i] = myrow[i * LPS * RS + cb];
            }
            block_rows_lockstep(cross, s, next_r);
            if (cross && m > 0) {
                store_col((1u << m) - 1u);
                wpos += m;
            }
        }
    }

    // ---- the level chain (find_best_match), all groups of the warp in lock-step ------------------------
    // want: this group searches.  in[]: 16 lookahead bytes.  L: usable lookahead (TAIL only; else lfull).
    // Returns len (>= 2, or 0) and leaves the candidate set of the last successful level in m.
    template <bool TAIL>
    __device__ __forceinline__ int search(bool want, const uint32_t (&in)[4], int L, int lfull, uint32_t (&m)[WPL]) const {
        uint32_t e[WPL], sh[WPL];
        row_of(m, in[0] & 0xFFu);
        row_of(e, (in[0] >> 8) & 0xFFu);
        shifted_small(sh, e, 1);
        uint32_t t = 0;
#pragma unroll
        for (int i = 0; i < WPL; i++) {
            m[i] &= sh[i];
            t |= m[i];
        }
        bool alive = want && t != 0u;
        if (TAIL) alive = alive && L >= 2 && L >= min_pat;
        uint32_t b = __ballot_sync(kFull, alive);
        if (b == 0u) return 0;
        alive = (b & gmask) != 0u;
        int len = alive ? 2 : 0;
#define TB_LEVEL(K)                                                                   \
    {                                                                                 \
        if ((K) == 15 && lfull < 16) goto done;                                       \
        const uint32_t c = (in[(K) >> 2] >> (8 * ((K) & 3))) & 0xFFu;                 \
        row_of(e, c);                                                                 \
        shifted_small(sh, e, (K));                                                    \
        t = 0;                                                                        \
        _Pragma("unroll") for (int i = 0; i < WPL; i++) t |= m[i] & sh[i];            \
        bool ok = alive && t != 0u;                                                   \
        if (TAIL) ok = ok && (K) < L;                                                 \
        b = __ballot_sync(kFull, ok);                                                 \
        if (b == 0u) goto done;                                                       \
        alive = (b & gmask) != 0u;                                                    \
        if (alive) {                                                                  \
            len++;                                                                    \
            _Pragma("unroll") for (int i = 0; i < WPL; i++) m[i] &= sh[i];            \
        }                                                                             \
    }
        TB_LEVEL(2) TB_LEVEL(3) TB_LEVEL(4) TB_LEVEL(5) TB_LEVEL(6) TB_LEVEL(7) TB_LEVEL(8) TB_LEVEL(9)
        TB_LEVEL(10) TB_LEVEL(11) TB_LEVEL(12) TB_LEVEL(13) TB_LEVEL(14) TB_LEVEL(15)
#undef TB_LEVEL
    done:
        return len;
    }

    // ---- one tamp_compressor_poll per active group (compressor.c:532-660) ------------------------------
    __device__ __forceinline__ void fetch_lookahead() {
        const uint8_t *b = ring + (p & (kRing - 4));
        const uint32_t a0 = *reinterpret_cast<const uint32_t *>(b), a1 = *reinterpret_cast<const uint32_t *>(b + 4),
                       a2 = *reinterpret_cast<const uint32_t *>(b + 8), a3 = *reinterpret_cast<const uint32_t *>(b + 12),
                       a4 = *reinterpret_cast<const uint32_t *>(b + 16);
        const int sh = (p & 3) * 8;
        in[0] = __funnelshift_r(a0, a1, sh);
        in[1] = __funnelshift_r(a1, a2, sh);
        in[2] = __funnelshift_r(a2, a3, sh);
        in[3] = __funnelshift_r(a3, a4, sh);
    }

    // SLOW: some group of the warp has fewer than 16 bytes of lookahead left (ring fill = N - p).
    template <bool SLOW>
    __device__ __forceinline__ void poll_step(int lfull, int ext_cap) {
        const int r = SLOW ? (N - p < 16 ? N - p : 16) : 16;
        const int L = r < lfull ? r : lfull;
        bool want = active;       // this group runs the match search in this step
        int rle_total = 0;        // > 0: short-run decision pending on the search result (compressor.c:490-503)
        int rle_avail = 0;
        uint32_t bits = 0;        // what this step emits (nb == 0: nothing) and appends to the window (wn == 0: nothing)
        int nb = 0, ws = 0, wn = 0;

        if (EXT) {
            if (active && ext_n) {  // extended-match continuation (compressor.c:442-469)
                want = false;
                int avail = r;
                bool emit = false;
                while (avail > 0) {
                    if (ext_pos + ext_n >= W || ext_n >= ext_cap) {
                        emit = true;
                        break;
                    }
                    const int maxp = ext_n + avail < ext_cap ? ext_n + avail : ext_cap;
                    int n = ext_n;
                    uint32_t m[WPL];
#pragma unroll
                    for (int i = 0; i < WPL; i++) m[i] = ext_set[i];
                    while (n < maxp) {
                        uint32_t e[WPL], sh[WPL];
                        row_of(e, T(ext_start + n));
                        shifted_any(sh, e, n);
                        uint32_t t = 0;
#pragma unroll
                        for (int i = 0; i < WPL; i++) t |= m[i] & sh[i];
                        if (!group_any(t != 0u)) break;
#pragma unroll
                        for (int i = 0; i < WPL; i++) m[i] &= sh[i];
                        n++;
                    }
                    if (n > ext_n) {
                        avail -= n - ext_n;
                        p += n - ext_n;
                        ext_pos = lowest_pos<false>(m);
#pragma unroll
                        for (int i = 0; i < WPL; i++) ext_set[i] = m[i];
                        const bool stopped_early = n < maxp;
                        ext_n = n;
                        if (stopped_early && avail > 0) {  // the next search cannot extend: emit now
                            emit = true;
                            break;
                        }
                        continue;
                    }
                    emit = true;
                    break;
                }
                if (emit) {  // write_extended_match_token, compressor.c:377-415
                    int xn;
                    const uint32_t x = exthuff_bits(ext_n - min_pat - 12, 3, xn);
                    const uint32_t sym = lut[kSymExt];
                    bits = ((((sym & 0xFFFFu) << xn) | x) << WBITS) | (uint32_t)ext_pos;
                    nb = (int)(sym >> 16) + xn + WBITS;
                    const int room = W - wpos;
                    ws = ext_start;
                    wn = ext_n < room ? ext_n : room;
                    ext_n = 0;
                }
            } else if (active && (rle != 0 || (in[0] & 0xFFu) == last)) {
                // RLE accumulation (compressor.c:471-523).  Otherwise avail = total = 0: the block is a no-op.
                int avail = 16;
                {
                    const uint32_t bl = last * 0x01010101u;
#pragma unroll
                    for (int i = 3; i >= 0; i--) {
                        const uint32_t x = in[i] ^ bl;
                        if (x) avail = 4 * i + ((__ffs(x) - 1) >> 3);
                    }
                    if (avail > r) avail = r;
                    if (avail > kRleMax - rle) avail = kRleMax - rle;
                }
                const int total = rle + avail;
                const bool ended = (avail < r) || (total >= kRleMax);
                if (!ended && total > 0) {
                    rle = total;
                    p += avail;
                    want = false;
                } else if (total >= 2) {
                    if (total == avail && total <= 6) {  // short run seen entirely in this poll: a longer match wins
                        rle_total = total;
                        rle_avail = avail;
                    } else {
                        want = false;
                        p += avail;
                        int xn;
                        const uint32_t x = exthuff_bits(total - 2, 4, xn);
                        const uint32_t sym = lut[kSymRle];
                        bits = ((sym & 0xFFFFu) << xn) | x;
                        nb = (int)(sym >> 16) + xn;
                        const int room = W - wpos;
                        const int nw = total < kRleWindowMax ? total : kRleWindowMax;
                        wn = nw < room ? nw : room;
                        ws = p - wn;  // the run's bytes are all equal: any wn of them
                        rle = 0;
                    }
                } else if (rle == 1) {  // lone run byte from an earlier poll
                    want = false;
                    bits = (1u << lbits) | last;
                    nb = lbits + 1;
                    ws = p - 1;
                    wn = 1;
                    rle = 0;
                }
            }
        }

        // ---- match search: every group, lock-step ----
        uint32_t m[WPL];
        int len;
        len = search<SLOW>(want, in, L, lfull, m);
        const int idx = lowest_pos<true>(m);

        if (EXT && rle_total) {
            if (len > rle_total) {
                rle = 0;  // the pattern match wins; falls through to the token below
            } else {
                want = false;
                p += rle_avail;
                int xn;
                const uint32_t x = exthuff_bits(rle_total - 2, 4, xn);
                const uint32_t sym = lut[kSymRle];
                bits = ((sym & 0xFFFFu) << xn) | x;
                nb = (int)(sym >> 16) + xn;
                const int room = W - wpos;
                wn = rle_total < room ? rle_total : room;  // rle_total <= 6 < kRleWindowMax
                ws = p - wn;
                rle = 0;
            }
        }
        if (want) {
            if (len < min_pat) {
                const uint32_t c = in[0] & 0xFFu;
                if (c >> lbits) {
                    res = kExcessBits;
                    p = N;  // ends the stream
                } else {
                    bits = (1u << lbits) | c;
                    nb = lbits + 1;
                    ws = p;
                    wn = 1;
                    p += 1;
                }
            } else if (EXT && len > min_pat + 11) {
                ext_n = len;
                ext_pos = idx;
                ext_start = p;
#pragma unroll
                for (int i = 0; i < WPL; i++) ext_set[i] = m[i];
                p += len;
            } else {
                const uint32_t e = lut[len - min_pat];
                bits = ((e & 0xFFFFu) << WBITS) | (uint32_t)idx;
                nb = (int)(e >> 16) + WBITS;
                ws = p;
                wn = len;
                p += len;
            }
        }
        put(bits, nb);
        fetch_lookahead();  // for the next step
        if (EXT && wn > 0) last = T(ws + wn - 1);
        __syncwarp();  // every lane's bitmap reads of this step precede the column update below
        window_write(ws, wn);
    }

    __device__ __forceinline__ void load_input(int from, int to) {  // [from, to) 16-byte aligned, within npad
        for (int off = from + gl * 16; off < to; off += LPS * 16) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + off));
            const int ro = off & (kRing - 1);
            *reinterpret_cast<uint4 *>(ring + ro) = v;
            if (ro < kMirror) *reinterpret_cast<uint4 *>(ring + kRing + ro) = v;
        }
    }
};

template <int WBITS, int LPS, bool EXT>
__global__ void __launch_bounds__(kWarpsPerCta * 32) k_group_compress(GroupCompArgs a) {
    using G = Geo<WBITS, LPS>;
    using S = Stream<WBITS, LPS, EXT>;
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / LPS, gl = lane % LPS;
    uint32_t *lut = reinterpret_cast<uint32_t *>(smem);
    if (threadIdx.x < 16) lut[threadIdx.x] = (uint32_t)kHuff.code[threadIdx.x] | ((uint32_t)kHuff.bits[threadIdx.x] << 16);
    uint8_t *base = smem + G::LUT_BYTES + (size_t)(warp * G::SPW + grp) * G::PER_STREAM;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(base + G::OFF_MBAR);

    S st;
    st.rows = reinterpret_cast<uint32_t *>(base);
    st.rowp = st.rows + gl * G::WPL;
    st.myrow = st.rows + gl * G::RS;
    st.ring = base + G::OFF_RING;
    st.scr = reinterpret_cast<uint32_t *>(base + G::OFF_SCR);
    st.lut = lut;
    st.gl = gl;
    st.gmask = LPS == 32 ? kFull : (((1u << LPS) - 1u) << (grp * LPS));
    st.lbits = a.literal;
    st.min_pat = min_pattern_size(WBITS, a.literal);
    const int lfull = EXT ? 16 : st.min_pat + 13;  // MAX_PATTERN_SIZE clipped by the 16-byte ring
    const int ext_cap = st.min_pat + 11 + kExtExtraMax;
    // keep >= 176 bytes of history (extended-match window writes re-read their source) and >= 64 of lookahead
    constexpr int kRefillBelow = kRing - 176 - LPS * 16;

    if (gl == 0) mbar_init(mbar, 1);
    __syncthreads();
    uint32_t phase = 0;

    st.active = false;
    st.N = 0; st.npad = 0; st.loaded = 0; st.p = 0; st.res = kOk;
    st.wpos = 0; st.cb = 0; st.blk_src = 0; st.last = 0;
    st.acc_lo = 0; st.acc_hi = 0; st.cnt = 0; st.ow = 0; st.myword = 0; st.out32 = nullptr; st.src = nullptr; st.trig_p = 0;
    st.rle = 0; st.ext_n = 0; st.ext_pos = 0; st.ext_start = 0;
#pragma unroll
    for (int i = 0; i < G::WPL; i++) st.ext_set[i] = 0;
#pragma unroll
    for (int i = 0; i < G::RPL; i++) st.old_r[i] = st.next_r[i] = 0;

    const uint64_t ngroups = (uint64_t)gridDim.x * kWarpsPerCta * G::SPW;
    uint64_t next = ((uint64_t)blockIdx.x * kWarpsPerCta + warp) * G::SPW + grp;
    uint64_t stream = 0;

    for (;;) {
        // Fast step: every group still has a full 16-byte lookahead and needs no service.
        if (!__any_sync(kFull, st.p >= st.trig_p)) {
            st.template poll_step<false>(lfull, ext_cap);
            continue;
        }
        // Slow step.  Group-divergent section first (group-scoped synchronisation only): finish a completed
        // stream, start the next, top up the input ring.
        for (;;) {
        if (!st.active) {
            if (next >= a.b.n_streams) break;
            // -- start a stream: dictionary bitmaps by TMA, head of the input by coalesced 128-bit loads --
            stream = next;
            next += ngroups;
            st.group_sync();
            if (gl == 0) {
                fence_proxy_async();
                mbar_expect_tx(mbar, G::ROW_BYTES);
                tma_load_1d(st.rows, a.dictrows, G::ROW_BYTES, mbar);
            }
            st.src = a.b.in + stream * a.b.in_stride;
            st.N = a.b.in_sizes ? (int)a.b.in_sizes[stream] : (int)a.b.in_stride;
            st.npad = (st.N + 15) & ~15;
            st.loaded = st.npad < kRing ? st.npad : kRing;
            st.load_input(0, st.loaded);
            mbar_wait(mbar, phase);
            phase ^= 1;
            st.group_sync();

            st.out32 = reinterpret_cast<uint32_t *>(a.b.out + stream * a.b.out_stride);
            st.ow = 0;
            st.cnt = 0;
            st.acc_lo = 0;
            st.acc_hi = 0;
            {
                const uint32_t header = ((uint32_t)(WBITS - 8) << 5) | ((uint32_t)(a.literal - 5) << 3) |
                                        ((a.flags & TB_F_CUSTOM_DICT) ? 4u : 0u) | (EXT ? 2u : 0u) |
                                        ((a.flags & TB_F_DICT_RESET) ? 1u : 0u);
                st.put(header, 8);
                if (a.flags & TB_F_DICT_RESET) st.put(0, 8);
            }
            st.wpos = 0;
            st.cb = 0;
            st.blk_src = 0;
#pragma unroll
            for (int i = 0; i < G::RPL; i++) st.old_r[i] = st.myrow[i * LPS * G::RS];
            st.block_rows(0, st.next_r);
            st.last = st.rows[G::WW] & 0xFFu;  // pad word of row 0 carries dictionary[W-1] (k_build_dictrows)
            st.p = 0;
            st.res = kOk;
            st.rle = 0;
            st.ext_n = 0;
            st.active = true;
        }
        if (st.p < st.N) break;
        {
            // -- flush (compressor.c:728-810) and stream epilogue --
            uint32_t out_bytes;
            if (st.res == kOk) {
                if (EXT) {
                    if (st.rle == 1) {
                        st.put((1u << st.lbits) | st.last, st.lbits + 1);
                    } else if (st.rle >= 2) {
                        int xn;
                        const uint32_t x = st.exthuff_bits(st.rle - 2, 4, xn);
                        const uint32_t sym = lut[kSymRle];
                        st.put(((sym & 0xFFFFu) << xn) | x, (int)(sym >> 16) + xn);
                    } else if (st.ext_n) {
                        int xn;
                        const uint32_t x = st.exthuff_bits(st.ext_n - st.min_pat - 12, 3, xn);
                        const uint32_t sym = lut[kSymExt];
                        st.put(((((sym & 0xFFFFu) << xn) | x) << WBITS) | (uint32_t)st.ext_pos,
                               (int)(sym >> 16) + xn + WBITS);
                    }
                }
                if (a.write_token && ((st.cnt & 7) || (a.flags & TB_F_DICT_RESET)))
                    st.put(kHuff.code[kSymFlush], kHuff.bits[kSymFlush]);
                out_bytes = st.ow * 4 + ((st.cnt + 7) >> 3);
            } else {
                // Error path: the reference has drained whole bytes of everything queued before the failing poll.
                out_bytes = st.ow * 4 + (st.cnt >> 3);
            }
            {
                const uint32_t parked = st.ow & (LPS - 1);
                if ((uint32_t)gl < parked) st.out32[st.ow - parked + gl] = __byte_perm(st.myword, 0, 0x0123);
                const uint32_t tail = out_bytes - st.ow * 4;  // < 4 bytes: the accumulator's pending bits, MSb first
                const uint32_t w0 = st.cnt ? st.acc_lo << (32 - st.cnt) : 0u;
                uint8_t *o8 = reinterpret_cast<uint8_t *>(st.out32 + st.ow);
                if ((uint32_t)gl < tail) o8[gl] = (uint8_t)(w0 >> (24 - 8 * gl));
            }
            if (gl == 0) {
                a.b.out_sizes[stream] = out_bytes;
                if (a.b.status) a.b.status[stream] = (int8_t)st.res;
            }
            st.active = false;
            st.N = 0;
            st.p = 0;
            st.rle = 0;
            st.ext_n = 0;
        }
        }
        if (st.active && st.loaded < st.npad && st.loaded - st.p < kRefillBelow) {
            st.group_sync();
            const int to = st.loaded + LPS * 16 < st.npad ? st.loaded + LPS * 16 : st.npad;
            st.load_input(st.loaded, to);
            st.loaded = to;
            st.group_sync();
        }
        // next slow step: when the lookahead starts to shrink (N - p < 16), or the ring needs topping up
        st.trig_p = !st.active ? 0x7fffffff
                    : st.loaded < st.npad ? st.loaded - kRefillBelow + 1
                                          : (st.p + 16 <= st.N ? st.N - 15 : st.p);
        __syncwarp();
        if (!__any_sync(kFull, st.active)) break;
        st.fetch_lookahead();
        st.template poll_step<true>(lfull, ext_cap);
    }
}

template <int WBITS, int LPS, bool EXT>
void launch_one(const GroupCompArgs &a, cudaStream_t st) {
    using G = Geo<WBITS, LPS>;
    static int blocks_per_sm = 0;
    static int sms = 0;
    const size_t smem = G::CTA_BYTES;
    if (!blocks_per_sm) {
        cudaFuncSetAttribute(k_group_compress<WBITS, LPS, EXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_group_compress<WBITS, LPS, EXT>,
                                                      kWarpsPerCta * 32, smem);
        if (blocks_per_sm < 1) blocks_per_sm = 1;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const uint64_t per_cta = (uint64_t)kWarpsPerCta * G::SPW;
    const uint64_t want = (a.b.n_streams + per_cta - 1) / per_cta;
    const uint64_t persistent = (uint64_t)sms * blocks_per_sm;  // grid = SM count x resident CTAs
    const unsigned grid = (unsigned)(want < persistent ? want : persistent);
    k_group_compress<WBITS, LPS, EXT><<<grid, kWarpsPerCta * 32, smem, st>>>(a);
    count_launch();
}

}  // namespace

int g_group_lps = 0;  // benchmark hook: lanes per stream for window 10 (0 = default)

bool launch_group_compress_batch(const CompBatchConf &cf, const uint8_t *d_dict, const BatchArgs &b, cudaStream_t st) {
    if (cf.window > 10 || (cf.flags & TB_F_LAZY)) return false;
    if (b.in_offsets) return false;  // strided layout only
    if ((b.in_stride & 15) || ((uintptr_t)b.in & 15) || (b.out_stride & 3) || ((uintptr_t)b.out & 3)) return false;
    if (b.in_stride > (1u << 30)) return false;
    // never OUTPUT_FULL in this kernel: require worst-case room (all literals + header + flush token)
    const uint64_t bound = 2 + (b.in_stride * (uint64_t)(cf.literal + 1) + 7) / 8 + 6;
    if (b.out_stride < ((bound + 3) & ~3ull)) return false;
    if (b.n_streams == 0) return true;

    const int W = 1 << cf.window;
    const bool ext = (cf.flags & TB_F_EXTENDED) != 0;
    GroupCompArgs a;
    a.b = b;
    a.literal = cf.literal;
    a.flags = cf.flags;
    a.write_token = cf.write_token;
    switch (cf.window) {
        case 8:
            a.dictrows = stage_dictrows(d_dict, W, Geo<8, 4>::RS, st);
            if (!a.dictrows) return false;
            ext ? launch_one<8, 4, true>(a, st) : launch_one<8, 4, false>(a, st);
            return true;
        case 9:
            a.dictrows = stage_dictrows(d_dict, W, Geo<9, 4>::RS, st);
            if (!a.dictrows) return false;
            ext ? launch_one<9, 4, true>(a, st) : launch_one<9, 4, false>(a, st);
            return true;
        case 10: {
            const int lps = g_group_lps ? g_group_lps : 8;
            const int rs = lps == 4 ? Geo<10, 4>::RS : lps == 16 ? Geo<10, 16>::RS : lps == 32 ? Geo<10, 32>::RS : Geo<10, 8>::RS;
            a.dictrows = stage_dictrows(d_dict, W, rs, st);
            if (!a.dictrows) return false;
            if (lps == 4)
                ext ? launch_one<10, 4, true>(a, st) : launch_one<10, 4, false>(a, st);
            else if (lps == 16)
                ext ? launch_one<10, 16, true>(a, st) : launch_one<10, 16, false>(a, st);
            else if (lps == 32)
                ext ? launch_one<10, 32, true>(a, st) : launch_one<10, 32, false>(a, st);
            else
                ext ? launch_one<10, 8, true>(a, st) : launch_one<10, 8, false>(a, st);
            return true;
        }
        default: return false;
    }
}

}  // namespace tb
