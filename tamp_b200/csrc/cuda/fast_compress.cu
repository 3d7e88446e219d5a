// Specialised batch compressor: one warp per stream, window <= 1024 bytes, whole streams from init to flush.
//
// What the reference does per token (tamp/_c_src/tamp/compressor.c:532-660): scan the whole W-byte
// window for the longest prefix of the 16-byte input ring (find_best_match,
// compressor_find_match_desktop.c:82-167), emit a literal or a Huffman(len)+index token, append the
// consumed bytes to the window.  The scan is 81-88 % of the reference's time.
//
// B200 formulation.  The window is never stored as bytes.  Instead the warp keeps, in shared memory,
// 32 position bitmaps of W bits each: row h (0..15) has bit x set iff the HIGH nibble of window[x] is
// h, row 16+l iff the LOW nibble is l.  W = 1024 bits = 32 words = one word per lane, so
//
//     E(c)        = row[c >> 4] & row[16 + (c & 15)]             positions holding byte c (1 word/lane)
//     M_0         = E(in[0])
//     M_k         = M_{k-1} & (E(in[k]) >> k)                     positions where k+1 bytes match
//
// and the longest match is the last non-empty M_k, its index the lowest set bit (= the reference's
// lowest-index tie-break and its early exit at the maximum length).  Shifting in zeros past bit W-1
// reproduces "a match never runs past the end of the buffer".  One level costs two conflict-free LDS,
// a shuffle for the neighbour word, a funnel shift and a vote — for ALL W positions at once — instead
// of the reference's byte-by-byte scan.
//
// Window updates are the transpose of that layout: a 32x32 bit-matrix transpose (5 shuffle stages)
// turns the next 32 bytes to be written into one word per row; lane r owns row r and merges the
// block into the window bitmap as tokens consume it (no atomics, no byte window).
// The seeded dictionary is staged per stream as its 32 pre-computed rows with one TMA bulk copy
// (cp.async.bulk, completion on an mbarrier) while the lanes fetch the stream's input with coalesced
// 128-bit loads.  Output bits are accumulated warp-uniformly and leave through a 128-byte staging
// line written back with one coalesced store per 32 words.
//
// Extended format (RLE + extended match, compressor.c:437-525) reuses the same bitmaps: an extended
// match continuation (find_extended_match, :297-333) is simply "keep iterating levels" on the
// candidate set of the match that started it.
//
// Bit-exactness notes: ring fill at poll entry is min(16, N - p) in both the compress and the flush
// phase because compress_cb only polls a full ring (compressor.c:709) — see DESIGN.md.
#include "../tb_wire.h"
#include "tb_cuda.h"
#include "tb_device_common.cuh"
#include "tb_ptx.cuh"

namespace tb {

namespace {

constexpr int kRingBytes = 1024;   // per-warp input ring (power of two)
constexpr int kRingMirror = 32;    // first bytes of the ring repeated behind it: unwrapped 20-byte lookahead reads
constexpr int kWarpsPerCta = 8;   // normal launches
constexpr int kWarpsPerCtaDeferred = 1;  // pick-up pass behind the position-parallel kernel: one-warp CTAs fit beside its CTA

template <int WBITS>
struct Geo {
    static constexpr int W = 1 << WBITS;
    static constexpr int WW = W / 32;          // words per bitmap row
    static constexpr int RS = WW + 1;          // padded row stride (odd => conflict-free column access)
    static constexpr int ROW_BYTES = ((32 * RS * 4) + 15) / 16 * 16;
    // rows | ring + mirror | 32 token records | 32-word bit staging | mbarrier
    static constexpr int OFF_RING = ROW_BYTES;
    static constexpr int OFF_RECS = OFF_RING + kRingBytes + kRingMirror;
    static constexpr int OFF_STAGE = OFF_RECS + 128;
    static constexpr int OFF_MBAR = OFF_STAGE + 128;
    static constexpr int PER_WARP = OFF_MBAR + 16;
};

struct FastCompArgs {
    BatchArgs b;
    const uint32_t *dictrows;
    int literal, flags, write_token;
    int only_deferred;  // skip streams whose out_sizes entry is not kDeferred
    int small_grid;     // one CTA per SM
};

// 32x32 bit-matrix transpose across the warp: on return bit i of lane r == bit r of lane i on entry.
__device__ __forceinline__ uint32_t transpose32(uint32_t x, int lane) {
    uint32_t t;
    t = __shfl_xor_sync(0xffffffffu, x, 16);
    x = (lane & 16) ? ((t >> 16) | (x & 0xFFFF0000u)) : ((x & 0x0000FFFFu) | (t << 16));
    t = __shfl_xor_sync(0xffffffffu, x, 8);
    x = (lane & 8) ? (((t >> 8) & 0x00FF00FFu) | (x & 0xFF00FF00u)) : ((x & 0x00FF00FFu) | ((t & 0x00FF00FFu) << 8));
    t = __shfl_xor_sync(0xffffffffu, x, 4);
    x = (lane & 4) ? (((t >> 4) & 0x0F0F0F0Fu) | (x & 0xF0F0F0F0u)) : ((x & 0x0F0F0F0Fu) | ((t & 0x0F0F0F0Fu) << 4));
    t = __shfl_xor_sync(0xffffffffu, x, 2);
    x = (lane & 2) ? (((t >> 2) & 0x33333333u) | (x & 0xCCCCCCCCu)) : ((x & 0x33333333u) | ((t & 0x33333333u) << 2));
    t = __shfl_xor_sync(0xffffffffu, x, 1);
    x = (lane & 1) ? (((t >> 1) & 0x55555555u) | (x & 0xAAAAAAAAu)) : ((x & 0x55555555u) | ((t & 0x55555555u) << 1));
    return x;
}

// ---- per-stream compressor -----------------------------------------------------------------------

template <int WBITS, bool EXT>
struct Stream {
    using G = Geo<WBITS>;
    static constexpr int WW = G::WW, RS = G::RS, MASK = G::W - 1;

    // shared memory views
    uint32_t *rows;        // [32][RS] nibble bitmaps of the window
    const uint32_t *rowp;  // rows + lane       (column `lane` of every row)
    uint32_t *myrow;       // rows + lane * RS  (row `lane`)
    uint8_t *ring;         // input ring (+ mirror)
    uint32_t *recs;        // 32 token records: bits << 5 | nbits
    uint32_t *stage;       // 32-word bit staging line
    int lane;
    uint32_t lane_valid, nb_mask;

    // input
    const uint8_t *in;
    int N, npad, loaded;

    // window-bitmap maintenance (block = 32 window positions = word `cb` of every row)
    int wpos, cb, blk_src;
    uint32_t old_r, next_r;
    uint32_t last;  // last byte written to the window (RLE reference byte; extended format only)

    // bit output: records are queued per token and packed 32 at a time (warp prefix sum)
    int nrec;        // queued records
    int pend_bits;   // bits already sitting in stage[] below the next record (always < 32)
    uint32_t *out32;
    uint32_t ow;     // words already stored to global

    int lbits, min_pat;

    // parse state
    int p, res;
    int rle, ext_n, ext_pos, ext_start;
    uint32_t ext_set;

    __device__ __forceinline__ uint32_t T(int pos) const { return ring[pos & (kRingBytes - 1)]; }

    // ---- bit output ---------------------------------------------------------------------------------
    // Static-Huffman bit-pack: each token queues (bits, nbits); every 32 tokens the warp prefix-sums the
    // lengths, every lane ORs its token into the MSb-first staging line, and whole words go out with one
    // coalesced store (bit writer of compressor.c:49-75, restated for 32 tokens at once).
    __device__ __forceinline__ void pack_and_store(int count) {
        __syncwarp();
        uint32_t rec = lane < count ? recs[lane] : 0u;
        int n = (int)(rec & 31u);
        uint32_t bits = rec >> 5;
        int incl = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const int total = pend_bits + __shfl_sync(0xffffffffu, incl, 31);
        if (n) {
            const int start = pend_bits + incl - n;  // bit offset of this token in the staging line
            const int w = start >> 5, o = start & 31;
            const uint64_t v = (uint64_t)bits << (64 - n - o);
            atomicOr(&stage[w], (uint32_t)(v >> 32));
            if ((uint32_t)v) atomicOr(&stage[w + 1], (uint32_t)v);
        }
        __syncwarp();
        const int nwords = total >> 5;
        const uint32_t mine = stage[lane];
        __syncwarp();
        if (lane < nwords) out32[ow + lane] = __byte_perm(mine, 0, 0x0123);
        const uint32_t carry = __shfl_sync(0xffffffffu, mine, nwords & 31);
        stage[lane] = (lane == 0 && (total & 31)) ? carry : 0u;
        ow += nwords;
        pend_bits = total & 31;
        nrec = 0;
        __syncwarp();
    }
    __device__ __forceinline__ void put(uint32_t bits, int n) {
        if (lane == 0) recs[nrec] = (bits << 5) | (uint32_t)n;
        if (++nrec == 32) pack_and_store(32);
    }
    __device__ __forceinline__ void put_exthuff(int v, int t) {
        int i = v >> t;
        put(((uint32_t)kHuff.code[i] << t) | (uint32_t)(v & ((1 << t) - 1)), kHuff.bits[i] - 1 + t);
    }

    // ---- bitmap primitives ----------------------------------------------------------------------------
    // E(c): positions of the window that hold byte c, this lane's word.
    __device__ __forceinline__ uint32_t row_of(uint32_t c) const {
        uint32_t e = rowp[(c >> 4) * RS] & rowp[(16 + (c & 15)) * RS];
        if (WW < 32) e &= lane_valid;
        return e;
    }
    // word `lane` of (E >> k), 0 < k < 32
    __device__ __forceinline__ uint32_t shifted_small(uint32_t e, int k) const {
        uint32_t nb = __shfl_down_sync(0xffffffffu, e, 1);
        if (WW == 32) nb &= nb_mask;
        return __funnelshift_r(e, nb, k);
    }
    // word `lane` of (E >> k), any k
    __device__ __forceinline__ uint32_t shifted_any(uint32_t e, int k) const {
        int s = k >> 5, src = lane + s;
        uint32_t a = __shfl_sync(0xffffffffu, e, src & 31);
        uint32_t b = __shfl_sync(0xffffffffu, e, (src + 1) & 31);
        if (src > 31) a = 0;
        if (src + 1 > 31) b = 0;
        return __funnelshift_r(a, b, k & 31);
    }
    __device__ __forceinline__ int lowest_pos(uint32_t m) const {
        uint32_t mine = m ? (uint32_t)(lane * 32 + __ffs(m) - 1) : 0xFFFFu;
        return (int)__reduce_min_sync(0xffffffffu, mine);
    }

    // ---- window update -----------------------------------------------------------------------------------
    // Row bits of input bytes [s, s+32) (bytes at or past N never reach the window: use 0).
    __device__ __forceinline__ uint32_t block_rows(int s) const {
        int q = s + lane;
        uint32_t x = 0;
        if (q < N) {
            uint32_t c = T(q);
            x = (1u << (c >> 4)) | (0x10000u << (c & 15));
        }
        return transpose32(x, lane);
    }

    // Append m input bytes starting at input position s to the window at wpos (destination wraps).
    __device__ __forceinline__ void window_write(int s, int m) {
        if (EXT) {
            if (m <= 0) return;
            last = T(s + m - 1);
        }
        __syncwarp();  // every lane's bitmap reads of this poll precede the row update below
        {
            const int off = wpos & 31, off2 = off + m;
            if (off2 < 32 && (!EXT || blk_src + off == s)) {  // common case: stays inside the pending block
                const uint32_t lm2 = (1u << off2) - 1u;
                myrow[cb] = (next_r & lm2) | (old_r & ~lm2);
                wpos += m;
                __syncwarp();
                return;
            }
        }
        while (m > 0) {
            const int off = wpos & 31;
            if (EXT && blk_src + off != s) {
                // the bytes about to land here are not the ones the pending block was built from
                const uint32_t lm = (1u << off) - 1u;  // off < 32
                old_r = (next_r & lm) | (old_r & ~lm);
                blk_src = s - off;
                next_r = (block_rows(blk_src) & ~lm) | (old_r & lm);
            }
            const int take = m < 32 - off ? m : 32 - off;
            const int off2 = off + take;
            const uint32_t lm2 = off2 == 32 ? 0xffffffffu : ((1u << off2) - 1u);
            myrow[cb] = (next_r & lm2) | (old_r & ~lm2);
            wpos = (wpos + take) & MASK;
            s += take;
            m -= take;
            if (off2 == 32) {
                cb = wpos >> 5;
                blk_src = s;
                old_r = myrow[cb];
                next_r = block_rows(s);
            }
        }
        __syncwarp();
    }

    // ---- find_best_match via bitmap levels -----------------------------------------------------------------
    // in[] = 16 lookahead bytes, L = usable length (only consulted when TAIL; otherwise L == LFULL).
    // Returns len (>= 2) or 0; idx = lowest index; mset = candidate set of the final level.
    template <bool TAIL>
    __device__ __forceinline__ int search(const uint32_t (&in)[4], int L, int lfull, int &idx, uint32_t &mset) const {
        if (TAIL && (L < 2 || L < min_pat)) return 0;
        uint32_t m = row_of(in[0] & 0xFFu) & shifted_small(row_of((in[0] >> 8) & 0xFFu), 1);
        if (!__any_sync(0xffffffffu, m != 0)) return 0;
        int len;
#define TB_LEVEL(K)                                                                   \
    {                                                                                 \
        if (TAIL && (K) >= L) { len = (K); goto found; }                              \
        if ((K) == 15 && lfull < 16) { len = 15; goto found; }                        \
        const uint32_t c = (in[(K) >> 2] >> (8 * ((K) & 3))) & 0xFFu;                 \
        const uint32_t mn = m & shifted_small(row_of(c), (K));                        \
        if (!__any_sync(0xffffffffu, mn != 0)) { len = (K); goto found; }             \
        m = mn;                                                                       \
    }
        TB_LEVEL(2) TB_LEVEL(3) TB_LEVEL(4) TB_LEVEL(5) TB_LEVEL(6) TB_LEVEL(7) TB_LEVEL(8) TB_LEVEL(9)
        TB_LEVEL(10) TB_LEVEL(11) TB_LEVEL(12) TB_LEVEL(13) TB_LEVEL(14) TB_LEVEL(15)
#undef TB_LEVEL
        len = 16;
    found:
        idx = lowest_pos(m);
        mset = m;
        return len;
    }

    __device__ __forceinline__ void put_literal(uint32_t c) { put((1u << lbits) | c, lbits + 1); }
    __device__ __forceinline__ void put_token(int len, int idx) {
        const int h = len - min_pat;
        put(((uint32_t)kHuff.code[h] << WBITS) | (uint32_t)idx, kHuff.bits[h] + WBITS);
    }
    __device__ __forceinline__ void put_ext_match() {  // write_extended_match_token, compressor.c:377-415
        put(kHuff.code[kSymExt], kHuff.bits[kSymExt]);
        put_exthuff(ext_n - min_pat - 12, 3);
        put((uint32_t)ext_pos, WBITS);
    }
    __device__ __forceinline__ void put_rle(int count) {  // write_rle_token, compressor.c:342-350
        put(kHuff.code[kSymRle], kHuff.bits[kSymRle]);
        put_exthuff(count - 2, 4);
    }

    // One tamp_compressor_poll (compressor.c:532-660) at input position p; ring fill = min(16, N - p).
    template <bool TAIL>
    __device__ __forceinline__ void poll(int lfull, int ext_cap) {
        const int r = TAIL ? N - p : 16;
        uint32_t in[4];
        {
            const uint8_t *b = ring + (p & (kRingBytes - 4));
            const uint32_t a0 = *reinterpret_cast<const uint32_t *>(b), a1 = *reinterpret_cast<const uint32_t *>(b + 4),
                           a2 = *reinterpret_cast<const uint32_t *>(b + 8), a3 = *reinterpret_cast<const uint32_t *>(b + 12),
                           a4 = *reinterpret_cast<const uint32_t *>(b + 16);
            const int sh = (p & 3) * 8;
            in[0] = __funnelshift_r(a0, a1, sh);
            in[1] = __funnelshift_r(a1, a2, sh);
            in[2] = __funnelshift_r(a2, a3, sh);
            in[3] = __funnelshift_r(a3, a4, sh);
        }
        const int L = TAIL ? (r < lfull ? r : lfull) : lfull;
        int idx = 0, len = 0;
        uint32_t mset = 0;
        bool have_match = false;

        if (EXT) {
            if (ext_n) {  // extended-match continuation (compressor.c:442-469)
                int avail = r;
                bool emit = false;
                while (avail > 0) {
                    if (ext_pos + ext_n >= G::W || ext_n >= ext_cap) {
                        emit = true;
                        break;
                    }
                    const int maxp = ext_n + avail < ext_cap ? ext_n + avail : ext_cap;
                    int n = ext_n;
                    uint32_t m = ext_set;
                    while (n < maxp) {
                        uint32_t mn = m & shifted_any(row_of(T(ext_start + n)), n);
                        if (!__any_sync(0xffffffffu, mn != 0)) break;
                        m = mn;
                        n++;
                    }
                    if (n > ext_n) {
                        avail -= n - ext_n;
                        p += n - ext_n;
                        ext_pos = lowest_pos(m);
                        ext_set = m;
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
                if (emit) {
                    put_ext_match();
                    const int room = G::W - wpos;
                    window_write(ext_start, ext_n < room ? ext_n : room);
                    ext_n = 0;
                }
                return;
            }
            // RLE accumulation (compressor.c:471-523).  Fast reject: nothing pending and the next byte
            // differs from the last window byte => avail = total = 0 and the whole block is a no-op.
            if (rle != 0 || (in[0] & 0xFFu) == last) {
                int avail = 16;
                {
                    const uint32_t bl = last * 0x01010101u;
#pragma unroll
                    for (int i = 3; i >= 0; i--) {
                        uint32_t x = in[i] ^ bl;
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
                    return;
                }
                if (total >= 2) {
                    bool use_rle = true;
                    if (total == avail && total <= 6) {
                        len = search<TAIL>(in, L, lfull, idx, mset);
                        if (len > total) {
                            use_rle = false;
                            have_match = true;
                            rle = 0;
                        }
                    }
                    if (use_rle) {
                        p += avail;
                        put_rle(total);
                        const int room = G::W - wpos;
                        const int nw = total < kRleWindowMax ? total : kRleWindowMax;
                        window_write(p - total, nw < room ? nw : room);
                        rle = 0;
                        return;
                    }
                } else if (rle == 1) {  // lone run byte from an earlier poll
                    put_literal(last);
                    window_write(p - 1, 1);
                    rle = 0;
                    return;
                }
            }
        }

        if (!have_match) len = search<TAIL>(in, L, lfull, idx, mset);

        if (len < min_pat) {
            const uint32_t c = in[0] & 0xFFu;
            if (c >> lbits) {
                res = kExcessBits;
                p = N;  // ends both poll loops
                return;
            }
            put_literal(c);
            window_write(p, 1);
            p += 1;
        } else if (EXT && len > min_pat + 11) {
            ext_n = len;
            ext_pos = idx;
            ext_start = p;
            ext_set = mset;
            p += len;
        } else {
            put_token(len, idx);
            window_write(p, len);
            p += len;
        }
    }

    __device__ __forceinline__ void refill() {
        int off = loaded + lane * 16;
        if (off < npad) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(in + off));
            const int ro = off & (kRingBytes - 1);
            *reinterpret_cast<uint4 *>(ring + ro) = v;
            if (ro < kRingMirror) *reinterpret_cast<uint4 *>(ring + kRingBytes + ro) = v;
        }
        loaded = loaded + 512 < npad ? loaded + 512 : npad;
        __syncwarp();
    }
};

template <int WBITS, bool EXT, int WPC>
__global__ void __launch_bounds__(WPC * 32) k_fast_compress(FastCompArgs a) {
    using G = Geo<WBITS>;
    using S = Stream<WBITS, EXT>;
#ifndef TB_EMU
    extern __shared__ __align__(128) uint8_t smem[];
#else
    uint8_t *smem = emu::g_smem;
#endif
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *base = smem + (size_t)warp * G::PER_WARP;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(base + G::OFF_MBAR);

    S st;
    st.rows = reinterpret_cast<uint32_t *>(base);
    st.rowp = st.rows + lane;
    st.myrow = st.rows + lane * G::RS;
    st.ring = base + G::OFF_RING;
    st.recs = reinterpret_cast<uint32_t *>(base + G::OFF_RECS);
    st.stage = reinterpret_cast<uint32_t *>(base + G::OFF_STAGE);
    st.lane = lane;
    st.lane_valid = lane < G::WW ? 0xffffffffu : 0u;
    st.nb_mask = lane == 31 ? 0u : 0xffffffffu;
    st.lbits = a.literal;
    st.min_pat = min_pattern_size(WBITS, a.literal);
    const int lfull = EXT ? 16 : st.min_pat + 13;  // MAX_PATTERN_SIZE clipped by the 16-byte ring
    const int ext_cap = st.min_pat + 11 + kExtExtraMax;

    if (lane == 0) mbar_init(mbar, 1);
    __syncwarp();
    uint32_t phase = 0;

    // Normal launches: warp w takes streams w, w + nwarps, ...  Pick-up launches (only_deferred): warp w scans the
    // sizes of streams [32 w, 32 w + 32), [32 (w + nwarps), ...) with one coalesced load per block and compresses
    // the marked ones.
    const uint64_t nwarps = (uint64_t)gridDim.x * WPC;
    const uint64_t span = a.only_deferred ? 32 : 1;
    for (uint64_t base = ((uint64_t)blockIdx.x * WPC + warp) * span; base < a.b.n_streams; base += nwarps * span) {
      uint32_t todo = 1u;
      if (a.only_deferred) {
          const uint64_t s = base + lane;
          todo = __ballot_sync(0xffffffffu, s < a.b.n_streams && a.b.out_sizes[s] == kDeferred);
      }
      while (todo) {
        const uint64_t stream = base + (uint64_t)(__ffs(todo) - 1);
        todo &= todo - 1u;
        // -- stage the dictionary bitmaps (TMA) and the head of the input (coalesced 128-bit loads) --
        __syncwarp();
        if (lane == 0) {
            fence_proxy_async();
            mbar_expect_tx(mbar, G::ROW_BYTES);
            tma_load_1d(st.rows, a.dictrows, G::ROW_BYTES, mbar);
        }
        st.in = a.b.in + stream * a.b.in_stride;
        st.N = a.b.in_sizes ? (int)a.b.in_sizes[stream] : (int)a.b.in_stride;
        st.npad = (st.N + 15) & ~15;
        st.loaded = st.npad < kRingBytes ? st.npad : kRingBytes;
        for (int off = lane * 16; off < st.loaded; off += 512) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(st.in + off));
            *reinterpret_cast<uint4 *>(st.ring + off) = v;
            if (off < kRingMirror) *reinterpret_cast<uint4 *>(st.ring + kRingBytes + off) = v;
        }
        st.stage[lane] = 0;
        mbar_wait(mbar, phase);
        phase ^= 1;
        __syncwarp();

        st.out32 = reinterpret_cast<uint32_t *>(a.b.out + stream * a.b.out_stride);
        st.ow = 0;
        st.nrec = 0;
        st.pend_bits = 0;
        {
            uint32_t header = ((uint32_t)(WBITS - 8) << 5) | ((uint32_t)(a.literal - 5) << 3) |
                              ((a.flags & TB_F_CUSTOM_DICT) ? 4u : 0u) | (EXT ? 2u : 0u) |
                              ((a.flags & TB_F_DICT_RESET) ? 1u : 0u);
            if (stream_appends(a.flags, stream)) {
                st.put(kAppendStart >> 16, 16);
            } else {
                st.put(header, 8);
                if (a.flags & TB_F_DICT_RESET) st.put(0, 8);
            }
        }
        st.wpos = 0;
        st.cb = 0;
        st.blk_src = 0;
        st.old_r = st.myrow[0];
        st.next_r = st.block_rows(0);
        st.last = st.rows[G::WW] & 0xFFu;  // pad word of row 0 carries dictionary[W-1] (k_build_dictrows)
        st.p = 0;
        st.res = kOk;
        st.rle = 0;
        st.ext_n = 0;
        st.ext_pos = 0;
        st.ext_start = 0;
        st.ext_set = 0;
        const int N = st.N;

        // Polls with a full 16-byte lookahead run in stretches that need no ring refill: the ring keeps
        // >= 256 bytes of lookahead and >= 256 bytes of history, so a stretch ends 256 bytes before the loaded
        // frontier (or 16 bytes before the end of the stream once everything is resident).  An error parks p
        // at the end of the stream, so the loops test one condition.
        while (st.p + 16 <= N) {
            const int stop = st.loaded < st.npad ? st.loaded - 256 : N - 16;
            while (st.p <= stop) st.template poll<false>(lfull, ext_cap);
            if (st.loaded < st.npad && st.p + 16 <= N) st.refill();
        }
        while (st.p < N) st.template poll<true>(lfull, ext_cap);

        // -- flush (compressor.c:728-810) ------------------------------------------------------------
        uint32_t out_bytes;
        if (st.res == kOk) {
            if (EXT) {
                if (st.rle == 1)
                    st.put_literal(st.last);
                else if (st.rle >= 2)
                    st.put_rle(st.rle);
                else if (st.ext_n)
                    st.put_ext_match();
            }
            st.pack_and_store(st.nrec);
            if (ends_with_flush(a.write_token, (uint32_t)st.pend_bits, a.flags, stream, (uint64_t)N)) {
                st.put(kHuff.code[kSymFlush], kHuff.bits[kSymFlush]);
                st.pack_and_store(st.nrec);
            }
            out_bytes = st.ow * 4 + ((st.pend_bits + 7) >> 3);
        } else {
            // Error path: the reference has drained whole bytes of everything queued before the failing poll.
            st.pack_and_store(st.nrec);
            out_bytes = st.ow * 4 + (st.pend_bits >> 3);
        }
        {
            const uint32_t tail = out_bytes - st.ow * 4;  // < 4 bytes left in stage[0], MSb first
            const uint32_t w0 = st.stage[0];
            uint8_t *o8 = reinterpret_cast<uint8_t *>(st.out32 + st.ow);
            if ((uint32_t)lane < tail) o8[lane] = (uint8_t)(w0 >> (24 - 8 * lane));
        }
        if (lane == 0) {
            a.b.out_sizes[stream] = out_bytes;
            if (a.b.status) a.b.status[stream] = (int8_t)st.res;
        }
      }
    }
}

// Dictionary bytes -> 32 nibble bitmaps of W bits, rows padded to RS words.
__global__ void k_build_dictrows(const uint8_t *dict, int W, uint32_t *rows_out, int rs, int total_words) {
    for (int t = threadIdx.x; t < total_words; t += blockDim.x) rows_out[t] = 0;
    __syncthreads();
    const int ww = W / 32;
    for (int t = threadIdx.x; t < 32 * ww; t += blockDim.x) {
        int r = t / ww, w = t % ww;
        uint32_t bits = 0;
        for (int i = 0; i < 32; i++) {
            uint32_t c = dict[32 * w + i];
            uint32_t nib = r < 16 ? (c >> 4) : (c & 15u);
            if (nib == (uint32_t)(r & 15)) bits |= 1u << i;
        }
        rows_out[r * rs + w] = bits;
    }
    // RLE's "previous byte" at stream start is the last dictionary byte (specification.rst:219-222);
    // it rides in row 0's padding word, which no lane ever reads as bitmap data.
    if (threadIdx.x == 0) rows_out[ww] = dict[W - 1];
}

#ifndef TB_EMU
// Device scratch for dictionary bitmaps: a small ring of slots so that back-to-back launches on different
// CUDA streams never share one.
constexpr int kDictSlots = 16, kDictSlotBytes = 5120;
uint8_t *g_dictrows = nullptr;
int g_dictslot = 0;
#endif  // TB_EMU

}  // namespace

#ifndef TB_EMU
// Builds the dictionary's nibble bitmaps (row stride rs words, rows padded to a multiple of 16 bytes in
// total) on `st`; returns the device pointer or nullptr.  Bitmaps of dictionaries inside the registered
// static range (the engine's seeded tables, which never change) are built once and kept; others go through
// a small ring of scratch slots, rebuilt per launch.
static const uint8_t *g_static_lo = nullptr, *g_static_hi = nullptr;
void register_static_dictionaries(const uint8_t *lo, size_t bytes) {
    g_static_lo = lo;
    g_static_hi = lo + bytes;
}

const uint32_t *stage_dictrows(const uint8_t *d_dict, int W, int rs, cudaStream_t st) {
    const int total_words = ((32 * rs * 4 + 15) / 16 * 16) / 4;
    if (total_words * 4 > kDictSlotBytes) return nullptr;
    struct Cached {
        const uint8_t *dict;
        int W, rs;
        uint32_t *rows;
    };
    static Cached cache[32];
    static int n_cached = 0;
    const bool is_static = d_dict >= g_static_lo && d_dict < g_static_hi;
    if (is_static) {
        for (int i = 0; i < n_cached; i++)
            if (cache[i].dict == d_dict && cache[i].W == W && cache[i].rs == rs) return cache[i].rows;
        if (n_cached < 32) {
            uint32_t *rows = nullptr;
            if (cudaMalloc(&rows, kDictSlotBytes) == cudaSuccess) {
                k_build_dictrows<<<1, 256, 0, st>>>(d_dict, W, rows, rs, total_words);
                count_launch();
                // later launches may run on other CUDA streams: make the table visible before anyone can use it
                cudaStreamSynchronize(st);
                cache[n_cached++] = Cached{d_dict, W, rs, rows};
                return rows;
            }
            cudaGetLastError();
        }
    }
    if (!g_dictrows && cudaMalloc(&g_dictrows, kDictSlots * kDictSlotBytes) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    uint32_t *slot = reinterpret_cast<uint32_t *>(g_dictrows + (size_t)(g_dictslot++ % kDictSlots) * kDictSlotBytes);
    k_build_dictrows<<<1, 256, 0, st>>>(d_dict, W, slot, rs, total_words);
    count_launch();
    return slot;
}

namespace {

template <int WBITS, bool EXT, int WPC>
void launch_wpc(const FastCompArgs &a, cudaStream_t st) {
    using G = Geo<WBITS>;
    static int blocks_per_sm = 0;
    static int sms = 0;
    const size_t smem = (size_t)G::PER_WARP * WPC;
    if (!blocks_per_sm) {
        cudaFuncSetAttribute(k_fast_compress<WBITS, EXT, WPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_fast_compress<WBITS, EXT, WPC>, WPC * 32, smem);
        if (blocks_per_sm < 1) blocks_per_sm = 1;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    uint64_t want = (a.b.n_streams + WPC - 1) / WPC;
    uint64_t persistent = (uint64_t)sms * (a.small_grid ? 1 : blocks_per_sm);  // grid = SM count x resident CTAs
    unsigned grid = (unsigned)(want < persistent ? want : persistent);
    k_fast_compress<WBITS, EXT, WPC><<<grid, WPC * 32, smem, st>>>(a);
    count_launch();
}

template <int WBITS, bool EXT>
void launch_one(const FastCompArgs &a, cudaStream_t st) {
    if (a.only_deferred)
        launch_wpc<WBITS, EXT, kWarpsPerCtaDeferred>(a, st);
    else
        launch_wpc<WBITS, EXT, kWarpsPerCta>(a, st);
}

}  // namespace

bool launch_fast_compress_batch(const CompBatchConf &cf, const uint8_t *d_dict, const BatchArgs &b, cudaStream_t st,
                                bool only_deferred, bool small_grid) {
    if (cf.window > 10 || (cf.flags & TB_F_LAZY)) return false;
    if (b.in_offsets) return false;  // strided layout only
    if ((b.in_stride & 15) || ((uintptr_t)b.in & 15) || (b.out_stride & 3) || ((uintptr_t)b.out & 3)) return false;
    if (b.in_stride > (1u << 30)) return false;
    // never OUTPUT_FULL in this kernel: require worst-case room (all literals + header + flush token)
    const uint64_t bound = 2 + (b.in_stride * (uint64_t)(cf.literal + 1) + 7) / 8 + 6;
    if (b.out_stride < ((bound + 3) & ~3ull)) return false;
    if (b.n_streams == 0) return true;

    const int W = 1 << cf.window;
    const uint32_t *slot = stage_dictrows(d_dict, W, W / 32 + 1, st);
    if (!slot) return false;

    FastCompArgs a;
    a.b = b;
    a.dictrows = slot;
    a.literal = cf.literal;
    a.flags = cf.flags;
    a.write_token = cf.write_token;
    a.only_deferred = only_deferred ? 1 : 0;
    a.small_grid = small_grid ? 1 : 0;
    switch (cf.window * 2 + ((cf.flags & TB_F_EXTENDED) ? 1 : 0)) {
        case 16: launch_one<8, false>(a, st); break;
        case 17: launch_one<8, true>(a, st); break;
        case 18: launch_one<9, false>(a, st); break;
        case 19: launch_one<9, true>(a, st); break;
        case 20: launch_one<10, false>(a, st); break;
        case 21: launch_one<10, true>(a, st); break;
        default: return false;
    }
    return true;
}

#endif  // TB_EMU

}  // namespace tb
