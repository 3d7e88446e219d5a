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

namespace tb {

namespace {

constexpr int kRingBytes = 1024;  // per-warp input ring (power of two)
constexpr int kWarpsPerCta = 8;

template <int WBITS>
struct Geo {
    static constexpr int W = 1 << WBITS;
    static constexpr int WW = W / 32;          // words per bitmap row
    static constexpr int RS = WW + 1;          // padded row stride (odd => conflict-free column access)
    static constexpr int ROW_BYTES = ((32 * RS * 4) + 15) / 16 * 16;
    static constexpr int PER_WARP = ROW_BYTES + kRingBytes + 128 + 16;  // rows | ring | out line | mbarrier
};

struct FastCompArgs {
    BatchArgs b;
    const uint32_t *dictrows;
    int literal, flags, write_token;
};

// ---- small PTX helpers ---------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

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
    static constexpr int W = G::W, WW = G::WW, RS = G::RS, MASK = G::W - 1;

    // shared memory views
    uint32_t *rows;          // [32][RS]
    const uint32_t *rowp;    // rows + lane            (column `lane` of every row)
    uint32_t *myrow;         // rows + lane * RS       (row `lane`)
    uint8_t *ring;           // input ring
    uint32_t *oline;         // 32-word output staging line
    int lane;
    uint32_t lane_valid, nb_mask;

    // input
    const uint8_t *in;
    int N, loaded;

    // window-bitmap maintenance (block = 32 window positions = word `cb` of every row)
    int wpos, cb, blk_src;
    uint32_t old_r, next_r;
    uint32_t last;  // last byte written to the window

    // bit output
    uint64_t acc;
    int nacc, on;
    uint32_t *out32;
    uint32_t ow;

    int lbits, min_pat;

    __device__ __forceinline__ uint32_t T(int pos) const { return ring[pos & (kRingBytes - 1)]; }

    __device__ __forceinline__ void put(uint32_t bits, int n) {
        acc |= (uint64_t)bits << (64 - nacc - n);
        nacc += n;
        if (nacc >= 32) {
            if (lane == 0) oline[on] = __byte_perm((uint32_t)(acc >> 32), 0, 0x0123);  // MSb-first byte order
            acc <<= 32;
            nacc -= 32;
            on++;
            if (on == 32) {
                __syncwarp();
                out32[ow + lane] = oline[lane];
                __syncwarp();
                ow += 32;
                on = 0;
            }
        }
    }
    __device__ __forceinline__ void put_exthuff(int v, int t) {
        int i = v >> t;
        put(((uint32_t)kHuff.code[i] << t) | (uint32_t)(v & ((1 << t) - 1)), kHuff.bits[i] - 1 + t);
    }

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
        uint32_t who = __ballot_sync(0xffffffffu, m != 0);
        int src = __ffs(who) - 1;
        uint32_t mm = __shfl_sync(0xffffffffu, m, src);
        return src * 32 + __ffs(mm) - 1;
    }

    // Block machinery ------------------------------------------------------------------------------
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
        if (m <= 0) return;
        last = T(s + m - 1);
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

    // find_best_match via bitmap levels.  in[] = 16 lookahead bytes, L = usable length.
    // Returns len (>= 2) or 0; idx = lowest index; mset = candidate set of the final level.
    __device__ __forceinline__ int search(const uint32_t (&in)[4], int L, int &idx, uint32_t &mset) const {
        if (L < 2 || L < min_pat) return 0;
        uint32_t m = row_of(in[0] & 0xFFu);
        int len = 1;
#pragma unroll
        for (int k = 1; k < 16; k++) {
            if (k >= L) break;
            uint32_t c = (in[k >> 2] >> (8 * (k & 3))) & 0xFFu;
            uint32_t mn = m & shifted_small(row_of(c), k);
            if (!__any_sync(0xffffffffu, mn != 0)) break;
            m = mn;
            len = k + 1;
        }
        if (len < 2) return 0;
        idx = lowest_pos(m);
        mset = m;
        return len;
    }
};

template <int WBITS, bool EXT>
__global__ void __launch_bounds__(kWarpsPerCta * 32) k_fast_compress(FastCompArgs a) {
    using G = Geo<WBITS>;
    using S = Stream<WBITS, EXT>;
    constexpr int W = G::W;
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *base = smem + (size_t)warp * G::PER_WARP;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(base + G::ROW_BYTES + kRingBytes + 128);

    S st;
    st.rows = reinterpret_cast<uint32_t *>(base);
    st.rowp = st.rows + lane;
    st.myrow = st.rows + lane * G::RS;
    st.ring = base + G::ROW_BYTES;
    st.oline = reinterpret_cast<uint32_t *>(base + G::ROW_BYTES + kRingBytes);
    st.lane = lane;
    st.lane_valid = lane < G::WW ? 0xffffffffu : 0u;
    st.nb_mask = lane == 31 ? 0u : 0xffffffffu;
    st.lbits = a.literal;
    st.min_pat = min_pattern_size(WBITS, a.literal);
    const int cap = EXT ? 16 : st.min_pat + 13;  // MAX_PATTERN_SIZE clipped by the 16-byte ring
    const int ext_cap = st.min_pat + 11 + kExtExtraMax;

    if (lane == 0) mbar_init(mbar, 1);
    __syncwarp();
    uint32_t phase = 0;

    const uint64_t nwarps = (uint64_t)gridDim.x * kWarpsPerCta;
    for (uint64_t stream = (uint64_t)blockIdx.x * kWarpsPerCta + warp; stream < a.b.n_streams; stream += nwarps) {
        // -- stage the dictionary bitmaps (TMA) and the head of the input (coalesced 128-bit loads) --
        __syncwarp();
        if (lane == 0) {
            fence_proxy_async();
            mbar_expect_tx(mbar, G::ROW_BYTES);
            tma_load_1d(st.rows, a.dictrows, G::ROW_BYTES, mbar);
        }
        st.in = a.b.in + stream * a.b.in_stride;
        st.N = a.b.in_sizes ? (int)a.b.in_sizes[stream] : (int)a.b.in_stride;
        const int npad = (st.N + 15) & ~15;
        st.loaded = npad < kRingBytes ? npad : kRingBytes;
        for (int off = lane * 16; off < st.loaded; off += 512)
            *reinterpret_cast<uint4 *>(st.ring + off) = __ldg(reinterpret_cast<const uint4 *>(st.in + off));
        mbar_wait(mbar, phase);
        phase ^= 1;
        __syncwarp();

        st.out32 = reinterpret_cast<uint32_t *>(a.b.out + stream * a.b.out_stride);
        st.ow = 0;
        st.on = 0;
        st.acc = 0;
        st.nacc = 0;
        {
            uint32_t header = ((uint32_t)(WBITS - 8) << 5) | ((uint32_t)(a.literal - 5) << 3) |
                              ((a.flags & TB_F_CUSTOM_DICT) ? 4u : 0u) | (EXT ? 2u : 0u) |
                              ((a.flags & TB_F_DICT_RESET) ? 1u : 0u);
            st.put(header, 8);
            if (a.flags & TB_F_DICT_RESET) st.put(0, 8);
        }
        st.wpos = 0;
        st.cb = 0;
        st.blk_src = 0;
        st.old_r = st.myrow[0];
        st.next_r = st.block_rows(0);
        st.last = st.rows[G::WW] & 0xFFu;  // pad word of row 0 carries dictionary[W-1] (k_build_dictrows)

        int p = 0;
        int res = kOk;
        int rle = 0, ext_n = 0, ext_pos = 0, ext_start = 0;
        uint32_t ext_set = 0;
        const int N = st.N;

        while (p < N) {
            // keep >= 256 bytes of lookahead and >= 256 bytes of history in the ring
            if (p + 256 > st.loaded && st.loaded < npad) {
                int off = st.loaded + lane * 16;
                if (off < npad)
                    *reinterpret_cast<uint4 *>(st.ring + (off & (kRingBytes - 1))) =
                        __ldg(reinterpret_cast<const uint4 *>(st.in + off));
                st.loaded = st.loaded + 512 < npad ? st.loaded + 512 : npad;
                __syncwarp();
            }
            const int r = N - p < 16 ? N - p : 16;  // ring fill at poll entry
            // 16 lookahead bytes, warp-uniform
            uint32_t in[4];
            {
                const uint32_t *r32 = reinterpret_cast<const uint32_t *>(st.ring);
                const int w0 = p >> 2, sh = (p & 3) * 8;
                uint32_t a0 = r32[w0 & 255], a1 = r32[(w0 + 1) & 255], a2 = r32[(w0 + 2) & 255],
                         a3 = r32[(w0 + 3) & 255], a4 = r32[(w0 + 4) & 255];
                in[0] = __funnelshift_r(a0, a1, sh);
                in[1] = __funnelshift_r(a1, a2, sh);
                in[2] = __funnelshift_r(a2, a3, sh);
                in[3] = __funnelshift_r(a3, a4, sh);
            }

            int idx = 0, len = 0;
            uint32_t mset = 0;
            bool have_match = false;

            if (EXT) {
                if (ext_n) {  // extended-match continuation (compressor.c:442-469)
                    int avail = r;
                    bool emit = false;
                    while (avail > 0) {
                        if (ext_pos + ext_n >= W || ext_n >= ext_cap) {
                            emit = true;
                            break;
                        }
                        int maxp = ext_n + avail < ext_cap ? ext_n + avail : ext_cap;
                        int n = ext_n;
                        uint32_t m = ext_set;
                        while (n < maxp) {
                            uint32_t mn = m & st.shifted_any(st.row_of(st.T(ext_start + n)), n);
                            if (!__any_sync(0xffffffffu, mn != 0)) break;
                            m = mn;
                            n++;
                        }
                        if (n > ext_n) {
                            avail -= n - ext_n;
                            p += n - ext_n;
                            ext_pos = st.lowest_pos(m);
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
                    if (emit) {  // write_extended_match_token, compressor.c:377-415
                        st.put(kHuff.code[kSymExt], kHuff.bits[kSymExt]);
                        st.put_exthuff(ext_n - st.min_pat - 12, 3);
                        st.put((uint32_t)ext_pos, WBITS);
                        int room = W - st.wpos;
                        st.window_write(ext_start, ext_n < room ? ext_n : room);
                        ext_n = 0;
                    }
                    continue;
                }
                // RLE accumulation (compressor.c:471-523)
                int avail = 0;
                {
                    const uint32_t bl = st.last * 0x01010101u;
                    avail = 16;
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
                    continue;
                }
                if (total >= 2) {
                    bool use_rle = true;
                    if (total == avail && total <= 6) {
                        len = st.search(in, r < cap ? r : cap, idx, mset);
                        if (len > total) {
                            use_rle = false;
                            have_match = true;
                            rle = 0;
                        }
                    }
                    if (use_rle) {  // write_rle_token, compressor.c:342-359
                        p += avail;
                        st.put(kHuff.code[kSymRle], kHuff.bits[kSymRle]);
                        st.put_exthuff(total - 2, 4);
                        int room = W - st.wpos;
                        int nw = total < kRleWindowMax ? total : kRleWindowMax;
                        st.window_write(p - total, nw < room ? nw : room);
                        rle = 0;
                        continue;
                    }
                } else if (rle == 1) {  // lone run byte from an earlier poll
                    st.put((1u << st.lbits) | st.last, st.lbits + 1);
                    st.window_write(p - 1, 1);
                    rle = 0;
                    continue;
                }
            }

            if (!have_match) len = st.search(in, r < cap ? r : cap, idx, mset);

            if (len < st.min_pat) {
                const uint32_t c = in[0] & 0xFFu;
                if (c >> st.lbits) {
                    res = kExcessBits;
                    break;
                }
                st.put((1u << st.lbits) | c, st.lbits + 1);
                st.window_write(p, 1);
                p += 1;
            } else if (EXT && len > st.min_pat + 11) {
                ext_n = len;
                ext_pos = idx;
                ext_start = p;
                ext_set = mset;
                p += len;
            } else {
                const int h = len - st.min_pat;
                st.put(((uint32_t)kHuff.code[h] << WBITS) | (uint32_t)idx, kHuff.bits[h] + WBITS);
                st.window_write(p, len);
                p += len;
            }
        }

        // -- flush (compressor.c:728-810) ------------------------------------------------------------
        uint32_t out_bytes;
        if (res == kOk) {
            if (EXT) {
                if (rle == 1) {
                    st.put((1u << st.lbits) | st.last, st.lbits + 1);
                } else if (rle >= 2) {
                    st.put(kHuff.code[kSymRle], kHuff.bits[kSymRle]);
                    st.put_exthuff(rle - 2, 4);
                } else if (ext_n) {
                    st.put(kHuff.code[kSymExt], kHuff.bits[kSymExt]);
                    st.put_exthuff(ext_n - st.min_pat - 12, 3);
                    st.put((uint32_t)ext_pos, WBITS);
                }
            }
            if (a.write_token && ((st.nacc & 7) || (a.flags & TB_F_DICT_RESET)))
                st.put(kHuff.code[kSymFlush], kHuff.bits[kSymFlush]);
            out_bytes = (st.ow + st.on) * 4 + ((st.nacc + 7) >> 3);
        } else {
            // Error path: the reference has drained whole bytes of everything queued before the failing poll.
            out_bytes = (st.ow + st.on) * 4 + (st.nacc >> 3);
        }
        __syncwarp();
        if (lane < st.on) st.out32[st.ow + lane] = st.oline[lane];
        {
            const uint32_t tail = out_bytes - (st.ow + st.on) * 4;
            uint8_t *o8 = reinterpret_cast<uint8_t *>(st.out32 + st.ow + st.on);
            if ((uint32_t)lane < tail) o8[lane] = (uint8_t)(st.acc >> (56 - 8 * lane));
        }
        if (lane == 0) {
            a.b.out_sizes[stream] = out_bytes;
            if (a.b.status) a.b.status[stream] = (int8_t)res;
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

// Device scratch for dictionary bitmaps: a small ring of slots so that back-to-back launches on different
// CUDA streams never share one.
constexpr int kDictSlots = 16, kDictSlotBytes = 4352;
uint8_t *g_dictrows = nullptr;
int g_dictslot = 0;

template <int WBITS, bool EXT>
void launch_one(const FastCompArgs &a, cudaStream_t st) {
    using G = Geo<WBITS>;
    static int blocks_per_sm = 0;
    static int sms = 0;
    const size_t smem = (size_t)G::PER_WARP * kWarpsPerCta;
    if (!blocks_per_sm) {
        cudaFuncSetAttribute(k_fast_compress<WBITS, EXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_fast_compress<WBITS, EXT>, kWarpsPerCta * 32,
                                                      smem);
        if (blocks_per_sm < 1) blocks_per_sm = 1;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    uint64_t want = (a.b.n_streams + kWarpsPerCta - 1) / kWarpsPerCta;
    uint64_t persistent = (uint64_t)sms * blocks_per_sm;  // grid = SM count x resident CTAs
    unsigned grid = (unsigned)(want < persistent ? want : persistent);
    k_fast_compress<WBITS, EXT><<<grid, kWarpsPerCta * 32, smem, st>>>(a);
    count_launch();
}

}  // namespace

bool launch_fast_compress_batch(const CompBatchConf &cf, const uint8_t *d_dict, const BatchArgs &b, cudaStream_t st) {
    if (cf.window > 10 || (cf.flags & TB_F_LAZY)) return false;
    if (b.in_offsets) return false;  // strided layout only
    if ((b.in_stride & 15) || ((uintptr_t)b.in & 15) || (b.out_stride & 3) || ((uintptr_t)b.out & 3)) return false;
    if (b.in_stride > (1u << 30)) return false;
    // never OUTPUT_FULL in this kernel: require worst-case room (all literals + header + flush token)
    const uint64_t bound = 2 + (b.in_stride * (uint64_t)(cf.literal + 1) + 7) / 8 + 6;
    if (b.out_stride < ((bound + 3) & ~3ull)) return false;
    if (b.n_streams == 0) return true;

    const int W = 1 << cf.window, rs = W / 32 + 1, total_words = ((32 * rs * 4 + 15) / 16 * 16) / 4;
    if (!g_dictrows && cudaMalloc(&g_dictrows, kDictSlots * kDictSlotBytes) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    uint32_t *slot = reinterpret_cast<uint32_t *>(g_dictrows + (size_t)(g_dictslot++ % kDictSlots) * kDictSlotBytes);
    k_build_dictrows<<<1, 1024, 0, st>>>(d_dict, W, slot, rs, total_words);
    count_launch();

    FastCompArgs a;
    a.b = b;
    a.dictrows = slot;
    a.literal = cf.literal;
    a.flags = cf.flags;
    a.write_token = cf.write_token;
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

}  // namespace tb
