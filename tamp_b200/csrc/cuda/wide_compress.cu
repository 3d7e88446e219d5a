// Specialised batch compressor for windows of 2 KiB .. 32 KiB (window bits 11..15): one CTA per stream.
//
// Same formulation as fast_compress.cu (nibble bitmaps of the window, AND/shift levels, bit-matrix
// transpose for window updates, TMA-staged dictionary rows, warp prefix-sum bit pack) scaled to rows of
// W/32 = 64 .. 1024 words: every thread owns WPL consecutive words of each row (128-bit LDS), a level is
//   E = row[hi] & row[lo]   (WPL words)        M &= (E >> k)   (funnel shifts; neighbour word by shuffle,
//                                                               warp-edge word re-read from shared memory)
// and "is any candidate left" is a CTA-wide OR (__syncthreads_or; a plain warp vote when one warp suffices).
// window = 15 keeps 32 x 4 KiB of bitmaps (128 KiB) in shared memory, one stream per SM, 8 warps scanning.
//
// Reference behaviour restated: tamp/_c_src/tamp/compressor.c:297-333, :342-415, :437-660, :728-810 and
// compressor_find_match_desktop.c:82-167 (see fast_compress.cu for the line-by-line notes).
#include "../tb_wire.h"
#include "tb_cuda.h"
#include "tb_device_common.cuh"
#include "tb_ptx.cuh"

namespace tb {

namespace {

constexpr int kRingBytes = 1024;
constexpr int kRingMirror = 32;

template <int WBITS, int NWARPS>
struct WGeo {
    static constexpr int W = 1 << WBITS;
    static constexpr int WW = W / 32;
    static constexpr int T = NWARPS * 32;
    static constexpr int WPL = WW / T;
    static constexpr int RS = WW + 4;  // row stride in words: keeps rows 16-byte aligned; pad words stay zero
    static constexpr int ROW_BYTES = 32 * RS * 4;
    static constexpr int OFF_RING = ROW_BYTES;
    static constexpr int OFF_RECS = OFF_RING + kRingBytes + kRingMirror;
    static constexpr int OFF_STAGE = OFF_RECS + 128;
    static constexpr int OFF_EROW = OFF_STAGE + 128;       // one bitmap row of scratch (extended-match levels)
    static constexpr int OFF_RED = OFF_EROW + (WW + 8) * 4;  // cross-warp reduction slots
    static constexpr int OFF_MBAR = OFF_RED + 64;
    static constexpr int SMEM = OFF_MBAR + 16;
    static_assert(WPL >= 2 && WPL <= 4, "thread owns 2..4 words of each row");
};

struct WideCompArgs {
    BatchArgs b;
    const uint32_t *dictrows;
    int literal, flags, write_token;
    int only_deferred;  // pick-up pass: just the streams whose out_sizes entry is kDeferred (left by hwalk_compress.cu)
};

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

template <int WBITS, bool EXT, int NWARPS>
struct WStream {
    using G = WGeo<WBITS, NWARPS>;
    static constexpr int WW = G::WW, RS = G::RS, WPL = G::WPL, MASK = G::W - 1;

    uint32_t *rows;
    const uint32_t *mycols;  // rows + tid * WPL   (this thread's words of every row)
    uint32_t *myrow;         // rows + lane * RS   (warp 0: row `lane`)
    uint8_t *ring;
    uint32_t *recs, *stage, *erow, *red;
    int tid, lane, warp;

    const uint8_t *in;
    int N, npad, loaded;

    int wpos, cb, blk_src;
    uint32_t old_r, next_r;  // warp 0 only
    uint32_t last;

    int nrec, pend_bits;  // pend_bits / ow are maintained by warp 0 only
    uint32_t *out32;
    uint32_t ow;

    int lbits, min_pat;
    int p, res;
    int rle, ext_n, ext_pos, ext_start;
    uint32_t ext_set[WPL];

    __device__ __forceinline__ void cta_sync() const {
        if (NWARPS == 1)
            __syncwarp();
        else
            __syncthreads();
    }
    __device__ __forceinline__ bool cta_any(bool pred) const {
        if (NWARPS == 1) return __any_sync(0xffffffffu, pred);
        return __syncthreads_or(pred ? 1 : 0) != 0;
    }
    __device__ __forceinline__ uint32_t T(int pos) const { return ring[pos & (kRingBytes - 1)]; }

    // ---- bit output (warp 0 packs; see fast_compress.cu) ---------------------------------------------
    __device__ __forceinline__ void pack_and_store(int count) {
        if (warp == 0) {
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
                const int start = pend_bits + incl - n;
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
            __syncwarp();
        }
        nrec = 0;
    }
    // Records carry up to 27 payload bits; tokens of wide windows can be longer (9 + 15 bits fits, the
    // 18-bit extended header and the 15-bit position are queued separately).
    __device__ __forceinline__ void put(uint32_t bits, int n) {
        if (tid == 0) recs[nrec] = (bits << 5) | (uint32_t)n;
        if (++nrec == 32) pack_and_store(32);
    }
    __device__ __forceinline__ void put_exthuff(int v, int t) {
        int i = v >> t;
        put(((uint32_t)kHuff.code[i] << t) | (uint32_t)(v & ((1 << t) - 1)), kHuff.bits[i] - 1 + t);
    }

    // ---- bitmap primitives ------------------------------------------------------------------------------
    __device__ __forceinline__ void row_of(uint32_t c, uint32_t (&e)[WPL]) const {
        const uint32_t *h = mycols + (c >> 4) * RS, *l = mycols + (16 + (c & 15)) * RS;
        if (WPL == 4) {
            const uint4 a = *reinterpret_cast<const uint4 *>(h), b = *reinterpret_cast<const uint4 *>(l);
            e[0] = a.x & b.x;
            e[1] = a.y & b.y;
            e[WPL - 2] = a.z & b.z;
            e[WPL - 1] = a.w & b.w;
        } else {
            const uint2 a = *reinterpret_cast<const uint2 *>(h), b = *reinterpret_cast<const uint2 *>(l);
            e[0] = a.x & b.x;
            e[1] = a.y & b.y;
        }
    }
    // first word of the NEXT thread's E(c): shuffle inside the warp, shared-memory re-read at the warp edge
    __device__ __forceinline__ uint32_t next_word_of(uint32_t c, uint32_t e0) const {
        uint32_t nb = __shfl_down_sync(0xffffffffu, e0, 1);
        if (lane == 31) nb = mycols[(c >> 4) * RS + WPL] & mycols[(16 + (c & 15)) * RS + WPL];  // pad words are zero
        return nb;
    }
    // m &= E(c) >> k for 0 < k < 32; returns whether this thread still has candidates
    __device__ __forceinline__ bool level_small(uint32_t c, int k, uint32_t (&m)[WPL]) const {
        uint32_t e[WPL];
        row_of(c, e);
        const uint32_t nb = next_word_of(c, e[0]);
        uint32_t any = 0;
#pragma unroll
        for (int j = 0; j < WPL; j++) {
            const uint32_t hi = j + 1 < WPL ? e[j + 1 < WPL ? j + 1 : 0] : nb;
            m[j] &= __funnelshift_r(e[j], hi, k);
            any |= m[j];
        }
        return any != 0;
    }
    // m &= E(c) >> k for any k (extended-match continuation): E goes through a scratch row in shared memory
    __device__ __forceinline__ bool level_any(uint32_t c, int k, uint32_t (&m)[WPL]) {
        uint32_t e[WPL];
        row_of(c, e);
#pragma unroll
        for (int j = 0; j < WPL; j++) erow[tid * WPL + j] = e[j];
        cta_sync();
        const int s = k >> 5, sh = k & 31;
        uint32_t any = 0;
#pragma unroll
        for (int j = 0; j < WPL; j++) {
            const int w = tid * WPL + j + s;
            const uint32_t a = w < WW ? erow[w] : 0u, b = w + 1 < WW ? erow[w + 1] : 0u;
            m[j] &= __funnelshift_r(a, b, sh);
            any |= m[j];
        }
        cta_sync();
        return any != 0;
    }
    __device__ __forceinline__ int lowest_pos(const uint32_t (&m)[WPL]) {
        uint32_t mine = 0xFFFFu;
#pragma unroll
        for (int j = WPL - 1; j >= 0; j--)
            if (m[j]) mine = (uint32_t)((tid * WPL + j) * 32 + __ffs(m[j]) - 1);
        mine = __reduce_min_sync(0xffffffffu, mine);
        if (NWARPS == 1) return (int)mine;
        if (lane == 0) red[warp] = mine;
        __syncthreads();
        uint32_t best = red[0];
#pragma unroll
        for (int w = 1; w < NWARPS; w++) best = min(best, red[w]);
        return (int)best;
    }

    // ---- window update (warp 0 owns the rows' current block) -------------------------------------------------
    __device__ __forceinline__ uint32_t block_rows(int s) const {
        int q = s + lane;
        uint32_t x = 0;
        if (q < N) {
            uint32_t c = T(q);
            x = (1u << (c >> 4)) | (0x10000u << (c & 15));
        }
        return transpose32(x, lane);
    }
    __device__ __forceinline__ void window_write(int s, int m) {
        if (m <= 0) return;
        if (EXT) last = T(s + m - 1);
        cta_sync();  // every thread's bitmap reads of this poll precede the row update below
        while (m > 0) {
            const int off = wpos & 31;
            const int take = m < 32 - off ? m : 32 - off;
            const int off2 = off + take;
            if (warp == 0) {
                if (EXT && blk_src + off != s) {
                    const uint32_t lm = (1u << off) - 1u;
                    old_r = (next_r & lm) | (old_r & ~lm);
                    next_r = (block_rows(s - off) & ~lm) | (old_r & lm);
                }
                const uint32_t lm2 = off2 == 32 ? 0xffffffffu : ((1u << off2) - 1u);
                myrow[cb] = (next_r & lm2) | (old_r & ~lm2);
            }
            if (EXT && blk_src + off != s) blk_src = s - off;
            wpos = (wpos + take) & MASK;
            s += take;
            m -= take;
            if (off2 == 32) {
                cb = wpos >> 5;
                blk_src = s;
                if (warp == 0) {
                    old_r = myrow[cb];
                    next_r = block_rows(s);
                }
            }
        }
        cta_sync();
    }

    // ---- find_best_match ---------------------------------------------------------------------------------------
    template <bool TAIL>
    __device__ __forceinline__ int search(const uint32_t (&in)[4], int L, int lfull, int &idx, uint32_t (&mset)[WPL]) {
        if (TAIL && (L < 2 || L < min_pat)) return 0;
        uint32_t m[WPL];
        row_of(in[0] & 0xFFu, m);
        if (!cta_any(level_small((in[0] >> 8) & 0xFFu, 1, m))) return 0;
        int len = 16;
#pragma unroll
        for (int k = 2; k < 16; k++) {
            if (TAIL && k >= L) {
                len = k;
                break;
            }
            if (k == 15 && lfull < 16) {
                len = 15;
                break;
            }
            uint32_t mn[WPL];
#pragma unroll
            for (int j = 0; j < WPL; j++) mn[j] = m[j];
            const uint32_t c = (in[k >> 2] >> (8 * (k & 3))) & 0xFFu;
            if (!cta_any(level_small(c, k, mn))) {
                len = k;
                break;
            }
#pragma unroll
            for (int j = 0; j < WPL; j++) m[j] = mn[j];
        }
        idx = lowest_pos(m);
#pragma unroll
        for (int j = 0; j < WPL; j++) mset[j] = m[j];
        return len;
    }

    __device__ __forceinline__ void put_literal(uint32_t c) { put((1u << lbits) | c, lbits + 1); }
    __device__ __forceinline__ void put_token(int len, int idx) {
        const int h = len - min_pat;
        put(((uint32_t)kHuff.code[h] << WBITS) | (uint32_t)idx, kHuff.bits[h] + WBITS);
    }
    __device__ __forceinline__ void put_ext_match() {
        put(kHuff.code[kSymExt], kHuff.bits[kSymExt]);
        put_exthuff(ext_n - min_pat - 12, 3);
        put((uint32_t)ext_pos, WBITS);
    }
    __device__ __forceinline__ void put_rle(int count) {
        put(kHuff.code[kSymRle], kHuff.bits[kSymRle]);
        put_exthuff(count - 2, 4);
    }

    __device__ __forceinline__ void refill() {
        int off = loaded + tid * 16;
        if (tid < 32 && off < npad) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(in + off));
            const int ro = off & (kRingBytes - 1);
            *reinterpret_cast<uint4 *>(ring + ro) = v;
            if (ro < kRingMirror) *reinterpret_cast<uint4 *>(ring + kRingBytes + ro) = v;
        }
        loaded = loaded + 512 < npad ? loaded + 512 : npad;
        cta_sync();
    }

    template <bool TAIL>
    __device__ __forceinline__ void poll(int lfull, int ext_cap) {
        if (p + 256 > loaded && loaded < npad) refill();
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
        uint32_t mset[WPL];
#pragma unroll
        for (int j = 0; j < WPL; j++) mset[j] = 0;
        bool have_match = false;

        if (EXT) {
            if (ext_n) {
                int avail = r;
                bool emit = false;
                while (avail > 0) {
                    if (ext_pos + ext_n >= G::W || ext_n >= ext_cap) {
                        emit = true;
                        break;
                    }
                    const int maxp = ext_n + avail < ext_cap ? ext_n + avail : ext_cap;
                    int n = ext_n;
                    uint32_t m[WPL];
#pragma unroll
                    for (int j = 0; j < WPL; j++) m[j] = ext_set[j];
                    while (n < maxp) {
                        uint32_t mn[WPL];
#pragma unroll
                        for (int j = 0; j < WPL; j++) mn[j] = m[j];
                        if (!cta_any(level_any(T(ext_start + n), n, mn))) break;
#pragma unroll
                        for (int j = 0; j < WPL; j++) m[j] = mn[j];
                        n++;
                    }
                    if (n > ext_n) {
                        avail -= n - ext_n;
                        p += n - ext_n;
                        ext_pos = lowest_pos(m);
#pragma unroll
                        for (int j = 0; j < WPL; j++) ext_set[j] = m[j];
                        const bool stopped_early = n < maxp;
                        ext_n = n;
                        if (stopped_early && avail > 0) {
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
                } else if (rle == 1) {
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
                return;
            }
            put_literal(c);
            window_write(p, 1);
            p += 1;
        } else if (EXT && len > min_pat + 11) {
            ext_n = len;
            ext_pos = idx;
            ext_start = p;
#pragma unroll
            for (int j = 0; j < WPL; j++) ext_set[j] = mset[j];
            p += len;
        } else {
            put_token(len, idx);
            window_write(p, len);
            p += len;
        }
    }
};

template <int WBITS, bool EXT, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32) k_wide_compress(WideCompArgs a) {
    using G = WGeo<WBITS, NWARPS>;
    using S = WStream<WBITS, EXT, NWARPS>;
#ifndef TB_EMU
    extern __shared__ __align__(128) uint8_t smem[];
#else  // tests/emu: the kernel stepped on the CPU (test infrastructure; see tests/emu/cuda_emu.h)
    uint8_t *smem = emu::g_smem;
#endif
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + G::OFF_MBAR);

    S st;
    st.tid = threadIdx.x;
    st.lane = threadIdx.x & 31;
    st.warp = threadIdx.x >> 5;
    st.rows = reinterpret_cast<uint32_t *>(smem);
    st.mycols = st.rows + st.tid * G::WPL;
    st.myrow = st.rows + st.lane * G::RS;
    st.ring = smem + G::OFF_RING;
    st.recs = reinterpret_cast<uint32_t *>(smem + G::OFF_RECS);
    st.stage = reinterpret_cast<uint32_t *>(smem + G::OFF_STAGE);
    st.erow = reinterpret_cast<uint32_t *>(smem + G::OFF_EROW);
    st.red = reinterpret_cast<uint32_t *>(smem + G::OFF_RED);
    st.lbits = a.literal;
    st.min_pat = min_pattern_size(WBITS, a.literal);
    const int lfull = EXT ? 16 : st.min_pat + 13;
    const int ext_cap = st.min_pat + 11 + kExtExtraMax;

    if (st.tid == 0) mbar_init(mbar, 1);
    st.cta_sync();
    uint32_t phase = 0;

    for (uint64_t stream = blockIdx.x; stream < a.b.n_streams; stream += gridDim.x) {
        if (a.only_deferred && a.b.out_sizes[stream] != kDeferred) continue;  // (the same word for the whole CTA)
        st.cta_sync();
        if (st.tid == 0) {
            fence_proxy_async();
            mbar_expect_tx(mbar, G::ROW_BYTES);
            tma_load_1d(st.rows, a.dictrows, G::ROW_BYTES, mbar);
        }
        st.in = a.b.in + stream * a.b.in_stride;
        st.N = a.b.in_sizes ? (int)a.b.in_sizes[stream] : (int)a.b.in_stride;
        st.npad = (st.N + 15) & ~15;
        st.loaded = st.npad < kRingBytes ? st.npad : kRingBytes;
        for (int off = st.tid * 16; off < st.loaded; off += G::T * 16) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(st.in + off));
            *reinterpret_cast<uint4 *>(st.ring + off) = v;
            if (off < kRingMirror) *reinterpret_cast<uint4 *>(st.ring + kRingBytes + off) = v;
        }
        if (st.tid < 32) st.stage[st.tid] = 0;
        mbar_wait(mbar, phase);
        phase ^= 1;
        st.cta_sync();

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
        st.old_r = 0;
        st.next_r = 0;
        if (st.warp == 0) {
            st.old_r = st.myrow[0];
            st.next_r = st.block_rows(0);
        }
        st.last = st.rows[G::WW + 1] & 0xFFu;  // pad word WW+1 of row 0 carries dictionary[W-1]
        st.p = 0;
        st.res = kOk;
        st.rle = 0;
        st.ext_n = 0;
        st.ext_pos = 0;
        st.ext_start = 0;
#pragma unroll
        for (int j = 0; j < G::WPL; j++) st.ext_set[j] = 0;
        const int N = st.N;

        while (st.p + 16 <= N && st.res == kOk) st.template poll<false>(lfull, ext_cap);
        while (st.p < N && st.res == kOk) st.template poll<true>(lfull, ext_cap);

        // flush (compressor.c:728-810); bit-output state lives in warp 0
        if (st.res == kOk && EXT) {
            if (st.rle == 1)
                st.put_literal(st.last);
            else if (st.rle >= 2)
                st.put_rle(st.rle);
            else if (st.ext_n)
                st.put_ext_match();
        }
        st.pack_and_store(st.nrec);
        if (st.warp == 0) {
            uint32_t out_bytes;
            if (st.res == kOk) {
                if (ends_with_flush(a.write_token, (uint32_t)st.pend_bits, a.flags, stream, (uint64_t)N)) {
                    if (st.lane == 0) st.recs[0] = ((uint32_t)kHuff.code[kSymFlush] << 5) | kHuff.bits[kSymFlush];
                    st.pack_and_store(1);
                }
                out_bytes = st.ow * 4 + ((st.pend_bits + 7) >> 3);
            } else {
                out_bytes = st.ow * 4 + (st.pend_bits >> 3);
            }
            const uint32_t tail = out_bytes - st.ow * 4;
            const uint32_t w0 = st.stage[0];
            uint8_t *o8 = reinterpret_cast<uint8_t *>(st.out32 + st.ow);
            if ((uint32_t)st.lane < tail) o8[st.lane] = (uint8_t)(w0 >> (24 - 8 * st.lane));
            if (st.lane == 0) {
                a.b.out_sizes[stream] = out_bytes;
                if (a.b.status) a.b.status[stream] = (int8_t)st.res;
            }
        }
    }
}

__global__ void k_build_dictrows_wide(const uint8_t *dict, int W, uint32_t *rows_out, int rs, int total_words) {
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
    if (threadIdx.x == 0) rows_out[ww + 1] = dict[W - 1];  // rides in a pad word no level ever reads
}

#ifndef TB_EMU
constexpr int kWideSlots = 8;
constexpr size_t kWideSlotBytes = 32 * (1024 + 4) * 4;
uint8_t *g_widerows = nullptr;
int g_wideslot = 0;

template <int WBITS, bool EXT, int NWARPS>
void launch_wide(const WideCompArgs &a, cudaStream_t st) {
    using G = WGeo<WBITS, NWARPS>;
    static int blocks_per_sm = 0, sms = 0;
    if (!blocks_per_sm) {
        cudaFuncSetAttribute(k_wide_compress<WBITS, EXT, NWARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_wide_compress<WBITS, EXT, NWARPS>, G::T, G::SMEM);
        if (blocks_per_sm < 1) blocks_per_sm = 1;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const uint64_t persistent = (uint64_t)sms * blocks_per_sm;
    const unsigned grid = (unsigned)(a.b.n_streams < persistent ? a.b.n_streams : persistent);
    k_wide_compress<WBITS, EXT, NWARPS><<<grid, G::T, G::SMEM, st>>>(a);
    count_launch();
}
#endif  // TB_EMU

}  // namespace

#ifndef TB_EMU
bool launch_wide_compress_batch(const CompBatchConf &cf, const uint8_t *d_dict, const BatchArgs &b, cudaStream_t st,
                                bool only_deferred) {
    if (cf.window < 11 || cf.window > 15 || (cf.flags & TB_F_LAZY)) return false;
    if (b.in_offsets) return false;
    if ((b.in_stride & 15) || ((uintptr_t)b.in & 15) || (b.out_stride & 3) || ((uintptr_t)b.out & 3)) return false;
    if (b.in_stride > (1u << 30)) return false;
    const uint64_t bound = 2 + (b.in_stride * (uint64_t)(cf.literal + 1) + 7) / 8 + 6;
    if (b.out_stride < ((bound + 3) & ~3ull)) return false;
    if (b.n_streams == 0) return true;
    if (!g_widerows && cudaMalloc(&g_widerows, kWideSlots * kWideSlotBytes) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    const int W = 1 << cf.window, rs = W / 32 + 4;
    uint32_t *slot = reinterpret_cast<uint32_t *>(g_widerows + (size_t)(g_wideslot++ % kWideSlots) * kWideSlotBytes);
    k_build_dictrows_wide<<<1, 1024, 0, st>>>(d_dict, W, slot, rs, 32 * rs);
    count_launch();
    WideCompArgs a;
    a.b = b;
    a.dictrows = slot;
    a.literal = cf.literal;
    a.flags = cf.flags;
    a.write_token = cf.write_token;
    a.only_deferred = only_deferred ? 1 : 0;
    const bool ext = (cf.flags & TB_F_EXTENDED) != 0;
    switch (cf.window) {
        case 11: ext ? launch_wide<11, true, 1>(a, st) : launch_wide<11, false, 1>(a, st); break;
        case 12: ext ? launch_wide<12, true, 1>(a, st) : launch_wide<12, false, 1>(a, st); break;
        case 13: ext ? launch_wide<13, true, 2>(a, st) : launch_wide<13, false, 2>(a, st); break;
        case 14: ext ? launch_wide<14, true, 4>(a, st) : launch_wide<14, false, 4>(a, st); break;
        default: ext ? launch_wide<15, true, 8>(a, st) : launch_wide<15, false, 8>(a, st); break;
    }
    return true;
}

#endif  // TB_EMU

}  // namespace tb
