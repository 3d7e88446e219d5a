// Batch decompressor for any window (8..15) with one WARP per stream and the window in shared memory.
// Default for windows 11..15 since round 2 (measured on B200: 40 / 23 GB/s at window 12 / 15 against 23 / 7 for the
// general kernel); fast_decompress.cu keeps windows <= 10.
//
// Why: the lane-per-stream kernel needs 32 windows per warp, which stops at 1 KiB windows; the general kernel keeps
// 2..32 KiB windows in global memory and walks them one byte at a time from a single thread.  Here the bit walk is
// warp-uniform (every lane holds the same bit buffer and decodes the same token — the frame is a strictly serial
// stream, decompressor.c:371-578) and the lanes share the COPY: a token's bytes are read from the shared-memory window
// by up to 32 lanes at once, then written to the window and to the output row (all reads before any write = the
// snapshot rule of tamp_window_copy, common.c:58-86).  Shared memory per stream is the window itself: 7 streams per SM
// at window 15, 56 at window 12.
//
// Semantics restated (whole-frame call, as fast_decompress.cu): header parsing and rejection
// (tamp_decompressor_read_header :276-297, populate_from_conf :304-329), literal / token / FLUSH / double-FLUSH reset
// (:466-514), OOB checks (:540-544, :231-236), RLE and extended match with their clipped window writes (:114-273),
// the partial token at the end of the output row (:547-562), INPUT_EXHAUSTED when a token's bits are not all there.
#include "../tb_wire.h"
#include "tb_cuda.h"
#include "tb_device_common.cuh"

namespace tb {

namespace {

struct WideDecArgs {
    BatchArgs b;
    const uint8_t *seed;    // 3 x 32 KiB seeded dictionaries (literal classes 5, 6, 7/8)
    const uint8_t *custom;  // caller dictionary or nullptr
    int window_bits_max;
    int only_deferred;      // pick-up pass: just the streams whose out_sizes entry is kDeferred (left by lsplit_decompress.cu)
};

// Warp-uniform bit reader: MSb-aligned unread bits in `bb`, `nb` of them valid.
struct BitReader {
    const uint8_t *in;
    uint32_t n, ip;
    uint64_t bb;
    int nb;
    __device__ __forceinline__ void refill() {  // decompressor.c:357-365, a word at a time where alignment allows
        while (nb <= 32 && ip < n) {
            if (ip + 4 <= n && ((reinterpret_cast<uintptr_t>(in) + ip) & 3) == 0) {
                const uint32_t w = *reinterpret_cast<const uint32_t *>(in + ip);
                bb |= (uint64_t)__byte_perm(w, 0, 0x0123) << (32 - nb);
                nb += 32;
                ip += 4;
            } else {
                bb |= (uint64_t)in[ip] << (56 - nb);
                nb += 8;
                ip += 1;
            }
        }
    }
    __device__ __forceinline__ void drop(int k) {
        bb <<= k;
        nb -= k;
    }
};

__global__ void __launch_bounds__(512) k_wide_decompress(WideDecArgs a) {
#ifndef TB_EMU
    extern __shared__ __align__(128) uint8_t smem[];
#else  // tests/emu: the kernel stepped on the CPU (test infrastructure; see tests/emu/cuda_emu.h)
    uint8_t *smem = emu::g_smem;
#endif
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    uint8_t *win = smem + ((size_t)warp << a.window_bits_max);
    const uint64_t nwarps = (uint64_t)gridDim.x * wpc;

    for (uint64_t stream = (uint64_t)blockIdx.x * wpc + warp; stream < a.b.n_streams; stream += nwarps) {
        if (a.only_deferred && a.b.out_sizes[stream] != kDeferred) continue;  // (the same word for the whole warp)
        BitReader r;
        r.in = a.b.in + (a.b.in_offsets ? a.b.in_offsets[stream] : stream * a.b.in_stride);
        r.n = a.b.in_sizes ? a.b.in_sizes[stream] : (uint32_t)a.b.in_stride;
        r.ip = 0;
        r.bb = 0;
        r.nb = 0;
        uint8_t *out = a.b.out + stream * a.b.out_stride;
        const uint32_t cap = (uint32_t)a.b.out_stride;
        uint32_t opos = 0;
        int status = kInputExhausted;
        bool active = false;
        int wbits = 10, lbits = 8, min_pat = 2;
        bool extended = false, dict_reset = false;
        const uint8_t *dict_src = a.seed + 2 * 32768;

        // ---- header (every lane reads the same bytes) -----------------------------------------------------------
        if (r.n != 0) {
            const uint32_t hs = frame_start(a.b.seg_header, stream, r.in, r.n), h = hs & 0xFFu;
            const uint32_t hdr = 1 + (h & 1u);
            if (r.n >= hdr) {
                if (hdr == 2 && (hs >> 8) != 0) {
                    status = kInvalidConf;
                } else {
                    wbits = (int)((h >> 5) & 7u) + 8;
                    lbits = (int)((h >> 3) & 3u) + 5;
                    extended = (h & 2u) != 0;
                    dict_reset = (h & 1u) != 0;
                    const bool use_custom = (h & 4u) != 0;
                    if (wbits > a.window_bits_max || (use_custom && !a.custom)) {
                        status = kInvalidConf;
                    } else {
                        min_pat = min_pattern_size(wbits, lbits);
                        const int seed_lit = extended ? lbits : 8;
                        dict_src = use_custom ? a.custom : a.seed + (seed_lit <= 5 ? 0 : seed_lit <= 6 ? 1 : 2) * 32768;
                        r.ip = hdr;
                        active = true;
                    }
                }
            }
        }
        const int W = 1 << wbits, mask = W - 1;
        const int max_plain_sym = extended ? kSymRle - 1 : kSymFlush - 1;
        int wpos = 0;
        bool last_flush = false;

        __syncwarp();
        if (active) {
            for (int i = lane * 16; i < W; i += 512)
                *reinterpret_cast<uint4 *>(win + i) = __ldg(reinterpret_cast<const uint4 *>(dict_src + i));
        }
        __syncwarp();

        while (active) {
            r.refill();
            if (r.nb == 0) break;  // frame fully consumed: INPUT_EXHAUSTED
            if (opos == cap) {
                status = kOutputFull;
                break;
            }
            const uint32_t top = (uint32_t)(r.bb >> 32);
            if (top >> 31) {  // literal (:466-482)
                if (r.nb < 1 + lbits) {
                    last_flush = false;
                    break;
                }
                uint32_t lits = (top << 1) >> (32 - lbits);
                r.drop(1 + lbits);
                last_flush = false;
                // literals right behind this one ride along (up to four per iteration): same bytes, fewer iterations
                int nlit = 1;
                while (nlit < 4) {
                    r.refill();
                    const uint32_t t2 = (uint32_t)(r.bb >> 32);
                    if (!(t2 >> 31) || r.nb < 1 + lbits || opos + (uint32_t)nlit >= cap) break;
                    lits |= ((t2 << 1) >> (32 - lbits)) << (8 * nlit);
                    r.drop(1 + lbits);
                    nlit++;
                }
                if (lane < nlit) {
                    const uint32_t c = (lits >> (8 * lane)) & 0xFFu;
                    win[(wpos + lane) & mask] = (uint8_t)c;
                    out[opos + lane] = (uint8_t)c;
                }
                wpos = (wpos + nlit) & mask;
                opos += (uint32_t)nlit;
                __syncwarp();
                continue;
            }
            const bool long_code = ((top >> 30) & 1u) != 0;
            const uint32_t e = long_code ? kHuff.lut[(top << 2) >> 25] : 0u;
            const int sym = long_code ? (int)(e & 15u) : 0;
            const int used = long_code ? 2 + (int)(e >> 4) : 2;
            if (r.nb < used) break;  // Huffman code incomplete

            if (sym <= max_plain_sym) {  // plain token (:516-572)
                last_flush = false;
                if (r.nb < used + wbits) break;  // offset not there: nothing is consumed, the frame ends here
                const int tlen = sym + min_pat;
                const int off = (int)((r.bb << used) >> (64 - wbits));
                if (off + tlen > W) {  // also covers off >= W (:540-544)
                    status = kOob;
                    break;
                }
                const uint32_t space = cap - opos;
                if ((uint32_t)tlen > space) {  // partial token at the end of the row: bytes only, no window update (:547-562)
                    if ((uint32_t)lane < space) out[opos + lane] = win[off + lane];
                    opos += space;
                    status = kOutputFull;
                    break;
                }
                r.drop(used + wbits);
                uint32_t byte = 0;
                if (lane < tlen) byte = win[off + lane];
                __syncwarp();  // snapshot: every source byte is read before any destination byte is written
                if (lane < tlen) {
                    win[(wpos + lane) & mask] = (uint8_t)byte;
                    out[opos + lane] = (uint8_t)byte;
                }
                wpos = (wpos + tlen) & mask;
                opos += (uint32_t)tlen;
                __syncwarp();
                continue;
            }
            if (sym == kSymFlush) {  // drop to the next byte boundary of the frame (:501-514)
                r.drop(used + ((r.nb - used) & 7));
                if (dict_reset && last_flush) {  // double FLUSH: re-seed the window
                    const int seed_lit = extended ? lbits : 8;
                    const uint8_t *sd = a.seed + (seed_lit <= 5 ? 0 : seed_lit <= 6 ? 1 : 2) * 32768;
                    __syncwarp();
                    for (int i = lane * 16; i < W; i += 512)
                        *reinterpret_cast<uint4 *>(win + i) = __ldg(reinterpret_cast<const uint4 *>(sd + i));
                    __syncwarp();
                    wpos = 0;
                }
                last_flush = true;
                continue;
            }
            // ---- RLE / extended match: the symbol is consumed (:521-526), then the count / size code ----------
            last_flush = false;
            uint64_t b2 = r.bb << used;
            int n2 = r.nb - used;
            const int trailing = sym == kSymRle ? 4 : 3;
            int value = -1;
            if (n2 >= 1 + trailing) {
                const uint32_t t2 = (uint32_t)(b2 >> 32);
                int s2 = -1, u2 = 0;
                if ((t2 >> 31) == 0) {
                    s2 = 0;
                    u2 = 1;
                } else {
                    const uint32_t e2 = kHuff.lut[(t2 << 1) >> 25];
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
            if (value < 0) break;  // the count / size code is not all there
            int xlen, off = 0, nwin;
            if (sym == kSymRle) {
                xlen = value + 2;
                nwin = xlen < kRleWindowMax ? xlen : kRleWindowMax;
                if (nwin > W - wpos) nwin = W - wpos;
            } else {
                xlen = value + min_pat + 12;
                if (n2 < wbits) {  // one more top-up for the offset (the reference parks the size, :216-223)
                    while (n2 <= 32 && r.ip < r.n) {
                        b2 |= (uint64_t)r.in[r.ip] << (56 - n2);
                        n2 += 8;
                        r.ip += 1;
                    }
                }
                if (n2 < wbits) break;
                off = (int)(b2 >> (64 - wbits));
                b2 <<= wbits;
                n2 -= wbits;
                if (off >= W || off + xlen > W) {  // :231-236
                    status = kOob;
                    break;
                }
                nwin = xlen < W - wpos ? xlen : W - wpos;
            }
            r.bb = b2;
            r.nb = n2;
            const uint32_t space = cap - opos;
            const uint32_t nout = (uint32_t)xlen < space ? (uint32_t)xlen : space;
            if (sym == kSymRle) {
                const uint32_t rsym = win[(wpos - 1) & mask];
                __syncwarp();
                for (uint32_t i = lane; i < nout; i += 32) out[opos + i] = (uint8_t)rsym;
                opos += nout;
                if (nout < (uint32_t)xlen) {
                    status = kOutputFull;
                    break;
                }
                if (lane < nwin) win[wpos + lane] = (uint8_t)rsym;
                wpos = (wpos + nwin) & mask;
            } else {
                // up to 134 bytes: five per lane, all read before anything is written
                uint32_t got[5];
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    const int i = lane + 32 * k;
                    got[k] = i < xlen ? win[off + i] : 0u;
                }
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    const int i = lane + 32 * k;
                    if ((uint32_t)i < nout) out[opos + i] = (uint8_t)got[k];
                }
                opos += nout;
                if (nout < (uint32_t)xlen) {
                    status = kOutputFull;
                    break;
                }
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    const int i = lane + 32 * k;
                    if (i < nwin) win[wpos + i] = (uint8_t)got[k];  // no wrap: clipped at the window's end
                }
                wpos = (wpos + nwin) & mask;
            }
            __syncwarp();
        }

        if (lane == 0) {
            a.b.out_sizes[stream] = opos;
            if (a.b.status) a.b.status[stream] = (int8_t)status;
        }
        __syncwarp();
    }
}

}  // namespace

#ifndef TB_EMU
bool launch_wide_decompress_batch(const uint8_t *d_seed, const uint8_t *d_custom, int window_bits_max, const BatchArgs &b,
                                  cudaStream_t st, bool only_deferred) {
    if (window_bits_max < 8 || window_bits_max > 15) return false;
    if (b.out_stride > 0xFFFFFFF0ull) return false;
    if (b.n_streams == 0) return true;
    static int sms = 0;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_wide_decompress, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        configured = true;
    }
    const size_t per_warp = (size_t)1 << window_bits_max;
    int wpc = (int)((224 * 1024) / per_warp);  // one CTA per SM holding as many windows as shared memory takes
    wpc = wpc > 16 ? 16 : wpc;                 // (small windows: several CTAs per SM instead of more warps per CTA)
    const int ctas_per_sm = (int)((224 * 1024) / (per_warp * wpc));
    const uint64_t want = (b.n_streams + wpc - 1) / wpc;
    uint64_t persistent = (uint64_t)sms * (ctas_per_sm < 4 ? ctas_per_sm : 4);
    if (persistent < 1) persistent = 1;
    WideDecArgs a;
    a.b = b;
    a.seed = d_seed;
    a.custom = d_custom;
    a.window_bits_max = window_bits_max;
    a.only_deferred = only_deferred ? 1 : 0;
    k_wide_decompress<<<(unsigned)(want < persistent ? want : persistent), wpc * 32, per_warp * wpc, st>>>(a);
    count_launch();
    return true;
}
#endif  // TB_EMU

}  // namespace tb
