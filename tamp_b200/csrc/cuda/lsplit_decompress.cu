// Split batch decompressor for frames of ANY window (8..15) and ANY output length: the decompressor of BASELINE.json
// config 3 (64 KiB streams, window 15), config 4 (4 KiB frames, window 10, decompress-only) and of every class of
// config 5 since round 2.
//
// Where it comes from.  split_decompress.cu separates the serial bit walk of a frame (tamp_decompressor_decompress_cb,
// decompressor.c:371-578: a token's length and the position of the next one depend on the bits alone) from the copies,
// but keeps the output row — which is the window while N <= W — in shared memory.  For longer rows the same holds with
// the window spelt out: the decoder writes every output byte to the window in order (decompressor.c:546-577,
// tamp_window_copy), so ring position p holds, for a token that starts at output position o,
//     output[base + p]       if p <  o mod W     (base = o - o mod W: this lap)
//     output[base + p - W]   if p >= o mod W     (the lap before; the dictionary's byte p while that is negative)
// — the frame's own output row, in global memory and L2-resident because it was written a moment ago, IS the history.
// No window copy exists anywhere; shared memory holds only the Huffman LUT and one 32 x 32 record tile per warp, so
// occupancy is no longer set by the window size (wide_decompress.cu: one 32 KiB window per warp, 7 warps per SM).
//
//   phase 1  PARSE, one lane per stream, registers only: up to kChunk tokens per lane -> 32-bit records (literal byte, or
//            kind | length | window offset) through the tile into a global scratch area;
//   phase 2  COPY, one warp per stream, for each of the warp's 32 streams: 32 tokens at a time, warp prefix sum of the
//            lengths gives every token its output position; a token whose source lies entirely before the group's output
//            (or in the dictionary) and does not straddle the write position is one unaligned 16-byte read and up to 16
//            byte stores, one token per lane; the others follow in order, lanes sharing the bytes — the group's own
//            output is mirrored in 512 bytes of shared memory, so a token fed by its neighbours does not wait for L2;
//   then the parse resumes where it stopped (its state never left the registers).
//
// Extended format: run / extended-match tokens are records like any other.  A run of more than 8 bytes, or a run /
// extended match clipped at the end of the window buffer, is written to the window only in part (decompressor.c:160-168,
// :250-258): from then on the window lags the output, and the first token behind it that is not a literal sends the
// stream to the pick-up pass.  So does everything else that is not a complete in-bounds token — FLUSH, dictionary_reset
// headers, hostile offsets, rows that fill up in the middle of a token: the stream is marked kDeferred and decoded from
// its start by fast_decompress.cu / wide_decompress.cu, so statuses and partial outputs stay the reference's.
#include "../tb_wire.h"
#include "tb_cuda.h"
#include "tb_device_common.cuh"
#include "tb_smem.cuh"

namespace tb {

namespace {

constexpr uint32_t kFullMask = 0xffffffffu;
constexpr int kLsWarps = 8;       // warps per CTA
constexpr int kChunk = 128;       // tokens per lane between two copy phases (records: 16 KiB per warp, L2-resident)
constexpr int LS_LUT = 0, LS_WARP0 = 128;
constexpr int kLsTile = 32 * 32 * 4;
constexpr int kStage = 32 * 16;   // the first kStage bytes of a group's own output are mirrored in shared memory
constexpr int kLsPerWarp = kLsTile + kStage + 16;  // (a 16-byte token may start at kStage - 1)
constexpr int kLsSmem = LS_WARP0 + kLsWarps * kLsPerWarp;

__device__ unsigned int d_lsplit_deferred_total = 0;

struct LsplitArgs {
    BatchArgs b;
    const uint8_t *seed;    // 3 x 32 KiB seeded dictionaries (literal classes 5, 6, 7/8)
    const uint8_t *custom;  // caller dictionary or nullptr
    int window_bits_max;
    uint32_t *scratch;      // [warps in the grid][32 streams][kChunk] token records
};

// record: bit 31 = token (else literal byte); bits 30..29 kind (0 plain match, 1 run, 2 extended match);
// bits 23..16 length; bits 14..0 window offset
constexpr uint32_t kKindPlain = 0, kKindRun = 1, kKindExt = 2;
__device__ __forceinline__ uint32_t ls_rec(uint32_t kind, uint32_t len, uint32_t off) { return 0x80000000u | (kind << 29) | (len << 16) | off; }

// 16 bytes at global address p (any alignment; the caller keeps 20 bytes behind p and 3 before it readable)
__device__ __forceinline__ void gload16(const uint8_t *p, uint32_t (&w)[4]) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const volatile uint32_t *q = reinterpret_cast<const volatile uint32_t *>(a & ~(uintptr_t)3);
    const int sh = (int)((a & 3u) << 3);
    const uint32_t a0 = q[0], a1 = q[1], a2 = q[2], a3 = q[3], a4 = q[4];
    w[0] = __funnelshift_r(a0, a1, sh);
    w[1] = __funnelshift_r(a1, a2, sh);
    w[2] = __funnelshift_r(a2, a3, sh);
    w[3] = __funnelshift_r(a3, a4, sh);
}

__global__ void __launch_bounds__(kLsWarps * 32) k_lsplit_decompress(LsplitArgs a) {
#ifndef TB_EMU
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t *sm = smem_raw;
#else
    uint8_t *sm = emu::g_smem;
#endif
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
#ifndef TB_EMU
    asm volatile("" : "+r"(sbase));
#endif
    const uint32_t sLut = sbase + LS_LUT, sTile = sbase + LS_WARP0 + warp * kLsPerWarp, sStage = sTile + kLsTile;
    const uint8_t *common = a.seed + 2 * 32768;  // the dictionary of v1 frames and of literal 7 / 8 (common.c:18-25)
    if (threadIdx.x < 128) sm[LS_LUT + threadIdx.x] = kHuff.lut[threadIdx.x];
    __syncthreads();

    uint32_t *myscratch = a.scratch + ((size_t)blockIdx.x * kLsWarps + warp) * (size_t)(kChunk * 32);
    uint32_t *lanescratch = myscratch + lane * kChunk;  // phase 1: this lane's stream
    const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t first = (uint64_t)blockIdx.x * blockDim.x + warp * 32;

    for (uint64_t batch = first; batch < a.b.n_streams; batch += nthreads) {
        const uint64_t stream = batch + lane;
        // ---- per-stream parse state (lane = stream) ----
        bool active = stream < a.b.n_streams, defer = false;
        const uint8_t *in = nullptr;
        uint32_t n = 0, ip = 0, cap = 0, opos = 0;
        int status = kInputExhausted, wbits = 10, lbits = 8, min_pat = 2;
        bool extended = false, lag = false;  // lag: the window no longer is the output (a token written to it in part)
        const uint8_t *dict = common;
        uint64_t bb = 0;  // MSb-aligned unread bits
        int nb = 0;
        if (active) {
            in = a.b.in + (a.b.in_offsets ? a.b.in_offsets[stream] : stream * a.b.in_stride);
            n = a.b.in_sizes ? a.b.in_sizes[stream] : (uint32_t)a.b.in_stride;
            cap = (uint32_t)a.b.out_stride;
            if (n == 0) {
                active = false;  // nothing to read: INPUT_EXHAUSTED, no output
            } else {
                // header (decompressor.c:276-329): anything unusual goes to the pick-up pass
                // (a dictionary_reset header is two bytes; the double FLUSH it allows is a FLUSH: deferred where it turns up)
                const uint32_t hs = frame_start(a.b.seg_header, stream, in, n), h = hs & 0xFFu;
                wbits = (int)((h >> 5) & 7u) + 8;
                lbits = (int)((h >> 3) & 3u) + 5;
                extended = (h & 2u) != 0;
                const bool use_custom = (h & 4u) != 0;
                const bool two = (h & 1u) != 0;
                if ((two && (n < 2 || (hs >> 8) != 0)) || wbits > a.window_bits_max || (use_custom && !a.custom)) {
                    defer = true;
                    active = false;
                } else {
                    min_pat = min_pattern_size(wbits, lbits);
                    const int seed_lit = extended ? lbits : 8;
                    dict = use_custom ? a.custom : a.seed + (seed_lit <= 5 ? 0 : seed_lit <= 6 ? 1 : 2) * 32768;
                    ip = two ? 2 : 1;
                    while (ip < n && ((reinterpret_cast<uintptr_t>(in) + ip) & 3) != 0) {  // ragged head
                        bb |= (uint64_t)in[ip] << (56 - nb);
                        nb += 8;
                        ip += 1;
                    }
                }
            }
        }
        const uint32_t W = 1u << wbits;
        const int max_plain_sym = extended ? kSymRle - 1 : kSymFlush - 1;
        uint32_t next_word = 0;  // the aligned word at in + ip, requested one refill ahead
        if (active && ip + 4 <= n) next_word = *reinterpret_cast<const uint32_t *>(in + ip);
        uint32_t copied = 0;     // output bytes of this lane's stream the copy phases have produced so far

        while (__any_sync(kFullMask, active || opos != copied)) {
            // ================= phase 1: parse up to kChunk tokens (lane = stream) =================
            uint32_t k = 0, mine = 0;  // iterations so far (warp-uniform); records of my stream in this chunk
            while (k < (uint32_t)kChunk && __any_sync(kFullMask, active)) {
                bool emit = false;
                uint32_t rec = 0;
                if (active) {
                    // top up the bit buffer (decompressor.c:357-365): whole aligned words, bytes in the frame's tail
                    if (nb <= 32) {
                        if (ip + 4 <= n) {
                            bb |= (uint64_t)__byte_perm(next_word, 0, 0x0123) << (32 - nb);
                            nb += 32;
                            ip += 4;
                            if (ip + 4 <= n) next_word = *reinterpret_cast<const uint32_t *>(in + ip);
                        } else {
                            while (ip < n) {  // at most 3 bytes
                                bb |= (uint64_t)in[ip] << (56 - nb);
                                nb += 8;
                                ip += 1;
                            }
                        }
                    }
                    const uint32_t top = (uint32_t)(bb >> 32);
                    const bool is_lit = (top >> 31) != 0;
                    const uint32_t e = smem::ld8(sLut + ((top << 2) >> 25));
                    const bool long_code = ((top >> 30) & 1u) != 0;
                    const int sym = long_code ? (int)(e & 15u) : 0;
                    const int used = long_code ? 2 + (int)(e >> 4) : 2;
                    const int need = is_lit ? 1 + lbits : used + wbits;  // <= 9 + 15 bits: inside `top`
                    const int tlen = is_lit ? 1 : sym + min_pat;
                    const uint32_t off = (top << used) >> (32 - wbits);
                    const bool shape = is_lit || (sym <= max_plain_sym && off + (uint32_t)tlen <= W && !lag);
                    if (nb >= need && shape && opos + (uint32_t)tlen <= cap) {
                        rec = is_lit ? (top << 1) >> (32 - lbits) : ls_rec(kKindPlain, (uint32_t)tlen, off);
                        emit = true;
                        bb <<= need;
                        nb -= need;
                        opos += (uint32_t)tlen;
                    } else {
                        // extended format: run / extended-match tokens (second Huffman code without the flag bit + raw bits)
                        if (!is_lit && extended && !lag && (sym == kSymRle || sym == kSymExt)) {
                            const bool is_run = sym == kSymRle;
                            const uint32_t t2 = top << used;  // (used <= 9)
                            const bool long2 = (t2 >> 31) != 0;
                            const uint32_t e2 = smem::ld8(sLut + ((t2 << 1) >> 25));
                            const int hv = long2 ? (int)(e2 & 15u) : 0;
                            const int used2 = long2 ? 1 + (int)(e2 >> 4) : 1;
                            const int tr = is_run ? 4 : 3;
                            const int raw = (hv << tr) + (int)((t2 << used2) >> (32 - tr));
                            const int bits_tok = used + used2 + tr + (is_run ? 0 : wbits);  // <= 7 + 8 + 3 + 15 = 33: from the 64-bit buffer
                            const int xlen = is_run ? raw + 2 : raw + min_pat + 12;
                            const uint32_t xoff = is_run ? 0u : (uint32_t)((bb << (used + used2 + tr)) >> (64 - wbits));
                            if (nb >= bits_tok && opos + (uint32_t)xlen <= cap && (is_run || xoff + (uint32_t)xlen <= W)) {
                                rec = ls_rec(is_run ? kKindRun : kKindExt, (uint32_t)xlen, xoff);
                                // the window takes min(count, 8) bytes of a run, and nothing past the end of its buffer
                                const uint32_t wroom = W - (opos & (W - 1u));
                                if ((is_run && xlen > kRleWindowMax) || (uint32_t)(is_run ? (xlen < kRleWindowMax ? xlen : kRleWindowMax) : xlen) > wroom)
                                    lag = true;
                                emit = true;
                                bb <<= bits_tok;
                                nb -= bits_tok;
                                opos += (uint32_t)xlen;
                            }
                        }
                        if (!emit) {
                            // the frame ends, or something the copy phase does not do (same order of checks as the reference's loop)
                            if (nb == 0) {
                                // frame fully consumed: INPUT_EXHAUSTED
                            } else if (opos == cap) {
                                status = kOutputFull;  // bits left but the row is full (decompressor.c:433-463)
                            } else if (nb < (is_lit ? need : used) || (!is_lit && sym <= max_plain_sym && !lag && nb < need)) {
                                // incomplete token at the end of the frame: nothing is consumed
                            } else if (!is_lit && sym == kSymFlush && ip == n && nb - used < 8) {
                                // the closing FLUSH of a frame written with write_token (compressor.c:784-794): the decoder
                                // drops the padding bits behind it (decompressor.c:501-514) and the input is exhausted
                            } else {
                                defer = true;  // FLUSH, OOB, a token that does not fit the row, a token behind a partly written one
                            }
                            active = false;
                        }
                    }
                }
                // records go through a 32 x 32 tile: row = token index, column = lane; a full tile leaves as 128 bytes per stream
                if (emit) {
                    smem::st32(sTile + ((k & 31u) << 7) + 4u * lane, rec);
                    mine = k + 1;
                }
                k++;
                if ((k & 31u) == 0) {
                    __syncwarp();
                    uint4 *dstp = reinterpret_cast<uint4 *>(lanescratch + (k - 32));
#pragma unroll
                    for (int r = 0; r < 8; r++) {
                        uint4 v;
                        v.x = smem::ld32(sTile + ((4 * r) << 7) + 4u * lane);
                        v.y = smem::ld32(sTile + ((4 * r + 1) << 7) + 4u * lane);
                        v.z = smem::ld32(sTile + ((4 * r + 2) << 7) + 4u * lane);
                        v.w = smem::ld32(sTile + ((4 * r + 3) << 7) + 4u * lane);
                        dstp[r] = v;
                    }
                    __syncwarp();
                }
            }
            {  // the last, partial tile
                __syncwarp();
                const uint32_t k0 = k & ~31u;
                if (k0 != k) {
                    uint4 *dstp = reinterpret_cast<uint4 *>(lanescratch + k0);
#pragma unroll
                    for (int r = 0; r < 8; r++) {
                        uint4 v;
                        v.x = smem::ld32(sTile + ((4 * r) << 7) + 4u * lane);
                        v.y = smem::ld32(sTile + ((4 * r + 1) << 7) + 4u * lane);
                        v.z = smem::ld32(sTile + ((4 * r + 2) << 7) + 4u * lane);
                        v.w = smem::ld32(sTile + ((4 * r + 3) << 7) + 4u * lane);
                        dstp[r] = v;
                    }
                }
                __syncwarp();
            }

            // ================= phase 2: copy this chunk's tokens (warp = stream) =================
            if (defer) {  // decoded again from its start by the pick-up pass
                mine = 0;
                copied = opos;
            }
            for (int s = 0; s < 32; s++) {
                const uint32_t cnt = __shfl_sync(kFullMask, mine, s);
                if (cnt == 0) continue;
                const uint64_t sid = batch + s;
                const uint32_t sW = 1u << __shfl_sync(kFullMask, wbits, s);
                uint32_t done = __shfl_sync(kFullMask, copied, s);  // output position of the group's first token
                const uint64_t dict_bits = (uint64_t)reinterpret_cast<uintptr_t>(dict);
                const uint8_t *s_dict = reinterpret_cast<const uint8_t *>((uintptr_t)(
                    ((uint64_t)__shfl_sync(kFullMask, (uint32_t)(dict_bits >> 32), s) << 32) | __shfl_sync(kFullMask, (uint32_t)dict_bits, s)));
                uint8_t *out = a.b.out + sid * a.b.out_stride;
                const uint32_t scap = (uint32_t)a.b.out_stride;
                const uint32_t *recs = myscratch + s * kChunk;
                uint32_t rec = lane < cnt ? recs[lane] : 0u;
                for (uint32_t k0 = 0; k0 < cnt; k0 += 32) {
                    const uint32_t nextrec = k0 + 32 + lane < cnt ? recs[k0 + 32 + lane] : 0u;  // requested a group ahead
                    const bool valid = k0 + lane < cnt;
                    const bool is_tok = (rec & 0x80000000u) != 0;
                    const uint32_t kind = (rec >> 29) & 3u;
                    const int len = valid ? (is_tok ? (int)((rec >> 16) & 0xFFu) : 1) : 0;
                    const uint32_t off = rec & 0x7FFFu;
                    int incl = len;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int t = __shfl_up_sync(kFullMask, incl, d);
                        if (lane >= d) incl += t;
                    }
                    const uint32_t dst = done + (uint32_t)(incl - len);  // this token's output position
                    // where the window's bytes off .. off + len - 1 are, as the token sees them
                    const uint32_t wm = dst & (sW - 1u), lapbase = dst - wm;
                    const bool this_lap = off + (uint32_t)len <= wm, prev_lap = off >= wm;
                    const int64_t s0 = this_lap ? (int64_t)lapbase + off : (int64_t)lapbase + off - (int64_t)sW;
                    const bool from_dict = prev_lap && s0 < 0;
                    // at once: a plain match (<= 16 bytes) that does not straddle the write position, whose bytes were produced
                    // before this group or are the dictionary's, with 20 readable bytes behind the source
                    const bool indep = is_tok && kind == kKindPlain && (this_lap || prev_lap) &&
                                       (from_dict ? off + 20u <= sW : ((uint32_t)s0 + (uint32_t)len <= done && (uint32_t)s0 + 20u <= scap));
                    const bool dep = valid && is_tok && !indep;
                    if (valid && !dep) {
                        uint32_t w[4] = {rec & 0xFFu, 0u, 0u, 0u};
                        if (is_tok) gload16(from_dict ? s_dict + off : out + (uint32_t)s0, w);
                        uint8_t *o = out + dst;
                        const uint32_t rel = dst - done;
                        if (rel < (uint32_t)kStage) {  // mirror
                            const uint32_t sa = sStage + rel;
                            smem::st8(sa, w[0]);
                            if (len > 1) smem::st8(sa + 1, w[0] >> 8);
                            if (len > 2) smem::st8(sa + 2, w[0] >> 16);
                            if (len > 3) smem::st8(sa + 3, w[0] >> 24);
                            if (len > 4) smem::st8(sa + 4, w[1]);
                            if (len > 5) smem::st8(sa + 5, w[1] >> 8);
                            if (len > 6) smem::st8(sa + 6, w[1] >> 16);
                            if (len > 7) smem::st8(sa + 7, w[1] >> 24);
                            if (len > 8) smem::st8(sa + 8, w[2]);
                            if (len > 9) smem::st8(sa + 9, w[2] >> 8);
                            if (len > 10) smem::st8(sa + 10, w[2] >> 16);
                            if (len > 11) smem::st8(sa + 11, w[2] >> 24);
                            if (len > 12) smem::st8(sa + 12, w[3]);
                            if (len > 13) smem::st8(sa + 13, w[3] >> 8);
                            if (len > 14) smem::st8(sa + 14, w[3] >> 16);
                            if (len > 15) smem::st8(sa + 15, w[3] >> 24);
                        }
                        o[0] = (uint8_t)w[0];
                        if (len > 1) o[1] = (uint8_t)(w[0] >> 8);
                        if (len > 2) o[2] = (uint8_t)(w[0] >> 16);
                        if (len > 3) o[3] = (uint8_t)(w[0] >> 24);
                        if (len > 4) o[4] = (uint8_t)w[1];
                        if (len > 5) o[5] = (uint8_t)(w[1] >> 8);
                        if (len > 6) o[6] = (uint8_t)(w[1] >> 16);
                        if (len > 7) o[7] = (uint8_t)(w[1] >> 24);
                        if (len > 8) o[8] = (uint8_t)w[2];
                        if (len > 9) o[9] = (uint8_t)(w[2] >> 8);
                        if (len > 10) o[10] = (uint8_t)(w[2] >> 16);
                        if (len > 11) o[11] = (uint8_t)(w[2] >> 24);
                        if (len > 12) o[12] = (uint8_t)w[3];
                        if (len > 13) o[13] = (uint8_t)(w[3] >> 8);
                        if (len > 14) o[14] = (uint8_t)(w[3] >> 16);
                        if (len > 15) o[15] = (uint8_t)(w[3] >> 24);
                    }
                    uint32_t deps = __ballot_sync(kFullMask, dep);
                    __syncwarp();
                    while (deps) {  // in order; the lanes share the token's bytes (a token's source never is its own output)
                        const int j = __ffs(deps) - 1;
                        deps &= deps - 1;
                        const uint32_t jdst = __shfl_sync(kFullMask, dst, j), joff = __shfl_sync(kFullMask, off, j);
                        const int jlen = __shfl_sync(kFullMask, len, j);
                        const bool jrun = __shfl_sync(kFullMask, kind, j) == kKindRun;
                        const uint32_t jwm = jdst & (sW - 1u), jbase = jdst - jwm;
                        const volatile uint8_t *vout = out;
                        // a byte produced inside this group comes from its mirror in shared memory (mirrored while the group's
                        // output fits kStage), anything older from the row itself
                        const uint32_t staged = (uint32_t)kStage;
                        for (int o = lane; o < jlen; o += 32) {
                            uint32_t b;
                            int64_t sp;
                            if (jrun) {  // the last byte written (the dictionary's last byte at the start of the stream)
                                sp = (int64_t)jdst - 1;
                            } else {
                                const uint32_t p = joff + (uint32_t)o;
                                sp = p < jwm ? (int64_t)jbase + p : (int64_t)jbase + p - (int64_t)sW;
                            }
                            if (sp < 0)
                                b = s_dict[jrun ? sW - 1u : joff + (uint32_t)o];
                            else if ((uint32_t)sp >= done && (uint32_t)sp - done < staged)
                                b = smem::ld8(sStage + ((uint32_t)sp - done));
                            else
                                b = vout[sp];
                            const uint32_t d = jdst + (uint32_t)o;
                            out[d] = (uint8_t)b;
                            if (d - done < staged) smem::st8(sStage + (d - done), b);
                        }
                        __syncwarp();
                    }
                    done += (uint32_t)__shfl_sync(kFullMask, incl, 31);
                    rec = nextrec;
                }
                if (lane == s) copied = done;
                __syncwarp();
            }
            __syncwarp();
        }

        // ---- results ----
        if (stream < a.b.n_streams) {
            if (defer) {
                a.b.out_sizes[stream] = kDeferred;
                atomicAdd(&d_lsplit_deferred_total, 1u);
            } else {
                a.b.out_sizes[stream] = opos;
                if (a.b.status) a.b.status[stream] = (int8_t)status;
            }
        }
        __syncwarp();
    }
}

}  // namespace

#ifndef TB_EMU
bool launch_lsplit_decompress_batch(const uint8_t *d_seed, const uint8_t *d_custom, int window_bits_max, const BatchArgs &b,
                                    cudaStream_t st) {
    if (window_bits_max < 8 || window_bits_max > 15) return false;
    if (b.out_stride > 0x7FFFFFF0ull || b.out_stride < 32) return false;
    if ((b.out_stride & 3) || (reinterpret_cast<uintptr_t>(b.out) & 3)) return false;  // aligned 32-bit reads of the history
    if (b.n_streams == 0) return true;
    static int blocks_per_sm = 0, sms = 0;
    if (!blocks_per_sm) {
        cudaFuncSetAttribute(k_lsplit_decompress, cudaFuncAttributeMaxDynamicSharedMemorySize, kLsSmem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_lsplit_decompress, kLsWarps * 32, kLsSmem);
        if (blocks_per_sm < 1) blocks_per_sm = 1;
        if (blocks_per_sm > 4) blocks_per_sm = 4;  // 32 warps per SM; more only spreads the record scratch over more of L2
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const uint64_t per_block = kLsWarps * 32;
    const uint64_t want = (b.n_streams + per_block - 1) / per_block;
    const uint64_t persistent = (uint64_t)sms * blocks_per_sm;
    const unsigned grid = (unsigned)(want < persistent ? want : persistent);
    // token records: stream-ordered scratch
    uint32_t *scratch = nullptr;
    const size_t scratch_bytes = (size_t)grid * kLsWarps * kChunk * 32 * sizeof(uint32_t);
    if (cudaMallocAsync(&scratch, scratch_bytes, st) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    LsplitArgs a;
    a.b = b;
    a.seed = d_seed;
    a.custom = d_custom;
    a.window_bits_max = window_bits_max;
    a.scratch = scratch;
    k_lsplit_decompress<<<grid, kLsWarps * 32, kLsSmem, st>>>(a);
    count_launch();
    cudaFreeAsync(scratch, st);
    // second pass: the window-keeping kernels pick up the streams marked kDeferred (usually none)
    static unsigned int *h_seen = nullptr;  // pinned mirror of d_lsplit_deferred_total
    static unsigned int last_seen = 0;
    if (!h_seen && cudaMallocHost(&h_seen, sizeof *h_seen) == cudaSuccess) *h_seen = 0;
    bool expect_work = true;
    if (h_seen) {
        const unsigned int now = *reinterpret_cast<volatile unsigned int *>(h_seen);
        expect_work = now != last_seen;
        last_seen = now;
    } else {
        cudaGetLastError();
    }
    bool ok;
    if (window_bits_max <= 10)
        ok = launch_fast_decompress_batch(d_seed, d_custom, window_bits_max, b, st, /*only_deferred=*/true, /*small_grid=*/!expect_work);
    else
        ok = launch_wide_decompress_batch(d_seed, d_custom, window_bits_max, b, st, /*only_deferred=*/true);
    if (h_seen) cudaMemcpyFromSymbolAsync(h_seen, d_lsplit_deferred_total, sizeof *h_seen, 0, cudaMemcpyDeviceToHost, st);
    return ok;
}
#endif  // TB_EMU

}  // namespace tb
