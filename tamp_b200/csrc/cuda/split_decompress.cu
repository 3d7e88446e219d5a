// Split batch decompressor for frames whose output fits the window (N <= W <= 1024, no wrap): the decompressor of the
// headline workload (BASELINE.json config 2) since round 2.
//
// Why.  The lane-per-stream decoder (fast_decompress.cu) keeps every lane's 1 KiB window in shared memory: 32 KiB per
// warp, 7 warps per SM, and the serial bit walk of a frame (tamp_decompressor_decompress_cb, decompressor.c:371-578)
// then runs at 1.75 warps per scheduler — latency-bound at 37 % of the issue slots (profiles/r01d_full_fd_summary.txt).
// The bit walk itself needs no window at all: a token's length and the position of the next token depend only on the
// bits (decompressor.c:466-545).  So the two halves are separated:
//
//   phase 1  PARSE, one lane per stream, registers only: 64-bit bit reader over the frame (aligned 32-bit loads, the next
//            word requested a refill ahead), Huffman LUT, one token per lock-step iteration -> a 16-bit record (literal
//            byte, or length + window offset).  Records pass through a 32 x 32 tile in shared memory (2 KiB per warp) and
//            leave as 64 contiguous bytes per stream into a global scratch area (L2-resident: it is re-read at once);
//   phase 2  COPY, one warp per stream, for each of the warp's 32 streams: 32 tokens at a time, warp prefix sum of the
//            lengths gives every token its output position; tokens whose source bytes were produced before this group
//            of 32 (or still are dictionary bytes) copy at once, one token per lane; the few that read bytes produced
//            inside the group follow in order, lanes sharing the bytes.  A token's source is one unaligned 16-byte read
//            (output row or dictionary, both in shared memory).  The output row is assembled in 1 KiB of shared memory
//            per warp and leaves with 128-bit stores.
//
// For N <= W the window never wraps: window position x holds output byte x once written and the dictionary byte before
// (tamp_window_copy's snapshot rule, common.c:58-86, reduces to "a source position at or past the token's own output
// position reads the dictionary").
//
// Extended format (decode_rle / decode_extended_match, decompressor.c:113-262).  An RLE token repeats the last byte
// written, an extended match copies up to 134 bytes from the window as it is before the token; both are complete
// in-bounds tokens like any other and travel as records of their own (an extended match as two: length, then window
// offset).  In the copy phase they take the in-order path, 32 bytes per step.  One thing breaks "window position =
// output position": an RLE token writes at most 8 bytes to the window (:160-168), so after a run of more than 8 the
// window lags the output.  Literals and further runs do not care; the first match token behind such a run sends the
// stream to fast_decompress.cu.
//
// Everything that is not a literal or a complete in-bounds token — FLUSH, dictionary_reset headers, hostile offsets,
// output rows that fill up, frames that produce more than W bytes — is left to fast_decompress.cu: the stream is marked
// kDeferred and picked up by a second launch, so statuses and partial outputs stay the reference's in every case.
#include "../tb_wire.h"
#include "tb_cuda.h"
#include "tb_device_common.cuh"
#include "tb_smem.cuh"

namespace tb {

namespace {

constexpr uint32_t kFull = 0xffffffffu;
constexpr int kRowBytes = 1024 + 32;              // output row + slack for the 16-byte source reads
constexpr int kSplitWarps = 16;                   // warps per CTA
constexpr int kMaxTok = 1024;                     // a token yields at least one byte and the output is at most W bytes
// shared memory: Huffman LUT, the common dictionary (+ slack), then per warp: output row, 32 x 32 token tile
constexpr int S_LUT = 0, S_DICT = 128, S_WARP0 = S_DICT + 1024 + 32;
constexpr int kTileBytes = 32 * 32 * 2;
constexpr int kPerWarp = kRowBytes + kTileBytes;
constexpr int kSplitSmem = S_WARP0 + kSplitWarps * kPerWarp;

__device__ unsigned int d_split_deferred_total = 0;

struct SplitDecArgs {
    BatchArgs b;
    const uint8_t *seed;    // 3 x 32 KiB seeded dictionaries (literal classes 5, 6, 7/8)
    const uint8_t *custom;  // caller dictionary or nullptr
    int window_bits_max;
    int aligned_io;         // out rows 16-byte aligned (128-bit stores allowed)
    uint16_t *scratch;      // [warps in the grid][32 streams][kMaxTok] token records
};


// record: bit 15 = match; match: (len - 2) << 10 | offset (len 2..16 -> 4 bits); literal: the byte.
// bits 15 + 14 = special: kind << 12 | payload: 0 = run (payload: count), 1 = extended match (payload: length; the record
// behind it, kind 2, carries the window offset), 3 = filler (an extended match never straddles a group of 32 records)
__device__ __forceinline__ uint32_t rec_match(int len, uint32_t off) { return 0x8000u | ((uint32_t)(len - 2) << 10) | off; }
__device__ __forceinline__ uint32_t rec_special(uint32_t kind, uint32_t payload) { return 0xC000u | (kind << 12) | payload; }
constexpr uint32_t kRecRun = 0, kRecExt = 1, kRecExtOff = 2, kRecFill = 3;

// The copy phase's rare tokens (extended format): a run (joff >= 0x10000: jlen copies of the last byte written, the
// dictionary's last byte at the start of the stream) or an extended match of more than 32 bytes.  A source byte below the
// token's own output position is output, the rest still is the dictionary's; the token's own bytes never feed it
// (tamp_window_copy's snapshot rule), so 32 bytes at a time is exact.  s_dict == nullptr: the dictionary is the shared one.
__device__ __noinline__ void copy_long_token(uint32_t sRow, uint32_t sDict, const uint8_t *s_dict, uint32_t jdst, uint32_t joff,
                                             int jlen, int lane, uint32_t W) {
    if (joff >= 0x10000u) {
        const uint32_t x = jdst ? jdst - 1u : W - 1u;
        const uint32_t b = jdst ? smem::ld8(sRow + x) : (!s_dict ? smem::ld8(sDict + x) : (uint32_t)__ldg(s_dict + x));
        for (int o = lane; o < jlen; o += 32) smem::st8(sRow + jdst + (uint32_t)o, b);
    } else {
        for (int o = lane; o < jlen; o += 32) {
            const uint32_t x = joff + (uint32_t)o;
            const uint32_t b = x < jdst ? smem::ld8(sRow + x) : (!s_dict ? smem::ld8(sDict + x) : (uint32_t)__ldg(s_dict + x));
            smem::st8(sRow + jdst + (uint32_t)o, b);
        }
    }
}

__global__ void __launch_bounds__(kSplitWarps * 32) k_split_decompress(SplitDecArgs a) {
#ifndef TB_EMU
    extern __shared__ __align__(128) uint8_t smem[];
#else
    uint8_t *smem = emu::g_smem;
#endif
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
#ifndef TB_EMU
    asm volatile("" : "+r"(sbase));
#endif
    const uint32_t sLut = sbase + S_LUT, sDict = sbase + S_DICT;
    const uint32_t sRow = sbase + S_WARP0 + warp * kPerWarp, sTile = sRow + kRowBytes;
    const uint8_t *row = smem + S_WARP0 + warp * kPerWarp;
    const uint8_t *common = a.seed + 2 * 32768;           // the dictionary of v1 frames and of literal 7 / 8 (common.c:18-25)
    if (threadIdx.x < 128) smem[S_LUT + threadIdx.x] = kHuff.lut[threadIdx.x];
    for (int i = threadIdx.x; i < (1024 + 32) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t *>(smem + S_DICT)[i] = i < 256 ? reinterpret_cast<const uint32_t *>(common)[i] : 0u;
    __syncthreads();

    uint16_t *myscratch = a.scratch + ((size_t)blockIdx.x * kSplitWarps + warp) * (size_t)(kMaxTok * 32);
    uint16_t *lanescratch = myscratch + lane * kMaxTok;   // phase 1: this lane's stream
    const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t first = (uint64_t)blockIdx.x * blockDim.x + warp * 32;

    for (uint64_t batch = first; batch < a.b.n_streams; batch += nthreads) {
        const uint64_t stream = batch + lane;
        // ================= phase 1: parse (lane = stream) =================
        bool active = stream < a.b.n_streams, defer = false;
        const uint8_t *in = nullptr;
        uint32_t n = 0, ip = 0, cap = 0, opos = 0;
        int status = kInputExhausted, wbits = 10, lbits = 8, min_pat = 2, max_plain_sym = kSymFlush - 1;
        const uint8_t *dict = common;
        uint64_t bb = 0;   // MSb-aligned unread bits
        int nb = 0;
        if (active) {
            in = a.b.in + (a.b.in_offsets ? a.b.in_offsets[stream] : stream * a.b.in_stride);
            n = a.b.in_sizes ? a.b.in_sizes[stream] : (uint32_t)a.b.in_stride;
            cap = (uint32_t)a.b.out_stride;
            // header (decompressor.c:276-329): anything unusual goes to the lane-per-stream kernel
            if (n == 0) {
                active = false;  // nothing to read: INPUT_EXHAUSTED, no output
            } else {
                // (a dictionary_reset header is two bytes; the double FLUSH it allows is a FLUSH: deferred where it turns up)
                const uint32_t hs = frame_start(a.b.seg_header, stream, in, n), h = hs & 0xFFu;
                wbits = (int)((h >> 5) & 7u) + 8;
                lbits = (int)((h >> 3) & 3u) + 5;
                const bool extended = (h & 2u) != 0, use_custom = (h & 4u) != 0;
                const bool two = (h & 1u) != 0;
                if ((two && (n < 2 || (hs >> 8) != 0)) || wbits > a.window_bits_max || wbits > 10 || (use_custom && !a.custom)) {
                    defer = true;
                    active = false;
                } else {
                    min_pat = min_pattern_size(wbits, lbits);
                    const int seed_lit = extended ? lbits : 8;
                    dict = use_custom ? a.custom : a.seed + (seed_lit <= 5 ? 0 : seed_lit <= 6 ? 1 : 2) * 32768;
                    max_plain_sym = extended ? kSymRle - 1 : kSymFlush - 1;
                    ip = two ? 2 : 1;
                    // ragged head: bytes up to the next aligned word of the frame
                    while (ip < n && ((reinterpret_cast<uintptr_t>(in) + ip) & 3) != 0) {
                        bb |= (uint64_t)in[ip] << (56 - nb);
                        nb += 8;
                        ip += 1;
                    }
                }
            }
        }
        const uint32_t W = 1u << wbits;
        const uint32_t room = cap < W ? cap : W;  // output beyond this needs the ring / the OUTPUT_FULL rules
        bool long_run = false;                    // a run of more than 8 bytes went by: the window lags the output
        uint32_t pending = 0;                     // second record of an extended match, due in the next iteration
        bool had_special = false;                 // the stream has run / extended-match records (the copy phase looks for them only then)
        uint32_t next_word = 0;                   // the aligned word at in + ip, requested one refill ahead
        if (active && ip + 4 <= n) next_word = *reinterpret_cast<const uint32_t *>(in + ip);
        uint32_t k = 0;  // tokens so far: the same in every lane that is still active
        while (__any_sync(kFull, active)) {
            bool emit = false;
            uint32_t rec = 0;
            if (active && pending) {
                rec = pending;
                pending = 0;
                emit = true;
            } else if (active) {
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
                const int need = is_lit ? 1 + lbits : used + wbits;
                const int tlen = is_lit ? 1 : sym + min_pat;
                const uint32_t off = (top << used) >> (32 - wbits);  // used + wbits <= 19 bits: all inside `top`
                const bool shape = is_lit || (sym <= max_plain_sym && off + (uint32_t)tlen <= W && !long_run);
                if (nb >= need && shape && opos + (uint32_t)tlen <= room) {
                    rec = is_lit ? (top << 1) >> (32 - lbits) : rec_match(tlen, off);
                    emit = true;
                    bb <<= need;
                    nb -= need;
                    opos += (uint32_t)tlen;
                } else {
                    // extended format: run / extended-match tokens (second Huffman code without the flag bit + raw bits)
                    if (!is_lit && max_plain_sym < kSymFlush - 1 && (sym == kSymRle || sym == kSymExt)) {
                        const bool is_run = sym == kSymRle;
                        const uint32_t t2 = top << used;  // (used <= 9)
                        const bool long2 = (t2 >> 31) != 0;
                        const uint32_t e2 = smem::ld8(sLut + ((t2 << 1) >> 25));
                        const int hv = long2 ? (int)(e2 & 15u) : 0;
                        const int used2 = long2 ? 1 + (int)(e2 >> 4) : 1;
                        const int tr = is_run ? 4 : 3;
                        const int raw = (hv << tr) + (int)((t2 << used2) >> (32 - tr));
                        const int bits_tok = used + used2 + tr + (is_run ? 0 : wbits);  // <= 9 + 8 + 4 or 7 + 8 + 3 + 10
                        const int xlen = is_run ? raw + 2 : raw + min_pat + 12;
                        const uint32_t xoff = is_run ? 0u : ((top << (used + used2 + tr)) >> (32 - wbits));
                        const bool fits = nb >= bits_tok && opos + (uint32_t)xlen <= room &&
                                          (is_run || (xoff + (uint32_t)xlen <= W && !long_run));
                        had_special = had_special || fits;
                        if (fits && !(!is_run && (k & 31u) == 31u)) {
                            rec = rec_special(is_run ? kRecRun : kRecExt, (uint32_t)xlen);
                            if (!is_run) pending = rec_special(kRecExtOff, xoff);
                            if (is_run && xlen > kRleWindowMax) long_run = true;
                            emit = true;
                            bb <<= bits_tok;
                            nb -= bits_tok;
                            opos += (uint32_t)xlen;
                        } else if (fits) {  // the two records of an extended match stay inside one group: filler first
                            rec = rec_special(kRecFill, 0u);
                            emit = true;
                        }
                    }
                    if (!emit) {
                        // the frame ends, or something the copy phase does not do (same order of checks as the reference's loop)
                        if (nb == 0) {
                            // frame fully consumed: INPUT_EXHAUSTED
                        } else if (opos == cap) {
                            status = kOutputFull;             // bits left but the row is full (decompressor.c:433-463)
                        } else if (nb < (is_lit ? need : used) || (!is_lit && sym <= max_plain_sym && nb < need)) {
                            // incomplete token at the end of the frame: nothing is consumed
                        } else if (!is_lit && sym == kSymFlush && ip == n && nb - used < 8) {
                            // the closing FLUSH of a frame written with write_token (compressor.c:784-794): the decoder
                            // drops the padding bits behind it (decompressor.c:501-514) and the input is exhausted
                        } else {
                            defer = true;                     // FLUSH, OOB, a token that does not fit, a match behind a long run
                        }
                        active = false;
                    }
                }
            }
            // records go through a 32 x 32 tile: row = token index, column = lane; a full tile leaves as 64 bytes per
            // stream, so that the copy phase reads a stream's records back to back
            if (emit) smem::st16(sTile + ((k & 31u) << 6) + 2u * lane, rec);
            k++;
            if ((k & 31u) == 0) {
                __syncwarp();
                uint4 v[4];
                uint32_t *vw = reinterpret_cast<uint32_t *>(v);
#pragma unroll
                for (int r = 0; r < 16; r++)
                    vw[r] = smem::ld16(sTile + ((2 * r) << 6) + 2u * lane) | (smem::ld16(sTile + ((2 * r + 1) << 6) + 2u * lane) << 16);
                uint4 *dstp = reinterpret_cast<uint4 *>(lanescratch + (k - 32));
#pragma unroll
                for (int r = 0; r < 4; r++) dstp[r] = v[r];
                __syncwarp();
            }
        }
        // the last, partial tile
        {
            __syncwarp();
            const uint32_t k0 = k & ~31u;
            if (k0 != k && k0 < (uint32_t)kMaxTok) {
                uint4 v[4];
                uint32_t *vw = reinterpret_cast<uint32_t *>(v);
#pragma unroll
                for (int r = 0; r < 16; r++)
                    vw[r] = smem::ld16(sTile + ((2 * r) << 6) + 2u * lane) | (smem::ld16(sTile + ((2 * r + 1) << 6) + 2u * lane) << 16);
                uint4 *dstp = reinterpret_cast<uint4 *>(lanescratch + k0);
#pragma unroll
                for (int r = 0; r < 4; r++) dstp[r] = v[r];
            }
            __syncwarp();
        }

        // ================= phase 2: copy (warp = stream) =================
        for (int s = 0; s < 32; s++) {
            const uint64_t sid = batch + s;
            if (sid >= a.b.n_streams) break;
            const bool s_defer = __shfl_sync(kFull, (int)defer, s) != 0;
            const bool s_special = __shfl_sync(kFull, (int)had_special, s) != 0;
            const uint32_t s_out = __shfl_sync(kFull, opos, s);
            const int s_status = __shfl_sync(kFull, status, s);
            const uint64_t dict_bits = (uint64_t)reinterpret_cast<uintptr_t>(dict);
            const uint8_t *s_dict = reinterpret_cast<const uint8_t *>((uintptr_t)(
                ((uint64_t)__shfl_sync(kFull, (uint32_t)(dict_bits >> 32), s) << 32) | __shfl_sync(kFull, (uint32_t)dict_bits, s)));
            if (s_defer) {
                if (lane == 0) {
                    a.b.out_sizes[sid] = kDeferred;
                    atomicAdd(&d_split_deferred_total, 1u);
                }
                continue;
            }
            const bool common_dict = s_dict == common;    // its bytes are in shared memory
            const uint16_t *recs = myscratch + s * kMaxTok;
            uint32_t done = 0;                            // output bytes of the groups before this one
            uint32_t rec = recs[lane];                    // (records past the stream's last token are never used: see `valid`)
            for (uint32_t k0 = 0; done < s_out; k0 += 32) {
                const uint32_t nextrec = k0 + 32 < (uint32_t)kMaxTok ? recs[k0 + 32 + lane] : 0u;  // requested a group ahead
                const bool is_match = (rec & 0x8000u) != 0, is_special = s_special && (rec & 0xC000u) == 0xC000u;
                int len0 = is_match ? (int)((rec >> 10) & 15u) + 2 : 1;
                uint32_t off = rec & 1023u;  // bit 16: a run (special records only)
                if (s_special) {  // (never in a v1 frame)
                    const uint32_t skind = (rec >> 12) & 3u;
                    const uint32_t behind = __shfl_down_sync(kFull, rec, 1);  // an extended match's window offset
                    if (is_special) {
                        len0 = skind <= kRecExt ? (int)(rec & 0xFFu) : 0;
                        off = skind == kRecRun ? 0x10000u : (behind & 1023u);
                    }
                }
                // tokens of this group: up to the one that completes the stream's output (the records behind it are stale)
                int incl = len0;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int t = __shfl_up_sync(kFull, incl, d);
                    if (lane >= d) incl += t;
                }
                const bool valid = done + (uint32_t)(incl - len0) < s_out;
                const int len = valid ? len0 : 0;
                const uint32_t dst = done + (uint32_t)(incl - len0);  // this token's output (= window) position
                // a source byte at x < dst is output byte x; at x >= dst it still is the dictionary's.  The token goes at once
                // if all of its bytes were produced before this group, or all still are the (shared-memory) dictionary's
                const bool from_row = off + (uint32_t)len <= done, from_dict = off >= dst && common_dict;
                const bool dep = valid && len0 > 0 && is_match && (is_special || !(from_row || from_dict));
                const bool now = valid && !dep && len0 > 0;
                uint32_t w0 = rec & 0xFFu, w1 = 0, w2 = 0, w3 = 0;
                if (now && is_match) {
                    const uint32_t sa = (from_row ? sRow : sDict) + off, q = sa & ~3u;
                    const int sh = (int)(sa << 3);
                    const uint32_t a0 = smem::ld32(q), a1 = smem::ld32(q + 4), a2 = smem::ld32(q + 8), a3 = smem::ld32(q + 12), a4 = smem::ld32(q + 16);
                    w0 = __funnelshift_r(a0, a1, sh);
                    w1 = __funnelshift_r(a1, a2, sh);
                    w2 = __funnelshift_r(a2, a3, sh);
                    w3 = __funnelshift_r(a3, a4, sh);
                }
                // (the 20 bytes read may reach into what other lanes of the group write — bytes behind the token's own, never
                // used; the barrier puts every read in front of every write all the same)
                __syncwarp();
                if (now) {
                    const uint32_t da = sRow + dst;
                    smem::st8(da, w0);
                    if (len > 1) smem::st8(da + 1, w0 >> 8);
                    if (len > 2) smem::st8(da + 2, w0 >> 16);
                    if (len > 3) smem::st8(da + 3, w0 >> 24);
                    if (len > 4) smem::st8(da + 4, w1);
                    if (len > 5) smem::st8(da + 5, w1 >> 8);
                    if (len > 6) smem::st8(da + 6, w1 >> 16);
                    if (len > 7) smem::st8(da + 7, w1 >> 24);
                    if (len > 8) smem::st8(da + 8, w2);
                    if (len > 9) smem::st8(da + 9, w2 >> 8);
                    if (len > 10) smem::st8(da + 10, w2 >> 16);
                    if (len > 11) smem::st8(da + 11, w2 >> 24);
                    if (len > 12) smem::st8(da + 12, w3);
                    if (len > 13) smem::st8(da + 13, w3 >> 8);
                    if (len > 14) smem::st8(da + 14, w3 >> 16);
                    if (len > 15) smem::st8(da + 15, w3 >> 24);
                }
                uint32_t deps = __ballot_sync(kFull, dep);
                __syncwarp();
                while (deps) {  // in order; the lanes share the token's bytes (all reads before the writes)
                    const int j = __ffs(deps) - 1;
                    deps &= deps - 1;
                    const uint32_t jdst = __shfl_sync(kFull, dst, j), joff = __shfl_sync(kFull, off, j);
                    const int jlen = __shfl_sync(kFull, len, j);
                    if (jlen <= 32 && joff < 0x10000u) {  // the usual case: a plain match fed by this group's bytes
                        uint32_t b = 0;
                        if (lane < jlen) {
                            const uint32_t x = joff + (uint32_t)lane;
                            b = x < jdst ? smem::ld8(sRow + x) : (common_dict ? smem::ld8(sDict + x) : (uint32_t)__ldg(s_dict + x));
                        }
                        __syncwarp();
                        if (lane < jlen) smem::st8(sRow + jdst + lane, b);
                    } else {  // runs, extended matches of more than 32 bytes: rare, kept out of line
                        copy_long_token(sRow, sDict, common_dict ? nullptr : s_dict, jdst, joff, jlen, lane,
                                        1u << __shfl_sync(kFull, wbits, s));
                    }
                    __syncwarp();
                }
                done += (uint32_t)__shfl_sync(kFull, valid ? incl : 0, 31 - __clz(__ballot_sync(kFull, valid)));
                rec = nextrec;
            }
            __syncwarp();
            // the row leaves: 128-bit stores where the layout allows, bytes otherwise
            uint8_t *out = a.b.out + sid * a.b.out_stride;
            if (a.aligned_io) {
                const uint32_t n16 = s_out & ~15u;
                for (uint32_t o = lane * 16; o < n16; o += 512)
                    *reinterpret_cast<uint4 *>(out + o) = *reinterpret_cast<const uint4 *>(row + o);
                for (uint32_t o = n16 + lane; o < s_out; o += 32) out[o] = row[o];
            } else {
                for (uint32_t o = lane; o < s_out; o += 32) out[o] = row[o];
            }
            if (lane == 0) {
                a.b.out_sizes[sid] = s_out;
                if (a.b.status) a.b.status[sid] = (int8_t)s_status;
            }
            __syncwarp();  // the row is reused by the next stream
        }
        __syncwarp();
    }
}

}  // namespace

#ifndef TB_EMU
bool launch_split_decompress_batch(const uint8_t *d_seed, const uint8_t *d_custom, int window_bits_max, const BatchArgs &b,
                                   cudaStream_t st) {
    if (window_bits_max > 10) return false;
    if (b.out_stride > 0xFFFFFFF0ull) return false;
    if (b.n_streams == 0) return true;
    static int blocks_per_sm = 0, sms = 0;
    const size_t smem = kSplitSmem;
    if (!blocks_per_sm) {
        cudaFuncSetAttribute(k_split_decompress, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_split_decompress, kSplitWarps * 32, smem);
        if (blocks_per_sm < 1) blocks_per_sm = 1;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const uint64_t per_block = kSplitWarps * 32;
    const uint64_t want = (b.n_streams + per_block - 1) / per_block;
    const uint64_t persistent = (uint64_t)sms * blocks_per_sm;
    const unsigned grid = (unsigned)(want < persistent ? want : persistent);
    // token records: stream-ordered scratch (the kernel before a later reuse has finished by then)
    uint16_t *scratch = nullptr;
    const size_t scratch_bytes = (size_t)grid * kSplitWarps * kMaxTok * 32 * sizeof(uint16_t);
    if (cudaMallocAsync(&scratch, scratch_bytes, st) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    SplitDecArgs a;
    a.b = b;
    a.seed = d_seed;
    a.custom = d_custom;
    a.window_bits_max = window_bits_max;
    a.aligned_io = ((b.out_stride & 15) == 0 && (reinterpret_cast<uintptr_t>(b.out) & 15) == 0) ? 1 : 0;
    a.scratch = scratch;
    k_split_decompress<<<grid, kSplitWarps * 32, smem, st>>>(a);
    count_launch();
    cudaFreeAsync(scratch, st);
    // second pass: the lane-per-stream kernel picks up the streams marked kDeferred (usually none)
    static unsigned int *h_seen = nullptr;  // pinned mirror of d_split_deferred_total
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
    const bool ok = launch_fast_decompress_batch(d_seed, d_custom, window_bits_max, b, st, /*only_deferred=*/true,
                                                 /*small_grid=*/!expect_work);
    if (h_seen) cudaMemcpyFromSymbolAsync(h_seen, d_split_deferred_total, sizeof *h_seen, 0, cudaMemcpyDeviceToHost, st);
    return ok;
}
#endif  // TB_EMU

}  // namespace tb
