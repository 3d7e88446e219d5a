// History-walk batch compressor, v1 format, ANY window (8..15) and ANY stream length: one CTA per stream.  The
// compressor of BASELINE.json config 3 (64 KiB streams, window 15), of the frames config 4 decodes (4 KiB, window 10)
// and of every class of config 5, since round 2.
//
// Where it comes from.  The segment-walk kernel (walk_compress.cu) evaluates find_best_match
// (compressor_find_match_desktop.c:82-167) only at the offsets the greedy parse (compressor.c:625-657) reaches, but it
// needs the whole stream and the dictionary side by side in shared memory: streams no longer than the window, windows
// up to 1 KiB.  What makes it work — in the v1 format every consumed byte is written to the window in order
// (compressor.c:652-657), so the window a poll at input offset q sees does not depend on the parse — holds for any
// length: ring position x holds the last byte written there,
//     window_q[x] = V[max { v < W + q : v = x (mod W) }],   V = dictionary ++ input  ("virtual time": input offset p is v = W + p)
// so a candidate is a point in time, alive for W steps (until its ring position is overwritten), and the classic
// time-ordered hash chain (zlib's prev[]) enumerates the candidates of q newest first; the walk stops at distance W.
//
// Data structures, all in shared memory, all indexed by time modulo R (C = chunk of input offsets in flight; R = the
// multiple of C that holds W + C + 32 bytes; input offset 0 is position 0, the dictionary sits in [R - W, R)):
//   HB  the last W + C (+ 32 lookahead) bytes of V — a match source is one unaligned 16-byte read; 32 bytes of
//       mirror behind the end make reads across the wrap contiguous;
//   LK  u16 per byte: distance to the previous position with the same bigram hash (0: none within W);
//   HD  u16 per hash: newest position + 1 (P1 only).
// The dictionary's bytes, links and heads are computed once per launch (k_hwalk_dict) and copied in per stream.
//
// Per chunk of C offsets:
//   P0  the chunk (+ 32 bytes of lookahead) by 128-bit loads;
//   P1  links of the chunk's offsets: __match_any_sync inside blocks of 32 offsets, the head table across blocks
//       (the warps take their turns at the table in time order; everything else runs in parallel), plus one bit per
//       offset "has a candidate at all";
//   P2  segment walk as in walk_compress.cu: a lane walks the greedy parse through a segment of SEG offsets from a
//       guessed entry, running the candidate chain of an offset only when the walk stands on it; then every segment
//       takes its left neighbour's exit as its entry and re-walks until it meets its old path; repeat until no entry
//       changes.  Segments are handed out through a queue, so a CTA may have fewer lanes than segments;
//   P4  static-Huffman bit pack, one lane per segment: exclusive scan of the segments' bit counts, tokens ORed into an
//       MSb-first staging line (write_to_bit_buffer / partial_flush, compressor.c:49-75), whole words leave, the
//       partial word is carried into the next chunk.
//
// A match candidate at time distance D from q (ring position r = (q - D) mod W) supplies its own bytes up to q, then
// the window continues with the bytes one lap older (ring positions >= q mod W have not been overwritten yet), and
// never runs past the end of the window buffer: length <= W - r (compressor_find_match_desktop.c:59).  Longest match
// wins, then the lowest ring position (:65-68) — a max over keys (len << 16 | ~r), order-independent.
//
// Streams whose chains are pathologically long (runs, short periods) are marked kDeferred and left to the bitmap
// kernels (fast_compress.cu / wide_compress.cu), whose cost does not depend on the data.
#include <stdio.h>
#include <stdlib.h>

#include "../tb_wire.h"
#include "tb_cuda.h"
#include "tb_device_common.cuh"
#include "tb_smem.cuh"
#include "history_ring.cuh"

namespace tb {

namespace {

__device__ unsigned int d_hwalk_deferred_total = 0;  // streams deferred so far (cumulative)

struct HwalkLayout {
    uint32_t W, C, R, HS, nseg;
    uint32_t oHB, oLK, oHD, oBEST, oHAS, oPATH, oSBIT, oEXIT, oENTRY, oSTAGE, oMISC, total;
};

__host__ __device__ inline HwalkLayout hwalk_layout(int wbits, int cbits, int hbits, int seg) {
    HwalkLayout L;
    L.W = 1u << wbits;
    L.C = 1u << cbits;
    L.R = ring_size(L.W, L.C);
    L.HS = 1u << hbits;
    L.nseg = L.C / (uint32_t)seg;
    L.oHB = 0;
    L.oLK = up16(L.R + kPad);
    L.oHD = L.oLK + 2u * L.R;
    L.oBEST = L.oHD + 2u * L.HS;
    L.oHAS = L.oBEST + 2u * L.C;
    L.oPATH = L.oHAS + up16(L.C / 8u);
    L.oSBIT = L.oPATH + up16(4u * L.nseg);
    L.oEXIT = L.oSBIT + up16(4u * L.nseg);
    L.oENTRY = L.oEXIT + 2u * up16(L.nseg);  // (two copies of the exits: a round reads the previous round's)
    L.oSTAGE = L.oENTRY + up16(L.nseg);
    L.oMISC = L.oSTAGE + up16(4u * (L.C * 9u / 32u + 4u));
    L.total = L.oMISC + 64u;
    return L;
}

// MISC words
enum { M_QUEUE = 0, M_BAIL = 1, M_PAIRS = 2, M_MISFIT = 3, M_CARRY = 4 };

struct HwalkArgs {
    BatchArgs b;
    const uint8_t *dict;          // W bytes
    const uint16_t *dict_links;   // W entries (k_hwalk_dict)
    const uint16_t *dict_heads;   // HS entries
    int window_bits, literal, flags, write_token;
    int chunk_bits, hash_bits;
    int budget;     // lock-step iterations a warp may spend on one round of one chunk before the stream is given up
    int max_pairs;  // same-bigram pairs inside blocks of 32 offsets, per chunk offset x 16, above which it is given up
};

// P1.  Links the bigrams of chunk offsets [0, cn) — time-linear positions pvs + p — into the chains; offsets >= nbig have
// no bigram (the stream ends).  All warps of the CTA; `hasw` (or 0) receives one bit per offset: has a candidate.
__device__ __forceinline__ uint32_t build_links(uint32_t sHB, uint32_t sLK, uint32_t sHD, uint32_t sHAS, int pvs, int cn, int nbig,
                                                int W, int R, int hbits, uint32_t qm0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t pairs = 0;
    for (int base = 0; base < cn; base += 32 * nwarps) {
        const int p = base + warp * 32 + lane;
        const bool inb = p < cn, valid = p < nbig;
        const uint32_t phys = (uint32_t)(pvs + p);
        const uint32_t b0 = inb ? smem::ld8(sHB + phys) : 0u, b1 = inb ? smem::ld8(sHB + phys + 1u) : 0u;
        const uint32_t h = valid ? bigram_hash(b0 | (b1 << 8), hbits) : (0x10000u | (uint32_t)lane);  // no bigram: matches nobody
        const uint32_t peers = __match_any_sync(kFull, h);
        const uint32_t lower = peers & lt;
        uint32_t link = lower ? (uint32_t)(lane - (31 - __clz(lower))) : 0u;
        const bool first = valid && !lower, last = valid && (peers >> lane) == 1u;
        pairs += valid ? __popc(lower) : 0;
        // the head table, in time order: one warp after the other
        for (int g = 0; g < nwarps; g++) {
            if (warp == g) {
                if (first) {
                    const uint32_t hv = smem::ld16(sHD + 2u * h);
                    if (hv) {
                        int d = (int)phys + 1 - (int)hv;
                        if (d <= 0) d += R;
                        link = d <= W ? (uint32_t)d : 0u;
                    }
                }
                __syncwarp();
                if (last) smem::st16(sHD + 2u * h, phys + 1u);
                __syncwarp();
            }
            if (nwarps > 1) __syncthreads();
        }
        if (inb) smem::st16(sLK + 2u * phys, valid ? link : 0u);
        if (sHAS) {
            // window position q-1 holds input[q-1] followed by bytes one lap older: no chain covers that bigram
            const uint32_t qmr = (qm0 + (uint32_t)p) & (uint32_t)(W - 1);
            bool strad = false;
            if (valid && qmr != 0u) {
                const uint32_t pprev = phys ? phys - 1u : (uint32_t)R - 1u;
                int pold = (int)phys - W;
                if (pold < 0) pold += R;
                strad = smem::ld8(sHB + pprev) == b0 && smem::ld8(sHB + (uint32_t)pold) == b1;
            }
            const uint32_t m = __ballot_sync(kFull, valid && (link != 0u || strad));
            if (lane == 0 && base + warp * 32 < cn) smem::st32(sHAS + 4u * (uint32_t)((base >> 5) + warp), m);
        }
    }
    return pairs;
}

template <int SEG>
__global__ void __maxnreg__(80) k_hwalk_compress(HwalkArgs a) {
#ifndef TB_EMU
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t *sm = smem_raw;
#else
    uint8_t *sm = emu::g_smem;
#endif
    constexpr uint32_t kSegMask = SEG == 32 ? kFull : ((1u << (SEG & 31)) - 1u);
    const int tid = threadIdx.x, lane = tid & 31, T = blockDim.x;
    const int wbits = a.window_bits, lbits = a.literal;
    const int W = 1 << wbits, C = 1 << a.chunk_bits;
    const HwalkLayout Lo = hwalk_layout(wbits, a.chunk_bits, a.hash_bits, SEG);
    const int R = (int)Lo.R;
    const int min_pat = min_pattern_size(wbits, lbits);
    const int max_len = min_pat + 13;
    uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
#ifndef TB_EMU
    asm volatile("" : "+r"(sbase));
#endif
    const uint32_t sHB = sbase + Lo.oHB, sLK = sbase + Lo.oLK, sHD = sbase + Lo.oHD, sBEST = sbase + Lo.oBEST;
    const uint32_t sHAS = sbase + Lo.oHAS, sPATH = sbase + Lo.oPATH, sSBIT = sbase + Lo.oSBIT, sEXIT = sbase + Lo.oEXIT;
    const uint32_t sENTRY = sbase + Lo.oENTRY, exit_stride = up16(Lo.nseg);
    uint32_t sExitNow = sEXIT;  // the copy of the exits the last round wrote
    uint32_t *stage = reinterpret_cast<uint32_t *>(sm + Lo.oSTAGE);
    uint32_t *misc = reinterpret_cast<uint32_t *>(sm + Lo.oMISC);
    const int stage_words = (int)(Lo.C * 9u / 32u + 4u);

    for (uint64_t stream = blockIdx.x; stream < a.b.n_streams; stream += gridDim.x) {
        const uint8_t *src = a.b.in + stream * a.b.in_stride;
        const int N = a.b.in_sizes ? (int)a.b.in_sizes[stream] : (int)a.b.in_stride;
        const int row16 = (int)a.b.in_stride;  // readable bytes of the row (a multiple of 16)
        uint32_t *out32 = reinterpret_cast<uint32_t *>(a.b.out + stream * a.b.out_stride);

        // ---- stream start: the dictionary is the history (times 0 .. W-1) ----------------------------------------
        __syncthreads();
        for (int i = tid; i < W / 16; i += T)
            reinterpret_cast<uint4 *>(sm + Lo.oHB + R - W)[i] = __ldg(reinterpret_cast<const uint4 *>(a.dict) + i);
        for (int i = tid; i < W / 8; i += T)
            reinterpret_cast<uint4 *>(sm + Lo.oLK + 2 * (R - W))[i] = __ldg(reinterpret_cast<const uint4 *>(a.dict_links) + i);
        for (int i = tid; i < (int)Lo.HS / 8; i += T)
            reinterpret_cast<uint4 *>(sm + Lo.oHD)[i] = __ldg(reinterpret_cast<const uint4 *>(a.dict_heads) + i);
        for (int i = tid; i < stage_words; i += T) stage[i] = 0u;
        if (tid == 0) {
            const uint32_t header = ((uint32_t)(wbits - 8) << 5) | ((uint32_t)(lbits - 5) << 3) |
                                    ((a.flags & TB_F_CUSTOM_DICT) ? 4u : 0u) | ((a.flags & TB_F_DICT_RESET) ? 1u : 0u);
            stage[0] = stream_appends(a.flags, stream) ? kAppendStart : header << 24;
        }
        uint32_t carry = (a.flags & TB_F_DICT_RESET) ? 16u : 8u;  // bits waiting in the staging line's first word(s)
        uint32_t ow = 0;          // whole words already written to the output row
        int res = kOk;
        int pvs = 0;              // time-linear position of the chunk's first offset
        int chunk_entry = 0;      // where the token that straddles the chunk boundary ends
        bool bail = false;
        __syncthreads();

        for (int cs = 0; cs < N && res == kOk; cs += C) {
            const int cn = N - cs < C ? N - cs : C;
            const int nbig = N - cs - 1 < cn ? N - cs - 1 : cn;  // offsets that have a bigram
            const int nseg = (cn + SEG - 1) / SEG;
            const uint32_t qm0 = (uint32_t)cs & (uint32_t)(W - 1);

            // ---- P0: the chunk and its lookahead --------------------------------------------------------------------
            for (int i = tid; i < (C + kPad) / 16; i += T) {
                const int goff = cs + 16 * i;
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (goff < row16 && goff < N + kPad) v = __ldg(reinterpret_cast<const uint4 *>(src + goff));
                const int ph = pvs + 16 * i;
                *reinterpret_cast<uint4 *>(sm + Lo.oHB + ph) = v;                    // (ph < R + 32: the mirror is writable)
                if (ph >= R) *reinterpret_cast<uint4 *>(sm + Lo.oHB + ph - R) = v;   // lookahead past the wrap
                if (ph < kPad) *reinterpret_cast<uint4 *>(sm + Lo.oHB + R + ph) = v; // mirror of the first bytes
            }
            if (tid == 0) {
                misc[M_PAIRS] = 0u;
                misc[M_MISFIT] = 0xFFFFFFFFu;
            }
            __syncthreads();

            // ---- P1: chain links of the chunk, "has a candidate" bits -----------------------------------------------
            {
                const uint32_t pairs = __reduce_add_sync(kFull, build_links(sHB, sLK, sHD, sHAS, pvs, cn, nbig, W, R, a.hash_bits, qm0));
                if (lane == 0 && pairs) atomicAdd(&misc[M_PAIRS], pairs);
            }
            __syncthreads();
            if ((uint64_t)misc[M_PAIRS] * 16u > (uint64_t)a.max_pairs * (uint32_t)cn) {
                bail = true;  // runs / short periods: the candidate walk is the slower way
                break;
            }

            // ---- P2: segment walk -----------------------------------------------------------------------------------
            for (int round = 0;; round++) {
                if (tid == 0) misc[M_QUEUE] = 0u;
                // exits: every round writes its own copy and reads the one before it (a segment handled this round
                // never sees an exit written this round)
                const uint32_t sExitPrev = sEXIT + ((uint32_t)(round + 1) & 1u) * exit_stride;
                sExitNow = sEXIT + ((uint32_t)round & 1u) * exit_stride;
                __syncthreads();
                bool changed = false, active = true, adv = true, haveseg = false, gave_up = false;
                int s = 0, segbase = 0, nvalid = 0, pn = 0, p = 0, L = 0, pq = 0, D = 0, ca = 0;
                uint32_t hasmask = 0, oldpath = 0, newmask = 0, validmask = 0, qmr = 0, ln = 0, bestkey = 0;
                uint32_t la[4] = {0, 0, 0, 0};
                int budget = a.budget;
                while (__any_sync(kFull, active)) {
                    if (--budget <= 0) {
                        gave_up = true;
                        break;
                    }
                    if (active && adv) {
                        if (!haveseg) {
                            s = (int)atomicAdd(&misc[M_QUEUE], 1u);
                            if (s >= nseg) {
                                active = false;
                            } else {
                                segbase = s * SEG;
                                nvalid = cn - segbase >= SEG ? SEG : cn - segbase;
                                validmask = nvalid >= 32 ? kFull : ((1u << nvalid) - 1u);
                                const int entry = s == 0 ? chunk_entry : (round == 0 ? 0 : (int)smem::ld8(sExitPrev + (uint32_t)s - 1u));
                                const bool same = round > 0 && entry == (int)smem::ld8(sENTRY + (uint32_t)s);
                                smem::st8(sExitNow + (uint32_t)s, round > 0 ? smem::ld8(sExitPrev + (uint32_t)s) : 0u);  // unless the walk finds a new one
                                if (!same) {
                                    smem::st8(sENTRY + (uint32_t)s, (uint32_t)entry);
                                    changed = round > 0;
                                    oldpath = round > 0 ? smem::ld32(sPATH + 4u * (uint32_t)s) : 0u;
                                    if (entry >= nvalid || ((oldpath >> entry) & 1u)) {
                                        // nothing new to walk: what the old walk visited before the entry is void
                                        smem::st32(sPATH + 4u * (uint32_t)s, oldpath & __funnelshift_lc(0u, kFull, entry));
                                    } else {
                                        haveseg = true;
                                        const uint32_t hw = smem::ld32(sHAS + 4u * (uint32_t)(segbase >> 5));
                                        hasmask = SEG == 32 ? hw : ((hw >> (segbase & 31)) & kSegMask);
                                        newmask = 0u;
                                        pn = entry;
                                    }
                                }
                            }
                        }
                        if (haveseg) {
                            // ---- between offsets: continue the walk at pn.  Offsets without a candidate are literals
                            // (one step each): the next stop is the first offset with a candidate or on the old path ----
                            const uint32_t hi = __funnelshift_lc(0u, kFull, pn);  // offsets >= pn (none if pn >= 32)
                            const uint32_t stops = (hasmask | oldpath) & hi;
                            const int t = __ffs(stops) - 1;
                            const uint32_t below = (1u << t) - 1u;  // (stops == 0: unused)
                            if (!stops) {  // literals to the end of the segment (or the token jumped past it)
                                smem::st32(sPATH + 4u * (uint32_t)s, newmask | (hi & validmask));
                                smem::st8(sExitNow + (uint32_t)s, (uint32_t)(pn > SEG ? pn - SEG : 0));
                                haveseg = false;
                            } else if ((oldpath >> t) & 1u) {  // merged with the previous walk: same path and exit from here on
                                smem::st32(sPATH + 4u * (uint32_t)s, newmask | (hi & below) | (oldpath & ~below));
                                haveseg = false;
                            } else {
                                newmask |= hi & below;
                                p = t;
                                const int rel = segbase + t;
                                pq = pvs + rel;
                                qmr = (qm0 + (uint32_t)rel) & (uint32_t)(W - 1);
                                L = N - cs - rel < max_len ? N - cs - rel : max_len;
                                smem::load16(sHB + (uint32_t)pq, la);
                                bestkey = 0u;
                                const uint32_t d1 = smem::ld16(sLK + 2u * (uint32_t)pq);
                                const uint32_t pprev = pq ? (uint32_t)pq - 1u : (uint32_t)R - 1u;
                                const bool strad = qmr != 0u && d1 != 1u && smem::ld8(sHB + pprev) == (la[0] & 0xFFu);
                                if (strad) {
                                    D = 1;
                                    ca = (int)pprev;
                                    ln = d1 ? d1 - 1u : 0u;
                                } else {
                                    D = (int)d1;
                                    ca = pq - (int)d1;
                                    if (ca < 0) ca += R;
                                    ln = d1 ? smem::ld16(sLK + 2u * (uint32_t)ca) : 0u;
                                }
                                adv = false;
                            }
                        }
                    }
                    if (active && !adv) {
                        // ---- up to two candidates of the offset (the link behind the second one is loaded ahead) ----
                        const bool alive_a = D != 0;
                        const int Db = D + (int)ln;
                        const bool alive_b = alive_a && ln != 0u && Db <= W;
                        int cb = ca - (int)ln;
                        if (cb < 0) cb += R;
                        const uint32_t l3 = alive_b ? smem::ld16(sLK + 2u * (uint32_t)cb) : 0u;
                        if (alive_a) {
                            const uint32_t key = eval_candidate(D, (uint32_t)ca, la, L, qmr, sHB, pq, W, R);
                            bestkey = key > bestkey ? key : bestkey;
                        }
                        if (alive_b) {
                            const uint32_t key = eval_candidate(Db, (uint32_t)cb, la, L, qmr, sHB, pq, W, R);
                            bestkey = key > bestkey ? key : bestkey;
                        }
                        const int Dc = Db + (int)l3;
                        if (alive_b && l3 != 0u && Dc <= W) {
                            D = Dc;
                            ca = cb - (int)l3;
                            if (ca < 0) ca += R;
                            ln = smem::ld16(sLK + 2u * (uint32_t)ca);
                        } else {  // chain exhausted: the match at this offset is known
                            const int len = (int)(bestkey >> 16);
                            const bool is_match = len >= min_pat;
                            if (is_match) smem::st16(sBEST + 2u * (uint32_t)(segbase + p), (~bestkey) & 0xFFFFu);
                            newmask |= 1u << p;
                            pn = p + (is_match ? len : 1);
                            adv = true;
                        }
                    }
                }
                const int any = __syncthreads_or(changed ? 1 : 0);
                if (__syncthreads_or(gave_up ? 1 : 0)) {
                    bail = true;
                    break;
                }
                if (round > 0 && !any) break;
            }
            if (bail) break;

            // ---- P4: bit pack, one lane per segment -----------------------------------------------------------------
            // (tokens of a segment: the set bits of its path; a token's length is the distance to the next token)
            uint32_t cut = 0xFFFFFFFFu;  // chunk offset of the first literal that does not fit (compressor.c:629-631)
            if (lbits < 8) {
                for (int s = tid; s < nseg; s += T) {
                    const int segbase = s * SEG;
                    const int nvalid = cn - segbase >= SEG ? SEG : cn - segbase;
                    const bool last_seg = cs + segbase + SEG >= N;
                    const int segend = last_seg ? nvalid : SEG + (int)smem::ld8(sExitNow + (uint32_t)s);
                    uint32_t m = smem::ld32(sPATH + 4u * (uint32_t)s) & (nvalid >= 32 ? kFull : ((1u << nvalid) - 1u));
                    while (m) {
                        const int t = __ffs(m) - 1;
                        m &= m - 1;
                        const int next = m ? __ffs(m) - 1 : segend;
                        if (next - t == 1 && (smem::ld8(sHB + (uint32_t)(pvs + segbase + t)) >> lbits)) {
                            atomicMin(&misc[M_MISFIT], (uint32_t)(segbase + t));
                            break;
                        }
                    }
                }
                __syncthreads();
                cut = misc[M_MISFIT];
            }
            for (int pass = 0; pass < 2; pass++) {
                for (int s = tid; s < nseg; s += T) {
                    const int segbase = s * SEG;
                    const int nvalid = cn - segbase >= SEG ? SEG : cn - segbase;
                    const bool last_seg = cs + segbase + SEG >= N;
                    const int segend = last_seg ? nvalid : SEG + (int)smem::ld8(sExitNow + (uint32_t)s);
                    uint32_t m = smem::ld32(sPATH + 4u * (uint32_t)s) & (nvalid >= 32 ? kFull : ((1u << nvalid) - 1u));
                    uint32_t pos = pass ? carry + smem::ld32(sSBIT + 4u * (uint32_t)s) : 0u;
                    while (m) {
                        const int t = __ffs(m) - 1;
                        m &= m - 1;
                        if ((uint32_t)(segbase + t) >= cut) break;
                        const int next = m ? __ffs(m) - 1 : segend;
                        const int len = next - t;
                        uint32_t bits;
                        int nb;
                        if (len == 1) {
                            bits = (1u << lbits) | smem::ld8(sHB + (uint32_t)(pvs + segbase + t));
                            nb = lbits + 1;
                        } else {
                            const int sym = len - min_pat;
                            bits = ((uint32_t)kHuff.code[sym] << wbits) | smem::ld16(sBEST + 2u * (uint32_t)(segbase + t));
                            nb = (int)kHuff.bits[sym] + wbits;
                        }
                        if (pass) {
                            const uint32_t wi = pos >> 5, o = pos & 31u;
                            const uint64_t sv = (uint64_t)bits << (64 - nb - (int)o);
                            atomicOr(&stage[wi], (uint32_t)(sv >> 32));
                            if ((uint32_t)sv) atomicOr(&stage[wi + 1], (uint32_t)sv);
                        }
                        pos += (uint32_t)nb;
                    }
                    if (!pass) smem::st32(sSBIT + 4u * (uint32_t)s, pos);
                }
                __syncthreads();
                if (!pass) {
                    // exclusive scan of the segments' bit counts by warp 0; the chunk's total goes to M_CARRY
                    if (tid < 32) {
                        uint32_t run = 0;
                        for (int b0 = 0; b0 < nseg; b0 += 32) {
                            const int i = b0 + lane;
                            const uint32_t v = i < nseg ? smem::ld32(sSBIT + 4u * (uint32_t)i) : 0u;
                            uint32_t incl = v;
#pragma unroll
                            for (int d = 1; d < 32; d <<= 1) {
                                const uint32_t u = __shfl_up_sync(kFull, incl, d);
                                if (lane >= d) incl += u;
                            }
                            if (i < nseg) smem::st32(sSBIT + 4u * (uint32_t)i, run + incl - v);
                            run += __shfl_sync(kFull, incl, 31);
                        }
                        if (lane == 0) misc[M_CARRY] = run;
                    }
                    __syncthreads();
                }
            }
            // whole words leave; the partial word is carried
            {
                const uint32_t total = carry + misc[M_CARRY];
                const uint32_t tw = total >> 5;
                for (uint32_t i = (uint32_t)tid; i < tw; i += (uint32_t)T) out32[ow + i] = __byte_perm(stage[i], 0, 0x0123);
                const uint32_t part = stage[tw];
                __syncthreads();
                for (uint32_t i = (uint32_t)tid; i < tw + 2u; i += (uint32_t)T) stage[i] = i == 0u ? part : 0u;
                ow += tw;
                carry = total & 31u;
                if (cut != 0xFFFFFFFFu) res = kExcessBits;
                chunk_entry = (int)smem::ld8(sExitNow + (uint32_t)(nseg - 1));
            }
            pvs += C;
            if (pvs == R) pvs = 0;
            __syncthreads();
        }

        if (bail) {
            if (tid == 0) {
                a.b.out_sizes[stream] = kDeferred;
                atomicAdd(&d_hwalk_deferred_total, 1u);
            }
            continue;
        }

        // ---- stream end: FLUSH token / padding (compressor.c:784-810) -------------------------------------------
        uint32_t nbits = carry;  // bits in the staging line (beyond the `ow` words already out)
        __syncthreads();
        uint32_t tail_bytes;
        if (res == kOk) {
            if (ends_with_flush(a.write_token, nbits, a.flags, stream, (uint64_t)N)) {
                if (tid == 0) {
                    const uint32_t wi = nbits >> 5, o = nbits & 31u;
                    const uint64_t sv = (uint64_t)kHuff.code[kSymFlush] << (64 - kHuff.bits[kSymFlush] - (int)o);
                    stage[wi] |= (uint32_t)(sv >> 32);
                    stage[wi + 1] |= (uint32_t)sv;
                }
                nbits += kHuff.bits[kSymFlush];
            }
            tail_bytes = (nbits + 7u) >> 3;
        } else {
            tail_bytes = nbits >> 3;  // the reference has drained whole bytes of everything before the failing poll
        }
        __syncthreads();
        if ((uint32_t)tid < tail_bytes)
            reinterpret_cast<uint8_t *>(out32 + ow)[tid] = (uint8_t)(stage[tid >> 2] >> (24 - 8 * (tid & 3)));
        if (tid == 0) {
            a.b.out_sizes[stream] = 4u * ow + tail_bytes;
            if (a.b.status) a.b.status[stream] = (int8_t)res;
        }
    }
}

// Links and heads of the dictionary (times 0 .. W-1), once per launch: one CTA.
// `bias` = R - W: where the dictionary sits in the compressor's history (heads hold positions + 1).
__global__ void __launch_bounds__(512) k_hwalk_dict(const uint8_t *dict, int wbits, int hbits, uint32_t bias, uint16_t *links,
                                                    uint16_t *heads) {
#ifndef TB_EMU
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t *sm = smem_raw;
#else
    uint8_t *sm = emu::g_smem;
#endif
    const int W = 1 << wbits, HS = 1 << hbits, tid = threadIdx.x, T = blockDim.x;
    const uint32_t oLK = up16((uint32_t)W + kPad), oHD = oLK + 2u * (uint32_t)W;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
    for (int i = tid; i < (W + kPad) / 4; i += T)
        reinterpret_cast<uint32_t *>(sm)[i] = i < W / 4 ? reinterpret_cast<const uint32_t *>(dict)[i] : 0u;
    for (int i = tid; i < HS / 2; i += T) reinterpret_cast<uint32_t *>(sm + oHD)[i] = 0u;
    __syncthreads();
    build_links(sbase, sbase + oLK, sbase + oHD, 0u, 0, W, W - 1, W, 2 * W, hbits, 0u);
    __syncthreads();
    for (int i = tid; i < W; i += T) links[i] = reinterpret_cast<const uint16_t *>(sm + oLK)[i];
    for (int i = tid; i < HS; i += T) {
        const uint32_t h = reinterpret_cast<const uint16_t *>(sm + oHD)[i];
        heads[i] = (uint16_t)(h ? h + bias : 0u);
    }
}

struct HwalkPlan {
    int cbits, hbits, seg, threads;
};

// chunk, hash table, segment length and CTA size per window (tuning hook: TAMP_B200_HWALK_PLAN="cbits,hbits,seg,threads")
inline HwalkPlan hwalk_plan(int wbits) {
#ifndef TB_EMU
    if (const char *e = getenv("TAMP_B200_HWALK_PLAN")) {
        HwalkPlan p;
        if (sscanf(e, "%d,%d,%d,%d", &p.cbits, &p.hbits, &p.seg, &p.threads) == 4 && p.cbits >= 8 && p.cbits <= 14 && p.hbits >= 8 &&
            p.hbits <= 14 && (p.seg == 16 || p.seg == 32) && p.threads >= 32 && p.threads <= 1024 && p.threads % 32 == 0 &&
            hwalk_layout(wbits, p.cbits, p.hbits, p.seg).total <= 227u * 1024u)
            return p;
    }
#endif
    switch (wbits) {
        case 8: return {10, 11, 32, 32};
        case 9: return {10, 11, 32, 32};
        case 10: return {11, 11, 32, 64};
        case 11: return {11, 12, 16, 128};
        case 12: return {12, 12, 16, 256};
        case 13: return {12, 13, 16, 256};
        case 14: return {12, 13, 16, 256};
        default: return {13, 13, 16, 512};
    }
}

}  // namespace

#ifndef TB_EMU
bool launch_hwalk_compress_batch(const CompBatchConf &cf, const uint8_t *d_dict, const BatchArgs &b, cudaStream_t st) {
    if (cf.window < 8 || cf.window > 15) return false;
    if (cf.flags & (TB_F_EXTENDED | TB_F_LAZY)) return false;  // v1 greedy only
    if (b.in_offsets) return false;                          // strided layout only
    if ((b.in_stride & 15) || ((uintptr_t)b.in & 15) || (b.out_stride & 3) || ((uintptr_t)b.out & 3)) return false;
    if (b.in_stride > (1u << 30)) return false;
    if ((uintptr_t)d_dict & 15) return false;
    const uint64_t bound = 2 + (b.in_stride * (uint64_t)(cf.literal + 1) + 7) / 8 + 6;
    if (b.out_stride < ((bound + 3) & ~3ull)) return false;  // never OUTPUT_FULL in this kernel
    if (b.n_streams == 0) return true;

    const HwalkPlan plan = hwalk_plan(cf.window);
    const HwalkLayout Lo = hwalk_layout(cf.window, plan.cbits, plan.hbits, plan.seg);
    const int W = 1 << cf.window, HS = 1 << plan.hbits;
    // the dictionary's links and heads: stream-ordered scratch
    uint16_t *tables = nullptr;
    if (cudaMallocAsync(&tables, (size_t)(W + HS) * sizeof(uint16_t), st) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    {
        const size_t dsm = up16((uint32_t)W + kPad) + 2u * (size_t)W + 2u * (size_t)HS;
        static bool dict_attr = false;
        if (!dict_attr) {
            cudaFuncSetAttribute(k_hwalk_dict, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 64 + 65536 + 32768);
            dict_attr = true;
        }
        k_hwalk_dict<<<1, W >= 4096 ? 512 : 128, dsm, st>>>(d_dict, cf.window, plan.hbits, Lo.R - Lo.W, tables, tables + W);
        count_launch();
    }
    HwalkArgs a;
    a.b = b;
    a.dict = d_dict;
    a.dict_links = tables;
    a.dict_heads = tables + W;
    a.window_bits = cf.window;
    a.literal = cf.literal;
    a.flags = cf.flags;
    a.write_token = cf.write_token;
    a.chunk_bits = plan.cbits;
    a.hash_bits = plan.hbits;
    const int nseg = (int)Lo.nseg, segs_per_lane = (nseg + plan.threads - 1) / plan.threads;
    a.budget = segs_per_lane * (1024 + W);
    a.max_pairs = 16;  // more than one same-bigram pair per offset inside blocks of 32: runs, periods below ~12
    static int sms = 0;
    static int occ[2][16];  // [seg == 32][window]
    int &blocks_per_sm = occ[plan.seg == 32][cf.window];
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(k_hwalk_compress<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(k_hwalk_compress<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    }
    if (!blocks_per_sm || getenv("TAMP_B200_HWALK_PLAN")) {
        if (plan.seg == 32)
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_hwalk_compress<32>, plan.threads, Lo.total);
        else
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_hwalk_compress<16>, plan.threads, Lo.total);
        if (blocks_per_sm < 1) blocks_per_sm = 1;
    }
    const uint64_t persistent = (uint64_t)sms * blocks_per_sm;
    const unsigned grid = (unsigned)(b.n_streams < persistent ? b.n_streams : persistent);
    if (plan.seg == 32)
        k_hwalk_compress<32><<<grid, plan.threads, Lo.total, st>>>(a);
    else
        k_hwalk_compress<16><<<grid, plan.threads, Lo.total, st>>>(a);
    count_launch();
    cudaFreeAsync(tables, st);
    // second pass: the bitmap kernels pick up the streams marked kDeferred (usually none)
    static unsigned int *h_seen = nullptr;  // pinned mirror of d_hwalk_deferred_total
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
    if (cf.window <= 10)
        ok = launch_fast_compress_batch(cf, d_dict, b, st, /*only_deferred=*/true, /*small_grid=*/!expect_work);
    else
        ok = launch_wide_compress_batch(cf, d_dict, b, st, /*only_deferred=*/true);
    if (h_seen) cudaMemcpyFromSymbolAsync(h_seen, d_hwalk_deferred_total, sizeof *h_seen, 0, cudaMemcpyDeviceToHost, st);
    return ok;
}
#endif  // TB_EMU

}  // namespace tb
