// Cooperative history-walk batch compressor, v1 format, windows 11..15 (any stream length): one CTA per stream, one
// WARP per walker.  The compressor of BASELINE.json config 3 (64 KiB streams, window 15) and of the wide classes of
// config 5 since round 2.
//
// Why a second history-walk kernel.  hwalk_compress.cu gives every LANE a segment of the greedy parse and lets it chase
// the time-ordered hash chain of each offset it visits.  That is the right shape while an offset has a handful of
// candidates (windows up to 1 KiB: ~1.6 per visited offset).  At window 15 a visited offset of the benchmark text has
// ~150 candidates on average and well over a thousand for the common bigrams; the chain is a linked list, so one lane
// walks it alone while the CTA waits at the chunk's barrier — measured 1.6 GB/s, 4.7 barrier-stalled warps per
// issuing one, independent of the CTA shape (profiles/r02_hwalk15_summary.txt).
//
// What changes.  The candidates of a bigram are kept as an ARRAY, not a list: per chunk, a counting sort by bigram hash
// of every position still alive (the last W + C bytes of V = dictionary ++ input) gives bucket h = POS[CUR[h-1] ..
// CUR[h]).  A poll at q (find_best_match, compressor_find_match_desktop.c:82-167) then is warp-wide: the 32 lanes take
// 32 entries of q's bucket at a time, drop the ones outside q's window (time distance D not in 1..W; the bucket
// is unordered, which is harmless: the best match is a max over keys len << 16 | ~ring position), compare 16 bytes
// each, and __reduce_max_sync picks the reference's answer (longest, then lowest window index: :59-68).  Window
// position q-1 (input[q-1] followed by bytes one lap older: no bucket has that bigram) rides along as entry -1.
// The greedy parse (compressor.c:625-657) is walked by warps instead of lanes: warp w walks segment w of the chunk from
// a guessed entry (0), then takes its left neighbour's exit as its entry and re-walks until it steps on an offset of
// its old path; repeat until no entry changes (the parse is a function of the offset, see walk_compress.cu).
//
// Per chunk of C offsets: P0 load (as in hwalk_compress.cu), P1 counting sort (two passes over the alive positions:
// shared-memory atomics, one CTA-wide scan), P2 warp walks, P4 bit pack (one lane per 32 offsets; a token's length
// is the distance to the next token: compressor.c:49-75 for the bit order), whole words out, partial word carried.
//
// Tried and dropped (round 2): bucket index = bigram hash << sb | hash of the third byte, a poll evaluating only its
// trigram's sub-bucket and scanning the rest of the bigram when that leaves it without a 3-byte match.  Bit-exact, 43 %
// fewer evaluations on text at window 15 — and 4 % faster at best (26.8 against 28.1 ms per 256 MiB with the extra
// entries and the scan in the loop; the plain loop: 25.7): the kernel waits at its barriers and on shared-memory latency
// (issue 43 %), not on the number of compares.
//
// Streams whose buckets are pathologically full (runs, short periods: every position in one bucket) are marked
// kDeferred and left to the bitmap kernel (wide_compress.cu), whose cost does not depend on the data.
#include <stdio.h>
#include <stdlib.h>

#include "../tb_wire.h"
#include "tb_cuda.h"
#include "tb_device_common.cuh"
#include "tb_smem.cuh"
#include "history_ring.cuh"

namespace tb {

namespace {

__device__ unsigned int d_cwalk_deferred_total = 0;  // streams deferred so far (cumulative)

struct CwalkLayout {
    uint32_t W, C, R, HS, NA, nwords;
    uint32_t oHB, oPOS, oCUR, oBEST, oPATH, oSBIT, oSTAGE, oWARP, oMISC, total;
};

__host__ __device__ inline CwalkLayout cwalk_layout(int wbits, int cbits, int hbits) {
    CwalkLayout L;
    L.W = 1u << wbits;
    L.C = 1u << cbits;
    L.R = ring_size(L.W, L.C);
    L.HS = 1u << hbits;
    L.NA = L.W + L.C;      // positions alive for some poll of the chunk
    L.nwords = L.C / 32u;  // path words
    L.oHB = 0;
    L.oPOS = up16(L.R + kPad);
    L.oCUR = L.oPOS + up16(2u * L.NA);
    L.oBEST = L.oCUR + 4u * L.HS;
    L.oPATH = L.oBEST + 2u * L.C;
    L.oSBIT = L.oPATH + up16(4u * L.nwords + 4u);  // (+1 word: the pack looks one word ahead)
    L.oSTAGE = L.oSBIT + up16(4u * L.nwords);
    L.oWARP = L.oSTAGE + up16(4u * (L.C * 9u / 32u + 4u));
    L.oMISC = L.oWARP + 4u * 256u;  // per walker (<= 64): entry, exit (two copies: a round reads the previous round's); scan scratch (32)
    L.total = L.oMISC + 64u;
    return L;
}

enum { CM_MISFIT = 1, CM_CARRY = 2 };

struct CwalkArgs {
    BatchArgs b;
    const uint8_t *dict;  // W bytes
    int window_bits, literal, flags, write_token;
    int chunk_bits, hash_bits;
    int budget;  // bucket entries (in units of 32) a warp may look at in one round of one chunk before the stream is given up
};

__device__ __forceinline__ void clear_bits(uint32_t sPATH, int a, int b) {  // bits [a, b) of the path, b - a <= 32
    if (a >= b) return;
    const int wa = a >> 5, wb = (b - 1) >> 5;
    const uint32_t ma = kFull << (a & 31), mb = kFull >> (31 - ((b - 1) & 31));
    if (wa == wb) {
        smem::st32(sPATH + 4u * (uint32_t)wa, smem::ld32(sPATH + 4u * (uint32_t)wa) & ~(ma & mb));
    } else {
        smem::st32(sPATH + 4u * (uint32_t)wa, smem::ld32(sPATH + 4u * (uint32_t)wa) & ~ma);
        smem::st32(sPATH + 4u * (uint32_t)wb, smem::ld32(sPATH + 4u * (uint32_t)wb) & ~mb);
    }
}

template <int REGS, int GL>
__global__ void __maxnreg__(REGS) k_cwalk_compress(CwalkArgs a) {
#ifndef TB_EMU
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t *sm = smem_raw;
#else
    uint8_t *sm = emu::g_smem;
#endif
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, T = blockDim.x, nwarps = T >> 5;
    const int wbits = a.window_bits, lbits = a.literal, hbits = a.hash_bits;
    const int W = 1 << wbits, C = 1 << a.chunk_bits;
    const CwalkLayout Lo = cwalk_layout(wbits, a.chunk_bits, hbits);
    const int R = (int)Lo.R, HS = (int)Lo.HS;
    const int min_pat = min_pattern_size(wbits, lbits);
    const int max_len = min_pat + 13;
    // a walker is GL lanes (a warp, or half a warp where a poll has fewer candidates than a warp has lanes)
    const int nwalk = T / GL, walker = tid / GL, glane = tid % GL;
    const uint32_t gmask = GL == 32 ? kFull : (0xFFFFu << (tid & 16));
    const int SEGW = C / nwalk;  // offsets per walker (a multiple of 32)
    uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
#ifndef TB_EMU
    asm volatile("" : "+r"(sbase));
#endif
    const uint32_t sHB = sbase + Lo.oHB, sPOS = sbase + Lo.oPOS, sCUR = sbase + Lo.oCUR, sBEST = sbase + Lo.oBEST;
    const uint32_t sPATH = sbase + Lo.oPATH, sSBIT = sbase + Lo.oSBIT;
    uint32_t *cur = reinterpret_cast<uint32_t *>(sm + Lo.oCUR);
    uint32_t *stage = reinterpret_cast<uint32_t *>(sm + Lo.oSTAGE);
    uint32_t *wentry = reinterpret_cast<uint32_t *>(sm + Lo.oWARP), *wexit2 = wentry + 64, *wscan = wentry + 192;  // exits: two copies of 64
    uint32_t *misc = reinterpret_cast<uint32_t *>(sm + Lo.oMISC);
    const int stage_words = (int)(Lo.C * 9u / 32u + 4u);

    for (uint64_t stream = blockIdx.x; stream < a.b.n_streams; stream += gridDim.x) {
        const uint8_t *src = a.b.in + stream * a.b.in_stride;
        const int N = a.b.in_sizes ? (int)a.b.in_sizes[stream] : (int)a.b.in_stride;
        const int row16 = (int)a.b.in_stride;  // readable bytes of the row (a multiple of 16)
        uint32_t *out32 = reinterpret_cast<uint32_t *>(a.b.out + stream * a.b.out_stride);

        // ---- stream start: the dictionary is the history (positions R - W .. R - 1) ----------------------------------
        __syncthreads();
        for (int i = tid; i < W / 16; i += T)
            reinterpret_cast<uint4 *>(sm + Lo.oHB + R - W)[i] = __ldg(reinterpret_cast<const uint4 *>(a.dict) + i);
        for (int i = tid; i < stage_words; i += T) stage[i] = 0u;
        if (tid == 0) {
            const uint32_t header = ((uint32_t)(wbits - 8) << 5) | ((uint32_t)(lbits - 5) << 3) |
                                    ((a.flags & TB_F_CUSTOM_DICT) ? 4u : 0u) | ((a.flags & TB_F_DICT_RESET) ? 1u : 0u);
            stage[0] = stream_appends(a.flags, stream) ? kAppendStart : header << 24;
        }
        uint32_t carry = (a.flags & TB_F_DICT_RESET) ? 16u : 8u;  // bits waiting in the staging line's first word(s)
        uint32_t ow = 0;      // whole words already written to the output row
        int res = kOk;
        int pvs = 0;          // ring position of the chunk's first offset
        int chunk_entry = 0;  // where the token that straddles the chunk boundary ends
        bool bail = false;
        __syncthreads();

        for (int cs = 0; cs < N && res == kOk; cs += C) {
            const int cn = N - cs < C ? N - cs : C;
            const int nwords = (cn + 31) >> 5;
            const uint32_t qm0 = (uint32_t)cs & (uint32_t)(W - 1);

            // ---- P0: the chunk and its lookahead; path and bucket counters cleared ------------------------------------
            for (int i = tid; i < (C + kPad) / 16; i += T) {
                const int goff = cs + 16 * i;
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (goff < row16 && goff < N + kPad) v = __ldg(reinterpret_cast<const uint4 *>(src + goff));
                const int ph = pvs + 16 * i;
                *reinterpret_cast<uint4 *>(sm + Lo.oHB + ph) = v;                    // (ph < R + 32: the mirror is writable)
                if (ph >= R) *reinterpret_cast<uint4 *>(sm + Lo.oHB + ph - R) = v;   // lookahead past the wrap
                if (ph < kPad) *reinterpret_cast<uint4 *>(sm + Lo.oHB + R + ph) = v; // mirror of the first bytes
            }
            for (int i = tid; i < HS; i += T) cur[i] = 0u;
            for (int i = tid; i <= (int)Lo.nwords; i += T) smem::st32(sPATH + 4u * (uint32_t)i, 0u);
            if (tid == 0) misc[CM_MISFIT] = 0xFFFFFFFFu;
            __syncthreads();

            // ---- P1: counting sort of the alive positions by bigram hash -----------------------------------------------
            // alive for some poll of the chunk: times cs - W .. cs + cn - 1 (the stream's last byte has no bigram)
            const int nalive = W + (N - cs - 1 < cn ? N - cs - 1 : cn);
            for (int pass = 0; pass < 2; pass++) {
                for (int i = tid; i < nalive; i += T) {
                    int ph = pvs - W + i;
                    if (ph < 0) ph += R;
                    const uint32_t h = bigram_hash(smem::ld8(sHB + (uint32_t)ph) | (smem::ld8(sHB + (uint32_t)ph + 1u) << 8), hbits);
                    const uint32_t slot = atomicAdd(&cur[h], 1u);
                    if (pass) smem::st16(sPOS + 2u * slot, (uint32_t)ph);
                }
                __syncthreads();
                if (!pass) {
                    // exclusive scan of the bucket sizes (HS / T consecutive counters per thread)
                    const int per = HS / T;  // (HS is a multiple of T)
                    uint32_t sum = 0;
                    for (int k = 0; k < per; k++) sum += cur[tid * per + k];
                    uint32_t incl = sum;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t u = __shfl_up_sync(kFull, incl, d);
                        if (lane >= d) incl += u;
                    }
                    if (lane == 31) wscan[warp] = incl;
                    __syncthreads();
                    uint32_t base = incl - sum;
                    for (int w = 0; w < warp; w++) base += wscan[w];
                    for (int k = 0; k < per; k++) {
                        const uint32_t c = cur[tid * per + k];
                        cur[tid * per + k] = base;
                        base += c;
                    }
                    __syncthreads();
                }
            }
            // now cur[h] = end of bucket h, cur[h - 1] (0 for h = 0) its start

            // ---- P2: warp walks -------------------------------------------------------------------------------------------
            uint32_t *wexit = wexit2;  // the copy of the exits the last round wrote
            for (int round = 0;; round++) {
                bool changed = false, gave_up = false;
                const uint32_t *wexit_prev = wexit2 + 64 * ((round + 1) & 1);
                wexit = wexit2 + 64 * (round & 1);
                const int segstart = walker * SEGW;
                const int segend = segstart + SEGW < cn ? segstart + SEGW : cn;
                if (segstart < cn) {
                    int entry = walker == 0 ? chunk_entry : 0;
                    if (walker > 0 && round > 0) entry = (int)wexit_prev[walker - 1];
                    if (glane == 0) wexit[walker] = round > 0 ? wexit_prev[walker] : 0u;  // unless the walk finds a new one
                    __syncwarp(gmask);
                    if (round == 0 || entry != (int)wentry[walker]) {
                        changed = round > 0;
                        __syncwarp(gmask);
                        if (glane == 0) {
                            wentry[walker] = (uint32_t)entry;
                            if (round > 0) clear_bits(sPATH, segstart, segstart + entry < segend ? segstart + entry : segend);
                        }
                        __syncwarp(gmask);
                        int t = segstart + entry;
                        bool merged = false;
                        int budget = a.budget < (1 << 29) ? a.budget * (32 / GL) : a.budget;  // (tests switch the give-up off with a huge budget)
                        while (t < segend) {
                            if (round > 0 && ((smem::ld32(sPATH + 4u * (uint32_t)(t >> 5)) >> (t & 31)) & 1u)) {
                                merged = true;  // on the old path: same tokens and exit from here on
                                break;
                            }
                            // ---- the poll at chunk offset t, by the walker ----
                            const int pq = pvs + t;
                            const uint32_t qmr = (qm0 + (uint32_t)t) & (uint32_t)(W - 1);
                            const int L = N - cs - t < max_len ? N - cs - t : max_len;
                            uint32_t la[4];
                            smem::load16(sHB + (uint32_t)pq, la);
                            uint32_t bestkey = 0u;
                            if (L >= 2) {
                                const uint32_t h = bigram_hash(la[0] & 0xFFFFu, hbits);
                                const int bstart = h ? (int)cur[h - 1] : 0, bend = (int)cur[h];
                                const uint32_t pprev = pq ? (uint32_t)pq - 1u : (uint32_t)R - 1u;
                                const bool strad = qmr != 0u && smem::ld8(sHB + pprev) == (la[0] & 0xFFu);
                                for (int i = bstart - 1 + glane; i < bend; i += GL) {
                                    int D;
                                    uint32_t ca;
                                    if (i < bstart) {  // entry -1: window position q-1
                                        D = strad ? 1 : 0;
                                        ca = pprev;
                                    } else {
                                        ca = smem::ld16(sPOS + 2u * (uint32_t)i);
                                        D = pq - (int)ca;
                                        if (D <= 0) D += R;
                                    }
                                    if (D >= 1 && D <= W) {
                                        const uint32_t key = eval_candidate(D, ca, la, L, qmr, sHB, pq, W, R);
                                        bestkey = key > bestkey ? key : bestkey;
                                    }
                                }
                                budget -= (bend - bstart + GL) / GL;
                                bestkey = __reduce_max_sync(gmask, bestkey);
                            }
                            const int len = (int)(bestkey >> 16);
                            const bool is_match = len >= min_pat;
                            const int tn = t + (is_match ? len : 1);
                            __syncwarp(gmask);  // (every lane has read the path bit of t)
                            if (glane == 0) {
                                if (is_match) smem::st16(sBEST + 2u * (uint32_t)t, (~bestkey) & 0xFFFFu);
                                const uint32_t wa = sPATH + 4u * (uint32_t)(t >> 5);
                                smem::st32(wa, smem::ld32(wa) | (1u << (t & 31)));
                                if (round > 0) clear_bits(sPATH, t + 1, tn < segend ? tn : segend);  // old tokens inside the new one
                            }
                            __syncwarp(gmask);
                            t = tn;
                            if (budget <= 0) break;
                        }
                        if (budget <= 0) {
                            gave_up = true;
                        } else if (!merged && glane == 0) {
                            wexit[walker] = (uint32_t)(t - segend);
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
            const int last_walker = (cn - 1) / SEGW;
            const int chunk_exit = (int)wexit[last_walker];  // where the chunk's last token ends, past the chunk

            // ---- P4: bit pack, one lane per path word -------------------------------------------------------------------
            // (a token's length is the distance to the next token; a token reaches at most 16 offsets ahead)
            const bool last_chunk = cs + C >= N;
            uint32_t cut = 0xFFFFFFFFu;  // chunk offset of the first literal that does not fit (compressor.c:629-631)
            for (int pass = lbits < 8 ? -1 : 0; pass < 2; pass++) {
                for (int j = tid; j < nwords; j += T) {
                    uint32_t m = smem::ld32(sPATH + 4u * (uint32_t)j);
                    const uint32_t mnext = j + 1 < nwords ? smem::ld32(sPATH + 4u * (uint32_t)(j + 1)) : 0u;
                    const int wordend = mnext ? 32 + __ffs(mnext) - 1 : (last_chunk ? cn - 32 * j : (C - 32 * j) + chunk_exit);
                    uint32_t pos = pass > 0 ? carry + smem::ld32(sSBIT + 4u * (uint32_t)j) : 0u;
                    while (m) {
                        const int t = __ffs(m) - 1;
                        m &= m - 1;
                        const int off = 32 * j + t;
                        if ((uint32_t)off >= cut) break;
                        const int next = m ? __ffs(m) - 1 : wordend;
                        const int len = next - t;
                        uint32_t bits;
                        int nb;
                        if (len == 1) {
                            const uint32_t c = smem::ld8(sHB + (uint32_t)(pvs + off));
                            if (pass < 0) {
                                if (c >> lbits) {
                                    atomicMin(&misc[CM_MISFIT], (uint32_t)off);
                                    break;
                                }
                                continue;
                            }
                            bits = (1u << lbits) | c;
                            nb = lbits + 1;
                        } else {
                            if (pass < 0) continue;
                            const int sym = len - min_pat;
                            bits = ((uint32_t)kHuff.code[sym] << wbits) | smem::ld16(sBEST + 2u * (uint32_t)off);
                            nb = (int)kHuff.bits[sym] + wbits;
                        }
                        if (pass > 0) {
                            const uint32_t wi = pos >> 5, o = pos & 31u;
                            const uint64_t sv = (uint64_t)bits << (64 - nb - (int)o);
                            atomicOr(&stage[wi], (uint32_t)(sv >> 32));
                            if ((uint32_t)sv) atomicOr(&stage[wi + 1], (uint32_t)sv);
                        }
                        pos += (uint32_t)nb;
                    }
                    if (pass == 0) smem::st32(sSBIT + 4u * (uint32_t)j, pos);
                }
                __syncthreads();
                if (pass < 0) cut = misc[CM_MISFIT];
                if (pass == 0) {
                    // exclusive scan of the words' bit counts by warp 0; the chunk's total goes to CM_CARRY
                    if (tid < 32) {
                        uint32_t run = 0;
                        for (int b0 = 0; b0 < nwords; b0 += 32) {
                            const int i = b0 + lane;
                            const uint32_t v = i < nwords ? smem::ld32(sSBIT + 4u * (uint32_t)i) : 0u;
                            uint32_t incl = v;
#pragma unroll
                            for (int d = 1; d < 32; d <<= 1) {
                                const uint32_t u = __shfl_up_sync(kFull, incl, d);
                                if (lane >= d) incl += u;
                            }
                            if (i < nwords) smem::st32(sSBIT + 4u * (uint32_t)i, run + incl - v);
                            run += __shfl_sync(kFull, incl, 31);
                        }
                        if (lane == 0) misc[CM_CARRY] = run;
                    }
                    __syncthreads();
                }
            }
            // whole words leave; the partial word is carried
            {
                const uint32_t total = carry + misc[CM_CARRY];
                const uint32_t tw = total >> 5;
                for (uint32_t i = (uint32_t)tid; i < tw; i += (uint32_t)T) out32[ow + i] = __byte_perm(stage[i], 0, 0x0123);
                const uint32_t part = stage[tw];
                __syncthreads();
                for (uint32_t i = (uint32_t)tid; i < tw + 2u; i += (uint32_t)T) stage[i] = i == 0u ? part : 0u;
                ow += tw;
                carry = total & 31u;
                if (cut != 0xFFFFFFFFu) res = kExcessBits;
                chunk_entry = chunk_exit;
            }
            pvs += C;
            if (pvs == R) pvs = 0;
            __syncthreads();
        }

        if (bail) {
            if (tid == 0) {
                a.b.out_sizes[stream] = kDeferred;
                atomicAdd(&d_cwalk_deferred_total, 1u);
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

struct CwalkPlan {
    int cbits, hbits, threads, gl;  // gl: lanes per walker (32 or 16)
};

// chunk, hash table and CTA size per window (tuning hook: TAMP_B200_CWALK_PLAN="cbits,hbits,threads")
inline CwalkPlan cwalk_plan(int wbits) {
#ifndef TB_EMU
    if (const char *e = getenv("TAMP_B200_CWALK_PLAN")) {
        CwalkPlan p;
        p.gl = 32;
        if (sscanf(e, "%d,%d,%d,%d", &p.cbits, &p.hbits, &p.threads, &p.gl) >= 3 && (p.gl == 16 || p.gl == 32) && p.cbits >= 10 && p.cbits <= 14 && p.hbits >= 10 &&
            p.hbits <= 14 && p.threads >= 32 && p.threads <= 1024 && (p.threads & (p.threads - 1)) == 0 &&
            (1 << p.hbits) >= p.threads && (1 << p.cbits) / (p.threads / p.gl) >= 32 &&
            cwalk_layout(wbits, p.cbits, p.hbits).total <= 227u * 1024u)
            return p;
    }
#endif
    switch (wbits) {
        // measured on B200, 256 MiB of text per class (GB/s): window 11: 15.6, 12: 18.1, 13: 19.5 with 256-thread CTAs (several
        // per SM); windows 14 / 15 (one CTA per SM whatever its size): 16.3 / 13.1 with 1024 threads at 64 registers against
        // 12.7 / 10.4 with 512; hash bits: the counting sort scans the table, 11 beats 12 beats 13 below window 14
        // half-warp walkers (a poll at windows 11 / 12 has fewer candidates than a warp has lanes): 17.5 against 15.1 GB/s at
        // window 11, 19.3 against 17.6 at 12, no gain at 13 (profiles/r02_cwalk_halfwarp_sweep.log)
        case 11: return {12, 11, 256, 16};
        case 12: return {12, 11, 256, 16};
        case 13: return {12, 11, 256, 32};
        case 14: return {13, 12, 1024, 32};
        default: return {13, 12, 1024, 32};
    }
}

}  // namespace

#ifndef TB_EMU
bool launch_cwalk_compress_batch(const CompBatchConf &cf, const uint8_t *d_dict, const BatchArgs &b, cudaStream_t st) {
    if (cf.window < 11 || cf.window > 15) return false;
    if (cf.flags & (TB_F_EXTENDED | TB_F_LAZY)) return false;  // v1 greedy only
    if (b.in_offsets) return false;                          // strided layout only
    if ((b.in_stride & 15) || ((uintptr_t)b.in & 15) || (b.out_stride & 3) || ((uintptr_t)b.out & 3)) return false;
    if (b.in_stride > (1u << 30)) return false;
    if ((uintptr_t)d_dict & 15) return false;
    const uint64_t bound = 2 + (b.in_stride * (uint64_t)(cf.literal + 1) + 7) / 8 + 6;
    if (b.out_stride < ((bound + 3) & ~3ull)) return false;  // never OUTPUT_FULL in this kernel
    if (b.n_streams == 0) return true;

    const CwalkPlan plan = cwalk_plan(cf.window);
    const CwalkLayout Lo = cwalk_layout(cf.window, plan.cbits, plan.hbits);
    CwalkArgs a;
    a.b = b;
    a.dict = d_dict;
    a.window_bits = cf.window;
    a.literal = cf.literal;
    a.flags = cf.flags;
    a.write_token = cf.write_token;
    a.chunk_bits = plan.cbits;
    a.hash_bits = plan.hbits;
    const int segw = (1 << plan.cbits) / (plan.threads / plan.gl);
    a.budget = 16 * segw;  // text at window 15: ~1 unit per offset walked; period 4: ~20
    static int sms = 0;
    static int occ[16];
    int &blocks_per_sm = occ[cf.window];
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(k_cwalk_compress<80, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(k_cwalk_compress<64, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(k_cwalk_compress<80, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    }
    if (!blocks_per_sm || getenv("TAMP_B200_CWALK_PLAN")) {
        if (plan.threads > 768)
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_cwalk_compress<64, 32>, plan.threads, Lo.total);
        else
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_cwalk_compress<80, 32>, plan.threads, Lo.total);
        if (blocks_per_sm < 1) blocks_per_sm = 1;
    }
    const uint64_t persistent = (uint64_t)sms * blocks_per_sm;
    const unsigned grid = (unsigned)(b.n_streams < persistent ? b.n_streams : persistent);
    if (plan.threads > 768)  // 1024 threads: 64 registers each
        k_cwalk_compress<64, 32><<<grid, plan.threads, Lo.total, st>>>(a);
    else if (plan.gl == 16)  // half-warp walkers
        k_cwalk_compress<80, 16><<<grid, plan.threads, Lo.total, st>>>(a);
    else
        k_cwalk_compress<80, 32><<<grid, plan.threads, Lo.total, st>>>(a);
    count_launch();
    // second pass: the bitmap kernel picks up the streams marked kDeferred (usually none)
    const bool ok = launch_wide_compress_batch(cf, d_dict, b, st, /*only_deferred=*/true);
    return ok;
}
#endif  // TB_EMU

}  // namespace tb
