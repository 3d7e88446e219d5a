// Segment-walk batch compressor, v1 and extended format, streams no longer than the window (N <= W <= 1024): the
// dominant kernel of the headline workload (BASELINE.json config 2) since round 2.
//
// Where it comes from.  The position-parallel kernel (ppar_compress.cu) computes find_best_match
// (compressor_find_match_desktop.c:82-167) for EVERY input offset and then walks the greedy parse over that table.
// Measured (profiles/r02_ppar_phases_r01kernel.txt): 74 % of its 22.5 k warp instructions per stream are the candidate
// walk, and the parse only ever looks at ~424 of the 1024 table entries — the offsets INSIDE a match are never polled
// (tamp_compressor_poll consumes match_size bytes, compressor.c:652-657), yet they are the ones with the longest
// candidate chains (3197 candidates per stream for all offsets, 643 for the polled ones; tools/model/ppar_model.c).
//
// What it does instead.  The window a poll at offset p sees is still parse-independent (v1, N <= W):
//     window_p[x] = x < p ? input[x] : dictionary[x]
// so a match can be evaluated at any offset at any time — but it is only evaluated where a walk actually arrives:
//
//   P1  hash chains over the bigrams of the input (as in ppar_compress.cu), plus one bit per offset: "has a candidate";
//   P2  segment walk.  Lane l owns offsets [32 l, 32 l + 32).  Every lane walks the greedy parse
//       p <- p + max(1, len(p)) through its own segment, starting from a guessed entry offset (0), and runs the
//       candidate chain of an offset only when its walk stands on it (offsets without a candidate are skipped with
//       one FFS: they are literals).  Then every lane takes its left neighbour's exit as its entry and repeats from
//       there until it steps on an offset its previous walk visited — from then on the two walks coincide
//       (the parse is a function of the offset) — or leaves the segment.  Repeat until no entry changes: lane 0 is
//       exact after round 1, lane l after round l + 1 at the latest; on text the walks merge within a token or two,
//       so 2 rounds settle it (tools/model/walk_model.c: 726 candidates evaluated per stream, 444 offsets);
//   P3  token list: the visited-offset masks of the lanes, by warp prefix sum;
//   P4  static-Huffman bit pack: 32 tokens at a time, warp prefix sum of the bit lengths, tokens ORed into an
//       MSb-first staging line, coalesced stores (write_to_bit_buffer / partial_flush / flush, compressor.c:49-75,
//       :728-810) — as in ppar_compress.cu.
//
// Candidate encoding (chain links, hash heads): 0 = none, 1..1024 = dictionary position d + 1, 1025..2047 = input
// offset x + 1025.  A chain runs through the input offsets (newest first) and on through the dictionary positions
// (highest first), so "alive for a poll at q" — any input offset, or a dictionary position >= q — is the single
// compare e > q, and the first dead entry ends the chain.
//
// Extended format (EXT; compressor.c:437-525, :342-415 — what conf == NULL selects).  Lookahead 16 instead of 15.  Until a
// run of MORE than 8 bytes has been emitted every consumed byte has been written, so the window still is the v1 window
// (an RLE token writes at most 8 bytes, :342-359; an extended match is never clipped while N <= W).  Offsets where a run
// starts (the byte equals the last one written and so does the next, or the input ends) or whose match is 14+ bytes
// long are SPECIAL: entered with no run pending — always the case, a run is counted to its end in one go — the outcome
// (run token, lone run byte at the end of the input, the plain step when a run of <= 6 loses to a longer match, or an
// extended match) is a function of the offset alone and starts a token there, so the segment walk carries over: the
// lane that stands on a special offset resolves it by itself.  A 16-byte match grows against the window as it was at
// its start, up to 133 bytes, longest first, then lowest position (poll_extended_handling / find_extended_match: the
// candidate set only shrinks, its lowest member is reported) — done inside the candidate compare of that offset.
// Tokens may now be long (runs: 241 bytes), so a walk may jump whole segments: a lane whose entry lies past its segment
// passes it on.  best[q] of a special token: length field 17 = run, 18 = lone run byte, 19 = extended match + its
// position; the byte count is the distance to the next token.  A stream that emits a run of more than 8 bytes before
// its end is left to the bitmap kernel like the streams with pathological chains.
//
// Streams whose chains are pathologically long (runs, short periods) are left to the bitmap kernel through the same
// pick-up pass as in ppar_compress.cu.
#include "../tb_wire.h"
#include "tb_cuda.h"
#include "tb_device_common.cuh"
#include "tb_smem.cuh"

namespace tb {

namespace {

constexpr int kMaxN = 1024;
constexpr int kPad = 32;            // readable slack behind the byte arrays (unaligned 20-byte reads)
constexpr int kHashBits = 11, kHashSize = 1 << kHashBits;
constexpr uint32_t kFull = 0xffffffffu;
constexpr int kMaxLenV1 = 15;       // v1: min_pattern_size (2) + 13
constexpr int kMaxLenExt = 16;      // extended format: the whole 16-byte input ring
constexpr int kExtCap = 2 + 11 + kExtExtraMax;  // longest extended match: min_pattern + 11 + 120
constexpr uint32_t kKindRun = 17u, kKindLone = 18u, kKindExt = 19u;  // length field of best[] for the special tokens
constexpr uint32_t kIn = 1025;      // input offset x is candidate x + kIn; dictionary position d is candidate d + 1
constexpr uint32_t kIdxMask = 0x7FFu, kCountMax = 31u;  // hash head: candidate | chain population << 11
constexpr int kMaxPairs = 8192;     // chain population above which a stream goes to the bitmap kernel (typical text: ~3000)
constexpr int kMaxWalkIters = 1500; // safety valve of the walk: beyond this the bitmap kernel is the cheaper one
constexpr int kStageWords = (kMaxN * 9 / 8 + 16 + 3) / 4;

// per-CTA shared memory: the dictionary side, shared by all warps
constexpr int D_BYTES = 0;                               // dictionary bytes
constexpr int D_LINK = D_BYTES + kMaxN + kPad;           // u16: chain links of the dictionary positions
constexpr int D_HEAD = D_LINK + 2 * kMaxN;               // u16: chain heads of the dictionary
constexpr int D_LUT = D_HEAD + 2 * kHashSize;
constexpr int D_END = D_LUT + 64;
// per-warp shared memory: the input side
constexpr int OFF_COMB = 0;                              // input bytes
constexpr int OFF_LINK = OFF_COMB + kMaxN + kPad;        // u16 link[p]: next candidate of the chain through p
constexpr int OFF_HEAD = OFF_LINK + 2 * kMaxN;           // u16 head[h] (P1); afterwards best / staging line
constexpr int PER_WARP = OFF_HEAD + 2 * kHashSize;
constexpr int OFF_BEST = OFF_HEAD;                       // u16 best[p] = len << 11 | priority (P2 onwards; over the dead hash table)
constexpr int OFF_STAGE = OFF_BEST + 2 * kMaxN;          // u32 stage[kStageWords]
constexpr int OFF_TOK = OFF_LINK;                        // u16 tok[]: token offsets (P3 onwards; over the dead links)
static_assert(OFF_STAGE + 4 * kStageWords <= PER_WARP, "best + staging line fit the dead hash table");
static_assert(D_END % 16 == 0 && OFF_LINK % 16 == 0 && OFF_HEAD % 16 == 0 && D_HEAD % 16 == 0 && PER_WARP % 16 == 0, "aligned regions");
// One CTA per SM with as many warps (= streams in flight) as shared memory takes, less 8 KiB for the pick-up pass of
// the bitmap kernel (one-warp CTAs, 6.4 KiB each), which must find room beside this CTA.
constexpr int kSmemBudget = 227 * 1024 - 8 * 1024;
constexpr int kWarps = (kSmemBudget - D_END) / PER_WARP < 32 ? (kSmemBudget - D_END) / PER_WARP : 32;
constexpr int CTA_BYTES = D_END + kWarps * PER_WARP;

__device__ unsigned int d_walk_deferred_total = 0;  // streams deferred so far (cumulative); see launch_walk_compress_batch

struct WalkArgs {
    BatchArgs b;
    const uint8_t *dict;
    int window_bits, literal, flags, write_token;
    int max_pairs;  // streams with more chain pairs than this are deferred
};

__device__ __forceinline__ uint32_t bigram_hash(uint32_t key16) { return (key16 * 2654435761u) >> (32 - kHashBits); }


// Link the n-1 bigrams of the n bytes at shared address sBytes into per-hash chains, newest first: link[p] = candidate
// code of the previous entry with the same bigram hash, or whatever head[h] held before (none / a dictionary position)
// for the first one.  Offset p is stored as candidate p + first.  Returns this lane's share of the number of (offset,
// earlier offset with the same hash) pairs; with HAS, lane b also gets the mask of block b's offsets that have a
// candidate for a poll there: a live chain entry, or window position p-1 (input[p-1] followed by DICTIONARY bytes —
// in no chain — which matches 2+ bytes iff input[p-1] == input[p] and dict[p] == input[p+1]).  Warp-cooperative; all
// arrays by .shared address.
template <bool HAS, bool EXT = false>
__device__ __forceinline__ uint32_t build_chains(uint32_t sBytes, int n, uint32_t first, uint32_t sHead, uint32_t sLink,
                                                 int lane, uint32_t sDict, uint32_t &hasmask, uint32_t dict_last = 0) {
    uint32_t pairs = 0;
    const uint32_t lt = (1u << lane) - 1u;
    for (int base = 0; base < n; base += 32) {
        const int p = base + lane;
        const bool valid = p + 1 < n;
        const uint32_t b0 = smem::ld8(sBytes + (uint32_t)p), b1 = smem::ld8(sBytes + (uint32_t)p + 1u);  // (slack behind the array)
        const uint32_t h = valid ? bigram_hash(b0 | (b1 << 8)) : (0x10000u | (uint32_t)lane);  // invalid lanes match nobody
        const uint32_t peers = __match_any_sync(kFull, h);
        const uint32_t hv = valid ? smem::ld16(sHead + 2u * h) : 0u;
        const uint32_t lower = peers & lt;
        uint32_t pv = lower ? first + (uint32_t)(base + 31 - __clz(lower)) : (hv & kIdxMask);
        const uint32_t count = (hv >> 11) + __popc(lower);  // input offsets before p on this chain
        if (!valid) pv = 0;
        if (p < n) smem::st16(sLink + 2u * (uint32_t)p, pv);
        if (valid) pairs += count;
        if (HAS) {
            const uint32_t prev = p >= 1 ? smem::ld8(sBytes + (uint32_t)p - 1u) : dict_last;
            const bool strad = p >= 1 && prev == b0 && smem::ld8(sDict + (uint32_t)p) == b1;
            // extended format: a run starts here (the walk must stop even if no match candidate exists)
            const bool runstart = EXT && p < n && b0 == prev && (p + 1 >= n || b1 == prev);
            const uint32_t m = __ballot_sync(kFull, (valid && (pv > (uint32_t)p || strad)) || runstart);
            if (lane == (base >> 5)) hasmask = m;
        }
        __syncwarp();
        if (valid && (peers >> lane) == 1u) {  // the block's last entry with this hash
            const uint32_t c = count + 1 < kCountMax ? count + 1 : kCountMax;
            smem::st16(sHead + 2u * h, (first + (uint32_t)p) | (c << 11));
        }
        __syncwarp();
    }
    return pairs;
}

template <bool EXT>
__global__ void __launch_bounds__(kWarps * 32) k_walk_compress(WalkArgs a) {
    constexpr int kMaxLen = EXT ? kMaxLenExt : kMaxLenV1;
#ifndef TB_EMU
    extern __shared__ __align__(128) uint8_t smem[];
#else
    uint8_t *smem = emu::g_smem;
#endif
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = 1 << a.window_bits;
    const int wbits = a.window_bits;
    const int lbits = a.literal;
    uint8_t *dictb = smem + D_BYTES;
    uint16_t *dhead = reinterpret_cast<uint16_t *>(smem + D_HEAD);
    uint32_t *lut = reinterpret_cast<uint32_t *>(smem + D_LUT);
    uint8_t *wbase = smem + D_END + warp * PER_WARP;
    uint8_t *comb = wbase + OFF_COMB;
    uint16_t *head = reinterpret_cast<uint16_t *>(wbase + OFF_HEAD);
    uint16_t *best = reinterpret_cast<uint16_t *>(wbase + OFF_BEST);
    uint16_t *tok = reinterpret_cast<uint16_t *>(wbase + OFF_TOK);
    uint32_t *stage = reinterpret_cast<uint32_t *>(wbase + OFF_STAGE);
    // shared addresses for candidate-indexed accesses, biased by the candidate encoding.  (The base goes through an
    // empty asm so that the compiler keeps it in a register instead of re-deriving it inside the loops.)
    uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
#ifndef TB_EMU
    asm volatile("" : "+r"(sbase));
#endif
    const uint32_t sIn = sbase + (uint32_t)(D_END + warp * PER_WARP + OFF_COMB);
    const uint32_t sBytesInM = sIn - kIn;
    const uint32_t sBytesDictM = sbase + (uint32_t)D_BYTES - 1u;
    const uint32_t sLinkIn = sbase + (uint32_t)(D_END + warp * PER_WARP + OFF_LINK);
    const uint32_t sLinkInM = sLinkIn - 2u * kIn;
    const uint32_t sHeadIn = sbase + (uint32_t)(D_END + warp * PER_WARP + OFF_HEAD);
    const uint32_t sBestIn = sHeadIn;  // best[] lies over the dead hash table
    const uint32_t sLinkDictM = sbase + (uint32_t)D_LINK - 2u;

    // ---- once per CTA: dictionary bytes, their chains, the chain heads -----------------------------------
    for (int i = threadIdx.x; i < (kMaxN + kPad) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t *>(dictb)[i] = i * 4 < W ? reinterpret_cast<const uint32_t *>(a.dict)[i] : 0u;
    for (int i = threadIdx.x; i < kHashSize / 2; i += blockDim.x) reinterpret_cast<uint32_t *>(dhead)[i] = 0u;
    if (threadIdx.x < 16) lut[threadIdx.x] = (uint32_t)kHuff.code[threadIdx.x] | ((uint32_t)kHuff.bits[threadIdx.x] << 16);
    __syncthreads();
    if (warp == 0) {
        uint32_t unused = 0;
        build_chains<false>(sbase + (uint32_t)D_BYTES, W, 1u, sbase + (uint32_t)D_HEAD, sbase + (uint32_t)D_LINK, lane,
                            sbase + (uint32_t)D_BYTES, unused);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kHashSize; i += blockDim.x) dhead[i] &= (uint16_t)kIdxMask;  // populations count input offsets only
    __syncthreads();

    const uint64_t nwarps = (uint64_t)gridDim.x * kWarps;
    for (uint64_t stream = (uint64_t)blockIdx.x * kWarps + warp; stream < a.b.n_streams; stream += nwarps) {
        const uint8_t *src = a.b.in + stream * a.b.in_stride;
        const int N = a.b.in_sizes ? (int)a.b.in_sizes[stream] : (int)a.b.in_stride;
        uint32_t *out32 = reinterpret_cast<uint32_t *>(a.b.out + stream * a.b.out_stride);

        // ---- P0: input by coalesced 128-bit loads; hash table := the dictionary's chain heads -------------
        __syncwarp();
        for (int off = lane * 16; off < N; off += 512)
            *reinterpret_cast<uint4 *>(comb + off) = __ldg(reinterpret_cast<const uint4 *>(src + off));
        for (int i = lane; i < 2 * kHashSize / 16; i += 32)
            reinterpret_cast<uint4 *>(head)[i] = reinterpret_cast<const uint4 *>(dhead)[i];
        __syncwarp();

        // ---- P1: hash chains over the input; which offsets have a candidate at all --------------------------
        uint32_t hasmask = 0;  // lane l: offsets of segment l with at least one candidate
        {
            const uint32_t pairs = __reduce_add_sync(kFull, build_chains<true, EXT>(sIn, N, kIn, sHeadIn, sLinkIn, lane, sbase + (uint32_t)D_BYTES, hasmask, dictb[W - 1]));
            if (pairs > (uint32_t)a.max_pairs) {
                // Chains this long make the candidate walk the slower way: leave the stream to the bitmap kernel
                // (launched right behind this one), whose cost does not depend on the data.
                if (lane == 0) {
                    a.b.out_sizes[stream] = kDeferred;
                    atomicAdd(&d_walk_deferred_total, 1u);
                }
                continue;
            }
        }
        // the hash table is dead: match table (0 = literal) and staging line start out as zeros
        for (int i = lane; i < (2 * kMaxN + 4 * kStageWords + 15) / 16; i += 32)
            reinterpret_cast<uint4 *>(best)[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncwarp();

        // ---- P2: segment walk ---------------------------------------------------------------------------------
        // best[q] holds the raw key of the match at q: len << 11 | (candidate ^ 1023) — the longest match wins, then the
        // lowest window index (input offsets, x + 1025, map to 2046 - x; dictionary positions, d + 1, to 1022 - d: any
        // input offset beats any dictionary position, lower beats higher) — i.e. the reference's tie-break and
        // early-exit result (compressor_find_match_desktop.c:59-68).  0 = literal.
        const int segbase = 32 * lane;
        const int nvalid = N - segbase >= 32 ? 32 : (N - segbase > 0 ? N - segbase : 0);  // offsets of my segment
        const uint32_t validmask = nvalid >= 32 ? kFull : ((1u << nvalid) - 1u);
        uint32_t path = 0;   // offsets of my segment the walk from `entry` visits (= where tokens start)
        int exitrel = 0;     // where that walk enters the next segment (0..14)
        int entry = 0;       // where the walk enters my segment
        int budget = kMaxWalkIters;
        bool giveup = false;  // EXT: a run of more than 8 bytes before the stream's end
        for (;;) {
            // a lane walks if its entry is not on the path it already knows (round 1: nothing is known)
            const bool walk = entry < nvalid && !((path >> entry) & 1u);
            if (!walk) path &= __funnelshift_lc(0u, kFull, entry);  // what the old walk visited before the entry is void
            if (EXT && entry >= 32) exitrel = entry - 32;           // a long token jumps the whole segment
            const uint32_t oldpath = walk ? path : 0u;
            uint32_t newmask = 0;
            bool active = walk, adv = walk;
            int pn = entry;            // adv: the walk continues at offset pn (relative; may lie past the segment)
            int p = 0, q = 0, L = 0;   // the offset being evaluated (relative / absolute), its lookahead
            uint32_t e = 0, ln = 0;    // current candidate, the one after it (its link is loaded one step ahead)
            uint32_t la[4] = {0, 0, 0, 0}, bestkey = 0;
            uint32_t xkey = 0;         // EXT: the grown 16-byte match (len << 16 | ~position)
            while (__any_sync(kFull, active) && --budget > 0) {
                if (active && adv) {
                    // ---- between offsets: continue the walk at pn.  Offsets without a candidate are literals (one
                    // step each), so the next stop is the first offset that has a candidate or that the previous walk
                    // visited ----
                    const uint32_t hi = __funnelshift_lc(0u, kFull, pn);  // offsets >= pn (none if pn >= 32)
                    const uint32_t stops = (hasmask | oldpath) & hi;
                    const int t = __ffs(stops) - 1;
                    const uint32_t below = (1u << t) - 1u;  // (stops == 0: t = -1, unused)
                    if (!stops) {  // literals to the end of the segment (or the token jumped past it)
                        path = newmask | (hi & validmask);
                        exitrel = pn > 32 ? pn - 32 : 0;
                        active = false;
                    } else if ((oldpath >> t) & 1u) {  // merged with the previous walk: same path and exit from here on
                        path = newmask | (hi & below) | (oldpath & ~below);
                        active = false;
                    } else {
                        newmask |= hi & below;
                        p = t;
                        q = segbase + t;
                        L = N - q < kMaxLen ? N - q : kMaxLen;
                        smem::load16(sIn + (uint32_t)q, la);
                        bestkey = 0;
                        xkey = 0;
                        // the chain of q, its second entry loaded ahead; window position q-1 holds input[q-1] followed by
                        // dictionary bytes: its bigram is not the input's, so the chain does not cover it: it goes first
                        // when its first byte fits
                        const uint32_t c1 = smem::ld16(sLinkIn + 2u * (uint32_t)q);
                        const uint32_t c2 = c1 > (uint32_t)q ? smem::ld16((c1 >= kIn ? sLinkInM : sLinkDictM) + 2u * c1) : 0u;
                        const bool strad = q >= 1 && smem::ld8(sIn + (uint32_t)q - 1u) == (la[0] & 0xFFu);
                        e = strad ? (uint32_t)q - 1u + kIn : c1;
                        ln = strad ? c1 : c2;
                        adv = false;
                    }
                }
                if (active && !adv) {
                    // ---- up to two candidates of the offset q (the link behind the second one is loaded ahead) ----
                    const uint32_t ca = e, cb = ln;
                    const bool alive_a = ca > (uint32_t)q, alive_b = alive_a && cb > (uint32_t)q;
                    const uint32_t c3 = alive_b ? smem::ld16((cb >= kIn ? sLinkInM : sLinkDictM) + 2u * cb) : 0u;
#pragma unroll
                    for (int k = 0; k < 2; k++) {
                        const uint32_t c = k ? cb : ca;
                        if (k ? alive_b : alive_a) {
                            // 16-byte compare against the pattern at q
                            const bool in_side = c >= kIn;
                            uint32_t w[4];
                            smem::load16((in_side ? sBytesInM : sBytesDictM) + c, w);
                            const uint32_t d0 = w[0] ^ la[0], d1 = w[1] ^ la[1], d2 = w[2] ^ la[2], d3 = w[3] ^ la[3];
                            uint32_t d = d0;
                            int nb = 0;
                            if (!d) { d = d1; nb = 4; }
                            if (!d) { d = d2; nb = 8; }
                            if (!d) { d = d3; nb = 12; }
                            int n = d ? nb + ((__ffs(d) - 1) >> 3) : 16;
                            // input side: the candidate's bytes are input up to q (then dictionary); dictionary side: a
                            // match never runs past the window end
                            const int lim0 = (int)((in_side ? (uint32_t)q + kIn : (uint32_t)W + 1u) - c);
                            const int lim = lim0 < L ? lim0 : L;
                            if (n >= lim) {
                                n = lim;
                                if (in_side && lim < L) {  // ran into q: the window continues with dictionary bytes
                                    const uint32_t x = c - kIn;
                                    while (n < L && smem::ld8(sBytesDictM + 1u + x + (uint32_t)n) == smem::ld8(sIn + (uint32_t)(q + n))) n++;
                                }
                            }
                            const uint32_t key = ((uint32_t)n << 11) | (c ^ 1023u);
                            bestkey = key > bestkey ? key : bestkey;
                            if (EXT && n == 16) {
                                // a full 16-byte match keeps growing against the window as it is now, up to 133 bytes
                                // (find_extended_match, compressor.c:297-333): window bytes below q are input, the rest dictionary
                                const int xw = (int)(in_side ? c - kIn : c - 1u);
                                const int cap = N - q < kExtCap ? N - q : kExtCap;
                                const int room = W - xw < cap ? W - xw : cap;
                                int nx = 16;
                                while (nx < room) {
                                    const int y = xw + nx;
                                    const uint32_t wb = y < q ? smem::ld8(sIn + (uint32_t)y) : smem::ld8(sBytesDictM + 1u + (uint32_t)y);
                                    if (wb != smem::ld8(sIn + (uint32_t)(q + nx))) break;
                                    nx++;
                                }
                                const uint32_t xk = ((uint32_t)nx << 16) | (0xFFFFu - (uint32_t)xw);
                                xkey = xk > xkey ? xk : xkey;
                            }
                        }
                    }
                    e = c3;  // 0 unless both were alive
                    if (e > (uint32_t)q) {
                        ln = smem::ld16((e >= kIn ? sLinkInM : sLinkDictM) + 2u * e);
                    } else {  // chain exhausted: the match at q is known
                        const int len = (int)(bestkey >> 11);
                        int step = len < 2 ? 1 : len;
                        if (EXT) {
                            const uint32_t lastb = q ? smem::ld8(sIn + (uint32_t)q - 1u) : (uint32_t)dictb[W - 1];  // last byte written (RLE reference)
                            const uint32_t b0 = la[0] & 0xFFu;
                            const bool runstart = b0 == lastb && (q + 1 >= N || ((la[0] >> 8) & 0xFFu) == lastb);
                            if (runstart) {
                                // RLE accumulation (compressor.c:471-523), the run counted to its end in one go
                                int pp = q, rle = 0;
                                for (;;) {
                                    const int r = N - pp < 16 ? N - pp : 16;
                                    uint32_t w[4];
                                    smem::load16(sIn + (uint32_t)pp, w);
                                    const uint32_t bl = lastb * 0x01010101u;
                                    int avail = 16;
#pragma unroll
                                    for (int i = 3; i >= 0; i--) {
                                        const uint32_t x = w[i] ^ bl;
                                        if (x) avail = 4 * i + ((__ffs(x) - 1) >> 3);
                                    }
                                    if (avail > r) avail = r;
                                    if (avail > kRleMax - rle) avail = kRleMax - rle;
                                    const int total = rle + avail;
                                    const bool ended = avail < r || total >= kRleMax;
                                    if (!ended && total > 0) {
                                        rle = total;
                                        pp += avail;
                                        if (pp < N) continue;
                                        // the input ends inside the run: flush drains it (compressor.c:750-770)
                                        bestkey = (rle == 1 ? kKindLone : kKindRun) << 11;
                                        step = N - q;
                                        break;
                                    }
                                    if (total >= 2) {
                                        bool use_rle = true;
                                        if (total == avail && total <= 6) use_rle = !(len > total);  // short run: a longer match wins
                                        if (use_rle) {
                                            // the window gets 8 bytes only: parse-dependent from here on unless the stream ends
                                            if (total > kRleWindowMax && pp + avail < N) giveup = true;
                                            bestkey = kKindRun << 11;
                                            step = pp + avail - q;
                                        }
                                    }
                                    break;
                                }
                            }
                            if ((bestkey >> 11) <= 16u && len > 2 + 11) {  // extended match (not taken over by a run)
                                int xlen = len;
                                uint32_t xpos = 1022u - (bestkey & 1023u);
                                if (len == 16 && (xkey >> 16) >= 16u) {
                                    xlen = (int)(xkey >> 16);
                                    xpos = 0xFFFFu - (xkey & 0xFFFFu);
                                }
                                bestkey = (kKindExt << 11) | xpos;
                                step = xlen;
                            }
                        }
                        smem::st16(sBestIn + 2u * (uint32_t)q, bestkey);
                        newmask |= 1u << p;
                        pn = p + step;
                        adv = true;
                    }
                }
            }
            if (budget <= 0) break;
            // my entry is my left neighbour's exit
            int prev_exit = __shfl_up_sync(kFull, exitrel, 1);
            if (lane == 0) prev_exit = 0;
            const bool changed = prev_exit != entry && nvalid > 0;
            entry = prev_exit;
            if (!__any_sync(kFull, changed)) break;
        }
        if (EXT && __any_sync(kFull, giveup)) budget = 0;
        if (budget <= 0) {  // pathological chains (or a long run mid-stream): the bitmap kernel takes the stream
            if (lane == 0) {
                a.b.out_sizes[stream] = kDeferred;
                atomicAdd(&d_walk_deferred_total, 1u);
            }
            continue;
        }
        __syncwarp();  // link[] is dead: the token list overwrites it

        // ---- P3: token list: lane b contributes the tokens of segment b --------------------------------------
        int ntok;
        {
            const uint32_t mymask = path & validmask;
            const int cnt = __popc(mymask);
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(kFull, incl, d);
                if (lane >= d) incl += t;
            }
            ntok = __shfl_sync(kFull, incl, 31);
            int ti = incl - cnt;
            for (uint32_t m = mymask; m; m &= m - 1) tok[ti++] = (uint16_t)(segbase + __ffs(m) - 1);
        }

        // ---- P4: bit pack, 32 tokens at a time: warp prefix sum of the bit lengths, every lane ORs its token into
        // the MSb-first staging line ----------------------------------------------------------------------------------
        __syncwarp();
        const uint32_t hdr_bits = (a.flags & TB_F_DICT_RESET) ? 16u : 8u;
        uint32_t nbits = hdr_bits;
        int res = kOk;
        if (lane == 0) {
            const uint32_t header = ((uint32_t)(wbits - 8) << 5) | ((uint32_t)(lbits - 5) << 3) |
                                    ((a.flags & TB_F_CUSTOM_DICT) ? 4u : 0u) | (EXT ? 2u : 0u) | ((a.flags & TB_F_DICT_RESET) ? 1u : 0u);
            stage[0] = stream_appends(a.flags, stream) ? kAppendStart : header << 24;
        }
        __syncwarp();
        for (int tb0 = 0; tb0 < ntok; tb0 += 32) {
            const int i = tb0 + lane;
            uint32_t bits = 0;
            int nb = 0;
            bool misfit = false;
            if (i < ntok) {
                const int q = (int)tok[i];
                const uint32_t v = best[q];
                const int len = (int)(v >> 11);  // EXT: 17 = run, 18 = lone run byte, 19 = extended match
                if (len < 2 || (EXT && len == (int)kKindLone)) {
                    const uint32_t c = comb[q];
                    misfit = lbits < 8 && (c >> lbits) && !(EXT && len == (int)kKindLone);  // the lone run byte is not checked (:512-523)
                    bits = (1u << lbits) | c;
                    nb = lbits + 1;
                } else if (!EXT || len <= 2 + 11) {
                    const uint32_t h = lut[len - 2];
                    bits = ((h & 0xFFFFu) << wbits) | (1022u - (v & 1023u));
                    nb = (int)(h >> 16) + wbits;
                } else {
                    // write_rle_token (:342-350): symbol 12 + exthuff(count - 2, 4 raw bits);
                    // write_extended_match_token (:387-398): symbol 13 + exthuff(len - 14, 3 raw bits) + position
                    const int span = (i + 1 < ntok ? (int)tok[i + 1] : N) - q;  // bytes the token covers
                    const bool is_rle = len == (int)kKindRun;
                    const int t = is_rle ? 4 : 3;
                    const int val = is_rle ? span - 2 : span - 14;
                    const uint32_t h = lut[val >> t];
                    const int xn = (int)(h >> 16) - 1 + t;
                    const uint32_t x = ((h & 0xFFFFu) << t) | (uint32_t)(val & ((1 << t) - 1));
                    const uint32_t sym = lut[is_rle ? kSymRle : kSymExt];
                    bits = ((sym & 0xFFFFu) << xn) | x;
                    nb = (int)(sym >> 16) + xn;
                    if (!is_rle) {
                        bits = (bits << wbits) | (v & 1023u);
                        nb += wbits;
                    }
                }
            }
            if (lbits < 8) {  // a literal that does not fit ends the stream (compressor.c:629-631)
                const uint32_t mis = __ballot_sync(kFull, misfit);
                if (mis) {
                    if (lane >= __ffs(mis) - 1) nb = 0;
                    res = kExcessBits;
                    ntok = 0;
                }
            }
            int incl = nb;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(kFull, incl, d);
                if (lane >= d) incl += t;
            }
            if (nb) {
                const uint32_t start = nbits + (uint32_t)(incl - nb);
                const uint32_t wi = start >> 5, o = start & 31u;
                const uint64_t sv = (uint64_t)bits << (64 - nb - (int)o);
                atomicOr(&stage[wi], (uint32_t)(sv >> 32));
                if ((uint32_t)sv) atomicOr(&stage[wi + 1], (uint32_t)sv);
            }
            nbits += (uint32_t)__shfl_sync(kFull, incl, 31);
        }
        __syncwarp();
        uint32_t out_bytes;
        if (res == kOk) {
            if (ends_with_flush(a.write_token, nbits, a.flags, stream, (uint64_t)N)) {  // compressor.c:784-794
                if (lane == 0) {
                    const uint32_t wi = nbits >> 5, o = nbits & 31u;
                    const uint64_t sv = (uint64_t)kHuff.code[kSymFlush] << (64 - kHuff.bits[kSymFlush] - (int)o);
                    stage[wi] |= (uint32_t)(sv >> 32);
                    stage[wi + 1] |= (uint32_t)sv;
                }
                nbits += kHuff.bits[kSymFlush];
            }
            out_bytes = (nbits + 7u) >> 3;
        } else {
            out_bytes = nbits >> 3;  // the reference has drained whole bytes of everything before the failing poll
        }
        __syncwarp();
        {
            const uint32_t nwords = out_bytes >> 2;
            for (uint32_t wi = lane; wi < nwords; wi += 32) out32[wi] = __byte_perm(stage[wi], 0, 0x0123);
            const uint32_t tail = out_bytes & 3u;
            if ((uint32_t)lane < tail)
                reinterpret_cast<uint8_t *>(out32 + nwords)[lane] = (uint8_t)(stage[nwords] >> (24 - 8 * lane));
        }
        if (lane == 0) {
            a.b.out_sizes[stream] = out_bytes;
            if (a.b.status) a.b.status[stream] = (int8_t)res;
        }
    }
}

}  // namespace

#ifndef TB_EMU
bool launch_walk_compress_batch(const CompBatchConf &cf, const uint8_t *d_dict, const BatchArgs &b, cudaStream_t st) {
    if (cf.window > 10) return false;
    if (cf.flags & TB_F_LAZY) return false;                  // greedy only (lazy matching: ppar_compress.cu)
    const bool ext = (cf.flags & TB_F_EXTENDED) != 0;
    if (b.in_offsets) return false;                          // strided layout only
    if (b.in_stride > (1u << cf.window)) return false;       // streams no longer than the window
    if ((b.in_stride & 15) || ((uintptr_t)b.in & 15) || (b.out_stride & 3) || ((uintptr_t)b.out & 3)) return false;
    if ((uintptr_t)d_dict & 3) return false;
    const uint64_t bound = 2 + (b.in_stride * (uint64_t)(cf.literal + 1) + 7) / 8 + 6;
    if (b.out_stride < ((bound + 3) & ~3ull)) return false;  // never OUTPUT_FULL in this kernel
    if (b.n_streams == 0) return true;

    WalkArgs a;
    a.b = b;
    a.dict = d_dict;
    a.window_bits = cf.window;
    a.literal = cf.literal;
    a.flags = cf.flags;
    a.write_token = cf.write_token;
    a.max_pairs = kMaxPairs;
    static int blocks_per_sm = 0, sms = 0;
    if (!blocks_per_sm) {
        cudaFuncSetAttribute(k_walk_compress<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CTA_BYTES);
        cudaFuncSetAttribute(k_walk_compress<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CTA_BYTES);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_walk_compress<true>, kWarps * 32, CTA_BYTES);
        if (blocks_per_sm < 1) blocks_per_sm = 1;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const uint64_t want = (b.n_streams + kWarps - 1) / kWarps;
    const uint64_t persistent = (uint64_t)sms * blocks_per_sm;  // grid = SM count x resident CTAs
    if (ext)
        k_walk_compress<true><<<(unsigned)(want < persistent ? want : persistent), kWarps * 32, CTA_BYTES, st>>>(a);
    else
        k_walk_compress<false><<<(unsigned)(want < persistent ? want : persistent), kWarps * 32, CTA_BYTES, st>>>(a);
    count_launch();
    // second pass: the bitmap kernel picks up the streams marked kDeferred (usually none; it then only scans the sizes)
    static unsigned int *h_seen = nullptr;  // pinned mirror of d_walk_deferred_total
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
    const bool ok = launch_fast_compress_batch(cf, d_dict, b, st, /*only_deferred=*/true, /*small_grid=*/!expect_work);
    if (h_seen) cudaMemcpyFromSymbolAsync(h_seen, d_walk_deferred_total, sizeof *h_seen, 0, cudaMemcpyDeviceToHost, st);
    return ok;
}
#endif  // TB_EMU

}  // namespace tb
