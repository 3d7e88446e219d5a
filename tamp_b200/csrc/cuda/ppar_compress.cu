// Position-parallel batch compressor for streams no longer than the window (N <= W <= 1024): v1 format, v1 with
// lazy matching, and the extended (v2) format; and, lap by lap, for v1 streams longer than the window.
//
// Why this exists.  In the v1 format every consumed input byte is appended to the window in order
// (tamp_compressor_poll, compressor.c:652-657), so the window a poll at input offset p sees does not depend on
// how the bytes before p were tokenised: for N <= W it is simply
//
//     window_p[x] = x < p ? input[x] : dictionary[x]                      (no wrap, window_pos == p)
//
// Hence find_best_match (compressor_find_match_desktop.c:82-167) can be evaluated for EVERY offset p
// independently — no serial dependency, no per-token scalar bookkeeping — and only the walk
// p <- p + len(p) and the bit packing remain sequential.  Instead of scanning the whole
// window per token (the reference; fast_compress.cu does it with bitmaps), each offset only visits the
// window positions that hold its first two bytes, found through hash chains:
//
//   P1  chain build: for blocks of 32 offsets (lane = offset) link each offset to the previous offset with
//       the same bigram hash (table lookup for earlier blocks, __match_any_sync inside the block).  The
//       table starts out holding the dictionary's own chain heads, so a chain runs through the input
//       offsets (newest first) and then on through the dictionary positions (highest first).  The chain
//       populations are summed on the way: streams with too many candidate pairs (runs, short periods) are
//       left to the bitmap kernel, whose cost does not depend on the data (pick-up pass);
//   P2  match table: offsets without any candidate are settled 32 at a time; the others go to persistent
//       lanes: a lane owns one offset p at a time and walks its candidates (x = p-1, then the chain), one
//       16-byte compare per iteration, keeping max length / lowest index (the reference's tie-break and
//       early-exit result).  Lanes that run out of candidates take the next queued offsets (ballot + popc),
//       so the lanes stay busy whatever the chain lengths;
//   P3  parse -> token list.  v1: greedy (literal if len < 2; tamp_compressor_poll's decision, :625-649) and
//       without a serial walk — per block of 32 offsets, pointer doubling in registers gives every offset the
//       set of offsets its walk visits inside the block and where it leaves it; 32 dependent lookups stitch the
//       blocks.  The extended format does the same with its run / extended-match offsets handled one at a time;
//       lazy matching walks serially (see the kernel's comment);
//   P4  static-Huffman bit pack: 32 tokens at a time, warp prefix sum of the bit lengths, tokens ORed into an
//       MSb-first staging line, coalesced stores (write_to_bit_buffer / partial_flush / flush, :49-75, :728-810).
//
// Streams LONGER than the window (kModeLaps / kModeLazyLaps, v1 only; kernel mode 4 until it has GPU numbers): the same machinery one
// lap of W offsets at a time.  At offset p of lap k the window holds this lap's bytes below p - base and the previous
// lap's bytes above — "previous lap" in place of "dictionary" — so each lap rebuilds the old side's chains from its
// bytes, runs P1..P4 on its W offsets (16 bytes of lookahead past the lap's end), and carries the walk's entry offset
// and the partial output word into the next lap.
//
// One warp per stream, one CTA per SM; the dictionary, its chain links and chain heads are staged once per CTA.
// Lookahead at offset p is min(15, N - p) bytes in v1 (min_pattern_size is 2 for every window <= 10, so
// MAX_PATTERN_SIZE = 15) and min(16, N - p) in the extended format: compress_cb / flush only ever poll a ring
// holding min(16, N - p) bytes.
#include "../tb_wire.h"
#include "tb_cuda.h"
#include "tb_device_common.cuh"

namespace tb {

namespace {

constexpr int kMaxN = 1024;
constexpr int kPad = 32;            // readable slack behind the byte arrays (unaligned 20-byte reads)
constexpr int kHashBits = 11, kHashSize = 1 << kHashBits;
constexpr uint32_t kFull = 0xffffffffu;
constexpr uint32_t kNone = 0xFFFFu;
constexpr uint32_t kFromW = 0xFFFFFFF0u;  // lazy + laps: "the next candidate is the chain start of offset W"
constexpr int kMaxLenV1 = 15;       // v1: min_pattern_size (2) + 13
constexpr int kMaxLenExt = 16;      // extended format: the 16-byte input ring is the limit
enum { kModeV1 = 0, kModeLazy = 1, kModeExt = 2, kModeLaps = 3, kModeLazyLaps = 5 };
constexpr int kExtCap = 2 + 11 + kExtExtraMax;  // longest extended match: min_pattern + 11 + 120
constexpr int kMaxPairs = 8192;     // chain population above which a stream goes to the bitmap kernel (typical text: ~3000)
constexpr int kRefillMin = 8;       // idle lanes that trigger handing out new offsets in P2
constexpr int kStageWords = (kMaxN * 9 / 8 + 16 + 3) / 4;

// Index space of candidates: [0, 1024) input offsets, [1024, 2048) dictionary positions (+ 1024).
// per-CTA shared memory: the dictionary side, shared by all warps
constexpr int D_BYTES = 0;                               // dictionary bytes
constexpr int D_LINK = D_BYTES + kMaxN + kPad;           // u16: chain links of the dictionary positions
constexpr int D_HEAD = D_LINK + 2 * kMaxN;               // u16: chain heads of the dictionary (encoded + 1024)
constexpr int D_LUT = D_HEAD + 2 * kHashSize;
constexpr int D_END = D_LUT + 64;
// per-warp shared memory: the input side
constexpr int OFF_COMB = 0;                              // input bytes
constexpr int OFF_LINK = OFF_COMB + kMaxN + kPad;        // u16 link[p]: next candidate of the chain through p
constexpr int OFF_HEAD = OFF_LINK + 2 * kMaxN;           // u16 head[h] (P1); afterwards best / exits / staging line
constexpr int PER_WARP_BASE = OFF_HEAD + 2 * kHashSize;
// the link array is dead after P2: P3's visit masks live there; the hash table is dead after P1
constexpr int OFF_VISIT = OFF_LINK;                      // u32 visit[16 * block + entry offset]
constexpr int OFF_BEST = OFF_HEAD;                       // u16 best[p] = len << 10 | index (P2 onwards)
constexpr int OFF_EXIT = OFF_BEST + 2 * kMaxN;           // u8 exit[16 * block + entry offset]
constexpr int OFF_STAGE = OFF_EXIT + kMaxN / 2;          // u32 stage[kStageWords]
constexpr int OFF_QUEUE = OFF_EXIT;                      // u16 queue[1024]: offsets with at least one candidate (P2, non-lazy)
constexpr int OFF_BEST_NEXT = PER_WARP_BASE;             // lazy matching only: u16 table of the p+1 matches
// laps only, behind the lazy table if there is one: the previous lap's bytes and the chain links over them
// extended format: the walk still follows chain links, so its token list goes over the (dead) work queue and the
// staging line — only needed once the walk is over — over the links
constexpr int OFF_TOK_EXT = OFF_QUEUE;
constexpr int OFF_STAGE_EXT = OFF_LINK;
// One CTA per SM with as many warps (= streams in flight) as shared memory takes, less 8 KiB: the pick-up pass of
// the bitmap kernel (one-warp CTAs, 6.4 KiB each) must find room beside this CTA, or it would wait for it to retire
// and serialise the chunks of the pipelined host path.
constexpr int kSmemBudget = 227 * 1024 - 8 * 1024;
template <int MODE>
struct Lay {
    static constexpr bool kLazy = MODE == kModeLazy || MODE == kModeLazyLaps;
    static constexpr bool kLaps = MODE == kModeLaps || MODE == kModeLazyLaps;
    static constexpr int OFF_OLD_BYTES = PER_WARP_BASE + (kLazy ? 2 * kMaxN : 0);
    static constexpr int OFF_OLD_LINK = OFF_OLD_BYTES + kMaxN + kPad;
    static constexpr int PER_WARP = PER_WARP_BASE + (kLazy ? 2 * kMaxN : 0) + (kLaps ? kMaxN + kPad + 2 * kMaxN : 0);
    static constexpr int kWarps = (kSmemBudget - D_END) / PER_WARP < 32 ? (kSmemBudget - D_END) / PER_WARP : 32;
    static constexpr int CTA_BYTES = D_END + kWarps * PER_WARP;
    static_assert(PER_WARP % 16 == 0, "aligned regions");
};
static_assert(D_END % 16 == 0 && OFF_LINK % 16 == 0 && OFF_HEAD % 16 == 0 && D_HEAD % 16 == 0, "aligned regions");
static_assert(2 * kMaxN + kMaxN / 2 + 4 * kStageWords <= 2 * kHashSize, "best + exits + staging line fit the dead hash table");
static_assert(OFF_QUEUE + 2 * kMaxN <= PER_WARP_BASE, "the work queue fits behind the match table");
static_assert(4 * kStageWords <= 2 * kMaxN, "the staging line fits the link array");

// Streams deferred so far (cumulative).  The host reads a pinned copy that trails by a launch or two and uses it
// only to size the pick-up pass: a full grid while deferrals are being seen, one warp per SM otherwise.
__device__ unsigned int d_deferred_total = 0;

struct PparArgs {
    BatchArgs b;
    const uint8_t *dict;
    int window_bits, literal, flags, write_token;
    int max_pairs;  // streams with more chain pairs than this are deferred
};

__device__ __forceinline__ uint32_t bigram_hash(uint32_t key16) { return (key16 * 2654435761u) >> (32 - kHashBits); }

// Explicit .shared loads: candidate-indexed accesses pick the input side or the dictionary side with a select of
// two 32-bit shared addresses (no generic-pointer arithmetic in the loop).
#ifndef TB_EMU
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
#else  // tests/emu: the kernel stepped on the CPU (test infrastructure; see tests/emu/cuda_emu.h)
inline uint32_t lds32(uint32_t a) { return *reinterpret_cast<const uint32_t *>(emu_shared_ptr(a)); }
inline uint32_t lds16(uint32_t a) { return *reinterpret_cast<const uint16_t *>(emu_shared_ptr(a)); }
#endif

// 16 bytes starting at shared byte address `sa` (any alignment; the arrays have kPad slack behind them).
__device__ __forceinline__ void load16(uint32_t sa, uint32_t (&w)[4]) {
    const uint32_t q = sa & ~3u;
    const int sh = (int)(sa & 3u) * 8;
    const uint32_t a0 = lds32(q), a1 = lds32(q + 4), a2 = lds32(q + 8), a3 = lds32(q + 12), a4 = lds32(q + 16);
    w[0] = __funnelshift_r(a0, a1, sh);
    w[1] = __funnelshift_r(a1, a2, sh);
    w[2] = __funnelshift_r(a2, a3, sh);
    w[3] = __funnelshift_r(a3, a4, sh);
}

// 4 bytes starting at shared byte address `sa` (any alignment).
__device__ __forceinline__ uint32_t load4(uint32_t sa) {
    const uint32_t q = sa & ~3u;
    return __funnelshift_r(lds32(q), lds32(q + 4), (int)(sa & 3u) * 8);
}

// Hash table entries: bits 0..10 = index of the newest entry with this hash (kHeadNone: none), bits 11..15 = how many
// INPUT offsets the chain holds so far (saturating) — the chain population, summed up per stream to spot inputs
// whose chains are so long (runs, short periods) that the bitmap kernel is the cheaper one.
constexpr uint32_t kHeadNone = 0x7FFu, kCountMax = 31u;

// Link the n-1 bigrams of bytes[0, n) into per-hash chains, newest first: link[p] = index of the
// previous entry with the same bigram hash, or whatever head[h] held before (none / a dictionary
// position) for the first one.  Stored indices are offset by `first`.  Returns this lane's share of the
// number of (offset, earlier offset with the same hash) pairs.  Warp-cooperative.
__device__ __forceinline__ uint32_t build_chains(const uint8_t *bytes, int n, int first, uint16_t *head, uint16_t *link,
                                                 int lane, int limit) {  // limit: entries link[] has (n or n - 1)
    uint32_t pairs = 0;
    for (int base = 0; base < n; base += 32) {
        const int p = base + lane;
        const bool valid = p + 1 < n;
        uint32_t h = 0x10000u | (uint32_t)lane;  // invalid lanes match nobody
        if (valid) h = bigram_hash((uint32_t)bytes[p] | ((uint32_t)bytes[p + 1] << 8));
        const uint32_t peers = __match_any_sync(kFull, h);
        uint32_t count = 0;
        if (valid) {
            const uint32_t hv = head[h];
            const uint32_t lower = peers & ((1u << lane) - 1u);
            const uint32_t older = hv & kHeadNone;
            const uint32_t pv = lower ? (uint32_t)(first + base + 31 - __clz(lower)) : (older == kHeadNone ? kNone : older);
            link[p] = (uint16_t)pv;
            count = (hv >> 11) + __popc(lower);  // input offsets before p on this chain
            pairs += count;
        } else if (p < n && p < limit) {
            link[p] = (uint16_t)kNone;
        }
        __syncwarp();
        if (valid && (peers >> lane) == 1u) {  // the block's last entry with this hash
            const uint32_t c = count + 1 < kCountMax ? count + 1 : kCountMax;
            head[h] = (uint16_t)((uint32_t)(first + p) | (c << 11));
        }
        __syncwarp();
    }
    return pairs;
}

// LAZY: the reference's lazy matching (compressor.c:176-189, :576-619; TAMP_LAZY_MATCHING builds with
// conf.lazy_matching): a match of 2..8 bytes is dropped for a literal when the search one byte further on — against
// the SAME window, i.e. before this byte is appended — finds a longer one that does not cover the window position
// being written; that match is then used at the next poll.  Needs a second table (matches of input[p+1..] in
// window_p) and a serial walk; both fall out of the same candidate machinery.
//
// EXT: the extended (v2) format (compressor.c:437-525, :342-415).  Its RLE and extended-match tokens write at most 8
// bytes / the rest of the window, so in general the window depends on the parse — but as long as no run longer than 8
// bytes has been emitted, every consumed byte has been written and the window IS the v1 window.  The match table
// (16-byte lookahead) is therefore valid until that first long run; streams that emit one before their end are left to
// the bitmap kernel.  Run counting against the previous byte, the short-run-versus-match rule and extended matches (a
// match of 14+ bytes keeps growing, up to 133 bytes, against the window as it was at its start) are handled one offset
// at a time; the plain steps in between are parsed like v1.
template <int MODE>
__global__ void __launch_bounds__(Lay<MODE>::kWarps * 32) k_ppar_compress(PparArgs a) {
    constexpr bool LAZY = Lay<MODE>::kLazy, EXT = MODE == kModeExt, LAPS = Lay<MODE>::kLaps;
    constexpr int kMaxLen = EXT ? kMaxLenExt : kMaxLenV1;
    constexpr int PER_WARP = Lay<MODE>::PER_WARP, kWarps = Lay<MODE>::kWarps;
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
    uint16_t *dlink = reinterpret_cast<uint16_t *>(smem + D_LINK);
    uint16_t *dhead = reinterpret_cast<uint16_t *>(smem + D_HEAD);
    uint32_t *lut = reinterpret_cast<uint32_t *>(smem + D_LUT);
    uint8_t *wbase = smem + D_END + warp * PER_WARP;
    uint8_t *comb = wbase + OFF_COMB;
    uint16_t *link = reinterpret_cast<uint16_t *>(wbase + OFF_LINK);
    uint16_t *head = reinterpret_cast<uint16_t *>(wbase + OFF_HEAD);
    uint16_t *best = reinterpret_cast<uint16_t *>(wbase + OFF_BEST);
    uint16_t *best_next = reinterpret_cast<uint16_t *>(wbase + OFF_BEST_NEXT);  // LAZY only
    uint8_t *oldb = wbase + Lay<MODE>::OFF_OLD_BYTES;                                   // LAPS only
    uint16_t *oldlink = reinterpret_cast<uint16_t *>(wbase + Lay<MODE>::OFF_OLD_LINK);  // LAPS only
    uint16_t *queue = reinterpret_cast<uint16_t *>(wbase + OFF_QUEUE);          // !LAZY only
    uint32_t *visit = reinterpret_cast<uint32_t *>(wbase + OFF_VISIT);
    uint8_t *exits = wbase + OFF_EXIT;
    // token list: over the (dead) visit masks / chain links; the extended-format walk still follows links, so there it
    // has its own region
    uint16_t *tok = reinterpret_cast<uint16_t *>(wbase + (EXT ? OFF_TOK_EXT : OFF_VISIT));
    uint32_t *stage = reinterpret_cast<uint32_t *>(wbase + (EXT ? OFF_STAGE_EXT : OFF_STAGE));
    // shared addresses for candidate-indexed accesses: index i < 1024 -> input side, else dictionary side.
    // (The base goes through an empty asm so that the compiler keeps it in a register instead of re-deriving
    // the shared window address — S2R + LEA — inside the loops.)
    uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("" : "+r"(sbase));
    const uint32_t sBytesIn = sbase + (uint32_t)(D_END + warp * PER_WARP + OFF_COMB);
    const uint32_t sBytesDict = sbase + (uint32_t)D_BYTES - kMaxN;
    const uint32_t sLinkIn = sbase + (uint32_t)(D_END + warp * PER_WARP + OFF_LINK);
    const uint32_t sLinkDict = sbase + (uint32_t)D_LINK - 2 * kMaxN;

    // ---- once per CTA: dictionary bytes, their chains, the chain heads -----------------------------------
    for (int i = threadIdx.x; i < (kMaxN + kPad) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t *>(dictb)[i] = i * 4 < W ? reinterpret_cast<const uint32_t *>(a.dict)[i] : 0u;
    for (int i = threadIdx.x; i < kHashSize / 2; i += blockDim.x) reinterpret_cast<uint32_t *>(dhead)[i] = kHeadNone * 0x10001u;
    if (threadIdx.x < 16) lut[threadIdx.x] = (uint32_t)kHuff.code[threadIdx.x] | ((uint32_t)kHuff.bits[threadIdx.x] << 16);
    __syncthreads();
    if (warp == 0) build_chains(dictb, W, kMaxN, dhead, dlink, lane, W);
    __syncthreads();
    for (int i = threadIdx.x; i < kHashSize; i += blockDim.x) dhead[i] &= (uint16_t)kHeadNone;  // populations count input offsets only
    __syncthreads();

    const uint64_t nwarps = (uint64_t)gridDim.x * kWarps;
    for (uint64_t stream = (uint64_t)blockIdx.x * kWarps + warp; stream < a.b.n_streams; stream += nwarps) {
        const uint8_t *src = a.b.in + stream * a.b.in_stride;
        const int N = a.b.in_sizes ? (int)a.b.in_sizes[stream] : (int)a.b.in_stride;
        uint32_t *out32 = reinterpret_cast<uint32_t *>(a.b.out + stream * a.b.out_stride);
        // LAPS: a stream longer than the window is parsed one lap of W offsets at a time.  In the v1 format the window at
        // offset p of lap k holds input[base .. p) below p - base and the previous lap's bytes above (the dictionary in
        // lap 0), i.e. the situation of a short stream with "previous lap" in place of "dictionary": same chains, same
        // candidate walk.  The old side's chains are rebuilt from its bytes at the start of each lap; the walk's entry
        // offset and the partial output word carry over.
        const uint8_t *oldbytes = dictb;                 // what window positions at or above the write position hold
        uint32_t sOldBytes = sBytesDict, sOldLink = sLinkDict;
        int entry = 0;                                   // first token start of this lap (a token may straddle the boundary)
        uint32_t carry_bits = 0, carry_word = 0, words_done = 0;
        bool carry_cached = false;   // LAZY + LAPS: the previous lap's last poll left a cached match ...
        uint32_t carry_match = 0;    // ... this one (its entry of the p + 1 table)
        for (int base = 0;; base += W) {
        const int rem = N - base;                        // bytes from this lap's first offset to the end of the stream
        const int nl = (LAPS && rem > W) ? W : rem;      // offsets parsed in this pass (all of them unless LAPS)
        const bool last_lap = !LAPS || rem <= W;

        // ---- P0: input by coalesced 128-bit loads; hash table := the dictionary's chain heads -------------
        __syncwarp();
        if (LAPS && base) {
            for (int off = lane * 16; off < W; off += 512)
                *reinterpret_cast<uint4 *>(oldb + off) = *reinterpret_cast<const uint4 *>(comb + off);
            for (int i = lane; i < kHashSize / 2; i += 32) reinterpret_cast<uint32_t *>(head)[i] = kHeadNone * 0x10001u;
            __syncwarp();
            build_chains(oldb, W, kMaxN, head, oldlink, lane, W);
            for (int i = lane; i < kHashSize / 2; i += 32) reinterpret_cast<uint32_t *>(head)[i] &= kHeadNone * 0x10001u;
            oldbytes = oldb;
            sOldBytes = sbase + (uint32_t)(D_END + warp * PER_WARP + Lay<MODE>::OFF_OLD_BYTES) - kMaxN;
            sOldLink = sbase + (uint32_t)(D_END + warp * PER_WARP + Lay<MODE>::OFF_OLD_LINK) - 2 * kMaxN;
        }
        {
            const int want = LAPS ? (rem < W + 16 ? rem : W + 16) : nl;  // a lap reads 16 bytes of lookahead past its end
            for (int off = lane * 16; off < want; off += 512)
                *reinterpret_cast<uint4 *>(comb + off) = __ldg(reinterpret_cast<const uint4 *>(src + base + off));
        }
        if (!LAPS || base == 0)
            for (int i = lane; i < 2 * kHashSize / 16; i += 32)
                reinterpret_cast<uint4 *>(head)[i] = reinterpret_cast<const uint4 *>(dhead)[i];
        __syncwarp();

        // ---- P1: hash chains over the input -----------------------------------------------------------------
        {
            // (a lap followed by more input also links its last offset: that bigram's second byte is the next lap's first)
            const int nchain = LAPS ? (rem < W + 1 ? rem : W + 1) : nl;
            const uint32_t pairs = __reduce_add_sync(kFull, build_chains(comb, nchain, 0, head, link, lane, LAPS ? W : nchain));
            if ((!LAPS || base == 0) && pairs > (uint32_t)a.max_pairs) {
                // Chains this long make the candidate walk the slower way: leave the stream to the bitmap kernel
                // (launched right behind this one), whose cost does not depend on the data.
                if (lane == 0) {
                    a.b.out_sizes[stream] = kDeferred;
                    atomicAdd(&d_deferred_total, 1u);
                }
                break;
            }
        }
        // LAZY + LAPS: the p + 1 item of the lap's last offset starts at q = W, which has no link entry: its chain starts at
        // the newest entry of its bigram's bucket (read now, the table is about to be reused)
        uint32_t link_at_w = kNone;
        if (LAZY && LAPS && rem > W + 1) {
            const uint32_t hv = head[bigram_hash((uint32_t)comb[W] | ((uint32_t)comb[W + 1] << 8))] & kHeadNone;
            link_at_w = hv == kHeadNone ? kNone : hv;
        }
        if (LAZY && LAPS) __syncwarp();

        // ---- P2: best match for every offset (persistent lanes).  A work item is (q, bnd): the pattern starts at
        // input offset q and window positions below bnd hold input bytes, the rest dictionary bytes.  The normal
        // table has bnd = q; lazy matching adds the items (q = p + 1, bnd = p) --------------------------------------
        {
            bool work = false;
            int q = 0, bnd = 0, L = 0;
            uint32_t cand = kNone, from = 0;   // current candidate; index whose link yields the next one
            uint32_t la[4] = {0, 0, 0, 0}, bestkey = 0;
            uint16_t *dst = best;              // where this item's result goes
            int next_item = 0;
            // lazy: nl items of the normal table, then one p + 1 item for every offset that has a byte after it
            const int nl2 = LAPS ? (nl < rem - 1 ? nl : rem - 1) : nl - 1;
            int n_items = LAZY ? (nl > 0 ? nl + nl2 : 0) : 0;
            if constexpr (!LAZY) {
                // About half of the offsets have no candidate at all (literals for sure): settle them here, 32 at a
                // time, and queue the others, so that the persistent-lane loop only hands out offsets with work.
                for (int blk = 0; blk < nl; blk += 32) {
                    const int pp = blk + lane;
                    bool has = false;
                    if (pp < nl) {
                        const uint32_t c = link[pp];
                        const bool chain = c != kNone && !(c >= (uint32_t)kMaxN && (int)(c & (kMaxN - 1)) < pp);
                        const bool strad = pp >= 1 && comb[pp - 1] == comb[pp];
                        has = rem - pp >= 2 && (chain || strad);
                        if (!has) best[pp] = 0;
                    }
                    const uint32_t m = __ballot_sync(kFull, has);
                    if (has) queue[n_items + __popc(m & ((1u << lane) - 1u))] = (uint16_t)pp;
                    n_items += __popc(m);
                }
                __syncwarp();
            }
            // a chain ends with kNone, or at the first dictionary position below bnd (descending order): those
            // window positions hold input bytes
            auto live = [&](uint32_t c) { return c != kNone && !(c >= (uint32_t)kMaxN && (int)(c & (kMaxN - 1)) < bnd); };
            for (;;) {
                const uint32_t idle = __ballot_sync(kFull, !work);
                if (idle == kFull && next_item >= n_items) break;
                if (next_item < n_items && (idle == kFull || __popc(idle) >= kRefillMin)) {
                    const int mine = next_item + __popc(idle & ((1u << lane) - 1u));
                    next_item += __popc(idle);
                    if (!work && mine < n_items) {
                        if (LAZY && mine >= nl) {
                            bnd = mine - nl;
                            q = bnd + 1;
                            dst = best_next + bnd;
                        } else {
                            bnd = q = LAZY ? mine : (int)queue[mine];
                            dst = best + q;
                        }
                        load16(sBytesIn + (uint32_t)q, la);
                        L = rem - q < kMaxLen ? rem - q : kMaxLen;
                        bestkey = 0;
                        // x = bnd-1 holds input[bnd-1] followed by dictionary bytes: its bigram is not the input's, so
                        // the chain does not cover it; try it first when its first byte fits
                        if (bnd >= 1 && comb[bnd - 1] == (la[0] & 0xFFu)) {
                            cand = (uint32_t)(bnd - 1);
                            from = (LAZY && LAPS && q >= W) ? kFromW : (uint32_t)q;
                        } else {
                            cand = (LAZY && LAPS && q >= W) ? link_at_w : (uint32_t)link[q];
                            from = cand;
                        }
                        work = L >= 2 && live(cand);
                        if (!work) *dst = 0;
                    }
                }
                if (work) {
                    const uint32_t x = cand;
                    const bool in_dict = x >= (uint32_t)kMaxN;
                    const int xw = (int)(x & (kMaxN - 1));              // window index of the candidate
                    const int room = W - xw < L ? W - xw : L;           // a match never runs past the window end
                    const int inp = bnd - xw > 0 ? bnd - xw : 0;        // input bytes at the candidate (chain entries at or past bnd: none)
                    const int lim = in_dict ? room : (inp < room ? inp : room);
                    uint32_t w[4];
                    load16((in_dict ? sOldBytes : sBytesIn) + x, w);
                    const uint32_t d0 = w[0] ^ la[0], d1 = w[1] ^ la[1], d2 = w[2] ^ la[2], d3 = w[3] ^ la[3];
                    uint32_t d = d0;
                    int nb = 0;
                    if (!d) { d = d1; nb = 4; }
                    if (!d) { d = d2; nb = 8; }
                    if (!d) { d = d3; nb = 12; }
                    int n = d ? nb + ((__ffs(d) - 1) >> 3) : 16;
                    if (n >= lim) {
                        n = lim;
                        if (!in_dict) {  // ran into bnd: the window continues with dictionary (LAPS: previous lap) bytes
                            while (n < room && oldbytes[xw + n] == comb[q + n]) n++;
                        }
                    }
                    if (n >= 2) {
                        const uint32_t key = ((uint32_t)n << 16) | (0xFFFFu - (uint32_t)xw);
                        bestkey = key > bestkey ? key : bestkey;
                    }
                    cand = (LAZY && LAPS && from == kFromW) ? link_at_w
                                                            : lds16((from >= (uint32_t)kMaxN ? sOldLink : sLinkIn) + 2u * from);
                    from = cand;
                    if (!live(cand)) {
                        const uint32_t len = bestkey >> 16;
                        *dst = (uint16_t)(len ? (len << 10) | (0xFFFFu - (bestkey & 0xFFFFu)) : 0u);
                        work = false;
                    }
                }
            }
        }
        __syncwarp();

        // ---- P3: greedy parse -> token list (entry: offset | table << 10 | forced literal << 11) ----------------
        int ntok = 0;
        bool defer = false;
        bool next_cached = false;   // LAZY + LAPS: the walk's state at the end of this lap
        uint32_t next_match = 0;
        if constexpr (!LAZY && !EXT) {
            // Per block of 32 offsets: where does a walk entering at offset q leave the block, and which offsets does
            // it visit on the way (pointer doubling, 5 rounds in registers); 32 dependent lookups stitch the blocks.
            const int nblocks = (nl + 31) >> 5;
            for (int b = 0; b < nblocks; b++) {
                const int q = 32 * b + lane;
                const uint32_t v = q < nl ? best[q] : 0u;
                const int len = (int)(v >> 10);
                int J = lane + (len < 2 ? 1 : len);     // next offset of the walk, block-relative (>= 32: outside)
                uint32_t M = 1u << lane;
#pragma unroll
                for (int r = 0; r < 5; r++) {
                    const uint32_t tM = __shfl_sync(kFull, M, J & 31);
                    const int tJ = __shfl_sync(kFull, J, J & 31);
                    if (J < 32) {
                        M |= tM;
                        J = tJ;
                    }
                }
                if (lane < 16) {  // a walk enters a block at most 14 offsets in (tokens are at most 15 bytes long)
                    visit[16 * b + lane] = M;
                    exits[16 * b + lane] = (uint8_t)(J - 32);
                }
            }
            __syncwarp();
            uint32_t mymask = 0;  // lane b: offsets of block b where a token starts
            {
                int e = LAPS ? entry : 0;
                for (int b = 0; b < nblocks; b++) {
                    const uint32_t m = visit[16 * b + e];
                    e = exits[16 * b + e];
                    if (lane == b) mymask = m;
                }
                if (LAPS) entry = e;  // where the walk enters the next lap (a full lap is a whole number of blocks)
                const int left = nl - 32 * lane;  // offsets at or past the end are not tokens
                if (left < 32) mymask = left > 0 ? mymask & ((1u << left) - 1u) : 0u;
            }
            __syncwarp();  // visit[] is dead: the token list overwrites it
            // token list: lane b contributes the tokens of block b
            const int cnt = __popc(mymask);
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(kFull, incl, d);
                if (lane >= d) incl += t;
            }
            ntok = __shfl_sync(kFull, incl, 31);
            int ti = incl - cnt;
            for (uint32_t m = mymask; m; m &= m - 1) tok[ti++] = (uint16_t)(32 * lane + __ffs(m) - 1);
        } else if constexpr (LAZY) {
            // Lazy matching makes the step at p depend on whether the previous poll left a cached match: a serial
            // walk, the same in every lane (compressor.c:576-619 with window_pos == p).
            int p = LAPS ? entry : 0;
            bool cached = LAPS && carry_cached;
            while (p < nl) {
                const uint32_t m = cached ? (LAPS && p == 0 ? carry_match : (uint32_t)best_next[p - 1]) : (uint32_t)best[p];
                const int len = (int)(m >> 10);
                const int r = rem - p < 16 ? rem - p : 16;  // bytes in the input ring at this poll
                bool forced = false;
                if (len >= 2 && len <= 8 && r > len + 2) {
                    const uint32_t nx = best_next[p];
                    const int nlen = (int)(nx >> 10), nidx = (int)(nx & 1023u);
                    // the literal about to be written at window position p must not land inside the cached match
                    forced = nlen > len && !(p >= nidx && p < nidx + nlen);
                }
                if (lane == 0) tok[ntok] = (uint16_t)((uint32_t)p | (cached ? 1u << 10 : 0u) | (forced ? 1u << 11 : 0u));
                ntok++;
                if (forced) {
                    p += 1;
                    cached = true;
                } else {
                    p += len < 2 ? 1 : len;
                    cached = false;
                }
            }
            if (LAPS) {  // what the next lap's first poll starts from (P4 of this lap still needs the old carry)
                entry = p - nl;
                next_cached = cached;
                next_match = cached ? (uint32_t)best_next[nl - 1] : 0u;
            }
        } else {
            // Extended format.  Offsets where a run may start (the byte equals the one before it) or whose match is long
            // enough for an extended match (14+) are SPECIAL; everywhere else the step is the v1 step.  Entered with no
            // run pending, the outcome of a special offset — run token, lone byte, extended match, or the plain step when
            // a short run loses to a match — is a function of the offset alone, and it always starts a token there.  So
            // the walk is: per block of 32 offsets, pointer doubling in registers over the plain offsets (specials stop
            // it), then a short stitch that handles the specials the walk actually reaches, one at a time.
            // best[q] of a special token is rewritten for P4: length field 14 = extended match (+ its position),
            // 17 = run, 18 = lone run byte; the byte count of a run / extended match is the distance to the next token.
            const uint32_t dict_last = dictb[W - 1];  // RLE reference byte at stream start (specification.rst:219-222)
            auto special = [&](const int q) -> int {  // returns the offset of the next token; warp-uniform
                int p = q, rle = 0;
                for (;;) {
                    const int r = N - p < 16 ? N - p : 16;
                    const uint32_t last = p ? comb[p - 1] : dict_last;  // last byte written to the window
                    if (rle != 0 || comb[p] == last) {  // RLE accumulation (compressor.c:471-523)
                        uint32_t w[4];
                        load16(sBytesIn + (uint32_t)p, w);
                        const uint32_t bl = last * 0x01010101u;
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
                            p += avail;
                            if (p < N) continue;
                            // the input ends inside the run: flush drains it (compressor.c:750-770)
                            __syncwarp();
                            if (lane == 0) best[q] = (uint16_t)((rle == 1 ? 18u : 17u) << 10);
                            return N;
                        }
                        if (total >= 2) {
                            bool use_rle = true;
                            if (total == avail && total <= 6) use_rle = !((int)(best[p] >> 10) > total);  // short run: a longer match wins
                            if (use_rle) {
                                if (total > kRleWindowMax && p + avail < N) {  // the window gets 8 bytes only: parse-dependent from here on
                                    defer = true;
                                    return N;
                                }
                                __syncwarp();
                                if (lane == 0) best[q] = (uint16_t)(17u << 10);
                                return p + avail;
                            }
                        }
                    }
                    break;  // plain step at p == q (no run was pending)
                }
                const uint32_t m = best[q];
                const int len = (int)(m >> 10);
                if (len <= 2 + 11) return q + (len < 2 ? 1 : len);
                // Extended match: the longest match of input[q...] in the window as it is now, up to 133 bytes,
                // lowest position on ties (poll_extended_handling / find_extended_match restated: the candidate
                // set only ever shrinks, its lowest member is reported).  Only a full 16-byte match can grow.
                int xlen = len;
                uint32_t xpos = m & 1023u;
                if (len == 16) {
                    const int cap = N - q < kExtCap ? N - q : kExtCap;
                    uint32_t bestkey = 0;
                    uint32_t c = (q >= 1 && comb[q - 1] == comb[q]) ? (uint32_t)(q - 1) : kNone;  // x = q-1 first
                    uint32_t from = (uint32_t)q;
                    bool first = c != kNone;
                    for (int guard = 0; guard < 2 * kMaxN + 1; guard++) {  // chains are strictly descending: bounded anyway
                        if (!first) {
                            c = lds16((from >= (uint32_t)kMaxN ? sLinkDict : sLinkIn) + 2u * from);
                            from = c;
                        }
                        first = false;
                        if (c == kNone || (c >= (uint32_t)kMaxN && (int)(c & (kMaxN - 1)) < q)) break;
                        const int xw = (int)(c & (kMaxN - 1));
                        const int room = W - xw < cap ? W - xw : cap;
                        // every lane compares 4 bytes: window bytes below q come from the input, the rest from the dictionary
                        int n = 0;
                        for (int base = 0; base < room; base += 128) {
                            const int j = base + 4 * lane;
                            uint32_t diff = 0;
                            if (j < room) {
                                uint32_t wv = 0;
#pragma unroll
                                for (int k = 0; k < 4; k++) {
                                    const int y = xw + j + k;
                                    const uint32_t byte = y < q ? comb[y] : dictb[y];
                                    wv |= byte << (8 * k);
                                }
                                diff = wv ^ load4(sBytesIn + (uint32_t)(q + j));
                                const int valid = room - j;  // bytes of this word inside the limit
                                if (valid < 4) diff |= 0xffffffffu << (8 * valid);
                            } else {
                                diff = 1u;
                            }
                            const uint32_t bad = __ballot_sync(kFull, diff != 0u);
                            if (bad) {
                                const int fl = __ffs(bad) - 1;
                                const uint32_t fd = __shfl_sync(kFull, diff, fl);
                                n = base + 4 * fl + ((__ffs(fd) - 1) >> 3);
                                break;
                            }
                            n = base + 128;
                        }
                        if (n > room) n = room;
                        const uint32_t key = ((uint32_t)n << 16) | (0xFFFFu - (uint32_t)xw);
                        bestkey = key > bestkey ? key : bestkey;
                    }
                    if ((int)(bestkey >> 16) >= 16) {
                        xlen = (int)(bestkey >> 16);
                        xpos = 0xFFFFu - (bestkey & 0xFFFFu);
                    }
                }
                __syncwarp();  // every lane has read best[q] above
                if (lane == 0) best[q] = (uint16_t)((14u << 10) | xpos);
                return q + xlen;
            };

            uint32_t mymask = 0;  // lane b: offsets of block b where a token starts
            const int nblocks = (N + 31) >> 5;
            int p = 0;            // where the walk stands (warp-uniform)
            for (int b = 0; b < nblocks && !defer; b++) {
                if (p >= 32 * (b + 1)) continue;  // a long token jumped the whole block
                const int q = 32 * b + lane;
                int J = lane + 1;
                uint32_t M = 1u << lane;
                if (q < N) {
                    const int len = (int)(best[q] >> 10);
                    const uint32_t last = q ? comb[q - 1] : dict_last;
                    bool run_start = comb[q] == last;
                    // a byte that repeats the one before it but is not followed by another one is a run of one, which
                    // the poll hands to the plain step (total < 2) unless the input ends there (lone byte): not special
                    run_start = run_start && (q + 1 >= N || comb[q + 1] == last);
                    if (run_start || len > 2 + 11) {
                        J = lane;  // special: the doubling stops here
                        M = 0u;
                    } else {
                        J = lane + (len < 2 ? 1 : len);
                    }
                }
#pragma unroll
                for (int r = 0; r < 5; r++) {
                    const uint32_t tM = __shfl_sync(kFull, M, J & 31);
                    const int tJ = __shfl_sync(kFull, J, J & 31);
                    if (J < 32) {
                        M |= tM;
                        J = tJ;
                    }
                }
                uint32_t bm = 0;
                while (p < 32 * (b + 1) && p < N) {
                    const int e = p - 32 * b;
                    bm |= __shfl_sync(kFull, M, e);
                    const int j = __shfl_sync(kFull, J, e);
                    if (j >= 32) {
                        p = 32 * b + j;
                        break;
                    }
                    bm |= 1u << j;  // a special offset always starts a token
                    p = special(32 * b + j);
                    if (defer) break;
                }
                if (lane == b) mymask = bm;
            }
            if (!defer) {
                const int rem = N - 32 * lane;  // offsets at or past N are not tokens
                if (rem < 32) mymask = rem > 0 ? mymask & ((1u << rem) - 1u) : 0u;
                __syncwarp();
                const int cnt = __popc(mymask);
                int incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int t = __shfl_up_sync(kFull, incl, d);
                    if (lane >= d) incl += t;
                }
                ntok = __shfl_sync(kFull, incl, 31);
                int ti = incl - cnt;
                for (uint32_t m = mymask; m; m &= m - 1) tok[ti++] = (uint16_t)(32 * lane + __ffs(m) - 1);
            }
        }

        if (EXT && defer) {
            if (lane == 0) {
                a.b.out_sizes[stream] = kDeferred;
                atomicAdd(&d_deferred_total, 1u);
            }
            break;
        }

        __syncwarp();
        for (int i = lane; i < kStageWords; i += 32) stage[i] = 0u;

        // ---- P4: bit pack, 32 tokens at a time: warp prefix sum of the bit lengths, every lane ORs its token into
        // the MSb-first staging line ----------------------------------------------------------------------------------
        __syncwarp();
        const uint32_t hdr_bits = (a.flags & TB_F_DICT_RESET) ? 16u : 8u;
        uint32_t nbits = (LAPS && base) ? carry_bits : hdr_bits;  // later laps continue the partial word of the one before
        int res = kOk;
        if (lane == 0) {
            const uint32_t header = ((uint32_t)(wbits - 8) << 5) | ((uint32_t)(lbits - 5) << 3) |
                                    ((a.flags & TB_F_CUSTOM_DICT) ? 4u : 0u) | (EXT ? 2u : 0u) | ((a.flags & TB_F_DICT_RESET) ? 1u : 0u);
            stage[0] = (LAPS && base) ? carry_word : stream_appends(a.flags, stream) ? kAppendStart : header << 24;
        }
        __syncwarp();
        for (int tb0 = 0; tb0 < ntok; tb0 += 32) {
            const int i = tb0 + lane;
            uint32_t bits = 0;
            int nb = 0;
            bool misfit = false;
            if (i < ntok) {
                const uint32_t e = tok[i];
                const int q = (int)(e & 1023u);
                if constexpr (EXT) {
                    const int span = (i + 1 < ntok ? (int)(tok[i + 1] & 1023u) : N) - q;  // bytes an RLE / extended match covers
                    const uint32_t v = best[q];
                    const uint32_t lf = v >> 10;  // 0/1 literal, 2..13 match, 14 extended match, 17 run, 18 lone run byte
                    if (lf < 2 || lf == 18) {
                        const uint32_t c = comb[q];
                        misfit = lf != 18 && lbits < 8 && (c >> lbits);  // the lone run byte is not checked (:512-523)
                        bits = (1u << lbits) | c;
                        nb = lbits + 1;
                    } else if (lf <= 2 + 11) {
                        const uint32_t h = lut[lf - 2];
                        bits = ((h & 0xFFFFu) << wbits) | (v & 1023u);
                        nb = (int)(h >> 16) + wbits;
                    } else {
                        // write_rle_token (:342-350): symbol 12 + exthuff(count - 2, 4 raw bits);
                        // write_extended_match_token (:387-398): symbol 13 + exthuff(len - 14, 3 raw bits) + position
                        const bool is_rle = lf == 17;
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
                } else {
                    const uint32_t v = (LAZY && (e & 1024u)) ? (LAPS && q == 0 ? carry_match : (uint32_t)best_next[q - 1]) : (uint32_t)best[q];
                    const int len = (LAZY && (e & 2048u)) ? 0 : (int)(v >> 10);
                    if (len < 2) {
                        const uint32_t c = comb[q];
                        misfit = lbits < 8 && (c >> lbits);
                        bits = (1u << lbits) | c;
                        nb = lbits + 1;
                    } else {
                        const uint32_t h = lut[len - 2];
                        bits = ((h & 0xFFFFu) << wbits) | (v & 1023u);
                        nb = (int)(h >> 16) + wbits;
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
        if (LAPS && !last_lap && res == kOk) {
            // more laps follow: the whole words leave now, the partial one and its bit count carry over
            const uint32_t nwords = nbits >> 5;
            for (uint32_t wi = lane; wi < nwords; wi += 32) out32[words_done + wi] = __byte_perm(stage[wi], 0, 0x0123);
            carry_word = stage[nwords];
            carry_bits = nbits & 31u;
            words_done += nwords;
            if (LAZY) {
                carry_cached = next_cached;
                carry_match = next_match;
            }
            __syncwarp();  // stage[] is rewritten by the next lap
            continue;
        }
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
            for (uint32_t wi = lane; wi < nwords; wi += 32) out32[words_done + wi] = __byte_perm(stage[wi], 0, 0x0123);
            const uint32_t tail = out_bytes & 3u;
            if ((uint32_t)lane < tail)
                reinterpret_cast<uint8_t *>(out32 + words_done + nwords)[lane] = (uint8_t)(stage[nwords] >> (24 - 8 * lane));
        }
        if (lane == 0) {
            a.b.out_sizes[stream] = 4u * words_done + out_bytes;
            if (a.b.status) a.b.status[stream] = (int8_t)res;
        }
        break;
        }  // laps
    }
}

}  // namespace

#ifndef TB_EMU
template <int MODE>
static void launch_variant(const PparArgs &a, cudaStream_t st) {
    static int blocks_per_sm = 0, sms = 0;
    constexpr int kWarps = Lay<MODE>::kWarps, kBytes = Lay<MODE>::CTA_BYTES;
    if (!blocks_per_sm) {
        cudaFuncSetAttribute(k_ppar_compress<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBytes);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_ppar_compress<MODE>, kWarps * 32, kBytes);
        if (blocks_per_sm < 1) blocks_per_sm = 1;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const uint64_t want = (a.b.n_streams + kWarps - 1) / kWarps;
    const uint64_t persistent = (uint64_t)sms * blocks_per_sm;  // grid = SM count x resident CTAs
    k_ppar_compress<MODE><<<(unsigned)(want < persistent ? want : persistent), kWarps * 32, kBytes, st>>>(a);
    count_launch();
}

bool launch_ppar_compress_batch(const CompBatchConf &cf, const uint8_t *d_dict, const BatchArgs &b, cudaStream_t st,
                                bool allow_laps) {
    if (cf.window > 10) return false;
    if ((cf.flags & TB_F_EXTENDED) && (cf.flags & TB_F_LAZY)) return false;  // that combination stays with the general kernel
    if (b.in_offsets) return false;                          // strided layout only
    const bool laps = b.in_stride > (1u << cf.window);       // streams longer than the window: the lap variant (v1 only)
    if (laps && (!allow_laps || (cf.flags & TB_F_EXTENDED) || b.in_stride > (1u << 30))) return false;
    if ((b.in_stride & 15) || ((uintptr_t)b.in & 15) || (b.out_stride & 3) || ((uintptr_t)b.out & 3)) return false;
    if ((uintptr_t)d_dict & 3) return false;
    const uint64_t bound = 2 + (b.in_stride * (uint64_t)(cf.literal + 1) + 7) / 8 + 6;
    if (b.out_stride < ((bound + 3) & ~3ull)) return false;  // never OUTPUT_FULL in this kernel
    if (b.n_streams == 0) return true;

    PparArgs a;
    a.b = b;
    a.dict = d_dict;
    a.window_bits = cf.window;
    a.literal = cf.literal;
    a.flags = cf.flags;
    a.write_token = cf.write_token;
    if (cf.flags & TB_F_LAZY) {
        // the bitmap kernel has no lazy matching: nothing could pick deferred streams up, so none are deferred
        a.max_pairs = 0x7fffffff;
        laps ? launch_variant<kModeLazyLaps>(a, st) : launch_variant<kModeLazy>(a, st);
        return true;
    }
    a.max_pairs = kMaxPairs;
    if (laps)
        launch_variant<kModeLaps>(a, st);
    else if (cf.flags & TB_F_EXTENDED)
        launch_variant<kModeExt>(a, st);
    else
        launch_variant<kModeV1>(a, st);
    // second pass: the bitmap kernel picks up the streams marked kDeferred (usually none; it then only scans the sizes)
    static unsigned int *h_seen = nullptr;  // pinned mirror of d_deferred_total
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
    if (h_seen) cudaMemcpyFromSymbolAsync(h_seen, d_deferred_total, sizeof *h_seen, 0, cudaMemcpyDeviceToHost, st);
    return ok;
}
#endif  // TB_EMU

}  // namespace tb
