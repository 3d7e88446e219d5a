// TEST INFRASTRUCTURE ONLY — never part of the product, never linked into libtamp_b200*.so.
//
// A small SIMT emulator that lets the kernel sources under tamp_b200/csrc/cuda/ be compiled with g++ and stepped
// on the CPU, so that `-m "not gpu"` tests can check the kernels' LOGIC (shared-memory region aliasing, queue
// indices, token lists, bit packing) against the oracle without a GPU.  It proves nothing about speed and is not a
// fallback: the library refuses to work without CUDA (tests/test_abi.py).
//
// Model: one CTA at a time; every CUDA thread is a fiber (ucontext) with its own stack.  A fiber runs until it
// reaches a warp collective (__shfl_sync, __ballot_sync, __match_any_sync, __reduce_*_sync, __syncwarp) or
// __syncthreads, deposits its operand and yields; the last participant to arrive publishes all operands and
// everybody continues.  Lanes of a warp therefore do NOT run in lock step between collectives — the scheduler
// visits runnable fibers in an order shuffled by a seed — so code that relies on implicit warp-synchronous
// execution (a missing __syncwarp between a shared-memory write and another lane's read) misbehaves here for some
// seeds, which is what a test wants.  A cycle in which no fiber makes progress is reported as a deadlock
// (divergent collectives).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <functional>
#include <vector>

namespace emu {

constexpr size_t kSmemBytes = 232448;  // 227 KiB
constexpr size_t kStackBytes = 96 * 1024;

struct Exchange {  // one in-flight collective of one (warp, mask)
    uint32_t mask = 0, arrived = 0, gen = 0;
    uint64_t vals[32];
    uint64_t snap[2][32];
};

struct Warp {
    std::vector<Exchange> slots;
};

struct Fiber {
    ucontext_t ctx;
    uint8_t *stack = nullptr;
    uint3 tid;
    int warp = 0, lane = 0;
    bool done = false;
    const char *where = "";  // the collective / barrier the fiber last entered (deadlock report)
    uint32_t where_mask = 0;
    unsigned or_calls = 0;
};

struct Cta {
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    uint32_t bar_arrived = 0, bar_gen = 0, alive = 0;
    uint64_t progress = 0;
    int or_acc[2] = {0, 0};
};

inline ucontext_t g_sched;
inline Cta *g_cta = nullptr;
inline Fiber *g_cur = nullptr;
inline uint3 g_block_idx, g_block_dim, g_grid_dim;
alignas(128) inline uint8_t g_smem[kSmemBytes];
inline uint32_t g_seg_header = 0;  // BatchArgs::seg_header of the next decompressor launches (emu_set_seg_header)
inline std::function<void()> g_body;
inline uint64_t g_rng = 1;

inline uint32_t rnd() {
    g_rng ^= g_rng << 13;
    g_rng ^= g_rng >> 7;
    g_rng ^= g_rng << 17;
    return (uint32_t)(g_rng >> 11);
}

inline void yield() { swapcontext(&g_cur->ctx, &g_sched); }

inline void fiber_main() {
    g_body();
    g_cur->done = true;
    g_cta->alive--;
    g_cta->progress++;
    if (g_cta->alive && g_cta->bar_arrived == g_cta->alive) {  // exited threads are not waited for
        g_cta->bar_arrived = 0;
        g_cta->bar_gen++;
    }
    yield();
}

// Deposit `v`, wait for every lane of `mask`, return the published operands of all 32 lanes.
inline const uint64_t *exchange(uint32_t mask, uint64_t v, const char *what = "warp collective") {
    Fiber *f = g_cur;
    f->where = what;
    f->where_mask = mask;
    Warp &w = g_cta->warps[f->warp];
    if (!((mask >> f->lane) & 1u)) {
        fprintf(stderr, "emu: lane %d calls a collective with mask %08x that excludes it\n", f->lane, mask);
        abort();
    }
    size_t si = 0;
    for (; si < w.slots.size(); si++)
        if (w.slots[si].mask == mask) break;
    if (si == w.slots.size()) {
        w.slots.emplace_back();
        w.slots.back().mask = mask;
    }
    Exchange *e = &w.slots[si];
    e->vals[f->lane] = v;
    e->arrived |= 1u << f->lane;
    const uint32_t my_gen = e->gen;
    if (e->arrived == mask) {
        memcpy(e->snap[my_gen & 1], e->vals, sizeof e->vals);
        e->arrived = 0;
        e->gen++;
        g_cta->progress++;
    } else {
        while (true) {
            yield();
            e = &g_cta->warps[f->warp].slots[si];  // the vector may have grown
            if (e->gen != my_gen) break;
        }
    }
    return g_cta->warps[f->warp].slots[si].snap[my_gen & 1];
}

inline void syncthreads() {
    Cta *c = g_cta;
    g_cur->where = "__syncthreads";
    const uint32_t my_gen = c->bar_gen;
    c->bar_arrived++;
    if (c->bar_arrived == c->alive) {
        c->bar_arrived = 0;
        c->bar_gen++;
        c->progress++;
        return;
    }
    while (c->bar_gen == my_gen) yield();
}

// Run `body` as a grid of CTAs, one after the other.  `seed` shuffles the order in which runnable fibers are visited.
inline void launch(unsigned grid, unsigned block, uint64_t seed, std::function<void()> body) {
    g_body = std::move(body);
    g_rng = seed * 0x9E3779B97F4A7C15ull + 1;
    g_block_dim = uint3{block, 1, 1};
    g_grid_dim = uint3{grid, 1, 1};
    for (unsigned b = 0; b < grid; b++) {
        Cta cta;
        cta.fibers.resize(block);
        cta.warps.resize((block + 31) / 32);
        cta.alive = block;
        g_cta = &cta;
        g_block_idx = uint3{b, 0, 0};
        for (auto &w : cta.warps) w.slots.reserve(8);
        for (unsigned t = 0; t < block; t++) {
            Fiber &f = cta.fibers[t];
            f.tid = uint3{t, 0, 0};
            f.warp = (int)(t / 32);
            f.lane = (int)(t % 32);
            f.stack = (uint8_t *)malloc(kStackBytes);
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack;
            f.ctx.uc_stack.ss_size = kStackBytes;
            f.ctx.uc_link = &g_sched;
            makecontext(&f.ctx, fiber_main, 0);
        }
        std::vector<unsigned> order(block);
        for (unsigned t = 0; t < block; t++) order[t] = t;
        while (cta.alive) {
            if (seed)
                for (unsigned t = block - 1; t > 0; t--) std::swap(order[t], order[rnd() % (t + 1)]);
            const uint64_t before = cta.progress;
            for (unsigned t : order) {
                Fiber &f = cta.fibers[t];
                if (f.done) continue;
                g_cur = &f;
                swapcontext(&g_sched, &f.ctx);
            }
            if (cta.progress == before && cta.alive) {
                fprintf(stderr, "emu: deadlock in block %u (%u threads alive; divergent collective or barrier)\n", b, cta.alive);
                for (unsigned t = 0; t < block; t++)
                    if (!cta.fibers[t].done && (t % 32 == 0 || cta.fibers[t].where != cta.fibers[t - 1].where))
                        fprintf(stderr, "  thread %u..: waiting in %s (mask %08x)\n", t, cta.fibers[t].where, cta.fibers[t].where_mask);
                abort();
            }
        }
        for (auto &f : cta.fibers) free(f.stack);
        g_cta = nullptr;
    }
}

}  // namespace emu

// ---- the CUDA device vocabulary the kernels use ------------------------------------------------------------------
#define threadIdx (emu::g_cur->tid)
#define blockIdx (emu::g_block_idx)
#define blockDim (emu::g_block_dim)
#define gridDim (emu::g_grid_dim)

#undef __launch_bounds__
#define __launch_bounds__(...)
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
// static __shared__ variables: one instance per kernel, shared by every fiber (CTAs run one after the other).
// Dynamic shared memory (extern __shared__) is emu::g_smem; the kernels select it under TB_EMU.
#undef __shared__
#define __shared__ static

inline void __syncthreads() { emu::syncthreads(); }
inline int __syncthreads_or(int pred) {
    // Two barriers around a per-CTA accumulator; consecutive calls alternate between two accumulators, so a thread
    // that is already in the next call cannot have its vote wiped by a late clear of this one.
    int *acc = emu::g_cta->or_acc + (emu::g_cur->or_calls++ & 1u);
    if (pred) *acc = 1;
    emu::syncthreads();
    const int r = *acc;
    emu::syncthreads();
    *acc = 0;
    return r;
}
inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::exchange(mask, 0, "__syncwarp"); }

template <typename T>
inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    static_assert(sizeof(T) <= 8, "shuffle operand");
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof v);
    const uint64_t *all = emu::exchange(mask, raw, "__shfl_sync");
    const int lane = emu::g_cur->lane;
    const int from = (lane & ~(width - 1)) | (src & (width - 1));
    T r;
    memcpy(&r, &all[from], sizeof r);
    return r;
}
template <typename T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof v);
    const uint64_t *all = emu::exchange(mask, raw, "__shfl_up_sync");
    const int lane = emu::g_cur->lane;
    const int from = lane - (int)delta;
    T r = v;
    if (from >= (lane & ~(width - 1))) memcpy(&r, &all[from], sizeof r);
    return r;
}
template <typename T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof v);
    const uint64_t *all = emu::exchange(mask, raw, "__shfl_down_sync");
    const int lane = emu::g_cur->lane;
    const int from = lane + (int)delta;
    T r = v;
    if (from <= (lane | (width - 1))) memcpy(&r, &all[from], sizeof r);
    return r;
}
template <typename T>
inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof v);
    const uint64_t *all = emu::exchange(mask, raw, "__shfl_xor_sync");
    const int from = emu::g_cur->lane ^ lanemask;
    (void)width;
    T r;
    memcpy(&r, &all[from], sizeof r);
    return r;
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
    const uint64_t *all = emu::exchange(mask, pred ? 1 : 0, "__ballot_sync");
    unsigned r = 0;
    for (int i = 0; i < 32; i++)
        if (((mask >> i) & 1u) && all[i]) r |= 1u << i;
    return r;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
template <typename T>
inline unsigned __match_any_sync(unsigned mask, T v) {
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof v);
    const uint64_t *all = emu::exchange(mask, raw, "__match_any_sync");
    unsigned r = 0;
    for (int i = 0; i < 32; i++)
        if (((mask >> i) & 1u) && all[i] == raw) r |= 1u << i;
    return r;
}
#define EMU_REDUCE(name, type, init, op)                          \
    inline type name(unsigned mask, type v) {                     \
        const uint64_t *all = emu::exchange(mask, (uint64_t)(int64_t)v, "__reduce_*_sync"); \
        type r = init;                                            \
        for (int i = 0; i < 32; i++)                              \
            if ((mask >> i) & 1u) {                               \
                const type x = (type)all[i];                      \
                r = op;                                           \
            }                                                     \
        return r;                                                 \
    }
EMU_REDUCE(__reduce_add_sync, unsigned, 0u, r + x)
EMU_REDUCE(__reduce_max_sync, unsigned, 0u, (x > r ? x : r))
EMU_REDUCE(__reduce_min_sync, unsigned, 0xffffffffu, (x < r ? x : r))
EMU_REDUCE(__reduce_or_sync, unsigned, 0u, r | x)
EMU_REDUCE(__reduce_and_sync, unsigned, 0xffffffffu, r &x)
inline int __reduce_add_sync(unsigned mask, int v) { return (int)__reduce_add_sync(mask, (unsigned)v); }
inline int __reduce_max_sync(unsigned mask, int v) {
    const uint64_t *all = emu::exchange(mask, (uint64_t)(int64_t)v, "__reduce_*_sync");
    int r = INT32_MIN;
    for (int i = 0; i < 32; i++)
        if ((mask >> i) & 1u) r = std::max(r, (int)(int64_t)all[i]);
    return r;
}
inline int __reduce_min_sync(unsigned mask, int v) {
    const uint64_t *all = emu::exchange(mask, (uint64_t)(int64_t)v, "__reduce_*_sync");
    int r = INT32_MAX;
    for (int i = 0; i < 32; i++)
        if ((mask >> i) & 1u) r = std::min(r, (int)(int64_t)all[i]);
    return r;
}
inline unsigned __activemask() { return 0xffffffffu; }

inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __ffsll(long long x) { return __builtin_ffsll(x); }
inline unsigned __brev(unsigned x) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
    return (unsigned)(((((uint64_t)hi) << 32) | lo) >> (sh & 31u));
}
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) {
    return (unsigned)((((((uint64_t)hi) << 32) | lo) << (sh & 31u)) >> 32);
}
inline unsigned __funnelshift_rc(unsigned lo, unsigned hi, unsigned sh) {
    sh = sh > 32 ? 32 : sh;
    return sh == 32 ? hi : __funnelshift_r(lo, hi, sh);
}
inline unsigned __funnelshift_lc(unsigned lo, unsigned hi, unsigned sh) {
    sh = sh > 32 ? 32 : sh;
    return sh == 32 ? lo : __funnelshift_l(lo, hi, sh);
}
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
    const uint64_t src = ((uint64_t)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned sel = (s >> (4 * i)) & 0xFu;
        unsigned b = (unsigned)(src >> (8 * (sel & 7u))) & 0xFFu;
        if (sel & 8u) b = (b & 0x80u) ? 0xFFu : 0u;
        r |= b << (8 * i);
    }
    return r;
}
inline unsigned __vcmpeq4(unsigned a, unsigned b) {
    unsigned r = 0;
    for (int i = 0; i < 4; i++)
        if (((a >> (8 * i)) & 0xFFu) == ((b >> (8 * i)) & 0xFFu)) r |= 0xFFu << (8 * i);
    return r;
}
inline unsigned __vcmpne4(unsigned a, unsigned b) { return ~__vcmpeq4(a, b); }
template <typename T>
inline T __ldg(const T *p) { return *p; }
template <typename T>
inline T __ldcs(const T *p) { return *p; }
template <typename T>
inline void __stcs(T *p, T v) { *p = v; }

template <typename T>
inline T atomicAdd(T *p, T v) { const T o = *p; *p = o + v; return o; }
template <typename T>
inline T atomicOr(T *p, T v) { const T o = *p; *p = o | v; return o; }
template <typename T>
inline T atomicAnd(T *p, T v) { const T o = *p; *p = o & v; return o; }
template <typename T>
inline T atomicMax(T *p, T v) { const T o = *p; *p = o > v ? o : v; return o; }
template <typename T>
inline T atomicMin(T *p, T v) { const T o = *p; *p = o < v ? o : v; return o; }
template <typename T>
inline T atomicExch(T *p, T v) { const T o = *p; *p = v; return o; }
template <typename T>
inline T atomicCAS(T *p, T cmp, T v) { const T o = *p; if (o == cmp) *p = v; return o; }

// Shared-state-space addresses are offsets into the emulated CTA's shared memory.
inline size_t __cvta_generic_to_shared(const void *p) { return (size_t)((const uint8_t *)p - emu::g_smem); }
inline void *emu_shared_ptr(uint32_t a) { return emu::g_smem + a; }

using std::max;
using std::min;
inline int min(int a, unsigned b) { return (unsigned)a < b ? a : (int)b; }
inline unsigned umin(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned umax(unsigned a, unsigned b) { return a > b ? a : b; }
