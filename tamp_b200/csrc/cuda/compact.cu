// Output compaction: the batch kernels write every stream into a fixed-stride row so that
// streams stay independent; storage and transport want contiguous frames.  This packs the rows of a finished batch
// into one buffer and produces the frame offsets (an exclusive prefix sum of the sizes), which is also the layout the
// decompressor accepts through TampB200Batch::in_offsets / in_sizes.
//
// Four small kernels: per-block sums of the sizes (1024 streams per block), a scan of the block sums, the offsets per
// block, and the copies on a grid of their own (a warp or a CTA per row by row length; aligned 32-bit stores, source
// words through a funnel shift, byte-exact ends).
#include "tb_cuda.h"

namespace tb {

namespace {

constexpr int kPerBlock = 1024, kThreads = 256;

__device__ __forceinline__ uint64_t warp_incl_scan(uint64_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

__global__ void __launch_bounds__(kThreads) k_block_sums(const uint32_t *sizes, uint64_t n, uint64_t *block_sums) {
    __shared__ uint64_t warp_sums[kThreads / 32];
    const uint64_t base = (uint64_t)blockIdx.x * kPerBlock;
    uint64_t s = 0;
    for (int i = threadIdx.x; i < kPerBlock; i += kThreads)
        if (base + i < n) s += sizes[base + i];
#pragma unroll
    for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t t = 0;
        for (int w = 0; w < kThreads / 32; w++) t += warp_sums[w];
        block_sums[blockIdx.x] = t;
    }
}

// Exclusive scan of the block sums in place (one block; the sums of a 2^20-stream batch are 1024 values).
__global__ void __launch_bounds__(1024) k_scan_block_sums(uint64_t *block_sums, uint64_t n_blocks, uint64_t *total) {
    __shared__ uint64_t warp_tot[32];
    __shared__ uint64_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint64_t base = 0; base < n_blocks; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const uint64_t v = i < n_blocks ? block_sums[i] : 0;
        const uint64_t incl = warp_incl_scan(v, lane);
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const uint64_t w = warp_tot[lane];
            const uint64_t wi = warp_incl_scan(w, lane);
            warp_tot[lane] = wi - w;  // exclusive offset of each warp
        }
        __syncthreads();
        const uint64_t carry = carry_s;
        if (i < n_blocks) block_sums[i] = carry + warp_tot[warp] + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_tot[warp] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
}

__global__ void __launch_bounds__(kThreads) k_pack_rows(const uint8_t *rows, uint64_t stride, const uint32_t *sizes,
                                                       uint64_t n, const uint64_t *block_offsets, uint8_t *packed,
                                                       uint64_t /*capacity*/, uint64_t *offsets) {
    __shared__ uint64_t local_off[kPerBlock];
    __shared__ uint64_t warp_tot[kThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t base = (uint64_t)blockIdx.x * kPerBlock;
    // exclusive scan of this block's 1024 sizes: 4 rounds of 256
    uint64_t carry = block_offsets[blockIdx.x];
    for (int r = 0; r < kPerBlock / kThreads; r++) {
        const uint64_t i = base + r * kThreads + threadIdx.x;
        const uint64_t v = i < n ? sizes[i] : 0;
        const uint64_t incl = warp_incl_scan(v, lane);
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        uint64_t before = 0, round_total = 0;
        for (int w = 0; w < kThreads / 32; w++) {
            if (w < warp) before += warp_tot[w];
            round_total += warp_tot[w];
        }
        local_off[r * kThreads + threadIdx.x] = carry + before + incl - v;
        carry += round_total;
        __syncthreads();
    }
    for (int i = threadIdx.x; i < kPerBlock; i += kThreads)
        if (base + i < n) offsets[base + i] = local_off[i];
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) offsets[n] = carry;
}

// Copy `sz` bytes (any alignment of either side) with a group of `nthreads` threads, thread `tid`: destination words are
// aligned 32-bit stores, source words come from aligned loads through a funnel shift; the bytes at both ends go singly.
__device__ __forceinline__ void copy_row(uint8_t *dst, const uint8_t *src, uint32_t sz, uint32_t tid, uint32_t nthreads) {
    const uint32_t head = (uint32_t)((4u - (uint32_t)((uintptr_t)dst & 3u)) & 3u) < sz ? (uint32_t)((4u - (uint32_t)((uintptr_t)dst & 3u)) & 3u) : sz;
    if (tid < head) dst[tid] = src[tid];
    dst += head;
    src += head;
    sz -= head;
    const uint32_t sh = (uint32_t)((uintptr_t)src & 3u) * 8u;
    uint32_t words = sz >> 2;
    if (sh && words) words -= 1;  // (the funnel reads the aligned word behind the one it completes: never past the row's last byte)
    uint32_t *d32 = reinterpret_cast<uint32_t *>(dst);
    if (sh == 0) {
        const uint32_t *s32 = reinterpret_cast<const uint32_t *>(src);
        for (uint32_t k = tid; k < words; k += nthreads) d32[k] = s32[k];
    } else {
        const uint32_t *s32 = reinterpret_cast<const uint32_t *>(src - (sh >> 3));
        for (uint32_t k = tid; k < words; k += nthreads) d32[k] = __funnelshift_r(s32[k], s32[k + 1], sh);
    }
    for (uint32_t k = (words << 2) + tid; k < sz; k += nthreads) dst[k] = src[k];
}

// The copies, separate from the offsets so that their grid follows the BYTES to move, not the number of rows: rows of up
// to kWarpRow bytes go one per warp, longer ones (64 KiB streams: ~22 KiB frames) one per CTA.
constexpr uint32_t kWarpRow = 2048;

__global__ void __launch_bounds__(kThreads) k_copy_rows(const uint8_t *rows, uint64_t stride, const uint32_t *sizes, uint64_t n,
                                                       const uint64_t *offsets, uint8_t *packed, uint64_t capacity, int cta_per_row) {
    if (cta_per_row) {
        for (uint64_t s = blockIdx.x; s < n; s += gridDim.x) {
            const uint32_t sz = sizes[s];
            const uint64_t off = offsets[s];
            if (off + sz > capacity) continue;  // does not fit: the caller sees offsets[n] > capacity
            copy_row(packed + off, rows + s * stride, sz, threadIdx.x, kThreads);
        }
    } else {
        const int lane = threadIdx.x & 31;
        const uint64_t nwarps = (uint64_t)gridDim.x * (kThreads / 32);
        for (uint64_t s = (uint64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); s < n; s += nwarps) {
            const uint32_t sz = sizes[s];
            const uint64_t off = offsets[s];
            if (off + sz > capacity) continue;
            copy_row(packed + off, rows + s * stride, sz, (uint32_t)lane, 32u);
        }
    }
}

}  // namespace

#ifndef TB_EMU
__global__ void k_offsets_to_sizes(const uint64_t *offsets, uint64_t n, uint32_t *sizes) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sizes[i] = (uint32_t)(offsets[i + 1] - offsets[i]);
}

void launch_offsets_to_sizes(const uint64_t *offsets, uint64_t n, uint32_t *sizes, cudaStream_t st) {
    if (!n) return;
    k_offsets_to_sizes<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(offsets, n, sizes);
    count_launch();
}

bool launch_compact(const uint8_t *rows, uint64_t stride, const uint32_t *sizes, uint64_t n, uint8_t *packed,
                    uint64_t capacity, uint64_t *offsets, cudaStream_t st) {
    const uint64_t n_blocks = (n + kPerBlock - 1) / kPerBlock;
    if (n_blocks > 0x7fffffffull) return false;
    if (n == 0) {
        cudaMemsetAsync(offsets, 0, sizeof(uint64_t), st);
        return true;
    }
    // block sums: stream-ordered scratch, private to this call (calls on different streams never share it)
    uint64_t *block_sums = nullptr;
    if (cudaMallocAsync(&block_sums, (n_blocks + 1) * sizeof(uint64_t), st) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    k_block_sums<<<(unsigned)n_blocks, kThreads, 0, st>>>(sizes, n, block_sums);
    k_scan_block_sums<<<1, 1024, 0, st>>>(block_sums, n_blocks, block_sums + n_blocks);
    k_pack_rows<<<(unsigned)n_blocks, kThreads, 0, st>>>(rows, stride, sizes, n, block_sums, packed, capacity, offsets);
    {
        static int sms = 0;
        if (!sms) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        }
        const int cta_per_row = stride > kWarpRow ? 1 : 0;
        const uint64_t want = cta_per_row ? n : (n + kThreads / 32 - 1) / (kThreads / 32);
        const uint64_t cap_grid = (uint64_t)sms * 8;
        k_copy_rows<<<(unsigned)(want < cap_grid ? want : cap_grid), kThreads, 0, st>>>(rows, stride, sizes, n, offsets, packed, capacity,
                                                                                        cta_per_row);
    }
    cudaFreeAsync(block_sums, st);
    count_launch();
    count_launch();
    count_launch();
    count_launch();
    return true;
}

#endif  // TB_EMU

}  // namespace tb
