// Kernels built on the general-purpose codec state machines (generic_comp.cuh / generic_dec.cuh):
//   * k_comp_job / k_dec_job ............. the per-call C API on a batch of one (state-in / state-out)
//   * k_generic_compress_batch ........... one warp per stream, any configuration
//   * k_generic_decompress_batch ......... one thread per stream, windows in global scratch
//   * k_synth ............................ synthetic input generator (bench / tests)
#include "generic_comp.cuh"
#include "generic_dec.cuh"
#include "synth.cuh"
#include "tb_cuda.h"

namespace tb {

// ---------------------------------------------------------------------------------------------
// Per-call API kernels
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(32) k_comp_job(TbCompJob *job, uint8_t *window, const uint8_t *in, uint8_t *out) {
#ifndef TB_EMU
    extern __shared__ __align__(16) uint8_t smem[];
#else  // tests/emu: the kernels stepped on the CPU (test infrastructure; see tests/emu/cuda_emu.h)
    uint8_t *smem = emu::g_smem;
#endif
    CompCtx c;
    ctx_from_state(c, job->st);
    c.win = smem;
    c.ring = smem + c.W;
    c.out = out;
    c.out_room = (size_t)job->out_cap;
    const int l = lane_id();
    for (int i = l * 16; i < c.W; i += 32 * 16)
        *reinterpret_cast<uint4 *>(c.win + i) = *reinterpret_cast<const uint4 *>(window + i);
    if (l < 16) c.ring[l] = job->st.input[l];
    __syncwarp();

    size_t consumed = 0;
    int res = run_op(c, (int)job->op, in, (size_t)job->in_size, job->write_token != 0, consumed);

    __syncwarp();
    for (int i = l * 16; i < c.W; i += 32 * 16)
        *reinterpret_cast<uint4 *>(window + i) = *reinterpret_cast<const uint4 *>(c.win + i);
    if (l < 16) job->st.input[l] = c.ring[l];
    if (l == 0) {
        ctx_to_state(c, job->st);
        job->res = res;
        job->out_written = c.written;
        job->in_consumed = consumed;
    }
}

__global__ void __launch_bounds__(32) k_dec_job(TbDecJob *job, uint8_t *window, const uint8_t *in, uint8_t *out) {
    if (threadIdx.x != 0) return;
    TbDecState s = job->st;
    DecIo io{in, (size_t)job->in_size, 0, out, (size_t)job->out_cap, 0};
    int res = dec_run(s, window, io);
    job->st = s;
    job->res = res;
    job->out_written = io.out_pos;
    job->in_consumed = io.in_pos;
}

#ifndef TB_EMU
void launch_comp_job(TbCompJob *d_job, uint8_t *d_window, const uint8_t *d_in, uint8_t *d_out, int window_bits,
                     cudaStream_t st) {
    size_t smem = ((size_t)1 << window_bits) + 16;
    k_comp_job<<<1, 32, smem, st>>>(d_job, d_window, d_in, d_out);
    count_launch();
}
#endif

#ifndef TB_EMU
void launch_dec_job(TbDecJob *d_job, uint8_t *d_window, const uint8_t *d_in, uint8_t *d_out, cudaStream_t st) {
    k_dec_job<<<1, 32, 0, st>>>(d_job, d_window, d_in, d_out);
    count_launch();
}
#endif

// ---------------------------------------------------------------------------------------------
// Generic batch compress: one warp per stream, WPC warps per CTA, window + ring in shared memory.
// ---------------------------------------------------------------------------------------------

__global__ void k_generic_compress_batch(CompBatchConf cf, const uint8_t *dict, BatchArgs b) {
#ifndef TB_EMU
    extern __shared__ __align__(16) uint8_t smem[];
#else  // tests/emu: the kernels stepped on the CPU (test infrastructure; see tests/emu/cuda_emu.h)
    uint8_t *smem = emu::g_smem;
#endif
    const int W = 1 << cf.window;
    const int warp = threadIdx.x >> 5, l = lane_id();
    const uint64_t stream = (uint64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (stream >= b.n_streams) return;
    uint8_t *base = smem + (size_t)warp * (W + 16);

    CompCtx c;
    ctx_init(c, cf.window, cf.literal, stream_appends(cf.flags, stream) ? cf.flags : cf.flags & ~TB_F_APPEND);
    c.win = base;
    c.ring = base + W;
    c.out = b.out + stream * b.out_stride;
    c.out_room = (size_t)b.out_stride;
    for (int i = l * 16; i < W; i += 32 * 16)
        *reinterpret_cast<uint4 *>(c.win + i) = __ldg(reinterpret_cast<const uint4 *>(dict + i));
    __syncwarp();

    const uint8_t *in = b.in + (b.in_offsets ? b.in_offsets[stream] : stream * b.in_stride);
    const size_t n = b.in_sizes ? (size_t)b.in_sizes[stream] : (size_t)b.in_stride;
    size_t consumed = 0;
    int res = run_op(c, TB_OP_COMPRESS_AND_FLUSH, in, n, cf.write_token != 0, consumed);
    if (l == 0) {
        b.out_sizes[stream] = (uint32_t)c.written;
        if (b.status) b.status[stream] = (int8_t)res;
    }
}

#ifndef TB_EMU
void launch_generic_compress_batch(const CompBatchConf &cf, const uint8_t *d_dict, const BatchArgs &b,
                                   cudaStream_t st) {
    if (b.n_streams == 0) return;
    const size_t per_warp = ((size_t)1 << cf.window) + 16;
    int wpc = (int)((48 * 1024) / per_warp);
    wpc = wpc < 1 ? 1 : (wpc > 4 ? 4 : wpc);
    const uint64_t blocks = (b.n_streams + wpc - 1) / wpc;
    k_generic_compress_batch<<<(unsigned)blocks, wpc * 32, per_warp * wpc, st>>>(cf, d_dict, b);
    count_launch();
}
#endif

// ---------------------------------------------------------------------------------------------
// Generic batch decompress: one thread per stream (grid-stride over streams); each thread owns a
// window slot in global scratch.  Header parsing restates tamp_decompressor_read_header
// (decompressor.c:276-297) and populate_from_conf (:304-329).
// ---------------------------------------------------------------------------------------------

__global__ void k_generic_decompress_batch(const uint8_t *seed, const uint8_t *custom, int window_bits_max,
                                           uint8_t *scratch, BatchArgs b) {
    const uint64_t slot = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t nslots = (uint64_t)gridDim.x * blockDim.x;
    uint8_t *win = scratch + (slot << window_bits_max);
    for (uint64_t stream = slot; stream < b.n_streams; stream += nslots) {
        const uint8_t *in = b.in + (b.in_offsets ? b.in_offsets[stream] : stream * b.in_stride);
        const size_t n = b.in_sizes ? (size_t)b.in_sizes[stream] : (size_t)b.in_stride;
        uint8_t *out = b.out + stream * b.out_stride;
        int res;
        size_t written = 0;
        do {
            if (n == 0) {
                res = kInputExhausted;
                break;
            }
            const uint32_t hs = frame_start(b.seg_header, stream, in, (uint32_t)n), h = hs & 0xFFu;
            const size_t hdr = 1 + (h & 1u);
            if (n < hdr) {
                res = kInputExhausted;
                break;
            }
            if (hdr == 2 && (hs >> 8) != 0) {
                res = kInvalidConf;
                break;
            }
            TbDecState s;
            s.bit_buffer = 0;
            s.window_pos = 0;
            s.bit_buffer_pos = 0;
            s.token_state = 0;
            s.pending_window_offset = 0;
            s.pending_match_size = 0;
            s.window_bits = (uint8_t)(((h >> 5) & 7u) + 8u);
            s.literal_bits = (uint8_t)(((h >> 3) & 3u) + 5u);
            s.flags = (uint8_t)(((h & 2u) ? TB_F_EXTENDED : 0) | ((h & 1u) ? TB_F_DICT_RESET : 0));
            s.min_pattern_size = (uint8_t)min_pattern_size(s.window_bits, s.literal_bits);
            s.skip_bytes = 0;
            s.window_bits_max = (uint8_t)window_bits_max;
            s.configured = 1;
            s.header_bytes_read = 0;
            s.last_was_flush = 0;
            const bool use_custom = (h >> 2) & 1u;
            if (s.window_bits > window_bits_max || (use_custom && custom == nullptr)) {
                res = kInvalidConf;
                break;
            }
            const int W = 1 << s.window_bits;
            const int seed_literal = (s.flags & TB_F_EXTENDED) ? s.literal_bits : 8;
            const uint8_t *src = use_custom ? custom : seed + (seed_literal <= 5 ? 0 : seed_literal <= 6 ? 1 : 2) * 32768;
            for (int i = 0; i < W; i += 16)
                *reinterpret_cast<uint4 *>(win + i) = __ldg(reinterpret_cast<const uint4 *>(src + i));
            DecIo io{in + hdr, n - hdr, 0, out, (size_t)b.out_stride, 0};
            res = dec_run(s, win, io);
            written = io.out_pos;
        } while (0);
        b.out_sizes[stream] = (uint32_t)written;
        if (b.status) b.status[stream] = (int8_t)res;
    }
}

#ifndef TB_EMU
uint64_t generic_decompress_slots(uint64_t n_streams, int window_bits_max) {
    // Bound scratch to 1 GiB: slots * (1 << window_bits_max) bytes.
    uint64_t cap = ((uint64_t)1 << 30) >> window_bits_max;
    uint64_t want = (n_streams + 127) / 128 * 128;
    uint64_t slots = want < cap ? want : cap;
    return slots < 128 ? 128 : slots;
}
#endif

#ifndef TB_EMU
void launch_generic_decompress_batch(const uint8_t *d_seed, const uint8_t *d_custom, int window_bits_max,
                                     uint8_t *d_scratch, uint64_t n_slots, const BatchArgs &b, cudaStream_t st) {
    if (b.n_streams == 0) return;
    k_generic_decompress_batch<<<(unsigned)(n_slots / 128), 128, 0, st>>>(d_seed, d_custom, window_bits_max, d_scratch,
                                                                         b);
    count_launch();
}
#endif

// ---------------------------------------------------------------------------------------------
// Synthetic input generator
// ---------------------------------------------------------------------------------------------

static __device__ SynthVocab g_vocab;

__global__ void k_synth(int kind, uint64_t first_k, uint64_t n_streams, uint64_t stream_len, uint8_t *out) {
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    synth_fill(kind, first_k + s, out + s * stream_len, stream_len, &g_vocab);
}

#ifndef TB_EMU
void launch_synth(int kind, uint64_t first_k, uint64_t n_streams, uint64_t stream_len, uint8_t *d_out,
                  cudaStream_t st) {
    static bool vocab_ready = false;
    if (!vocab_ready) {
        SynthVocab v;
        synth_build_vocab(v);
        cudaMemcpyToSymbol(g_vocab, &v, sizeof v);
        vocab_ready = true;
    }
    if (n_streams == 0) return;
    k_synth<<<(unsigned)((n_streams + 127) / 128), 128, 0, st>>>(kind, first_k, n_streams, stream_len, d_out);
    count_launch();
}
#endif

}  // namespace tb
