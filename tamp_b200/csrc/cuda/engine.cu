// tamp-b200 engine: the thin extern "C" shim between the host C API layer and the CUDA kernels.
//
// Owns everything the API itself may not own (the Tamp C API has init but no destructor and never
// allocates per object): one engine per process with a mutex, a CUDA stream, device
// staging buffers that grow on demand, and the seeded-dictionary tables.  Callers may release the GIL
// / use several host threads: entry points serialise on the engine mutex.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <mutex>
#include <vector>

#include "tamp/compressor.h"
#include "tamp_b200.h"
#include "tb_cuda.h"

namespace tb {

static std::mutex g_mu;         // engine state, every kernel launch sequence, the per-call API
static std::mutex g_dir_mu[2];  // one pipelined host-pointer batch per direction (0 = compress, 1 = decompress) at a time
static std::atomic<uint64_t> g_launches{0};
static char g_err[512] = "";
static int g_kernel_mode = 0;
static std::atomic<uint64_t> g_h2d{0}, g_d2h{0};  // bytes moved by the host-pointer batch entry points

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

struct DevBuf {
    uint8_t *p = nullptr;
    size_t cap = 0;
    bool ensure(size_t n) {
        if (n <= cap) return true;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n < 4096 ? 4096 : n + n / 4;
        if (cudaMalloc(&p, want) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        cap = want;
        return true;
    }
};

// Stream-ordered scratch of one asynchronous call: allocated and freed on the caller's stream (cudaMallocAsync /
// cudaFreeAsync), so that calls enqueued on different streams never share a buffer.
struct StreamTemp {
    uint8_t *p = nullptr;
    cudaStream_t st = nullptr;
    bool alloc(size_t n, cudaStream_t s) {
        st = s;
        if (cudaMallocAsync(&p, n < 256 ? 256 : n, s) != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            return false;
        }
        return true;
    }
    ~StreamTemp() {
        if (p) cudaFreeAsync(p, st);
    }
};

constexpr int kSlots = 8;  // most chunks in flight per direction of the pipelined host-pointer path (in use: kDefaultSlots)
constexpr int kDefaultSlots = 3;
constexpr uint64_t kDefaultChunkBytes = (uint64_t)48 << 20;

struct Engine {
    bool ready = false;
    int device = -1;
    cudaStream_t stream = nullptr;
    DevBuf job, window, in, out;           // per-call API staging
    DevBuf b_in, b_out, b_meta;            // host-batch API staging (single-shot path)
    struct Slot {                          // host-batch API staging (pipelined path): one chunk in flight each
        DevBuf in, out, meta;
        DevBuf packed, offs;               // contiguous frames of the chunk + their offsets (packed output / packed input)
        uint8_t *h_meta = nullptr;         // pinned mirror of meta: small arrays never stall the pipeline
        size_t h_meta_cap = 0;
        uint64_t *h_offs = nullptr;        // pinned: a chunk's frame offsets relative to its first frame (packed input)
        size_t h_offs_cap = 0;
        cudaStream_t st = nullptr;
        cudaEvent_t ev = nullptr;
        bool busy = false;
        uint64_t first = 0, count = 0;
    } slot[2][kSlots];                          // [direction]: a compress call and a decompress call may be in flight together
    DevBuf custom_dict[2];                 // aligned copy of a caller-supplied dictionary, per direction (host-pointer path)
    uint8_t *seed = nullptr;               // 3 x 32 KiB seeded dictionaries (literal classes 5, 6, 7/8)
};
static Engine g_eng;

static bool cuda_ok(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return true;
    tb_set_error("%s: %s", what, cudaGetErrorString(e));
    cudaGetLastError();
    return false;
}

static bool engine_init_locked() {
    Engine &E = g_eng;
    if (E.ready) {
        // the engine's tables live on the device that was current at first use: one process drives one GPU
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess || cur != E.device) {
            cudaGetLastError();
            tb_set_error("the engine is bound to CUDA device %d but device %d is current: use one process per GPU", E.device, cur);
            return false;
        }
        return true;
    }
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        tb_set_error("no CUDA device available: tamp-b200 has no CPU fallback");
        return false;
    }
    int dev = 0;
    if (!cuda_ok(cudaGetDevice(&dev), "cudaGetDevice")) return false;
    cudaDeviceProp prop;
    if (!cuda_ok(cudaGetDeviceProperties(&prop, dev), "cudaGetDeviceProperties")) return false;
    if (prop.major < 10) {
        tb_set_error("device %d is sm_%d%d; this build contains sm_100a code only", dev, prop.major, prop.minor);
        return false;
    }
    if (!cuda_ok(cudaStreamCreateWithFlags(&E.stream, cudaStreamNonBlocking), "cudaStreamCreate")) return false;
    if (!cuda_ok(cudaMalloc(&E.seed, 3 * 32768), "cudaMalloc(seed)")) return false;
    static unsigned char host_seed[3 * 32768];
    tamp_initialize_dictionary(host_seed, 32768, 5);  // common.c dictionary init is kept on the host
    tamp_initialize_dictionary(host_seed + 32768, 32768, 6);
    tamp_initialize_dictionary(host_seed + 65536, 32768, 8);
    if (!cuda_ok(cudaMemcpy(E.seed, host_seed, sizeof host_seed, cudaMemcpyHostToDevice), "seed upload")) return false;
    register_static_dictionaries(E.seed, sizeof host_seed);
    {   // per-call scratch is stream-ordered (cudaMallocAsync): keep freed blocks in the pool instead of returning them to
        // the driver at every synchronisation
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    E.device = dev;
    E.ready = true;
    return true;
}

static const uint8_t *seed_table(int literal_for_seed) {
    return g_eng.seed + (literal_for_seed <= 5 ? 0 : literal_for_seed <= 6 ? 1 : 2) * 32768;
}

}  // namespace tb

using namespace tb;

extern "C" {

void tb_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tb::g_err, sizeof tb::g_err, fmt, ap);
    va_end(ap);
}

const char *tamp_b200_last_error(void) { return tb::g_err; }
uint64_t tamp_b200_launch_count(void) { return tb::g_launches.load(); }
void tamp_b200_copy_bytes(uint64_t *h2d, uint64_t *d2h) {
    if (h2d) *h2d = tb::g_h2d.load();
    if (d2h) *d2h = tb::g_d2h.load();
}
const char *tamp_b200_version(void) { return "tamp-b200 0.1 (sm_100a)"; }
void tamp_b200_set_kernel_mode(int mode) {
    tb::g_kernel_mode = mode;
}

int tamp_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

tamp_res tamp_b200_set_device(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_eng.ready && g_eng.device != device) {
        tb_set_error("engine already bound to device %d (one process per GPU)", g_eng.device);
        return TAMP_ERROR;
    }
    if (!cuda_ok(cudaSetDevice(device), "cudaSetDevice")) return TAMP_ERROR;
    return engine_init_locked() ? TAMP_OK : TAMP_ERROR;
}

// ---- per-call API: batch of one, state-in / state-out -------------------------------------------

int tb_engine_run_comp_job(TbCompJob *job, unsigned char *window, const unsigned char *in, unsigned char *out) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!engine_init_locked()) return -1;
    Engine &E = g_eng;
    cudaSetDevice(E.device);
    const size_t W = (size_t)1 << job->st.window_bits;
    // No call can emit more than ~9 bits per pending input byte plus a few bytes of queued state.
    const size_t bound = (job->in_size + 64) * 9 / 8 + 64;
    const size_t out_cap = job->out_cap < bound ? (size_t)job->out_cap : bound;
    if (!E.job.ensure(sizeof(TbCompJob)) || !E.window.ensure(W) || !E.in.ensure(job->in_size + 16) ||
        !E.out.ensure(out_cap + 16)) {
        tb_set_error("device allocation failed");
        return -1;
    }
    TbCompJob staged = *job;
    staged.out_cap = out_cap;
    cudaStream_t st = E.stream;
    bool ok = cuda_ok(cudaMemcpyAsync(E.job.p, &staged, sizeof staged, cudaMemcpyHostToDevice, st), "H2D job") &&
              cuda_ok(cudaMemcpyAsync(E.window.p, window, W, cudaMemcpyHostToDevice, st), "H2D window");
    if (ok && job->in_size) ok = cuda_ok(cudaMemcpyAsync(E.in.p, in, job->in_size, cudaMemcpyHostToDevice, st), "H2D in");
    if (!ok) return -1;
    launch_comp_job(reinterpret_cast<TbCompJob *>(E.job.p), E.window.p, E.in.p, E.out.p, job->st.window_bits, st);
    if (!cuda_ok(cudaGetLastError(), "launch k_comp_job")) return -1;
    if (!cuda_ok(cudaMemcpyAsync(&staged, E.job.p, sizeof staged, cudaMemcpyDeviceToHost, st), "D2H job")) return -1;
    if (!cuda_ok(cudaMemcpyAsync(window, E.window.p, W, cudaMemcpyDeviceToHost, st), "D2H window")) return -1;
    if (!cuda_ok(cudaStreamSynchronize(st), "k_comp_job")) return -1;
    if (staged.out_written) {
        if (!cuda_ok(cudaMemcpy(out, E.out.p, staged.out_written, cudaMemcpyDeviceToHost), "D2H out")) return -1;
    }
    staged.out_cap = job->out_cap;
    *job = staged;
    return 0;
}

int tb_engine_run_dec_job(TbDecJob *job, unsigned char *window, const unsigned char *in, unsigned char *out) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!engine_init_locked()) return -1;
    Engine &E = g_eng;
    cudaSetDevice(E.device);
    const size_t W = (size_t)1 << job->st.window_bits;
    // A token yields at most 241 bytes (RLE) from >= 2 bits; cap the device buffer by what the input
    // can possibly expand to, plus whatever a resumed token still owes.
    const size_t bound = (job->in_size + 8) * 8 / 2 * 241 + 512;
    const size_t out_cap = job->out_cap < bound ? (size_t)job->out_cap : bound;
    if (!E.job.ensure(sizeof(TbDecJob)) || !E.window.ensure(W) || !E.in.ensure(job->in_size + 16) ||
        !E.out.ensure(out_cap + 16)) {
        tb_set_error("device allocation failed");
        return -1;
    }
    TbDecJob staged = *job;
    staged.out_cap = out_cap;
    cudaStream_t st = E.stream;
    bool ok = cuda_ok(cudaMemcpyAsync(E.job.p, &staged, sizeof staged, cudaMemcpyHostToDevice, st), "H2D job") &&
              cuda_ok(cudaMemcpyAsync(E.window.p, window, W, cudaMemcpyHostToDevice, st), "H2D window");
    if (ok && job->in_size) ok = cuda_ok(cudaMemcpyAsync(E.in.p, in, job->in_size, cudaMemcpyHostToDevice, st), "H2D in");
    if (!ok) return -1;
    launch_dec_job(reinterpret_cast<TbDecJob *>(E.job.p), E.window.p, E.in.p, E.out.p, st);
    if (!cuda_ok(cudaGetLastError(), "launch k_dec_job")) return -1;
    if (!cuda_ok(cudaMemcpyAsync(&staged, E.job.p, sizeof staged, cudaMemcpyDeviceToHost, st), "D2H job")) return -1;
    if (!cuda_ok(cudaMemcpyAsync(window, E.window.p, W, cudaMemcpyDeviceToHost, st), "D2H window")) return -1;
    if (!cuda_ok(cudaStreamSynchronize(st), "k_dec_job")) return -1;
    if (staged.out_written) {
        if (!cuda_ok(cudaMemcpy(out, E.out.p, staged.out_written, cudaMemcpyDeviceToHost), "D2H out")) return -1;
    }
    staged.out_cap = job->out_cap;
    *job = staged;
    return 0;
}

// ---- batch API -------------------------------------------------------------------------------------

size_t tamp_b200_compress_bound(const TampConf *conf, size_t n) {
    const size_t literal = conf ? conf->literal : 8;
    // header (<= 2) + all-literal payload + FLUSH token / padding slack (compressor.h:159-181 sizing notes)
    return 2 + (n * (literal + 1) + 7) / 8 + 6;
}

static bool conf_to_batch(const TampConf *conf, CompBatchConf &cf, bool write_token) {
    TampConf d;
    memset(&d, 0, sizeof d);
    d.window = 10;
    d.literal = 8;
    d.extended = 1;
    if (!conf) conf = &d;
    if (conf->window < 8 || conf->window > 15 || conf->literal < 5 || conf->literal > 8) return false;
    // append mode (compressor.c:209): every stream of the batch starts with a FLUSH instead of a header
    if (conf->append && (!conf->dictionary_reset || conf->use_custom_dictionary)) return false;
    cf.window = conf->window;
    cf.literal = conf->literal;
    cf.flags = (conf->extended ? TB_F_EXTENDED : 0) | (conf->dictionary_reset ? TB_F_DICT_RESET : 0) |
               (conf->use_custom_dictionary ? TB_F_CUSTOM_DICT : 0) | (conf->append ? TB_F_APPEND : 0);
#if TAMP_LAZY_MATCHING
    if (conf->lazy_matching) cf.flags |= TB_F_LAZY;
#endif
    cf.write_token = write_token ? 1 : 0;
    return true;
}

static BatchArgs to_args(const TampB200Batch *b) {
    BatchArgs a;
    a.in = b->in;
    a.in_offsets = b->in_offsets;
    a.in_sizes = b->in_sizes;
    a.in_stride = b->in_stride;
    a.out = b->out;
    a.out_stride = b->out_stride;
    a.out_sizes = b->out_sizes;
    a.status = b->status;
    a.n_streams = b->n_streams;
    return a;
}

static tamp_res compress_device_locked(const CompBatchConf &cf, const unsigned char *d_dictionary,
                                       const BatchArgs &a, cudaStream_t st, bool dict_staged = false) {
    Engine &E = g_eng;
    const uint8_t *dict;
    StreamTemp dict_copy;  // aligned copy of the caller's dictionary, private to this call (freed in stream order)
    if (cf.flags & TB_F_CUSTOM_DICT) {
        if (!d_dictionary) {
            tb_set_error("use_custom_dictionary set but no dictionary given");
            return TAMP_INVALID_CONF;
        }
        const size_t W = (size_t)1 << cf.window;
        if (!dict_staged) {
            if (!dict_copy.alloc(W, st)) return TAMP_ERROR;
            if (!cuda_ok(cudaMemcpyAsync(dict_copy.p, d_dictionary, W, cudaMemcpyDefault, st), "dictionary copy"))
                return TAMP_ERROR;
            dict = dict_copy.p;
        } else {
            dict = d_dictionary;  // the host-pointer path staged an aligned copy once for all of its chunks
        }
    } else {
        dict = seed_table((cf.flags & TB_F_EXTENDED) ? cf.literal : 8);
    }
    bool done = false;
    // kernel modes (test / benchmark hook): 0 = specialised kernels (the default dispatch: segment walk for v1 streams
    // no longer than a window <= 1 KiB, history walk for every other v1 batch, position-parallel kernel for the
    // extended format / lazy matching, then the bitmap kernels), 1 = general kernels only, 2 = bitmap kernels only,
    // 4 = the round-1 dispatch: no walk kernels, position-parallel compressor without its lap variant
    if (g_kernel_mode == 0) done = launch_walk_compress_batch(cf, dict, a, st);
    if (g_kernel_mode == 0 && !done) done = launch_cwalk_compress_batch(cf, dict, a, st);
    if ((g_kernel_mode == 0 || g_kernel_mode == 7) && !done) done = launch_hwalk_compress_batch(cf, dict, a, st);  // (7: test hook, history walk for every v1 batch)
    if (!done && (g_kernel_mode == 0 || g_kernel_mode == 4)) done = launch_ppar_compress_batch(cf, dict, a, st, g_kernel_mode == 0);
    if (g_kernel_mode != 1 && !done) done = launch_fast_compress_batch(cf, dict, a, st);
    if (g_kernel_mode != 1 && !done) done = launch_wide_compress_batch(cf, dict, a, st);
    if (!done) launch_generic_compress_batch(cf, dict, a, st);
    return cuda_ok(cudaGetLastError(), "compress batch launch") ? TAMP_OK : TAMP_ERROR;
}

static tamp_res decompress_device_locked(const unsigned char *d_dictionary, int window_bits_max, const BatchArgs &a,
                                         cudaStream_t st, bool dict_staged = false) {
    Engine &E = g_eng;
    const uint8_t *custom = nullptr;
    StreamTemp dict_copy, windows;  // private to this call, freed in stream order
    if (d_dictionary) {
        const size_t W = (size_t)1 << window_bits_max;  // the caller's dictionary holds 1 << window_bits_max bytes (tamp_b200.h)
        if (!dict_staged) {
            if (!dict_copy.alloc(W, st)) return TAMP_ERROR;
            if (!cuda_ok(cudaMemcpyAsync(dict_copy.p, d_dictionary, W, cudaMemcpyDefault, st), "dictionary copy"))
                return TAMP_ERROR;
            custom = dict_copy.p;
        } else {
            custom = d_dictionary;  // (staged by the host-pointer path)
        }
    }
    bool done = false;
    // (kernel mode 0: split parse / copy decompressor first, when no row can outgrow the window; 2 and 4: without it)
    if (g_kernel_mode == 0 && a.out_stride <= ((uint64_t)1 << (window_bits_max < 10 ? window_bits_max : 10)))
        done = launch_split_decompress_batch(E.seed, custom, window_bits_max, a, st);
    // (long split decompressor: windows 11..15 and enough streams to fill the GPU with one LANE per stream in its parse
    // phase; measured on B200: 172 GB/s against 23 for 2^16 streams of 64 KiB at window 15, but 19 against 23 for 2^12
    // of them, and 82 against 149 at window 10 / 4 KiB frames, where the windows fit shared memory: fast_decompress.cu)
    if (g_kernel_mode == 0 && !done && window_bits_max >= 11 && a.n_streams >= 8192)
        done = launch_lsplit_decompress_batch(E.seed, custom, window_bits_max, a, st);
    if (g_kernel_mode == 6 && !done) done = launch_lsplit_decompress_batch(E.seed, custom, window_bits_max, a, st);  // (test hook: always)
    if (g_kernel_mode != 1 && !done) done = launch_fast_decompress_batch(E.seed, custom, window_bits_max, a, st);
    if (g_kernel_mode != 1 && !done) done = launch_wide_decompress_batch(E.seed, custom, window_bits_max, a, st);
    if (!done) {
        const uint64_t slots = generic_decompress_slots(a.n_streams, window_bits_max);
        if (!windows.alloc((size_t)slots << window_bits_max, st)) {
            tb_set_error("scratch allocation failed");
            return TAMP_ERROR;
        }
        launch_generic_decompress_batch(E.seed, custom, window_bits_max, windows.p, slots, a, st);
    }
    return cuda_ok(cudaGetLastError(), "decompress batch launch") ? TAMP_OK : TAMP_ERROR;
}

tamp_res tamp_b200_compress_batch_device(const TampConf *conf, const unsigned char *dictionary,
                                         const TampB200Batch *batch, bool write_token, void *cuda_stream) {
    CompBatchConf cf;
    if (!batch || !conf_to_batch(conf, cf, write_token)) return TAMP_INVALID_CONF;
    std::lock_guard<std::mutex> lk(g_mu);
    if (!engine_init_locked()) return TAMP_ERROR;
    return compress_device_locked(cf, dictionary, to_args(batch), (cudaStream_t)cuda_stream);
}

tamp_res tamp_b200_decompress_batch_device(const unsigned char *dictionary, uint8_t window_bits_max,
                                           const TampB200Batch *batch, void *cuda_stream) {
    if (!batch || window_bits_max < 8 || window_bits_max > 15) return TAMP_INVALID_CONF;
    std::lock_guard<std::mutex> lk(g_mu);
    if (!engine_init_locked()) return TAMP_ERROR;
    return decompress_device_locked(dictionary, window_bits_max, to_args(batch), (cudaStream_t)cuda_stream);
}

// Pipelined host-pointer path (strided layouts): the batch is cut into chunks that flow through three
// slots, each with its own CUDA stream, so that chunk i's kernel overlaps chunk i+1's H2D copy and chunk
// i-1's D2H copy (both PCIe directions busy).  Only the bytes that carry data cross the bus: compressed
// rows travel as 2-D copies whose width is the longest row of the chunk, not the worst-case stride.
// Packed output of a host-pointer compress call: contiguous frames + offsets instead of fixed-stride rows.
struct PackedOut {
    unsigned char *packed;
    uint64_t capacity;
    uint64_t *offsets;   // n_streams + 1
    uint64_t total = 0;  // bytes of the chunks finished so far (chunks finish in order)
    bool overflow = false;
};

static bool pipe_finish_slot(Engine::Slot &S, bool compress, const TampB200Batch *b, PackedOut *po = nullptr) {
    if (!S.busy) return true;
    bool ok = cuda_ok(cudaEventSynchronize(S.ev), "chunk kernel");
    const uint64_t c = S.count;
    uint32_t *h_osz = b->out_sizes + S.first;
    memcpy(h_osz, S.h_meta + c * 4, c * 4);
    if (b->status) memcpy(b->status + S.first, S.h_meta + c * 8, c);
    unsigned char *h_out = b->out + S.first * b->out_stride;
    if (ok && compress && po) {
        // sizes arrived with the event: the chunk's frames are contiguous on the device (compacted behind the kernel)
        uint64_t run = 0;
        for (uint64_t i = 0; i < c; i++) {
            po->offsets[S.first + i] = po->total + run;
            run += h_osz[i];
        }
        if (po->total + run > po->capacity) {
            po->overflow = true;
            ok = false;
        } else if (run) {
            ok = cuda_ok(cudaMemcpyAsync(po->packed + po->total, S.packed.p, run, cudaMemcpyDeviceToHost, S.st), "D2H frames");
        }
        g_d2h += run;
        po->total += run;
    } else if (ok && compress) {
        // sizes arrived with the event; ship rows no wider than the longest one
        uint32_t mx = 0;
        for (uint64_t i = 0; i < c; i++) mx = h_osz[i] > mx ? h_osz[i] : mx;
        size_t width = ((size_t)mx + 63) & ~(size_t)63;
        if (width > b->out_stride) width = b->out_stride;
        if (width)
            ok = cuda_ok(cudaMemcpy2DAsync(h_out, b->out_stride, S.out.p, b->out_stride, width, c,
                                           cudaMemcpyDeviceToHost, S.st), "D2H rows");
        g_d2h += width * c;
        // no host wait here: whatever reuses this slot is enqueued on the same CUDA stream, behind this copy
    }
    S.busy = false;
    return ok;
}

static tamp_res host_batch_pipelined(bool compress, const CompBatchConf &cf, const unsigned char *dictionary,
                                     uint8_t wbits_max, const TampB200Batch *b, PackedOut *po = nullptr) {
    Engine &E = g_eng;
    const int dir = compress ? 0 : 1;
    Engine::Slot (&slots)[kSlots] = E.slot[dir];
    DevBuf &staged_dict = E.custom_dict[dir];
    const uint64_t n = b->n_streams;
    const bool packed_in = b->in_offsets != nullptr;  // (decompress) contiguous frames, ascending offsets: checked by the caller
    // bytes of input per stream: the stride, or the mean frame size of a packed input
    const uint64_t in_per_stream =
        packed_in ? (b->in_offsets[n - 1] + b->in_sizes[n - 1] - b->in_offsets[0]) / n + 1 : b->in_stride;
    // chunk size: ~48 MiB of input+output per slot, at least 1024 streams, at most n
    const uint64_t per_stream = in_per_stream + b->out_stride + 16;
    int nslots = kDefaultSlots;  // (tuning hooks, read per call: nothing is kept between calls or shared between the directions)
    uint64_t chunk_bytes = kDefaultChunkBytes;
    if (const char *e = getenv("TAMP_B200_SLOTS")) {
        const int v = atoi(e);
        if (v >= 2 && v <= kSlots) nslots = v;
    }
    if (const char *e = getenv("TAMP_B200_CHUNK_MIB")) {
        const int v = atoi(e);
        if (v >= 1 && v <= 512) chunk_bytes = (uint64_t)v << 20;
    }
    uint64_t chunk = chunk_bytes / per_stream;
    chunk = chunk < 1024 ? 1024 : chunk;
    chunk = (chunk + 255) & ~(uint64_t)255;
    if (chunk > n) chunk = n;
    for (int si = 0; si < nslots; si++) {
        Engine::Slot &S = slots[si];
        if (!S.st && !cuda_ok(cudaStreamCreateWithFlags(&S.st, cudaStreamNonBlocking), "slot stream")) return TAMP_ERROR;
        if (!S.ev && !cuda_ok(cudaEventCreateWithFlags(&S.ev, cudaEventDisableTiming), "slot event")) return TAMP_ERROR;
        if ((!packed_in && !S.in.ensure(chunk * b->in_stride + 16)) || !S.out.ensure(chunk * b->out_stride + 16) ||
            !S.meta.ensure(chunk * 9 + 64) || (po && (!S.packed.ensure(chunk * b->out_stride + 16) || !S.offs.ensure((chunk + 1) * 8))) ||
            (packed_in && !S.offs.ensure((chunk + 1) * 8))) {
            tb_set_error("device allocation failed (pipelined slots)");
            return TAMP_ERROR;
        }
        if (S.h_meta_cap < chunk * 9 + 64) {
            if (S.h_meta) cudaFreeHost(S.h_meta);
            S.h_meta = nullptr;
            S.h_meta_cap = 0;
            if (!cuda_ok(cudaMallocHost(&S.h_meta, chunk * 9 + 64), "pinned meta")) return TAMP_ERROR;
            S.h_meta_cap = chunk * 9 + 64;
        }
        if (packed_in && S.h_offs_cap < chunk * 8) {
            if (S.h_offs) cudaFreeHost(S.h_offs);
            S.h_offs = nullptr;
            S.h_offs_cap = 0;
            if (!cuda_ok(cudaMallocHost(&S.h_offs, chunk * 8), "pinned offsets")) return TAMP_ERROR;
            S.h_offs_cap = chunk * 8;
        }
        S.busy = false;
    }
    bool dict_staged = false;
    if (dictionary) {
        const size_t W = (size_t)1 << (compress ? cf.window : wbits_max);
        if (!staged_dict.ensure(W)) return TAMP_ERROR;
        if (!cuda_ok(cudaMemcpy(staged_dict.p, dictionary, W, cudaMemcpyDefault), "dictionary copy")) return TAMP_ERROR;
        dict_staged = true;
    }
    bool ok = true;
    uint64_t idx = 0;
    tamp_res failed = TAMP_OK;
    static const bool trace = getenv("TAMP_B200_TRACE") != nullptr;  // (debug aid: where a host-pointer call spends its time)
    double t_wait = 0, t_enq = 0, t_lock = 0;
    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    };
    for (uint64_t first = 0; first < n && ok; first += chunk, idx++) {
        Engine::Slot &S = slots[idx % nslots];
        const auto t0 = std::chrono::steady_clock::now();
        ok = pipe_finish_slot(S, compress, b, po);
        t_wait += since(t0);
        const auto t1 = std::chrono::steady_clock::now();
        if (!ok) break;
        const uint64_t c = n - first < chunk ? n - first : chunk;
        S.first = first;
        S.count = c;
        uint32_t *d_isz = reinterpret_cast<uint32_t *>(S.meta.p);
        uint32_t *d_osz = reinterpret_cast<uint32_t *>(S.meta.p + c * 4);
        int8_t *d_stat = reinterpret_cast<int8_t *>(S.meta.p + c * 8);
        uint64_t *d_offs = reinterpret_cast<uint64_t *>(S.offs.p);
        if (packed_in) {
            // the chunk's frames are one contiguous range of the caller's buffer; offsets become relative to its first frame
            const uint64_t base = b->in_offsets[first];
            const uint64_t extent = b->in_offsets[first + c - 1] + b->in_sizes[first + c - 1] - base;
            if (!S.in.ensure(extent + 32)) {
                tb_set_error("device allocation failed (packed input chunk)");
                ok = false;
                break;
            }
            for (uint64_t i = 0; i < c; i++) S.h_offs[i] = b->in_offsets[first + i] - base;
            memcpy(S.h_meta, b->in_sizes + first, c * 4);
            ok = cuda_ok(cudaMemcpyAsync(d_isz, S.h_meta, c * 4, cudaMemcpyHostToDevice, S.st), "H2D sizes") &&
                 cuda_ok(cudaMemcpyAsync(d_offs, S.h_offs, c * 8, cudaMemcpyHostToDevice, S.st), "H2D offsets");
            if (ok && extent)
                ok = cuda_ok(cudaMemcpyAsync(S.in.p, b->in + base, extent, cudaMemcpyHostToDevice, S.st), "H2D frames");
            g_h2d += extent + c * 12;
        } else {
            const unsigned char *h_in = b->in + first * b->in_stride;
            // input rows: only as wide as the longest row of the chunk when per-row sizes are known
            size_t width = b->in_stride;
            if (b->in_sizes) {
                uint32_t mx = 0;
                for (uint64_t i = 0; i < c; i++) mx = b->in_sizes[first + i] > mx ? b->in_sizes[first + i] : mx;
                width = ((size_t)mx + 63) & ~(size_t)63;
                if (width > b->in_stride) width = b->in_stride;
                memcpy(S.h_meta, b->in_sizes + first, c * 4);
                ok = cuda_ok(cudaMemcpyAsync(d_isz, S.h_meta, c * 4, cudaMemcpyHostToDevice, S.st), "H2D sizes");
                g_h2d += c * 4;
            }
            g_h2d += width * c;
            if (ok && width == b->in_stride)
                ok = cuda_ok(cudaMemcpyAsync(S.in.p, h_in, c * b->in_stride, cudaMemcpyHostToDevice, S.st), "H2D rows");
            else if (ok && width)
                ok = cuda_ok(cudaMemcpy2DAsync(S.in.p, b->in_stride, h_in, b->in_stride, width, c, cudaMemcpyHostToDevice,
                                               S.st), "H2D rows");
        }
        if (!ok) break;
        BatchArgs a;
        a.in = S.in.p;
        a.in_offsets = packed_in ? d_offs : nullptr;
        a.in_sizes = b->in_sizes ? d_isz : nullptr;
        a.in_stride = b->in_stride;
        a.out = S.out.p;
        a.out_stride = b->out_stride;
        a.out_sizes = d_osz;
        a.status = d_stat;
        a.n_streams = c;
        tamp_res r;
        {
            const auto tl = std::chrono::steady_clock::now();
            std::lock_guard<std::mutex> launch_lock(g_mu);  // launch sequences (and their launcher-static state) are serialised
            t_lock += since(tl);
            CompBatchConf cfc = cf;  // (TB_F_APPEND_TAIL: stream 0 of the BATCH keeps its header, not stream 0 of every chunk)
            if (first != 0) cfc.flags &= ~TB_F_APPEND_TAIL;
            r = compress ? compress_device_locked(cfc, dict_staged ? staged_dict.p : nullptr, a, S.st, dict_staged)
                         : decompress_device_locked(dict_staged ? staged_dict.p : nullptr, wbits_max, a, S.st, dict_staged);
            if (r == TAMP_OK && po &&
                !launch_compact(S.out.p, b->out_stride, d_osz, c, S.packed.p, c * b->out_stride, d_offs, S.st)) {
                tb_set_error("compaction scratch allocation failed");
                r = TAMP_ERROR;
            }
        }
        if (r != TAMP_OK) {  // (the other slots may still be copying into the caller's buffers: drain below)
            failed = r;
            ok = false;
            break;
        }
        // sizes + status land in the slot's pinned mirror (one copy: they are adjacent in meta)
        ok = cuda_ok(cudaMemcpyAsync(S.h_meta + c * 4, d_osz, c * 5, cudaMemcpyDeviceToHost, S.st), "D2H sizes");
        (void)d_stat;
        g_d2h += c * 4 + (b->status ? c : 0);
        if (ok && !compress) {  // decompressed rows are (nearly) full: one contiguous copy
            ok = cuda_ok(cudaMemcpyAsync(b->out + first * b->out_stride, S.out.p, c * b->out_stride,
                                         cudaMemcpyDeviceToHost, S.st), "D2H rows");
            g_d2h += c * b->out_stride;
        }
        ok = ok && cuda_ok(cudaEventRecord(S.ev, S.st), "event record");
        S.busy = ok;
        t_enq += since(t1);
    }
    const double t_loop = since(t_begin);
    for (uint64_t k = idx >= (uint64_t)nslots ? idx - nslots : 0; k < idx; k++)  // the chunks still in flight, oldest first (packed output grows in order)
        ok = pipe_finish_slot(slots[k % nslots], compress, b, po) && ok;
    for (int si = 0; si < nslots; si++) ok = cuda_ok(cudaStreamSynchronize(slots[si].st), "pipeline drain") && ok;
    if (po) po->offsets[n] = po->total;
    if (trace)
        fprintf(stderr, "[tamp_b200] %s%s: %llu streams, %llu chunks of %llu: loop %.2f ms (waiting for slots %.2f, enqueue %.2f of which "
                        "launch lock %.2f), drain %.2f ms\n", compress ? "compress" : "decompress", po ? " packed" : (packed_in ? " packed-in" : ""),
                (unsigned long long)n, (unsigned long long)idx, (unsigned long long)chunk, t_loop, t_wait, t_enq, t_lock, since(t_begin) - t_loop);
    if (failed != TAMP_OK) return failed;
    if (po && po->overflow) {
        tb_set_error("packed output does not fit its buffer (%llu bytes)", (unsigned long long)po->capacity);
        return TAMP_OUTPUT_FULL;
    }
    return ok ? TAMP_OK : TAMP_ERROR;
}

// Host-pointer variants: stage the batch through engine-owned device buffers.
// Layout of the meta buffer: [in_offsets n*8][in_sizes n*4][out_sizes n*4][status n].
static tamp_res host_batch(bool compress, const TampConf *conf, const unsigned char *dictionary, uint8_t wbits_max,
                           const TampB200Batch *b, bool write_token, PackedOut *po = nullptr, int extra_flags = 0) {
    CompBatchConf cf{};
    if (!b) return TAMP_INVALID_CONF;
    if (compress && !conf_to_batch(conf, cf, write_token)) return TAMP_INVALID_CONF;
    cf.flags |= extra_flags;  // (TB_F_APPEND | TB_F_APPEND_TAIL: the segments of one stream)
    if (!compress && (wbits_max < 8 || wbits_max > 15)) return TAMP_INVALID_CONF;
    std::unique_lock<std::mutex> lk(g_mu);
    if (g_eng.ready) cudaSetDevice(g_eng.device);  // host pointers only: the calling thread adopts the engine's device
    if (!engine_init_locked()) return TAMP_ERROR;
    Engine &E = g_eng;
    cudaSetDevice(E.device);
    const uint64_t n = b->n_streams;
    if (n == 0) return TAMP_OK;
    // packed input (contiguous frames + offsets) takes the pipelined path when the frames lie in ascending order
    bool packed_ok = !compress && b->in_offsets && b->in_sizes && n >= 4096;
    if (packed_ok)
        for (uint64_t i = 1; i < n && packed_ok; i++)
            packed_ok = b->in_offsets[i] >= b->in_offsets[i - 1] + b->in_sizes[i - 1] &&
                        b->in_offsets[i] - b->in_offsets[i - 1] <= ((uint64_t)1 << 20);
    if (po && (b->in_offsets || n < 1)) return TAMP_INVALID_CONF;  // packed output: strided input only
    if (getenv("TAMP_B200_TRACE") && b->in_offsets)
        fprintf(stderr, "[tamp_b200] %s with in_offsets: n %llu, in_sizes %p, pipelined %d (off[0] %llu off[1] %llu size[0] %u)\n",
                compress ? "compress" : "decompress", (unsigned long long)n, (const void *)b->in_sizes, (int)packed_ok,
                (unsigned long long)b->in_offsets[0], (unsigned long long)(n > 1 ? b->in_offsets[1] : 0), b->in_sizes ? b->in_sizes[0] : 0u);
    if ((!b->in_offsets && (n >= 4096 || po)) || packed_ok) {
        // pipelined path: its staging slots are per direction, so one compress call and one decompress call may be in
        // flight together (two host threads): the H2D-heavy call and the D2H-heavy call then keep both PCIe directions busy
        lk.unlock();
        std::lock_guard<std::mutex> dir_lock(g_dir_mu[compress ? 0 : 1]);
        cudaSetDevice(E.device);
        return host_batch_pipelined(compress, cf, dictionary, wbits_max, b, po);
    }
    // total input extent
    uint64_t in_bytes = 0;
    if (b->in_offsets || b->in_sizes) {
        for (uint64_t i = 0; i < n; i++) {
            uint64_t off = b->in_offsets ? b->in_offsets[i] : i * b->in_stride;
            uint64_t sz = b->in_sizes ? b->in_sizes[i] : b->in_stride;
            if (off + sz > in_bytes) in_bytes = off + sz;
        }
    } else {
        in_bytes = n * b->in_stride;
    }
    const uint64_t out_bytes = n * b->out_stride;
    const uint64_t meta_bytes = n * 17;
    if (!E.b_in.ensure(in_bytes + 16) || !E.b_out.ensure(out_bytes + 16) || !E.b_meta.ensure(meta_bytes + 64)) {
        tb_set_error("device allocation failed (%llu in, %llu out)", (unsigned long long)in_bytes,
                     (unsigned long long)out_bytes);
        return TAMP_ERROR;
    }
    cudaStream_t st = E.stream;
    uint64_t *d_off = reinterpret_cast<uint64_t *>(E.b_meta.p);
    uint32_t *d_isz = reinterpret_cast<uint32_t *>(E.b_meta.p + n * 8);
    uint32_t *d_osz = reinterpret_cast<uint32_t *>(E.b_meta.p + n * 12);
    int8_t *d_stat = reinterpret_cast<int8_t *>(E.b_meta.p + n * 16);
    bool ok = true;
    if (in_bytes) ok = ok && cuda_ok(cudaMemcpyAsync(E.b_in.p, b->in, in_bytes, cudaMemcpyHostToDevice, st), "H2D in");
    g_h2d += in_bytes + (b->in_offsets ? n * 8 : 0) + (b->in_sizes ? n * 4 : 0);
    g_d2h += out_bytes + n * 4 + (b->status ? n : 0);
    if (b->in_offsets)
        ok = ok && cuda_ok(cudaMemcpyAsync(d_off, b->in_offsets, n * 8, cudaMemcpyHostToDevice, st), "H2D offsets");
    if (b->in_sizes)
        ok = ok && cuda_ok(cudaMemcpyAsync(d_isz, b->in_sizes, n * 4, cudaMemcpyHostToDevice, st), "H2D sizes");
    if (!ok) return TAMP_ERROR;
    BatchArgs a;
    a.in = E.b_in.p;
    a.in_offsets = b->in_offsets ? d_off : nullptr;
    a.in_sizes = b->in_sizes ? d_isz : nullptr;
    a.in_stride = b->in_stride;
    a.out = E.b_out.p;
    a.out_stride = b->out_stride;
    a.out_sizes = d_osz;
    a.status = d_stat;
    a.n_streams = n;
    tamp_res r = compress ? compress_device_locked(cf, dictionary, a, st)
                          : decompress_device_locked(dictionary, wbits_max, a, st);
    if (r != TAMP_OK) return r;
    ok = cuda_ok(cudaMemcpyAsync(b->out, E.b_out.p, out_bytes, cudaMemcpyDeviceToHost, st), "D2H out") &&
         cuda_ok(cudaMemcpyAsync(b->out_sizes, d_osz, n * 4, cudaMemcpyDeviceToHost, st), "D2H sizes");
    if (ok && b->status) ok = cuda_ok(cudaMemcpyAsync(b->status, d_stat, n, cudaMemcpyDeviceToHost, st), "D2H status");
    ok = ok && cuda_ok(cudaStreamSynchronize(st), "batch kernel");
    return ok ? TAMP_OK : TAMP_ERROR;
}

tamp_res tamp_b200_compress_batch(const TampConf *conf, const unsigned char *dictionary, const TampB200Batch *batch,
                                  bool write_token) {
    return host_batch(true, conf, dictionary, 0, batch, write_token);
}

tamp_res tamp_b200_compress_batch_packed(const TampConf *conf, const unsigned char *dictionary, const TampB200Batch *batch,
                                         bool write_token, unsigned char *packed, uint64_t packed_capacity, uint64_t *offsets) {
    if (!batch || !offsets || (!packed && packed_capacity) || !batch->out_sizes) return TAMP_INVALID_CONF;
    PackedOut po;
    po.packed = packed;
    po.capacity = packed_capacity;
    po.offsets = offsets;
    if (batch->n_streams == 0) {
        offsets[0] = 0;
        return TAMP_OK;
    }
    // rows of the worst-case size are staged on the device only: the caller's batch->out / out_stride are not used
    TampB200Batch b = *batch;
    b.out = nullptr;
    b.out_stride = (tamp_b200_compress_bound(conf, (size_t)batch->in_stride) + 15) & ~(size_t)15;
    return host_batch(true, conf, dictionary, 0, &b, write_token, &po);
}

tamp_res tamp_b200_decompress_batch(const unsigned char *dictionary, uint8_t window_bits_max,
                                    const TampB200Batch *batch) {
    return host_batch(false, nullptr, dictionary, window_bits_max, batch, false);
}

tamp_res tamp_b200_compact_batch_device(const TampB200Batch *batch, unsigned char *packed, uint64_t packed_capacity,
                                        uint64_t *offsets, void *cuda_stream) {
    if (!batch || !offsets || (!packed && packed_capacity) || !batch->out_sizes) return TAMP_INVALID_CONF;
    std::lock_guard<std::mutex> lk(g_mu);
    if (!engine_init_locked()) return TAMP_ERROR;
    if (!launch_compact(batch->out, batch->out_stride, batch->out_sizes, batch->n_streams, packed, packed_capacity,
                        offsets, (cudaStream_t)cuda_stream)) {
        tb_set_error("compaction scratch allocation failed");
        return TAMP_ERROR;
    }
    return cuda_ok(cudaGetLastError(), "compact launch") ? TAMP_OK : TAMP_ERROR;
}

// ---- ONE long stream as a batch of segments ---------------------------------------------------
// dictionary_reset + append mode are the format's own mechanism for cutting a stream into independently (de)codable
// pieces (compressor.c:227-234, :847-881; decompressor.c:501-514).  Segment 0 is a dictionary_reset stream, every later
// segment an append-mode stream (a FLUSH padded to 16 bits where the header would be), each closed by
// flush(write_token = true): the concatenation is byte for byte what ONE reference compressor writes when
// tamp_compressor_reset_dictionary() is called after every segment_size input bytes, and any Tamp decompressor reads it
// as one stream.  With the segment offsets as an index the decompressor works segment-parallel as well.

uint64_t tamp_b200_segment_count(uint64_t in_size, uint64_t segment_size) {
    if (!segment_size) return 0;
    return in_size ? (in_size + segment_size - 1) / segment_size : 1;
}

uint64_t tamp_b200_segmented_bound(const TampConf *conf, uint64_t in_size, uint64_t segment_size) {
    return tamp_b200_segment_count(in_size, segment_size) * (uint64_t)tamp_b200_compress_bound(conf, (size_t)segment_size);
}

static bool segment_conf(const TampConf *conf, uint64_t segment_size, CompBatchConf &cf) {
    TampConf c;
    memset(&c, 0, sizeof c);
    c.window = 10;
    c.literal = 8;
    c.extended = 1;
    if (conf) c = *conf;
    // (a reset re-seeds the dictionary, so a custom dictionary cannot carry over; append is what this call adds itself)
    if (c.use_custom_dictionary || c.append) return false;
    if (segment_size == 0 || (segment_size & 15) || segment_size > ((uint64_t)1 << 30)) return false;
    c.dictionary_reset = 1;
    return conf_to_batch(&c, cf, /*write_token=*/true);
}

tamp_res tamp_b200_compress_segmented_device(const TampConf *conf, const unsigned char *in, uint64_t in_size,
                                             uint64_t segment_size, unsigned char *out, uint64_t out_capacity,
                                             uint64_t *seg_offsets, uint64_t *out_size, void *cuda_stream) {
    CompBatchConf cf;
    if (out_size) *out_size = 0;
    if ((!in && in_size) || (!out && out_capacity) || !segment_conf(conf, segment_size, cf)) return TAMP_INVALID_CONF;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const uint64_t nseg = tamp_b200_segment_count(in_size, segment_size), nfull = in_size / segment_size;
    const uint64_t rem = in_size - nfull * segment_size;
    const uint64_t stride = ((uint64_t)tamp_b200_compress_bound(conf, (size_t)segment_size) + 15) & ~(uint64_t)15;
    StreamTemp rows, meta, tail, offs;
    std::vector<int8_t> h_status(nseg);
    uint64_t total = 0;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (!engine_init_locked()) return TAMP_ERROR;
        // worst-case rows on the device, the sizes / statuses behind them; the last, shorter segment through a padded copy
        // (the kernels read their input 16 bytes at a time)
        if (!rows.alloc(nseg * stride, st) || !meta.alloc(nseg * 5 + 16, st) || (!seg_offsets && !offs.alloc((nseg + 1) * 8, st)) ||
            (nseg > nfull && !tail.alloc(segment_size, st))) {
            tb_set_error("scratch allocation failed");
            return TAMP_ERROR;
        }
        uint32_t *d_osz = reinterpret_cast<uint32_t *>(meta.p);
        uint32_t *d_tail_size = d_osz + nseg;
        int8_t *d_stat = reinterpret_cast<int8_t *>(meta.p + nseg * 4 + 16);
        uint64_t *d_offs = seg_offsets ? seg_offsets : reinterpret_cast<uint64_t *>(offs.p);
        BatchArgs a;
        a.in_offsets = nullptr;
        a.in_stride = segment_size;
        a.out_stride = stride;
        if (nfull) {
            a.in = in;
            a.in_sizes = nullptr;
            a.out = rows.p;
            a.out_sizes = d_osz;
            a.status = d_stat;
            a.n_streams = nfull;
            CompBatchConf c2 = cf;
            c2.flags |= TB_F_APPEND | TB_F_APPEND_TAIL;
            const tamp_res r = compress_device_locked(c2, nullptr, a, st);
            if (r != TAMP_OK) return r;
        }
        if (nseg > nfull) {
            const uint32_t rem32 = (uint32_t)rem;
            bool ok = cuda_ok(cudaMemsetAsync(tail.p, 0, segment_size, st), "tail segment") &&
                      cuda_ok(cudaMemcpyAsync(d_tail_size, &rem32, 4, cudaMemcpyHostToDevice, st), "tail size");
            if (ok && rem) ok = cuda_ok(cudaMemcpyAsync(tail.p, in + nfull * segment_size, rem, cudaMemcpyDeviceToDevice, st), "tail segment");
            if (!ok) return TAMP_ERROR;
            a.in = tail.p;
            a.in_sizes = d_tail_size;
            a.out = rows.p + nfull * stride;
            a.out_sizes = d_osz + nfull;
            a.status = d_stat + nfull;
            a.n_streams = 1;
            CompBatchConf c2 = cf;
            if (nfull) c2.flags |= TB_F_APPEND;
            const tamp_res r = compress_device_locked(c2, nullptr, a, st);
            if (r != TAMP_OK) return r;
        }
        if (!launch_compact(rows.p, stride, d_osz, nseg, out, out_capacity, d_offs, st)) {
            tb_set_error("compaction scratch allocation failed");
            return TAMP_ERROR;
        }
        if (!cuda_ok(cudaMemcpyAsync(h_status.data(), d_stat, nseg, cudaMemcpyDeviceToHost, st), "D2H status") ||
            !cuda_ok(cudaMemcpyAsync(&total, d_offs + nseg, 8, cudaMemcpyDeviceToHost, st), "D2H total"))
            return TAMP_ERROR;
    }
    if (!cuda_ok(cudaStreamSynchronize(st), "segmented compress")) return TAMP_ERROR;
    for (uint64_t i = 0; i < nseg; i++)
        if (h_status[i] != TAMP_OK) return (tamp_res)h_status[i];
    if (out_size) *out_size = total;
    return total > out_capacity ? TAMP_OUTPUT_FULL : TAMP_OK;
}

tamp_res tamp_b200_decompress_segmented_device(const unsigned char *in, const uint64_t *seg_offsets, uint64_t n_segments,
                                               uint64_t segment_size, uint8_t window_bits_max, unsigned char *out,
                                               uint64_t out_capacity, uint64_t *out_size, void *cuda_stream) {
    if (out_size) *out_size = 0;
    if (!n_segments) return TAMP_OK;
    if (!in || !seg_offsets || (!out && out_capacity) || window_bits_max < 8 || window_bits_max > 15 || segment_size < 32 ||
        (segment_size & 15) || segment_size > ((uint64_t)1 << 30))
        return TAMP_INVALID_CONF;
    const uint64_t nfull = n_segments - 1;  // every segment but the last one holds segment_size bytes
    if (nfull * segment_size > out_capacity) return TAMP_OUTPUT_FULL;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    StreamTemp meta, row;
    std::vector<uint32_t> h_osz(n_segments);
    std::vector<int8_t> h_status(n_segments);
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (!engine_init_locked()) return TAMP_ERROR;
    }
    // the configuration of the whole stream: segment 0's header byte (two small reads, outside the engine lock)
    uint64_t off0 = 0;
    uint8_t h0 = 0;
    if (!cuda_ok(cudaMemcpyAsync(&off0, seg_offsets, 8, cudaMemcpyDeviceToHost, st), "D2H offset") ||
        !cuda_ok(cudaStreamSynchronize(st), "D2H offset") ||
        !cuda_ok(cudaMemcpyAsync(&h0, in + off0, 1, cudaMemcpyDeviceToHost, st), "D2H header") ||
        !cuda_ok(cudaStreamSynchronize(st), "D2H header"))
        return TAMP_ERROR;
    if (nfull && !(h0 & 1u)) {
        tb_set_error("segments behind the first need a dictionary_reset stream (header bit 0)");
        return TAMP_INVALID_CONF;
    }
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (!meta.alloc(n_segments * 9 + 16, st) || !row.alloc(segment_size, st)) {
            tb_set_error("scratch allocation failed");
            return TAMP_ERROR;
        }
        uint32_t *d_isz = reinterpret_cast<uint32_t *>(meta.p), *d_osz = d_isz + n_segments;
        int8_t *d_stat = reinterpret_cast<int8_t *>(meta.p + n_segments * 8);
        launch_offsets_to_sizes(seg_offsets, n_segments, d_isz, st);
        BatchArgs a;
        a.in = in;
        a.in_stride = 0;
        a.out_stride = segment_size;
        if (nfull) {  // straight into the caller's buffer: row i is bytes [i * segment_size, (i + 1) * segment_size)
            a.in_offsets = seg_offsets;
            a.in_sizes = d_isz;
            a.out = out;
            a.out_sizes = d_osz;
            a.status = d_stat;
            a.n_streams = nfull;
            a.seg_header = 0x100u | h0;
            const tamp_res r = decompress_device_locked(nullptr, window_bits_max, a, st);
            if (r != TAMP_OK) return r;
        }
        // the last segment through a row of its own: the caller's buffer may end before segment_size bytes
        a.in_offsets = seg_offsets + nfull;
        a.in_sizes = d_isz + nfull;
        a.out = row.p;
        a.out_sizes = d_osz + nfull;
        a.status = d_stat + nfull;
        a.n_streams = 1;
        a.seg_header = nfull ? (0x300u | h0) : 0u;
        const tamp_res r = decompress_device_locked(nullptr, window_bits_max, a, st);
        if (r != TAMP_OK) return r;
        if (!cuda_ok(cudaMemcpyAsync(h_osz.data(), d_osz, n_segments * 4, cudaMemcpyDeviceToHost, st), "D2H sizes") ||
            !cuda_ok(cudaMemcpyAsync(h_status.data(), d_stat, n_segments, cudaMemcpyDeviceToHost, st), "D2H status"))
            return TAMP_ERROR;
    }
    if (!cuda_ok(cudaStreamSynchronize(st), "segmented decompress")) return TAMP_ERROR;
    for (uint64_t i = 0; i < nfull; i++) {
        // a full segment ends with INPUT_EXHAUSTED, or with OUTPUT_FULL in front of its closing FLUSH (decompressor.c:433-463)
        if (h_status[i] < 0) return (tamp_res)h_status[i];
        if (h_osz[i] != segment_size) {
            tb_set_error("segment %llu holds %u bytes, not the segment size", (unsigned long long)i, h_osz[i]);
            return TAMP_ERROR;
        }
    }
    if (h_status[nfull] < 0) return (tamp_res)h_status[nfull];
    const uint64_t room = out_capacity - nfull * segment_size;
    const uint64_t take = h_osz[nfull] < room ? h_osz[nfull] : room;
    if (take && (!cuda_ok(cudaMemcpyAsync(out + nfull * segment_size, row.p, take, cudaMemcpyDeviceToDevice, st), "last segment") ||
                 !cuda_ok(cudaStreamSynchronize(st), "last segment")))
        return TAMP_ERROR;
    if (out_size) *out_size = nfull * segment_size + take;
    return h_osz[nfull] > room ? TAMP_OUTPUT_FULL : TAMP_OK;
}

// Host-pointer forms.  Compress, 64 full segments or more: the pipelined packed path of tamp_b200_compress_batch_packed
// (chunks of segments flow through the staging slots, contiguous frames come back), the last, shorter segment through a
// call of its own.  Otherwise, and for decompress: one copy in, the device call, one copy out.
static tamp_res compress_segmented_pipelined(const TampConf *conf, const unsigned char *in, uint64_t in_size,
                                             uint64_t segment_size, unsigned char *out, uint64_t out_capacity,
                                             uint64_t *seg_offsets, uint64_t *out_size) {
    TampConf c;
    memset(&c, 0, sizeof c);
    c.window = 10;
    c.literal = 8;
    c.extended = 1;
    if (conf) c = *conf;
    c.dictionary_reset = 1;
    const uint64_t nseg = tamp_b200_segment_count(in_size, segment_size), nfull = in_size / segment_size;
    const uint64_t stride = ((uint64_t)tamp_b200_compress_bound(&c, (size_t)segment_size) + 15) & ~(uint64_t)15;
    std::vector<uint64_t> offs(nseg + 1);
    std::vector<uint32_t> sizes(nseg);
    std::vector<int8_t> stat(nseg);
    TampB200Batch b;
    memset(&b, 0, sizeof b);
    b.in = in;
    b.in_stride = segment_size;
    b.out_stride = stride;
    b.out_sizes = sizes.data();
    b.status = stat.data();
    b.n_streams = nfull;
    PackedOut po;
    po.packed = out;
    po.capacity = out_capacity;
    po.offsets = offs.data();
    tamp_res r = host_batch(true, &c, nullptr, 0, &b, /*write_token=*/true, &po, TB_F_APPEND | TB_F_APPEND_TAIL);
    if (r == TAMP_OUTPUT_FULL && out_size) *out_size = tamp_b200_segmented_bound(&c, in_size, segment_size);  // (room that suffices)
    if (r != TAMP_OK) return r;
    for (uint64_t i = 0; i < nfull; i++)
        if (stat[i] != TAMP_OK) return (tamp_res)stat[i];
    uint64_t total = po.total;
    if (nseg > nfull) {  // the last segment: an append-mode stream of its own, in_sizes says how long it is
        c.append = 1;
        std::vector<unsigned char> row(stride);
        uint32_t rem = (uint32_t)(in_size - nfull * segment_size), produced = 0;
        int8_t st = 0;
        TampB200Batch t;
        memset(&t, 0, sizeof t);
        t.in = in + nfull * segment_size;
        t.in_sizes = &rem;
        t.in_stride = segment_size;
        t.out = row.data();
        t.out_stride = stride;
        t.out_sizes = &produced;
        t.status = &st;
        t.n_streams = 1;
        r = host_batch(true, &c, nullptr, 0, &t, /*write_token=*/true);
        if (r != TAMP_OK) return r;
        if (st != TAMP_OK) return (tamp_res)st;
        if (total + produced > out_capacity) {
            if (out_size) *out_size = total + produced;
            return TAMP_OUTPUT_FULL;
        }
        memcpy(out + total, row.data(), produced);
        total += produced;
        offs[nseg] = total;
    }
    if (seg_offsets) memcpy(seg_offsets, offs.data(), (nseg + 1) * sizeof(uint64_t));
    if (out_size) *out_size = total;
    return TAMP_OK;
}

tamp_res tamp_b200_compress_segmented(const TampConf *conf, const unsigned char *in, uint64_t in_size, uint64_t segment_size,
                                      unsigned char *out, uint64_t out_capacity, uint64_t *seg_offsets, uint64_t *out_size) {
    if (out_size) *out_size = 0;
    CompBatchConf cf;
    if ((!in && in_size) || (!out && out_capacity) || !segment_conf(conf, segment_size, cf)) return TAMP_INVALID_CONF;
    if (in_size / segment_size >= 64)
        return compress_segmented_pipelined(conf, in, in_size, segment_size, out, out_capacity, seg_offsets, out_size);
    cudaStream_t st;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (!engine_init_locked()) return TAMP_ERROR;
        st = g_eng.stream;
    }
    const uint64_t nseg = tamp_b200_segment_count(in_size, segment_size);
    const uint64_t bound = tamp_b200_segmented_bound(conf, in_size, segment_size);
    const uint64_t d_cap = out_capacity < bound ? out_capacity : bound;
    StreamTemp d_in, d_out, d_offs;
    if (!d_in.alloc(in_size + 16, st) || !d_out.alloc(d_cap + 16, st) || !d_offs.alloc((nseg + 1) * 8, st)) {
        tb_set_error("staging allocation failed");
        return TAMP_ERROR;
    }
    if (in_size && !cuda_ok(cudaMemcpyAsync(d_in.p, in, in_size, cudaMemcpyHostToDevice, st), "H2D in")) return TAMP_ERROR;
    g_h2d += in_size;
    uint64_t total = 0;
    const tamp_res r = tamp_b200_compress_segmented_device(conf, d_in.p, in_size, segment_size, d_out.p, d_cap,
                                                           reinterpret_cast<uint64_t *>(d_offs.p), &total, st);
    if (out_size) *out_size = total;
    if (r != TAMP_OK) return r;
    bool ok = cuda_ok(cudaMemcpyAsync(out, d_out.p, total, cudaMemcpyDeviceToHost, st), "D2H out");
    if (ok && seg_offsets) ok = cuda_ok(cudaMemcpyAsync(seg_offsets, d_offs.p, (nseg + 1) * 8, cudaMemcpyDeviceToHost, st), "D2H offsets");
    ok = ok && cuda_ok(cudaStreamSynchronize(st), "segmented compress");
    g_d2h += total + (seg_offsets ? (nseg + 1) * 8 : 0);
    return ok ? TAMP_OK : TAMP_ERROR;
}

tamp_res tamp_b200_decompress_segmented(const unsigned char *in, const uint64_t *seg_offsets, uint64_t n_segments,
                                        uint64_t segment_size, uint8_t window_bits_max, unsigned char *out,
                                        uint64_t out_capacity, uint64_t *out_size) {
    if (out_size) *out_size = 0;
    if (!n_segments) return TAMP_OK;
    if (!in || !seg_offsets || (!out && out_capacity)) return TAMP_INVALID_CONF;
    for (uint64_t i = 0; i < n_segments; i++)
        if (seg_offsets[i + 1] < seg_offsets[i] || seg_offsets[i + 1] - seg_offsets[i] > 0xFFFFFFFFull) return TAMP_INVALID_CONF;
    cudaStream_t st;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (!engine_init_locked()) return TAMP_ERROR;
        st = g_eng.stream;
    }
    const uint64_t in_size = seg_offsets[n_segments];
    const uint64_t want = n_segments * segment_size;
    const uint64_t d_cap = out_capacity < want ? out_capacity : want;
    StreamTemp d_in, d_out, d_offs;
    if (!d_in.alloc(in_size + 16, st) || !d_out.alloc(d_cap + 16, st) || !d_offs.alloc((n_segments + 1) * 8, st)) {
        tb_set_error("staging allocation failed");
        return TAMP_ERROR;
    }
    if (!cuda_ok(cudaMemcpyAsync(d_in.p, in, in_size, cudaMemcpyHostToDevice, st), "H2D in") ||
        !cuda_ok(cudaMemcpyAsync(d_offs.p, seg_offsets, (n_segments + 1) * 8, cudaMemcpyHostToDevice, st), "H2D offsets"))
        return TAMP_ERROR;
    g_h2d += in_size + (n_segments + 1) * 8;
    uint64_t total = 0;
    const tamp_res r = tamp_b200_decompress_segmented_device(d_in.p, reinterpret_cast<const uint64_t *>(d_offs.p), n_segments,
                                                             segment_size, window_bits_max, d_out.p, d_cap, &total, st);
    if (out_size) *out_size = total;
    if (r != TAMP_OK && r != TAMP_OUTPUT_FULL) return r;
    if (total && (!cuda_ok(cudaMemcpyAsync(out, d_out.p, total, cudaMemcpyDeviceToHost, st), "D2H out") ||
                  !cuda_ok(cudaStreamSynchronize(st), "segmented decompress")))
        return TAMP_ERROR;
    g_d2h += total;
    return r;
}

tamp_res tamp_b200_synth_device(int kind, uint64_t first_k, uint64_t n_streams, uint64_t stream_len,
                                unsigned char *d_out, void *cuda_stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!engine_init_locked()) return TAMP_ERROR;
    launch_synth(kind, first_k, n_streams, stream_len, d_out, (cudaStream_t)cuda_stream);
    return cuda_ok(cudaGetLastError(), "synth launch") ? TAMP_OK : TAMP_ERROR;
}

}  // extern "C"
