/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * Threaded CPU batch driver used (a) as the `cpu_baseline` / `--impl reference` leg of bench.py and
 * (b) by the parity tests to produce expected bytes for many streams quickly.
 *
 * Built twice from this one file (oracle/Makefile):
 *   -DHARNESS_REF=1   linked with the UNMODIFIED reference C sources -> oracle/_ref/libharness_ref.so
 *                     per stream: tamp_compressor_init + tamp_compressor_compress_and_flush(write_token=false),
 *                     tamp_decompressor_init(conf=NULL) + tamp_decompressor_decompress — the call sequence of
 *                     devices/common/tamp_bench.c:113-120,:156-162 and fuzz/fuzz_round_trip.c:43-55.
 *   -DHARNESS_PORT=1  linked with oracle/tamp_oracle.c          -> oracle/_build/libharness_port.so
 *
 * Work is handed out in blocks of streams from an atomic counter; each worker owns its codec state and
 * window, re-initialised per stream exactly like the GPU path does.
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "synth.h"

#if HARNESS_REF
#include "tamp/compressor.h"
#include "tamp/decompressor.h"
#else
#include "tamp_oracle.h"
#endif

typedef struct {
    int op; /* 0 = compress, 1 = decompress, 2 = generate */
    int window, literal, extended, kind;
    const uint8_t *in;
    size_t in_stride;
    const uint32_t *in_sizes; /* NULL => every stream is in_stride bytes */
    uint8_t *out;
    size_t out_stride;
    uint32_t *out_sizes;
    int8_t *status;
    size_t n_streams;
    uint64_t first_k;
    atomic_size_t next;
    const SynthVocab *vocab;
} Job;

enum { BLOCK = 16 };

static void one_compress(Job *j, size_t s, uint8_t *window) {
    const uint8_t *src = j->in + s * j->in_stride;
    size_t n = j->in_sizes ? j->in_sizes[s] : j->in_stride;
    uint8_t *dst = j->out + s * j->out_stride;
#if HARNESS_REF
    TampCompressor c;
    TampConf conf;
    memset(&conf, 0, sizeof conf);
    conf.window = (uint16_t)j->window;
    conf.literal = (uint16_t)j->literal;
    conf.extended = (uint16_t)j->extended;
    size_t written = 0;
    tamp_res r = tamp_compressor_init(&c, &conf, window);
    if (r == TAMP_OK) r = tamp_compressor_compress_and_flush(&c, dst, j->out_stride, &written, src, n, NULL, false);
    j->out_sizes[s] = (uint32_t)written;
    if (j->status) j->status[s] = r;
#else
    (void)window;
    OracleConf conf = {j->window, j->literal, 0, j->extended, 0, 0};
    long r = oracle_compress(&conf, NULL, src, n, dst, j->out_stride, 0);
    j->out_sizes[s] = r < 0 ? 0u : (uint32_t)r;
    if (j->status) j->status[s] = (int8_t)(r < 0 ? (r == -100 ? 1 : r) : 0);
#endif
}

static void one_decompress(Job *j, size_t s, uint8_t *window) {
    const uint8_t *src = j->in + s * j->in_stride;
    size_t n = j->in_sizes ? j->in_sizes[s] : j->in_stride;
    uint8_t *dst = j->out + s * j->out_stride;
#if HARNESS_REF
    TampDecompressor d;
    size_t written = 0;
    tamp_res r = tamp_decompressor_init(&d, NULL, window, (uint8_t)j->window);
    if (r == TAMP_OK) r = tamp_decompressor_decompress(&d, dst, j->out_stride, &written, src, n, NULL);
    j->out_sizes[s] = (uint32_t)written;
    if (j->status) j->status[s] = r;
#else
    (void)window;
    int st = 0;
    long w = oracle_decompress(NULL, j->window, src, n, dst, j->out_stride, &st);
    j->out_sizes[s] = (uint32_t)w;
    if (j->status) j->status[s] = (int8_t)st;
#endif
}

static void *worker(void *arg) {
    Job *j = (Job *)arg;
    uint8_t *window = (uint8_t *)malloc(1u << 15);
    for (;;) {
        size_t b = atomic_fetch_add(&j->next, BLOCK);
        if (b >= j->n_streams) break;
        size_t e = b + BLOCK < j->n_streams ? b + BLOCK : j->n_streams;
        for (size_t s = b; s < e; s++) {
            if (j->op == 0)
                one_compress(j, s, window);
            else if (j->op == 1)
                one_decompress(j, s, window);
            else
                synth_fill(j->kind, j->first_k + s, j->out + s * j->out_stride, j->out_stride, j->vocab);
        }
    }
    free(window);
    return NULL;
}

static double run(Job *j, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t tid[256];
    struct timespec t0, t1;
    atomic_init(&j->next, 0);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < threads; t++) pthread_create(&tid[t], NULL, worker, j);
    for (int t = 0; t < threads; t++) pthread_join(tid[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* Returns wall seconds. */
double harness_compress(int window, int literal, int extended, const uint8_t *in, size_t in_stride,
                        const uint32_t *in_sizes, size_t n_streams, uint8_t *out, size_t out_stride,
                        uint32_t *out_sizes, int8_t *status, int threads) {
    Job j;
    memset(&j, 0, sizeof j);
    j.op = 0;
    j.window = window;
    j.literal = literal;
    j.extended = extended;
    j.in = in;
    j.in_stride = in_stride;
    j.in_sizes = in_sizes;
    j.out = out;
    j.out_stride = out_stride;
    j.out_sizes = out_sizes;
    j.status = status;
    j.n_streams = n_streams;
    return run(&j, threads);
}

double harness_decompress(int window_bits_max, const uint8_t *in, size_t in_stride, const uint32_t *in_sizes,
                          size_t n_streams, uint8_t *out, size_t out_stride, uint32_t *out_sizes, int8_t *status,
                          int threads) {
    Job j;
    memset(&j, 0, sizeof j);
    j.op = 1;
    j.window = window_bits_max;
    j.in = in;
    j.in_stride = in_stride;
    j.in_sizes = in_sizes;
    j.out = out;
    j.out_stride = out_stride;
    j.out_sizes = out_sizes;
    j.status = status;
    j.n_streams = n_streams;
    return run(&j, threads);
}

double harness_generate(int kind, uint64_t first_k, size_t n_streams, size_t stream_len, uint8_t *out, int threads) {
    static SynthVocab vocab;
    static int have_vocab = 0;
    if (!have_vocab) {
        synth_build_vocab(&vocab);
        have_vocab = 1;
    }
    Job j;
    memset(&j, 0, sizeof j);
    j.op = 2;
    j.kind = kind;
    j.first_k = first_k;
    j.out = out;
    j.out_stride = stream_len;
    j.n_streams = n_streams;
    j.vocab = &vocab;
    return run(&j, threads);
}

const char *harness_kind(void) {
#if HARNESS_REF
    return "reference";
#else
    return "port";
#endif
}
