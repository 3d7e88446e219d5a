/* What trigram chains would buy the segment walk ("next"): for the offsets the greedy parse polls, the
 * number of window positions that share the poll's first two bytes (what k_walk_compress evaluates today: its chains link
 * equal bigrams) against the number that share its first three, how many polls end with a 2-byte match (the ones that
 * would need the first-occurrence fallback), and the lock-step cost of both (ceil(candidates / 2) iterations per poll).
 *   gcc -O2 -I../../oracle -o /tmp/trigram_model trigram_model.c && /tmp/trigram_model [n_streams] [N] [wbits] [kind] */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "synth.h"

static void seed_dict(uint8_t *d, int n) {
    static const char chars[] = " \x000ei>to<ans\nr/.";
    uint32_t s = 3758097560u;
    for (int i = 0; i < n; i += 8) {
        uint32_t x = s; x ^= x << 13; x ^= x >> 17; x ^= x << 5; s = x;
        for (int j = 0; j < 8; j++) d[i + j] = (uint8_t)chars[(x >> (4 * j)) & 15];
    }
}

int main(int argc, char **argv) {
    int ns = argc > 1 ? atoi(argv[1]) : 300, N = argc > 2 ? atoi(argv[2]) : 1024, wbits = argc > 3 ? atoi(argv[3]) : 10;
    int kind = argc > 4 ? atoi(argv[4]) : 0, W = 1 << wbits;
    SynthVocab vocab; synth_build_vocab(&vocab);
    uint8_t *dict = malloc(W + 64), *in = malloc(N + 64);
    seed_dict(dict, W);
    double polls = 0, with2 = 0, c2 = 0, c3 = 0, it2 = 0, it3 = 0, len2 = 0, lits = 0, fallback_polls = 0;
    for (int k = 0; k < ns; k++) {
        memset(in, 0, N + 64);
        synth_fill(kind, k, in, N, &vocab);
        for (int q = 0; q < N;) {  /* the greedy parse (streams no longer than the window: window[x] = x < q ? in[x] : dict[x]) */
            int L = N - q < 15 ? N - q : 15, best = 0, n2 = 0, n3 = 0;
            if (L >= 2) for (int x = 0; x < W - 1; x++) {
                int room = W - x < L ? W - x : L, n = 0;
                while (n < room && (x + n < q ? in[x + n] : dict[x + n]) == in[q + n]) n++;
                if (n >= 2) n2++;
                if (n >= 3) n3++;
                if (n > best) best = n;
            }
            polls++;
            if (n2) { with2++; c2 += n2; it2 += (n2 + 1) / 2; }
            c3 += n3; it3 += (n3 + 1) / 2;
            if (best == 2) { len2++; }
            if (n2 && !n3) fallback_polls++;
            if (best < 2) lits++;
            q += best < 2 ? 1 : best;
        }
    }
    printf("N=%d window=%d kind=%d, per stream: polls %.1f (literals %.1f), polls with a bigram candidate %.1f\n", N, wbits, kind,
           polls / ns, lits / ns, with2 / ns);
    printf("  bigram chains : %.1f candidates, %.1f lane-iterations of two candidates\n", c2 / ns, it2 / ns);
    printf("  trigram chains: %.1f candidates, %.1f lane-iterations; polls that end with a 2-byte match %.1f (%.1f of them have no trigram candidate at all)\n",
           c3 / ns, it3 / ns, len2 / ns, fallback_polls / ns);
    return 0;
}
