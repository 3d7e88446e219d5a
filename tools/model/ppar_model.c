/* Design-space model for the position-parallel compressor: counts, on the bench's G_text streams,
 * what the candidate walk of P2 has to do under different chain keys and hand-out policies.  Test/analysis
 * infrastructure only (it includes the oracle's generator); nothing in the product links it.
 *   gcc -O2 -I../../oracle -o /tmp/ppar_model ppar_model.c && /tmp/ppar_model [n_streams] [N] [wbits] [kind]  */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "synth.h"

static uint32_t xs_state = 3758097560u;
static void seed_dict(uint8_t *d, int n) {  /* common.c:37-52 restated (literal 7/8 or v1 table) */
    static const char chars[] = " \x000ei>to<ans\nr/.";
    xs_state = 3758097560u;
    for (int i = 0; i < n; i += 8) {
        uint32_t x = xs_state; x ^= x << 13; x ^= x >> 17; x ^= x << 5; xs_state = x;
        for (int j = 0; j < 8; j++) d[i + j] = (uint8_t)chars[(x >> (4 * j)) & 15];
    }
}

int main(int argc, char **argv) {
    int ns = argc > 1 ? atoi(argv[1]) : 2000, N = argc > 2 ? atoi(argv[2]) : 1024, wbits = argc > 3 ? atoi(argv[3]) : 10;
    int kind = argc > 4 ? atoi(argv[4]) : 0;
    int W = 1 << wbits;
    SynthVocab vocab; synth_build_vocab(&vocab);
    uint8_t *dict = malloc(W + 64), *in = malloc(N + 64);
    seed_dict(dict, W);
    double tot_in = 0, tot_dict = 0, tot_items = 0, tot_tri_in = 0, tot_tri_dict = 0, tot_overlap = 0, tot_ge8 = 0, tot_ge4 = 0;
    double tot_visited = 0, tot_vis_cand = 0, tot_first_occ = 0, tot_need_bigram = 0;
    double hist[17] = {0};
    double it_by_T[33] = {0}, refills_by_T[33] = {0};
    double tri_it[33] = {0}, tri_ref[33] = {0};
    for (int k = 0; k < ns; k++) {
        memset(in, 0, N + 64);
        synth_fill(kind, k, in, N, &vocab);
        static int ncand[32768], ntri[32768], bestlen[32768];
        for (int q = 0; q < N; q++) {
            int L = N - q < 15 ? N - q : 15;
            ncand[q] = ntri[q] = 0; bestlen[q] = 0;
            if (L < 2) continue;
            int first = 1, bl = 0;
            for (int x = 0; x < W - 1; x++) {
                /* window at poll q: input below q, dictionary from q on */
                uint8_t b0 = x < q ? in[x] : dict[x], b1 = x + 1 < q ? in[x + 1] : dict[x + 1];
                if (b0 != in[q] || b1 != in[q + 1]) continue;
                int room = W - x < L ? W - x : L, n = 2;
                while (n < room && (x + n < q ? in[x + n] : dict[x + n]) == in[q + n]) n++;
                if (n > bl) bl = n;
                /* chain candidates: input-side need the input's own bigram at x (x + 1 < q always true here unless x == q-1) */
                int in_side = x < q;
                if (in_side && x == q - 1) { /* straddling candidate: not in any chain */ continue; }
                ncand[q]++;
                if (in_side) { tot_in++; if (first) first = 0; if (n >= q - x) tot_overlap++; } else tot_dict++;
                if (n >= 8) tot_ge8++;
                if (n >= 4) tot_ge4++;
                hist[n]++;
                /* trigram chains: the candidate shares 3 bytes (in its own array) */
                if (L >= 3) {
                    uint8_t b2 = in_side ? in[x + 2] : dict[x + 2];
                    if (x + 2 < W && b2 == in[q + 2] && (!in_side || x + 2 < N)) { ntri[q]++; if (in_side) tot_tri_in++; else tot_tri_dict++; }
                }
            }
            bestlen[q] = bl;
            if (first) tot_first_occ++;
            if (bl == 2) tot_need_bigram++;
        }
        /* greedy walk */
        for (int p = 0; p < N;) { tot_visited++; tot_vis_cand += ncand[p]; p += bestlen[p] < 2 ? 1 : bestlen[p]; }
        /* persistent-lane schedule */
        for (int tri = 0; tri < 2; tri++) {
            int *cnt = tri ? ntri : ncand;
            static int items[32768]; int ni = 0;
            for (int q = 0; q < N; q++) if (cnt[q] > 0) items[ni++] = cnt[q];
            if (!tri) tot_items += ni;
            for (int T = 1; T <= 32; T++) {
                int left[32] = {0}, next = 0, iters = 0, refills = 0;
                for (;;) {
                    int idle = 0; for (int l = 0; l < 32; l++) idle += left[l] == 0;
                    if (idle == 32 && next >= ni) break;
                    if (next < ni && (idle == 32 || idle >= T)) { refills++; for (int l = 0; l < 32 && next < ni; l++) if (!left[l]) left[l] = items[next++]; }
                    for (int l = 0; l < 32; l++) if (left[l]) left[l]--;
                    iters++;
                }
                if (tri) { tri_it[T] += iters; tri_ref[T] += refills; } else { it_by_T[T] += iters; refills_by_T[T] += refills; }
            }
        }
    }
    printf("streams %d N %d W %d kind %d\n", ns, N, W, kind);
    printf("bigram chains : input-side %.0f dict-side %.0f per stream; items (offsets with a candidate) %.0f\n", tot_in / ns, tot_dict / ns, tot_items / ns);
    printf("trigram chains: input-side %.0f dict-side %.0f per stream; offsets whose best is exactly 2: %.0f; bigram first occurrences %.0f\n", tot_tri_in / ns, tot_tri_dict / ns, tot_need_bigram / ns, tot_first_occ / ns);
    printf("candidates with n>=4: %.1f%%  n>=8: %.1f%%  overlap (n >= distance): %.2f%%\n", 100 * tot_ge4 / (tot_in + tot_dict), 100 * tot_ge8 / (tot_in + tot_dict), 100 * tot_overlap / (tot_in + tot_dict));
    printf("greedy walk visits %.0f offsets/stream, their candidates %.0f\n", tot_visited / ns, tot_vis_cand / ns);
    printf("len hist:"); for (int i = 2; i <= 15; i++) printf(" %d:%.1f%%", i, 100 * hist[i] / (tot_in + tot_dict)); printf("\n");
    printf("T  : bigram iters refills | trigram iters refills\n");
    for (int T = 1; T <= 32; T += (T < 12 ? 1 : 4)) printf("%2d : %6.1f %5.1f | %6.1f %5.1f\n", T, it_by_T[T] / ns, refills_by_T[T] / ns, tri_it[T] / ns, tri_ref[T] / ns);
    return 0;
}
