/* Model of the "segment walk" parse: every lane walks its own segment of S offsets from a
 * guessed entry, evaluating the match at an offset only when its walk reaches it; entries are then corrected from the
 * neighbour's exit until nothing changes (walks from different entries merge quickly).  Counts lock-step iterations
 * (one candidate compare or one zero-candidate step per lane per iteration), evaluated candidates and rounds.
 *   gcc -O2 -I../../oracle -o /tmp/walk_model walk_model.c && /tmp/walk_model [n_streams] [N] [wbits] [kind] [S] */
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
#define MAXN 4096
static int ncand[MAXN], blen[MAXN];
int ZC = 1, DYN = 0, CAP = 1 << 30, KPER = 1; double heavy_n = 0, heavy_c = 0;
int main(int argc, char **argv) {
    ZC = argc > 6 ? atoi(argv[6]) : 1; DYN = argc > 7 ? atoi(argv[7]) : 0; CAP = argc > 8 ? atoi(argv[8]) : 1 << 30; KPER = argc > 9 ? atoi(argv[9]) : 1;
    int ns = argc > 1 ? atoi(argv[1]) : 500, N = argc > 2 ? atoi(argv[2]) : 1024, wbits = argc > 3 ? atoi(argv[3]) : 10;
    int kind = argc > 4 ? atoi(argv[4]) : 0, S = argc > 5 ? atoi(argv[5]) : 32;
    int W = 1 << wbits, nseg = (N + S - 1) / S;
    SynthVocab vocab; synth_build_vocab(&vocab);
    uint8_t *dict = malloc(W + 64), *in = malloc(N + 64);
    seed_dict(dict, W);
    double t_iters = 0, t_evalc = 0, t_evalo = 0, t_rounds = 0, t_itersA = 0, max_rounds = 0, t_ideal = 0, t_visit = 0;
    double hist_rounds[40] = {0};
    for (int k = 0; k < ns; k++) {
        memset(in, 0, N + 64);
        synth_fill(kind, k, in, N, &vocab);
        for (int q = 0; q < N; q++) {
            int L = N - q < 15 ? N - q : 15, bl = 0, nc = 0;
            if (L >= 2) for (int x = 0; x < W - 1; x++) {
                uint8_t b0 = x < q ? in[x] : dict[x], b1 = x + 1 < q ? in[x + 1] : dict[x + 1];
                if (b0 != in[q] || b1 != in[q + 1]) continue;
                int room = W - x < L ? W - x : L, n = 2;
                while (n < room && (x + n < q ? in[x + n] : dict[x + n]) == in[q + n]) n++;
                if (n > bl) bl = n;
                if (x != q - 1) nc++;
            }
            ncand[q] = nc; blen[q] = bl;
        }
        { int v = 0; for (int p = 0; p < N; p += blen[p] < 2 ? 1 : blen[p]) v++; t_visit += v; }
        static char evaluated[MAXN], onpath[MAXN];
        static int entry[MAXN], exitv[MAXN];
        memset(evaluated, 0, sizeof evaluated); memset(onpath, 0, sizeof onpath);
        for (int i = 0; i < nseg; i++) { entry[i] = 0; exitv[i] = -1; }
        int rounds = 0, iters = 0, evalc = 0, evalo = 0, lanes = 32;
        for (;;) {
            /* one round: every segment whose entry changed (or first round) walks */
            int changed = 0;
            /* segments are processed by lane (seg % lanes), sequentially per lane; lock-step cost = max over lanes */
            static int lanecost[64]; memset(lanecost, 0, sizeof lanecost);
            static int newexit[MAXN];
            for (int i = 0; i < nseg; i++) {
                newexit[i] = exitv[i];
                if (rounds > 0 && exitv[i] >= 0 && onpath[i * S + entry[i]]) continue;  /* entry already on the marked path */
                if (i * S + entry[i] >= N) { newexit[i] = 0; continue; }
                int p = i * S + entry[i], cost = 0, merged = 0;
                /* unmark the old path below?  old marks before the merge point are stale: rebuild */
                static char np[MAXN]; memset(np + i * S, 0, S);
                while (p < (i + 1) * S && p < N) {
                    if (exitv[i] >= 0 && onpath[p]) { merged = 1; break; }
                    if (!evaluated[p]) { evaluated[p] = 1; cost += ncand[p] > 0 ? (ncand[p] > CAP ? (heavy_n++, heavy_c += ncand[p], 1) : (ncand[p] + KPER - 1) / KPER) : ZC; evalc += ncand[p]; evalo++; }
                    else cost += ncand[p] > 0 ? 1 : ZC;
                    np[p] = 1;
                    p += blen[p] < 2 ? 1 : blen[p];
                }
                if (merged) { for (int x = i * S; x < p; x++) onpath[x] = np[x]; }
                else { for (int x = i * S; x < (i + 1) * S && x < MAXN; x++) onpath[x] = np[x]; newexit[i] = p - (i + 1) * S; if (p >= N) newexit[i] = 0; }
                if (DYN) { int b = 0; for (int l = 1; l < lanes; l++) if (lanecost[l] < lanecost[b]) b = l; lanecost[b] += cost; } else lanecost[i % lanes] += cost;
            }
            int mx = 0; for (int l = 0; l < lanes; l++) if (lanecost[l] > mx) mx = lanecost[l];
            iters += mx; if (rounds == 0) t_itersA += mx;
            rounds++;
            for (int i = 0; i < nseg; i++) exitv[i] = newexit[i];
            for (int i = 1; i < nseg; i++) if (entry[i] != exitv[i - 1]) { entry[i] = exitv[i - 1]; changed = 1; }
            if (!changed) break;
            if (rounds > 35) { printf("no convergence\n"); break; }
        }
        /* check: marked path equals the true walk */
        { int p = 0; while (p < N) { if (!onpath[p]) { printf("WRONG path at stream %d offset %d\n", k, p); break; } p += blen[p] < 2 ? 1 : blen[p]; } }
        t_iters += iters; t_evalc += evalc; t_evalo += evalo; t_rounds += rounds; if (rounds > max_rounds) max_rounds = rounds;
        hist_rounds[rounds < 39 ? rounds : 39]++;
        t_ideal += (evalc + (double)evalo * 0.3) / 32;
    }
    printf("S=%d N=%d kind=%d: visited %.0f, evaluated offsets %.0f, evaluated candidates %.0f, lock-step iterations %.1f (round A %.1f), rounds %.2f (max %.0f)\n",
           S, N, kind, t_visit / ns, t_evalo / ns, t_evalc / ns, t_iters / ns, t_itersA / ns, t_rounds / ns, max_rounds);
    printf("heavy offsets (> %d candidates) per stream %.1f, their candidates %.0f\n", CAP, heavy_n / ns, heavy_c / ns);
    printf("rounds hist:"); for (int i = 1; i < 40; i++) if (hist_rounds[i]) printf(" %d:%.0f", i, hist_rounds[i]); printf("\n");
    return 0;
}
