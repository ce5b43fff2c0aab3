/* Smallest possible GPU check of a launch-planning change (no Python start-up: a few seconds of box time):
 *   mini_check M N   scores one random M x N pair (semiglobal Gotoh) with the engine's own launch plan and with forced
 *   2 and 3 warps per scheduler; the three scores must agree; prints the device time of each.
 * build: gcc -O2 -Iinclude tools/mini_check.c -Lanyseq_b200/_build -lanyseq_b200 -Wl,-rpath,'$ORIGIN' -o anyseq_b200/_build/mini_check */
#include <stdio.h>
#include <stdlib.h>
#include "anyseq.h"

int main(int argc, char** argv)
{
    const int m = argc > 1 ? atoi(argv[1]) : 1000000, n = argc > 2 ? atoi(argv[2]) : 1000000;
    char* q = malloc(m);
    char* s = malloc(n);
    unsigned long long x = 88172645463325252ull;
    for (int i = 0; i < m; ++i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; q[i] = "ACGT"[x & 3]; }
    for (int j = 0; j < n; ++j) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; s[j] = (j < m && (x & 15)) ? q[j] : "ACGT"[(x >> 8) & 3]; }
    anyseq_ctx* ctx;
    if (anyseq_ctx_create(0, &ctx) != ANYSEQ_OK) { fprintf(stderr, "no device: %s\n", anyseq_last_error()); return 2; }
    anyseq_scoring sc = {ANYSEQ_SEMIGLOBAL, 2, -1, -2, -1};
    long long ref = 0;
    int bad = 0;
    for (int pass = 0; pass < 3; ++pass) {
        const int bps = pass == 0 ? 0 : pass + 1;
        anyseq_ctx_tune(ctx, 0, 0, bps, 20000);
        anyseq_result r;
        float best = 1e30f;
        for (int rep = 0; rep < 2; ++rep) {
            if (anyseq_score(ctx, &sc, q, m, s, n, &r) != ANYSEQ_OK) { fprintf(stderr, "error: %s\n", anyseq_last_error()); return 3; }
            if (r.kernel_ms < best) best = r.kernel_ms;
        }
        if (pass == 0) ref = r.score;
        if (r.score != ref) bad = 1;
        printf("%d x %d warps/scheduler %s: %.2f ms %.1f GCUPS score %lld end (%d, %d)\n", m, n, bps == 0 ? "auto" : (bps == 2 ? "2" : "3"),
               best, (double)m * n / best / 1e6, (long long)r.score, r.end_i, r.end_j);
    }
    anyseq_ctx_destroy(ctx);
    printf(bad ? "MISMATCH\n" : "scores agree\n");
    return bad;
}
