/*
 * fullsize_check.c -- TEST INFRASTRUCTURE ONLY (part of oracle/; see anyseq_oracle.h for the rules).
 *
 * A second, independent CPU implementation of the score-only recurrences the oracle restates
 * (relax_global / relax_local of src/align.impala:46-79 with the build-defined Gotoh extension of
 * anyseq_oracle.c, result extraction of src/scoring.impala:29-77), written to be FAST enough for the
 * full 4.6 Mbp x 4.6 Mbp size of BASELINE.json configs[1] on a handful of host cores, which the scalar
 * block wavefront of anyseq_oracle.c (the faithful restatement) is not (about 4 h here).
 *
 * How it differs from the restatement -- on purpose, so that agreement means something:
 *   - rows are relaxed whole-vector-at-a-time: H without the horizontal gap first
 *     (M = max(diag + sigma, F)), then E by a prefix maximum over A[j] = M[j] + go - j*ge
 *     (valid because E(i,j) = max_k<j (H(i,k) + go + (j-1-k) ge) and the E term inside H(i,k) is
 *     dominated: E(i,k) + go <= E(i,k) + ge), then H = max(M, E);
 *   - tiles of FS_TH rows x FS_TW columns in a tile anti-diagonal wavefront (OpenMP), border vectors
 *     row H/F, column H/E, one corner per tile column.
 * It is validated against oracle_score_affine / oracle_score_linear on every size the scalar path finishes
 * (tests/test_oracle.py::test_fullsize_check_equals_oracle) and then used by tools/c2_check.py to freeze the
 * full-size result in tests/golden/.  Never linked into the product.
 */
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "anyseq_oracle.h"

#define FS_NEG_INF (-(1 << 30))
#define VL 16
typedef int32_t v16 __attribute__((vector_size(64), aligned(4)));
typedef int32_t v16a __attribute__((vector_size(64)));

static inline v16a vmax(v16a a, v16a b) { const v16a m = a > b; return (a & m) | (b & ~m); }
static inline v16a vset(int32_t x) { return (v16a){x, x, x, x, x, x, x, x, x, x, x, x, x, x, x, x}; }

/* lanes shifted up by S (lane l <- lane l-S), vacated lanes filled from `fill` */
#define SHIFT_UP(x, fill, S)                                                                              \
    __builtin_shuffle((fill), (x),                                                                        \
                      (v16a){16 - (S) + 0 < 16 ? 0 : 16 + 0 - (S), 1 < (S) ? 1 : 16 + 1 - (S),            \
                             2 < (S) ? 2 : 16 + 2 - (S), 3 < (S) ? 3 : 16 + 3 - (S),                      \
                             4 < (S) ? 4 : 16 + 4 - (S), 5 < (S) ? 5 : 16 + 5 - (S),                      \
                             6 < (S) ? 6 : 16 + 6 - (S), 7 < (S) ? 7 : 16 + 7 - (S),                      \
                             8 < (S) ? 8 : 16 + 8 - (S), 9 < (S) ? 9 : 16 + 9 - (S),                      \
                             10 < (S) ? 10 : 16 + 10 - (S), 11 < (S) ? 11 : 16 + 11 - (S),                \
                             12 < (S) ? 12 : 16 + 12 - (S), 13 < (S) ? 13 : 16 + 13 - (S),                \
                             14 < (S) ? 14 : 16 + 14 - (S), 15 < (S) ? 15 : 16 + 15 - (S)})

#define FS_TW 2048
#define FS_TH 512

typedef struct {
    int mode, same, diff, go, ge;   /* go = gap_init + gap_extend */
    const uint8_t* q;
    const uint8_t* s;
    int m, n, ntj;
    int32_t *rowH, *rowF;           /* [npad] bottom border of the tiles done so far */
    int32_t *colH, *colE;           /* [m]    right border of the tiles done so far */
    int32_t* corner;                /* [ntj]  H(i0-1, j0-1) of the next tile of a tile column */
    int32_t* sx;                    /* [npad] subject symbols as int32 (0x7fff past the end: matches nothing) */
    int32_t* tilebest;              /* [ntj]  local: running maximum per tile column */
} fs_ctx;

static void fs_tile(const fs_ctx* c, int bi, int bj)
{
    const int i0 = bi * FS_TH, i1 = i0 + FS_TH < c->m ? i0 + FS_TH : c->m;
    const int j0 = bj * FS_TW;
    const int nv = FS_TW / VL;
    const int local = c->mode == ORACLE_LOCAL;
    int32_t* rowH = c->rowH + j0;
    int32_t* rowF = c->rowF + j0;
    const int32_t* sx = c->sx + j0;
    const v16a vge = vset(c->ge), vgo = vset(c->go), vsame = vset(c->same), vdiff = vset(c->diff);
    const v16a vneg = vset(FS_NEG_INF), vzero = vset(0);
    const v16a lane_ge = (v16a){0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15} * vge;
    int32_t diag = c->corner[bj];
    c->corner[bj] = c->colH[i1 - 1];              /* corner of the tile below, before this tile overwrites it */
    v16a best = vset(ORACLE_SCORE_MIN);
    const int ncols_valid = (c->n - j0) < FS_TW ? (c->n - j0) : FS_TW;
    for (int i = i0; i < i1; ++i) {
        const int32_t hleft = c->colH[i], eleft = c->colE[i];
        const v16a vq = vset((int32_t)c->q[i]);
        /* E(i, j0) = max(E(i, j0-1) + ge, H(i, j0-1) + go); carry: E[jrel] = excl[jrel] + (jrel-1) ge */
        int32_t e0 = eleft + c->ge > hleft + c->go ? eleft + c->ge : hleft + c->go;
        int32_t carry = e0 + c->ge;
        v16a prev_up = vset(diag);                /* lane 15 = H(i-1, j0-1) */
        v16a hlast = vzero, elast = vzero;
        for (int v = 0; v < nv; ++v) {
            const v16a hup = *(const v16*)(rowH + v * VL);
            const v16a fup = *(const v16*)(rowF + v * VL);
            const v16a sv = *(const v16*)(sx + v * VL);
            const v16a hd = __builtin_shuffle(prev_up, hup, (v16a){15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30});   /* H(i-1, j-1) */
            prev_up = hup;
            const v16a eq = (sv == vq);
            const v16a sig = (vsame & eq) | (vdiff & ~eq);
            const v16a f = vmax(fup + vge, hup + vgo);
            v16a mm = vmax(hd + sig, f);
            if (local) mm = vmax(mm, vzero);
            /* A[j] = M[j] + go - jrel*ge, inclusive prefix maximum inside the vector */
            const v16a jge = vset(v * VL * c->ge) + lane_ge;       /* jrel * ge */
            v16a a = mm + vgo - jge;
            a = vmax(a, SHIFT_UP(a, vneg, 1));
            a = vmax(a, SHIFT_UP(a, vneg, 2));
            a = vmax(a, SHIFT_UP(a, vneg, 4));
            a = vmax(a, SHIFT_UP(a, vneg, 8));
            const v16a excl = vmax(SHIFT_UP(a, vneg, 1), vset(carry));
            const v16a e = excl + jge - vge;                       /* + (jrel - 1) ge */
            const v16a h = vmax(mm, e);
            carry = a[15] > carry ? a[15] : carry;
            *(v16*)(rowH + v * VL) = h;
            *(v16*)(rowF + v * VL) = f;
            if (local) {
                /* columns past the matrix edge never feed valid ones (left -> right only) but must not count */
                if ((v + 1) * VL <= ncols_valid) best = vmax(best, h);
                else for (int l = 0; l < VL; ++l) if (v * VL + l < ncols_valid && h[l] > best[l]) best[l] = h[l];
            }
            hlast = h; elast = e;
        }
        diag = hleft;
        /* right border for the next tile column; for the last tile column: the matrix's last column n-1 */
        if (bj == c->ntj - 1) {
            c->colH[i] = rowH[c->n - 1 - j0];
        } else {
            c->colH[i] = hlast[15];
            c->colE[i] = elast[15];
        }
    }
    if (local) {
        int32_t b = ORACLE_SCORE_MIN;
        for (int l = 0; l < VL; ++l) b = best[l] > b ? best[l] : b;
        if (b > c->tilebest[bj]) c->tilebest[bj] = b;
    }
}

/* score + end cell of the reference's get_score_pos(): global (m-1, n-1); semiglobal: last row first
 * (candidate -1 = border value 0 first, lowest index), last column only if strictly greater
 * (src/scoring.impala:39-77); local: score only (pos = -1, the end cell follows the block-slot rule of the
 * restatement and is not reproduced here). */
oracle_result fullsize_score(int mode, const uint8_t* q, int m, const uint8_t* s, int n,
                             int same, int diff, int gap_init, int gap_extend, int threads)
{
    oracle_result res = {0, -1, -1};
    if (m <= 0 || n <= 0) return res;
    fs_ctx c;
    memset(&c, 0, sizeof(c));
    c.mode = mode; c.same = same; c.diff = diff; c.ge = gap_extend; c.go = gap_init + gap_extend;
    c.q = q; c.s = s; c.m = m; c.n = n;
    const int ntj = (n + FS_TW - 1) / FS_TW, nti = (m + FS_TH - 1) / FS_TH;
    c.ntj = ntj;
    const size_t npad = (size_t)ntj * FS_TW;
    c.rowH = malloc(sizeof(int32_t) * npad); c.rowF = malloc(sizeof(int32_t) * npad);
    c.sx = malloc(sizeof(int32_t) * npad);
    c.colH = malloc(sizeof(int32_t) * (size_t)m); c.colE = malloc(sizeof(int32_t) * (size_t)m);
    c.corner = malloc(sizeof(int32_t) * (size_t)ntj); c.tilebest = malloc(sizeof(int32_t) * (size_t)ntj);
    const int glob = mode == ORACLE_GLOBAL;
    for (size_t j = 0; j < npad; ++j) {
        c.rowH[j] = glob ? gap_init + (int32_t)(j + 1) * gap_extend : 0;
        c.rowF[j] = FS_NEG_INF;
        c.sx[j] = j < (size_t)n ? (int32_t)s[j] : 0x7fff;     /* never equals a byte */
    }
    for (int i = 0; i < m; ++i) { c.colH[i] = glob ? gap_init + (i + 1) * gap_extend : 0; c.colE[i] = FS_NEG_INF; }
    for (int bj = 0; bj < ntj; ++bj) {
        const long long jl = (long long)bj * FS_TW - 1;          /* column left of the tile */
        c.corner[bj] = (glob && jl >= 0) ? gap_init + (int32_t)(jl + 1) * gap_extend : 0;
        c.tilebest[bj] = ORACLE_SCORE_MIN;
    }
    for (int d = 0; d < nti + ntj - 1; ++d) {
        const int blo = d - (ntj - 1) > 0 ? d - (ntj - 1) : 0;
        const int bhi = d < nti - 1 ? d : nti - 1;
        #pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
        for (int bi = blo; bi <= bhi; ++bi) fs_tile(&c, bi, d - bi);
    }
    /* rowH[0..n) = H(m-1, j); colH[0..m) = H(i, n-1) */
    if (mode == ORACLE_LOCAL) {
        int32_t b = ORACLE_SCORE_MIN;
        for (int bj = 0; bj < ntj; ++bj) b = c.tilebest[bj] > b ? c.tilebest[bj] : b;
        res.score = b;
    } else if (mode == ORACLE_GLOBAL) {
        res.score = c.rowH[n - 1]; res.pos_i = m - 1; res.pos_j = n - 1;
    } else {
        int32_t best = 0, bj_ = -1;
        for (int j = 0; j < n; ++j) if (c.rowH[j] > best) { best = c.rowH[j]; bj_ = j; }
        res.score = best; res.pos_i = m - 1; res.pos_j = bj_;
        int32_t cb = 0, ci = -1;
        for (int i = 0; i < m; ++i) if (c.colH[i] > cb) { cb = c.colH[i]; ci = i; }
        if (cb > best) { res.score = cb; res.pos_i = ci; res.pos_j = n - 1; }
    }
    free(c.rowH); free(c.rowF); free(c.sx); free(c.colH); free(c.colE); free(c.corner); free(c.tilebest);
    return res;
}
