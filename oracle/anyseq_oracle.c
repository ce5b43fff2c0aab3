/*
 * anyseq_oracle.c -- TEST INFRASTRUCTURE ONLY.  See anyseq_oracle.h for the
 * usage contract and the parity-pinning status.
 *
 * CPU restatement of the reference's CPU path.  Every function cites the
 * reference file:line it follows (paths relative to /root/reference/).
 * Written from the reference's behaviour, not translated from its source
 * (the reference is Impala; this is plain C99 + OpenMP).
 */
#include "anyseq_oracle.h"

#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define PRED_NONE   0   /* src/align.impala:37-40 */
#define PRED_GAP_Q  1
#define PRED_GAP_S  2
#define PRED_NO_GAP 3

#define GAP_SYM   '_'   /* src/traceback.impala:1-2 */
#define EMPTY_SYM ' '

/* half of "minus infinity" for the affine E/F borders: survives any number of
 * realistic +ge additions without wrapping. */
#define NEG_INF (-(1 << 30))

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin3(int a, int b, int c) { return imin(imin(a, b), c); }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }   /* src/utils.impala:10-12 */

/* A reference "Vector": logical indices -1 .. length-1, stored with a +1 shift
 * (src/dynprog.impala:173-176,194-199). */
typedef struct { int32_t* mem; int length; } ivec;

static ivec ivec_new(int length)
{
    ivec v;
    v.length = length;
    v.mem = (int32_t*)calloc((size_t)(length > 0 ? length : 0) + 2, sizeof(int32_t));
    return v;
}
static void ivec_free(ivec v) { free(v.mem); }
#define V(v, i) ((v).mem[(i) + 1])

int32_t oracle_next_pow_2(int32_t i)   /* src/utils.impala:19-28 */
{
    if (i == 0) return 0;
    int32_t n = i - 1, r = 1;
    while (n > 0) { n >>= 1; r <<= 1; }
    return r;
}

/* init_scores_global / init_scores_local: src/align.impala:85-86 */
static inline int32_t init_score(int mode, int gap, int i)
{
    return mode == ORACLE_GLOBAL ? (i + 1) * gap : 0;
}
/* init_predc_*: src/align.impala:88-90 */
static inline uint8_t init_predc_rows(int mode, int i)
{
    if (mode != ORACLE_GLOBAL) return PRED_NONE;
    return i == -1 ? PRED_NONE : PRED_GAP_S;
}
static inline uint8_t init_predc_cols(int mode, int i)
{
    if (mode != ORACLE_GLOBAL) return PRED_NONE;
    return i == -1 ? PRED_NONE : PRED_GAP_Q;
}

/* relax_global / relax_local: src/align.impala:46-79.  Strict '>' gives the
 * tie priority NO_GAP > GAP_Q > GAP_S; local clamps only when score < 0. */
static inline int32_t relax_cell(int local, uint8_t q, uint8_t s,
                                 int32_t no_gap, int32_t gap_q, int32_t gap_s,
                                 int same, int diff, int gap, uint8_t* pred)
{
    int32_t score = no_gap + (q == s ? same : diff);
    uint8_t p = PRED_NO_GAP;
    int32_t qg = gap_q + gap;
    if (qg > score) { score = qg; p = PRED_GAP_Q; }
    int32_t sg = gap_s + gap;
    if (sg > score) { score = sg; p = PRED_GAP_S; }
    if (local && 0 > score) { score = 0; p = PRED_NONE; }
    *pred = p;
    return score;
}

/* reduce_max: src/utils.impala:30-49 -> iteration_reduction
 * src/iteration_cpu.impala:205-250: 64 contiguous chunks, strict '>' inside a
 * chunk and across chunks => lowest index attaining the maximum. */
void oracle_reduce_max(const int32_t* vec, int offset, int length,
                       int32_t* score_out, int32_t* index_out)
{
    enum { NB = 64 };
    int bs = ceil_div(length, NB);
    int32_t ps[NB], pi[NB];
    for (int b = 0; b < NB; ++b) {
        int offs = offset + b * bs;
        int len = imin(bs, length - b * bs);
        int32_t sc = ORACLE_SCORE_MIN, ix = -1;
        for (int i = offs; i < offs + len; ++i) {
            int32_t v = vec[i];
            if (v > sc) { sc = v; ix = i; }
        }
        ps[b] = sc; pi[b] = ix;
    }
    int32_t sc = ps[0], ix = pi[0];
    for (int b = 1; b < NB; ++b)
        if (ps[b] > sc) { sc = ps[b]; ix = pi[b]; }
    *score_out = sc; *index_out = ix;
}

/* ------------------------------------------------------------------------
 * score(): src/align.impala:218-235
 *   storage      create_scoring_matrix_linmem   src/scoring.impala:218-259
 *   loop nest    iteration                      src/iteration_cpu.impala:15-57
 *   block access get_linmem_iteration_acc_device / get_local_...
 *                                               src/scoring_cpu.impala:1-85
 *   result       get_{global,semiglobal,local}_scoring_linmem
 *                                               src/scoring.impala:29-137
 * ---------------------------------------------------------------------- */
oracle_result oracle_score_linear(int mode,
                                  const uint8_t* q, int m,
                                  const uint8_t* s, int n,
                                  int same, int diff, int gap,
                                  int threads, int block_w, int block_h)
{
    const int BW = block_w > 0 ? block_w : ORACLE_BLOCK_W;
    const int BH = block_h > 0 ? block_h : ORACLE_BLOCK_H;
    const int local = (mode == ORACLE_LOCAL);
    oracle_result res = { ORACLE_SCORE_MIN, -1, -1 };
    if (threads < 1) threads = 1;

    const int nbi = ceil_div(m, BH), nbj = ceil_div(n, BW);
    ivec col = ivec_new(m), row = ivec_new(n), cor = ivec_new(imax(nbj - 1, 0));

    /* src/scoring.impala:224-242 */
    V(col, -1) = init_score(mode, gap, n - 1);
    for (int i = 0; i < m; ++i) V(col, i) = init_score(mode, gap, i);
    V(row, -1) = init_score(mode, gap, m - 1);
    for (int j = 0; j < n; ++j) V(row, j) = init_score(mode, gap, j);
    for (int k = 0; k < nbj; ++k) V(cor, k - 1) = init_score(mode, gap, k * BW - 1);

    /* local: per-slot maxima, src/scoring.impala:83-89 ; size = ceil(n/BW)
     * src/scoring_cpu.impala:37 */
    ivec mx = ivec_new(nbj), mxi = ivec_new(nbj), mxj = ivec_new(nbj);
    for (int k = 0; k < nbj; ++k) V(mx, k) = ORACLE_SCORE_MIN;

    const int maxb = imin(nbi, nbj);
    const int diags = (m > 0 && n > 0) ? nbi + nbj - 1 : 0;
    for (int d = 0; d < diags; ++d) {
        const int nb = imin3(d + 1, maxb, diags - d);
        #pragma omp parallel for num_threads(threads) schedule(static)
        for (int dj = 0; dj < nb; ++dj) {
            const int bi = imin(d, nbi - 1) - dj;
            const int bj = imax(d - nbi + 1, 0) + dj;
            const int oi = bi * BH, oj = bj * BW;
            const int h = imin(BH, m - oi), w = imin(BW, n - oj);

            /* block prologue: src/scoring_cpu.impala:11-14 */
            int32_t no_gap = V(cor, bj - 1);
            int32_t gap_q = 0;
            V(cor, bj - 1) = V(col, oi + h - 1);

            int32_t bmax = ORACLE_SCORE_MIN; int bpi = 0, bpj = 0;
            for (int i = 0; i < h; ++i) {
                gap_q = V(col, oi + i);                         /* update_begin_line */
                const uint8_t qc = q[oi + i];
                for (int j = 0; j < w; ++j) {
                    const int32_t gap_s = V(row, oj + j);
                    uint8_t p;
                    const int32_t sc = relax_cell(local, qc, s[oj + j], no_gap, gap_q, gap_s,
                                                  same, diff, gap, &p);
                    no_gap = gap_s;                              /* write(): :20-24 */
                    gap_q = sc;
                    V(row, oj + j) = sc;
                    if (local && sc > bmax) { bmax = sc; bpi = i; bpj = j; }   /* :48-54 */
                }
                no_gap = V(col, oi + i);                         /* update_end_line */
                V(col, oi + i) = V(row, oj + w - 1);
            }
            if (local) {                                         /* block_end: :56-73, slot = block_dia_j */
                if (bmax > V(mx, dj)) {
                    V(mx, dj) = bmax; V(mxi, dj) = bpi + oi; V(mxj, dj) = bpj + oj;
                }
            }
        }
    }

    if (mode == ORACLE_GLOBAL) {                                 /* src/scoring.impala:33-34 */
        res.score = V(col, m - 1); res.pos_i = m - 1; res.pos_j = n - 1;
    } else if (mode == ORACLE_SEMIGLOBAL) {                      /* :46-64 */
        int32_t sc, ix;
        oracle_reduce_max(&V(row, 0), -1, n + 1, &sc, &ix);
        if (sc > res.score) { res.score = sc; res.pos_i = m - 1; res.pos_j = ix; }
        oracle_reduce_max(&V(col, 0), -1, m + 1, &sc, &ix);
        if (sc > res.score) { res.score = sc; res.pos_i = ix; res.pos_j = n - 1; }
    } else {                                                     /* :103-110 */
        int32_t sc, ix;
        oracle_reduce_max(&V(mx, 0), 0, nbj, &sc, &ix);
        res.score = sc;
        if (ix >= 0) { res.pos_i = V(mxi, ix); res.pos_j = V(mxj, ix); }
    }
    ivec_free(col); ivec_free(row); ivec_free(cor);
    ivec_free(mx); ivec_free(mxi); ivec_free(mxj);
    return res;
}

/* ------------------------------------------------------------------------
 * Gotoh affine in the same block wavefront.  BUILD-DEFINED, see header.
 * Extra boundary state: E per row crosses block columns (ecol), F per column
 * crosses block rows (frow).
 * ---------------------------------------------------------------------- */
oracle_result oracle_score_affine(int mode,
                                  const uint8_t* q, int m,
                                  const uint8_t* s, int n,
                                  int same, int diff, int gi, int ge,
                                  int threads, int block_w, int block_h)
{
    const int BW = block_w > 0 ? block_w : ORACLE_BLOCK_W;
    const int BH = block_h > 0 ? block_h : ORACLE_BLOCK_H;
    const int local = (mode == ORACLE_LOCAL);
    const int go = gi + ge;
    oracle_result res = { ORACLE_SCORE_MIN, -1, -1 };
    if (threads < 1) threads = 1;

    const int nbi = ceil_div(m, BH), nbj = ceil_div(n, BW);
    ivec col = ivec_new(m), row = ivec_new(n), cor = ivec_new(imax(nbj - 1, 0));
    ivec ecol = ivec_new(m), frow = ivec_new(n);
    #define AINIT(i) (mode == ORACLE_GLOBAL ? ((i) < 0 ? 0 : gi + ((i) + 1) * ge) : 0)
    V(col, -1) = AINIT(n - 1);
    for (int i = 0; i < m; ++i) { V(col, i) = AINIT(i); V(ecol, i) = NEG_INF; }
    V(row, -1) = AINIT(m - 1);
    for (int j = 0; j < n; ++j) { V(row, j) = AINIT(j); V(frow, j) = NEG_INF; }
    for (int k = 0; k < nbj; ++k) V(cor, k - 1) = AINIT(k * BW - 1);

    ivec mx = ivec_new(nbj), mxi = ivec_new(nbj), mxj = ivec_new(nbj);
    for (int k = 0; k < nbj; ++k) V(mx, k) = ORACLE_SCORE_MIN;

    const int maxb = imin(nbi, nbj);
    const int diags = (m > 0 && n > 0) ? nbi + nbj - 1 : 0;
    for (int d = 0; d < diags; ++d) {
        const int nb = imin3(d + 1, maxb, diags - d);
        #pragma omp parallel for num_threads(threads) schedule(static)
        for (int dj = 0; dj < nb; ++dj) {
            const int bi = imin(d, nbi - 1) - dj;
            const int bj = imax(d - nbi + 1, 0) + dj;
            const int oi = bi * BH, oj = bj * BW;
            const int h = imin(BH, m - oi), w = imin(BW, n - oj);

            int32_t diag = V(cor, bj - 1);
            V(cor, bj - 1) = V(col, oi + h - 1);
            int32_t bmax = ORACLE_SCORE_MIN; int bpi = 0, bpj = 0;
            for (int i = 0; i < h; ++i) {
                int32_t hleft = V(col, oi + i);
                int32_t e = V(ecol, oi + i);
                const uint8_t qc = q[oi + i];
                for (int j = 0; j < w; ++j) {
                    const int32_t up = V(row, oj + j);
                    e = imax(e + ge, hleft + go);
                    const int32_t f = imax(V(frow, oj + j) + ge, up + go);
                    int32_t sc = diag + (qc == s[oj + j] ? same : diff);
                    if (e > sc) sc = e;            /* tie priority diag > E > F */
                    if (f > sc) sc = f;
                    if (local && 0 > sc) sc = 0;
                    diag = up; hleft = sc;
                    V(row, oj + j) = sc; V(frow, oj + j) = f;
                    if (local && sc > bmax) { bmax = sc; bpi = i; bpj = j; }
                }
                diag = V(col, oi + i);
                V(col, oi + i) = hleft;
                V(ecol, oi + i) = e;
            }
            if (local && bmax > V(mx, dj)) {
                V(mx, dj) = bmax; V(mxi, dj) = bpi + oi; V(mxj, dj) = bpj + oj;
            }
        }
    }
    #undef AINIT

    if (mode == ORACLE_GLOBAL) {
        res.score = V(col, m - 1); res.pos_i = m - 1; res.pos_j = n - 1;
    } else if (mode == ORACLE_SEMIGLOBAL) {
        int32_t sc, ix;
        oracle_reduce_max(&V(row, 0), -1, n + 1, &sc, &ix);
        if (sc > res.score) { res.score = sc; res.pos_i = m - 1; res.pos_j = ix; }
        oracle_reduce_max(&V(col, 0), -1, m + 1, &sc, &ix);
        if (sc > res.score) { res.score = sc; res.pos_i = ix; res.pos_j = n - 1; }
    } else {
        int32_t sc, ix;
        oracle_reduce_max(&V(mx, 0), 0, nbj, &sc, &ix);
        res.score = sc;
        if (ix >= 0) { res.pos_i = V(mxi, ix); res.pos_j = V(mxj, ix); }
    }
    ivec_free(col); ivec_free(row); ivec_free(cor); ivec_free(ecol); ivec_free(frow);
    ivec_free(mx); ivec_free(mxi); ivec_free(mxj);
    return res;
}

/* ------------------------------------------------------------------------
 * Independent textbook DPs: the pins for the restatements above.  Straight
 * from the recurrences; no block protocol, no corner
 * vectors, no shared code with the functions above beyond imax().
 * ---------------------------------------------------------------------- */
int32_t textbook_score_linear(int mode, const uint8_t* q, int m, const uint8_t* s, int n,
                              int same, int diff, int gap)
{
    int32_t* prev = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n + 1));
    int32_t* cur = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n + 1));
    const int glob = mode == ORACLE_GLOBAL, loc = mode == ORACLE_LOCAL;
    for (int j = 0; j <= n; ++j) prev[j] = glob ? j * gap : 0;
    int32_t best_local = ORACLE_SCORE_MIN;        /* max over computed cells only */
    int32_t best_lastcol = prev[n];               /* H(-1, n-1) takes part (semiglobal) */
    for (int i = 1; i <= m; ++i) {
        cur[0] = glob ? i * gap : 0;
        for (int j = 1; j <= n; ++j) {
            int32_t v = prev[j - 1] + (q[i - 1] == s[j - 1] ? same : diff);
            v = imax(v, cur[j - 1] + gap);
            v = imax(v, prev[j] + gap);
            if (loc) v = imax(v, 0);
            cur[j] = v;
            if (v > best_local) best_local = v;
        }
        if (cur[n] > best_lastcol) best_lastcol = cur[n];
        int32_t* t = prev; prev = cur; cur = t;
    }
    int32_t r;
    if (glob) r = prev[n];
    else if (loc) r = best_local;
    else {
        r = best_lastcol;
        for (int j = 0; j <= n; ++j) r = imax(r, prev[j]);   /* last row incl. H(m-1,-1) */
    }
    free(prev); free(cur);
    return r;
}

int32_t textbook_score_affine(int mode, const uint8_t* q, int m, const uint8_t* s, int n,
                              int same, int diff, int gi, int ge)
{
    const int glob = mode == ORACLE_GLOBAL, loc = mode == ORACLE_LOCAL;
    size_t sz = sizeof(int32_t) * (size_t)(n + 1);
    int32_t *Hp = (int32_t*)malloc(sz), *Hc = (int32_t*)malloc(sz), *F = (int32_t*)malloc(sz);
    Hp[0] = 0; F[0] = NEG_INF;
    for (int j = 1; j <= n; ++j) { Hp[j] = glob ? gi + j * ge : 0; F[j] = NEG_INF; }
    int32_t best_local = ORACLE_SCORE_MIN, best_lastcol = Hp[n];
    for (int i = 1; i <= m; ++i) {
        Hc[0] = glob ? gi + i * ge : 0;
        int32_t E = NEG_INF;
        for (int j = 1; j <= n; ++j) {
            E = imax(E + ge, Hc[j - 1] + gi + ge);
            F[j] = imax(F[j] + ge, Hp[j] + gi + ge);
            int32_t v = Hp[j - 1] + (q[i - 1] == s[j - 1] ? same : diff);
            v = imax(v, imax(E, F[j]));
            if (loc) v = imax(v, 0);
            Hc[j] = v;
            if (v > best_local) best_local = v;
        }
        if (Hc[n] > best_lastcol) best_lastcol = Hc[n];
        int32_t* t = Hp; Hp = Hc; Hc = t;
    }
    int32_t r;
    if (glob) r = Hp[n];
    else if (loc) r = best_local;
    else {
        r = best_lastcol;
        for (int j = 0; j <= n; ++j) r = imax(r, Hp[j]);
    }
    free(Hp); free(Hc); free(F);
    return r;
}

/* ------------------------------------------------------------------------
 * Splits: src/traceback_lintime.impala:1-42
 * ---------------------------------------------------------------------- */
typedef struct { ivec v; int num_blocks; int bpp; } splits_t;

static void splits_dims(const splits_t* sp, int part, int* off, int* height)
{
    int start = part * sp->bpp - 1;
    int end = imin((part + 1) * sp->bpp - 1, sp->num_blocks - 1);
    *off = V(sp->v, start);
    *height = V(sp->v, end) - *off;
}
static void splits_set(splits_t* sp, int part, int pos)
{
    V(sp->v, part * sp->bpp + sp->bpp / 2 - 1) = pos;
}

/* One Hirschberg level: traceback_lintime_step src/align.impala:273-290
 *   storage   create_scoring_hb_matrix_linmem   src/scoring.impala:261-317
 *   loop nest iteration_partitioned             src/iteration_cpu.impala:59-119
 *   accessor  get_iteration_acc_hb_device       src/scoring_cpu.impala:87-123
 *   reversed  get_sequence_acc_half             src/traceback_lintime.impala:137-148
 *   split     hb_sum                            src/traceback_lintime.impala:44-135
 */
static int lintime_step(int mode, const uint8_t* q, int m, const uint8_t* s, int n,
                        int same, int diff, int gap,
                        int part_width, splits_t* sp, int max_height, int threads)
{
    const int local = (mode == ORACLE_LOCAL);
    const int half = part_width / 2;
    const int num_halfs = (n + half - 1) / part_width * 2;
    const int bw = imin(ORACLE_BLOCK_W, half);
    const int BH = ORACLE_BLOCK_H;

    /* create_scoring_hb_matrix_linmem */
    const int nbj = ceil_div(n, bw);
    ivec colL = ivec_new(m), colR = ivec_new(m), row = ivec_new(n), cor = ivec_new(imax(nbj - 1, 0));
    const int bpp_sc = part_width / bw;
    for (int b = 0; b < nbj; ++b) {
        int part = b / bpp_sc, block = b % bpp_sc, off, ph;
        splits_dims(sp, part, &off, &ph);
        int part_blocks = imin(bpp_sc, nbj - part * bpp_sc);
        for (int i = block; i < ph; i += part_blocks) {
            V(colL, off + i) = init_score(mode, gap, i);
            V(colR, off + i) = init_score(mode, gap, i);
        }
    }
    for (int j = 0; j < n; ++j) V(row, j) = init_score(mode, gap, j % half);
    for (int k = 0; k < nbj; ++k) V(cor, k - 1) = init_score(mode, gap, (k * bw) % half - 1);

    /* iteration_partitioned */
    const int hnbj = half / bw;
    const int hnbi = ceil_div(max_height, BH);
    const int hmax = imin(hnbi, hnbj);
    const int diags = hnbi + hnbj - 1;
    for (int d = 0; d < diags; ++d) {
        const int hnb = imin3(d + 1, hmax, diags - d);
        #pragma omp parallel for num_threads(threads) schedule(static)
        for (int dj = 0; dj < hnb * num_halfs; ++dj) {
            const int hi = dj / hnb;
            const int left = (hi % 2 == 0);
            const int hdj = dj % hnb;
            const int hbi = imin(d, hnbi - 1) - hdj;
            const int hbj = imax(d - hnbi + 1, 0) + hdj;
            const int hoj = hi * half;
            int hoi, hh;
            splits_dims(sp, hi / 2, &hoi, &hh);
            const int oi = hoi + hbi * BH;
            const int oj = hoj + hbj * bw;
            const int hw = imin(half, n - hoj);
            const int height = imin(BH, hh - hbi * BH);
            const int width = imin(bw, n - oj);
            /* sequence accessors: forward base or reversed base */
            const int qbase = left ? hoi + hbi * BH : hoi + hh - hbi * BH - 1;
            const int sbase = left ? hoj + hbj * bw : hoj + hw - hbj * bw - 1;
            const int dir = left ? 1 : -1;
            if (width > 0) {
                ivec col = left ? colL : colR;
                const int bj = oj / bw;
                int32_t no_gap = V(cor, bj - 1);
                int32_t gap_q = 0;
                /* executed by the reference even for height <= 0 (then oi+height-1
                 * is the part's last row or -1): harmless, kept literal. */
                V(cor, bj - 1) = V(col, oi + height - 1);
                for (int i = 0; i < height; ++i) {
                    gap_q = V(col, oi + i);
                    const uint8_t qc = q[qbase + dir * i];
                    for (int j = 0; j < width; ++j) {
                        const int32_t gap_s = V(row, oj + j);
                        uint8_t p;
                        const int32_t sc = relax_cell(local, qc, s[sbase + dir * j],
                                                      no_gap, gap_q, gap_s, same, diff, gap, &p);
                        no_gap = gap_s; gap_q = sc; V(row, oj + j) = sc;
                    }
                    no_gap = V(col, oi + i);
                    V(col, oi + i) = V(row, oj + width - 1);
                }
            }
        }
    }

    /* hb_sum */
    const int parts = num_halfs / 2;
    const int bw2 = imin(ORACLE_BLOCK_W, half * 2);
    const int bpp2 = half * 2 / bw2;
    const int nblk = parts * bpp2;
    int32_t* bmax = (int32_t*)malloc(sizeof(int32_t) * (size_t)imax(nblk, 1));
    int32_t* bind = (int32_t*)malloc(sizeof(int32_t) * (size_t)imax(nblk, 1));
    for (int block = 0; block < nblk; ++block) {
        int part = block / bpp2, pb = block % bpp2, poff, len;
        splits_dims(sp, part, &poff, &len);
        int32_t mxv = ORACLE_SCORE_MIN, idx = -1;
        if (pb == 0 && len > 0) {
            int lhw = half;
            int rhw = imin(half, n - (part * 2 + 1) * half);
            mxv = init_score(mode, gap, lhw - 1) + V(colR, poff + len - 1);
            idx = -1;
            int32_t last = V(colL, poff + len - 1) + init_score(mode, gap, rhw - 1);
            if (last > mxv) { mxv = last; idx = len - 1; }
        }
        for (int i = pb; i < len - 1; i += bpp2) {
            int32_t val = V(colL, poff + i) + V(colR, poff + len - i - 2);
            if (val > mxv) { mxv = val; idx = i; }
        }
        bmax[block] = mxv; bind[block] = idx;
    }
    ivec heights = ivec_new(parts * 2 + 1);
    /* all reads of the splits happen before any write within a level in the
     * reference too: set_split_position touches only mid-part slots, which
     * get_part_dimensions of this level never reads. */
    for (int part = 0; part < parts; ++part) {
        int bo = part * bpp2, off, height;
        splits_dims(sp, part, &off, &height);
        int32_t mxv = bmax[bo], idx = bind[bo];
        for (int i = 1; i < bpp2; ++i)
            if (bmax[bo + i] > mxv) { mxv = bmax[bo + i]; idx = bind[bo + i]; }
        splits_set(sp, part, off + idx + 1);
        V(heights, part * 2) = idx + 1;
        V(heights, part * 2 + 1) = height - idx - 1;
        if (part == parts - 1) V(heights, parts * 2) = m - (off + height);
    }
    int32_t mh, mhi;
    oracle_reduce_max(&V(heights, 0), 0, heights.length, &mh, &mhi);
    free(bmax); free(bind);
    ivec_free(heights); ivec_free(colL); ivec_free(colR); ivec_free(row); ivec_free(cor);
    return mh;
}

/* traceback_offset: src/traceback.impala:47-80.  pred is the blockwise
 * predecessor band viewed at (row offset prow, column offset 0):
 * element (i,j) lives at ((i+prow+1)*mem_w + j + 1)  src/dynprog.impala:117-122 */
static void traceback_block(const uint8_t* pred, int mem_w, int prow,
                            const uint8_t* q, const uint8_t* s, int off_q, int off_s,
                            uint8_t* out_q, uint8_t* out_s, int end_i, int end_j)
{
    #define P(i, j) pred[(size_t)((i) + prow + 1) * (size_t)mem_w + (size_t)((j) + 1)]
    int i = end_i, j = end_j;
    uint8_t p = P(i, j);
    const int ob = off_q + off_s;
    while (p != PRED_NONE) {
        uint8_t sq = GAP_SYM, ss = GAP_SYM;
        int out_pos = i + j + 1;
        if (p == PRED_NO_GAP || p == PRED_GAP_S) { sq = q[off_q + i]; --i; }
        if (p == PRED_NO_GAP || p == PRED_GAP_Q) { ss = s[off_s + j]; --j; }
        out_q[ob + out_pos] = sq;
        out_s[ob + out_pos] = ss;
        p = P(i, j);
    }
    #undef P
}

/* traceback_lintime: src/align.impala:237-271 ; final pass
 * traceback_lintime_trace :292-311 with iteration_blockwise
 * (src/iteration_cpu.impala:121-157), get_iteration_acc_tb_device
 * (src/scoring_cpu.impala:125-157), get_traceback_acc
 * (src/mapping_cpu.impala:67-84), predecessors_blockwise
 * (src/predecessors.impala:36-46), iteration_tb (src/iteration_cpu.impala:159-173). */
int32_t oracle_traceback_lintime(int mode,
                                 const uint8_t* q, int m,
                                 const uint8_t* s, int n,
                                 int same, int diff, int gap,
                                 uint8_t* out_q, uint8_t* out_s,
                                 int32_t* splits_out, int threads)
{
    const int local = (mode == ORACLE_LOCAL);
    const int MPW = ORACLE_MIN_PART_W;
    if (threads < 1) threads = 1;

    /* create_traceback_module: src/traceback.impala:14-23 */
    memset(out_q, EMPTY_SYM, (size_t)(m + n));
    memset(out_s, EMPTY_SYM, (size_t)(m + n));

    int part_width = oracle_next_pow_2(n);
    int max_height = m;

    splits_t sp;
    sp.num_blocks = ceil_div(n, MPW);
    sp.v = ivec_new(sp.num_blocks);
    sp.bpp = part_width / MPW;
    V(sp.v, -1) = 0;                       /* i*num_blocks-1 for i = 0,1 */
    V(sp.v, sp.num_blocks - 1) = m;

    while (part_width > MPW) {
        max_height = lintime_step(mode, q, m, s, n, same, diff, gap, part_width, &sp, max_height, threads);
        part_width /= 2;
        sp.bpp /= 2;
    }

    /* traceback_lintime_trace */
    const int nbj = sp.num_blocks;
    const int mem_w = MPW + 1;
    const size_t mem_h = (size_t)m + (size_t)nbj;           /* (m + nbj - 1) + 1 */
    uint8_t* pred = (uint8_t*)calloc(mem_h * (size_t)mem_w, 1);

    #pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for (int bj = 0; bj < nbj; ++bj) {
        const int oj = bj * MPW;
        int oi, h;
        splits_dims(&sp, bj, &oi, &h);
        const int w = imin(MPW, n - oj);
        const int prow = oi + bj;
        #define P(i, j) pred[(size_t)((i) + prow + 1) * (size_t)mem_w + (size_t)((j) + 1)]
        for (int j = -1; j < w; ++j) P(-1, j) = init_predc_cols(mode, j);
        for (int i = 0; i < h; ++i) P(i, -1) = init_predc_rows(mode, i);
        int32_t rowbuf[ORACLE_MIN_PART_W + 1];
        #define R(j) rowbuf[(j) + 1]
        for (int j = -1; j < w; ++j) R(j) = init_score(mode, gap, j);
        int32_t no_gap = init_score(mode, gap, -1), gap_q = 0;
        for (int i = 0; i < h; ++i) {
            gap_q = init_score(mode, gap, i);
            const uint8_t qc = q[oi + i];
            for (int j = 0; j < w; ++j) {
                const int32_t gap_s = R(j);
                uint8_t p;
                const int32_t sc = relax_cell(local, qc, s[oj + j], no_gap, gap_q, gap_s,
                                              same, diff, gap, &p);
                no_gap = gap_s; gap_q = sc; R(j) = sc;
                P(i, j) = p;
            }
            no_gap = init_score(mode, gap, i);
        }
        #undef R
        #undef P
    }

    /* iteration_tb: one walk per 128-column block, each from its own
     * bottom-right cell (h-1, w-1) until PRED_NONE */
    #pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for (int b = 0; b < nbj; ++b) {
        int oi, h;
        splits_dims(&sp, b, &oi, &h);
        const int oj = b * MPW;
        const int w = imin(MPW, n - oj);
        traceback_block(pred, mem_w, oi + b, q, s, oi, oj, out_q, out_s, h - 1, w - 1);
    }

    if (splits_out)
        for (int k = -1; k < nbj; ++k) splits_out[k + 1] = V(sp.v, k);

    free(pred);
    ivec_free(sp.v);

    /* quirk Q1: score of the never-relaxed scoring object
     * (src/align.impala:244,264 ; src/scoring.impala:33,46-64,87-89,104) */
    if (mode == ORACLE_GLOBAL) return init_score(mode, gap, m - 1);
    if (mode == ORACLE_SEMIGLOBAL) return 0;
    return ORACLE_SCORE_MIN;
}

int64_t oracle_alignment_column_score(const uint8_t* aq, const uint8_t* as, int len,
                                      int same, int diff, int gap)
{
    int64_t t = 0;
    for (int k = 0; k < len; ++k) {
        uint8_t a = aq[k], b = as[k];
        if (a == EMPTY_SYM && b == EMPTY_SYM) continue;
        if (a == GAP_SYM || b == GAP_SYM) t += gap;
        else t += (a == b) ? same : diff;
    }
    return t;
}

uint64_t oracle_fnv1a64(const uint8_t* p, int64_t n)
{
    uint64_t h = 1469598103934665603ULL;
    for (int64_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ULL; }
    return h;
}

/* ========================================================================
 * Affine (Gotoh) linear-space traceback.  BUILD-DEFINED: the reference has no
 * affine path at all (src/align.impala:153-166 is an uncalled stub), so
 * *** parity is unpinned vs the reference ***.  The driver keeps the shape of
 * traceback_lintime (src/align.impala:237-311: Splits over 128-column blocks,
 * halves relaxed forward / on reversed sequences, hb_sum candidate scan order,
 * final blockwise pass + per-block walk); what is new is what Gotoh needs
 * (Myers & Miller 1988, adapted to splitting the SUBJECT):
 *   - every half also reports E (horizontal-gap state) at its last column;
 *   - a split vertex has a type: H (ordinary) or E (a horizontal gap runs
 *     through it); candidates per row i, scanned H first then E, strict '>':
 *         H:  LH(i) + RH(len-i-2)
 *         E:  LE(i) + RE(len-i-2) - gi          (the gap is opened once)
 *   - a block whose start (end) vertex has type E gets a free gap opening on
 *     its top border (on the top border of its reversed problem);
 *   - final blocks keep 4 predecessor bits per cell (H source, E ext, F ext) and
 *     are walked with a 3-state machine; a block with end type E starts the
 *     walk in state E if E(end) - gi > H(end).
 * For the global scheme the result is an optimal Gotoh alignment (tested:
 * column score under affine costs == textbook optimum).  Semiglobal / local
 * keep the reference's fragment semantics (zero borders, walk until NONE).
 * ====================================================================== */
typedef struct { int32_t* H; int32_t* E; } hecol;

/* Gotoh over rows [0,h) x cols [0,w) of (q,s) read with direction dir from the given
 * bases; top border opening cost open_top; writes H and E of the last column per row */
static void affine_half(int mode, const uint8_t* q, int qbase, const uint8_t* s, int sbase, int dir,
                        int h, int w, int same, int diff, int gi, int ge, int open_top,
                        int32_t* outH, int32_t* outE)
{
    const int glob = mode == ORACLE_GLOBAL, loc = mode == ORACLE_LOCAL;
    const int go = gi + ge;
    int32_t* Hrow = (int32_t*)malloc(sizeof(int32_t) * (size_t)(w + 1));
    int32_t* Frow = (int32_t*)malloc(sizeof(int32_t) * (size_t)(w + 1));
    for (int j = 0; j < w; ++j) { Hrow[j] = glob ? open_top + (j + 1) * ge : 0; Frow[j] = NEG_INF; }
    int32_t diag0 = 0;                                   /* H(-1,-1) */
    for (int i = 0; i < h; ++i) {
        int32_t hleft = glob ? gi + (i + 1) * ge : 0;    /* H(i,-1) */
        int32_t e = NEG_INF;
        int32_t diag = diag0;
        diag0 = hleft;
        const uint8_t qc = q[qbase + dir * i];
        for (int j = 0; j < w; ++j) {
            const int32_t up = Hrow[j];
            e = imax(e + ge, hleft + go);
            const int32_t f = imax(Frow[j] + ge, up + go);
            int32_t sc = diag + (qc == s[sbase + dir * j] ? same : diff);
            if (e > sc) sc = e;
            if (f > sc) sc = f;
            if (loc && 0 > sc) sc = 0;
            diag = up; hleft = sc; Hrow[j] = sc; Frow[j] = f;
        }
        outH[i] = hleft;
        outE[i] = e;
    }
    free(Hrow); free(Frow);
}

int32_t oracle_traceback_lintime_affine(int mode,
                                        const uint8_t* q, int m, const uint8_t* s, int n,
                                        int same, int diff, int gi, int ge,
                                        uint8_t* out_q, uint8_t* out_s,
                                        int32_t* splits_out, int32_t* types_out, int threads)
{
    const int glob = mode == ORACLE_GLOBAL, loc = mode == ORACLE_LOCAL;
    const int MPW = ORACLE_MIN_PART_W;
    const int go = gi + ge;
    if (threads < 1) threads = 1;
    memset(out_q, EMPTY_SYM, (size_t)(m + n));
    memset(out_s, EMPTY_SYM, (size_t)(m + n));

    int part_width = oracle_next_pow_2(n);
    splits_t sp;
    sp.num_blocks = ceil_div(n, MPW);
    sp.v = ivec_new(sp.num_blocks);
    sp.bpp = part_width / MPW;
    V(sp.v, -1) = 0;
    V(sp.v, sp.num_blocks - 1) = m;
    ivec vt = ivec_new(sp.num_blocks);           /* vertex types per splits slot: 0 = H, 1 = E */

    int32_t* LH = (int32_t*)malloc(sizeof(int32_t) * (size_t)(m + 1));
    int32_t* LE = (int32_t*)malloc(sizeof(int32_t) * (size_t)(m + 1));
    int32_t* RH = (int32_t*)malloc(sizeof(int32_t) * (size_t)(m + 1));
    int32_t* RE = (int32_t*)malloc(sizeof(int32_t) * (size_t)(m + 1));

    while (part_width > MPW) {
        const int half = part_width / 2;
        const int num_halfs = (n + half - 1) / part_width * 2;
        const int parts = num_halfs / 2;
        #pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
        for (int hh = 0; hh < num_halfs; ++hh) {
            const int p = hh / 2, left = (hh % 2 == 0);
            int off, len;
            splits_dims(&sp, p, &off, &len);
            const int start_idx = p * sp.bpp - 1;
            const int end_idx = imin((p + 1) * sp.bpp - 1, sp.num_blocks - 1);
            const int c_left = 2 * p * half, c_right = c_left + half;
            const int rhw = imin(half, n - c_right);
            if (len <= 0) continue;
            if (left) {
                const int open_top = V(vt, start_idx) ? 0 : gi;
                affine_half(mode, q, off, s, c_left, +1, len, half, same, diff, gi, ge, open_top, LH + off, LE + off);
            } else {
                const int open_top = V(vt, end_idx) ? 0 : gi;
                affine_half(mode, q, off + len - 1, s, c_right + rhw - 1, -1, len, rhw, same, diff, gi, ge, open_top,
                            RH + off, RE + off);
            }
        }
        /* hb_sum with the reference's candidate scan order, two candidates per row (H then E) */
        const int bw2 = imin(ORACLE_BLOCK_W, half * 2);
        const int bpp2 = half * 2 / bw2;
        for (int p = 0; p < parts; ++p) {
            int off, len;
            splits_dims(&sp, p, &off, &len);
            const int start_idx = p * sp.bpp - 1;
            const int end_idx = imin((p + 1) * sp.bpp - 1, sp.num_blocks - 1);
            const int rhw = imin(half, n - (2 * p + 1) * half);
            const int open_l = V(vt, start_idx) ? 0 : gi;       /* top border of the left half */
            const int open_r = V(vt, end_idx) ? 0 : gi;         /* top border of the reversed right half */
            int32_t best = ORACLE_SCORE_MIN, bidx = -1, btype = 0;
            const int32_t* lh = LH + off; const int32_t* le = LE + off;
            const int32_t* rh = RH + off; const int32_t* re = RE + off;
            for (int pb = 0; pb < bpp2; ++pb) {
                int32_t mxv = ORACLE_SCORE_MIN, idx = -1, typ = 0;
                if (pb == 0 && len > 0) {
                    /* all rows to the right half: the left half is one horizontal gap */
                    const int32_t bl = glob ? open_l + half * ge : 0;
                    mxv = bl + rh[len - 1]; idx = -1; typ = 0;
                    if (glob) { int32_t v = bl + re[len - 1] - gi; if (v > mxv) { mxv = v; typ = 1; } }
                    /* all rows to the left half */
                    const int32_t br = glob ? open_r + rhw * ge : 0;
                    int32_t v = lh[len - 1] + br;
                    if (v > mxv) { mxv = v; idx = len - 1; typ = 0; }
                    if (glob) { v = le[len - 1] + br - gi; if (v > mxv) { mxv = v; idx = len - 1; typ = 1; } }
                }
                for (int i = pb; i < len - 1; i += bpp2) {
                    int32_t v = lh[i] + rh[len - i - 2];
                    if (v > mxv) { mxv = v; idx = i; typ = 0; }
                    v = le[i] + re[len - i - 2] - gi;
                    if (v > mxv) { mxv = v; idx = i; typ = 1; }
                }
                if (pb == 0 || mxv > best) { best = mxv; bidx = idx; btype = typ; }
            }
            splits_set(&sp, p, off + bidx + 1);
            V(vt, p * sp.bpp + sp.bpp / 2 - 1) = btype;
        }
        part_width /= 2;
        sp.bpp /= 2;
    }

    /* final pass: every 128-column block, 4 predecessor bits per cell, 3-state walk */
    const int nbj = sp.num_blocks;
    #pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for (int b = 0; b < nbj; ++b) {
        int oi, h;
        splits_dims(&sp, b, &oi, &h);
        const int start_idx = b * sp.bpp - 1;
        const int end_idx = imin((b + 1) * sp.bpp - 1, nbj - 1);
        const int oj = b * MPW;
        const int w = imin(MPW, n - oj);
        const int ts = V(vt, start_idx), te = V(vt, end_idx);
        const int open_top = ts ? 0 : gi;
        uint8_t* pred = (uint8_t*)malloc((size_t)imax(h, 1) * (size_t)w);
        int32_t Hrow[ORACLE_MIN_PART_W], Frow[ORACLE_MIN_PART_W];
        for (int j = 0; j < w; ++j) { Hrow[j] = glob ? open_top + (j + 1) * ge : 0; Frow[j] = NEG_INF; }
        int32_t diag0 = 0, h_end = (w > 0 && h == 0) ? Hrow[w - 1] : 0, e_end = NEG_INF;
        for (int i = 0; i < h; ++i) {
            int32_t hleft = glob ? gi + (i + 1) * ge : 0;
            int32_t e = NEG_INF, diag = diag0;
            diag0 = hleft;
            const uint8_t qc = q[oi + i];
            for (int j = 0; j < w; ++j) {
                const int32_t up = Hrow[j];
                int eext = 0, fext = 0;
                int32_t eo = hleft + go;
                if (e + ge > eo) { eo = e + ge; eext = 1; }
                e = eo;
                int32_t fo = up + go;
                if (Frow[j] + ge > fo) { fo = Frow[j] + ge; fext = 1; }
                int32_t sc = diag + (qc == s[oj + j] ? same : diff);
                int src = PRED_NO_GAP;
                if (e > sc) { sc = e; src = PRED_GAP_Q; }
                if (fo > sc) { sc = fo; src = PRED_GAP_S; }
                if (loc && 0 > sc) { sc = 0; src = PRED_NONE; }
                pred[(size_t)i * w + j] = (uint8_t)(src | (eext << 2) | (fext << 3));
                diag = up; hleft = sc; Hrow[j] = sc; Frow[j] = fo;
            }
            h_end = hleft; e_end = e;
        }
        /* walk */
        int i = h - 1, j = w - 1;
        int state = 0;                                  /* 0 = H, 1 = E, 2 = F */
        if (te && h > 0 && w > 0 && e_end - gi > h_end) state = 1;
        const int ob = oi + oj;
        for (;;) {
            if (i < 0 && j < 0) break;
            if (i < 0) {                                /* top border: horizontal gap (global only) */
                if (!glob) break;
                out_q[ob + i + j + 1] = GAP_SYM; out_s[ob + i + j + 1] = s[oj + j]; --j; continue;
            }
            if (j < 0) {                                /* left border: vertical gap (global only) */
                if (!glob) break;
                out_q[ob + i + j + 1] = q[oi + i]; out_s[ob + i + j + 1] = GAP_SYM; --i; continue;
            }
            const uint8_t p = pred[(size_t)i * w + j];
            if (state == 0) {
                const int src = p & 3;
                if (src == PRED_NONE) break;
                if (src == PRED_NO_GAP) {
                    out_q[ob + i + j + 1] = q[oi + i]; out_s[ob + i + j + 1] = s[oj + j]; --i; --j;
                } else state = (src == PRED_GAP_Q) ? 1 : 2;
            } else if (state == 1) {
                out_q[ob + i + j + 1] = GAP_SYM; out_s[ob + i + j + 1] = s[oj + j];
                state = ((p >> 2) & 1) ? 1 : 0; --j;
            } else {
                out_q[ob + i + j + 1] = q[oi + i]; out_s[ob + i + j + 1] = GAP_SYM;
                state = ((p >> 3) & 1) ? 2 : 0; --i;
            }
        }
        free(pred);
    }
    if (splits_out) for (int k = -1; k < nbj; ++k) splits_out[k + 1] = V(sp.v, k);
    if (types_out) for (int k = -1; k < nbj; ++k) types_out[k + 1] = V(vt, k);
    free(LH); free(LE); free(RH); free(RE);
    ivec_free(sp.v); ivec_free(vt);
    /* the value the reference-shaped driver returns (never-relaxed scoring object, quirk Q1) */
    if (glob) return gi + m * ge;
    if (mode == ORACLE_SEMIGLOBAL) return 0;
    return ORACLE_SCORE_MIN;
}

/* ======================================================================
 * traceback_full: src/align.impala:190-216 -- the whole predecessor matrix
 * (full_predecessors, src/predecessors.impala:11-34: border row/column from
 * init_predc_rows/cols), relaxed with the scheme's relax (the order of the
 * relaxation does not influence the predecessors), then ONE walk
 * (traceback_offset, src/traceback.impala:47-80) from scoring.get_score_pos()
 * (src/scoring.impala:33-34, 46-64, 103-110).  Unlike traceback_lintime the
 * scoring object HAS been relaxed, so the returned value is the real score.
 * Element (i,j), i,j >= -1, lives at (i+1)*(n+1) + (j+1).
 * ====================================================================== */
int32_t oracle_traceback_full(int mode, const uint8_t* q, int m, const uint8_t* s, int n,
                              int same, int diff, int gap, uint8_t* out_q, uint8_t* out_s,
                              int32_t* start_out, int threads)
{
    const int local = mode == ORACLE_LOCAL;
    for (int k = 0; k < m + n; ++k) { out_q[k] = EMPTY_SYM; out_s[k] = EMPTY_SYM; }   /* src/traceback.impala:20-23 */
    const oracle_result r = oracle_score_linear(mode, q, m, s, n, same, diff, gap, threads, 0, 0);
    if (m <= 0 || n <= 0) { if (start_out) { start_out[0] = 0; start_out[1] = 0; } return r.score; }
    const size_t pitch = (size_t)n + 1;
    uint8_t* pred = (uint8_t*)malloc(((size_t)m + 1) * pitch);
    #define PF(i, j) pred[(size_t)((i) + 1) * pitch + (size_t)((j) + 1)]
    for (int i = -1; i < m; ++i) PF(i, -1) = init_predc_rows(mode, i);
    for (int j = 0; j < n; ++j) PF(-1, j) = init_predc_cols(mode, j);
    int32_t* row = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    for (int j = 0; j < n; ++j) row[j] = init_score(mode, gap, j);
    for (int i = 0; i < m; ++i) {
        int32_t no_gap = init_score(mode, gap, i - 1);
        int32_t gap_q = init_score(mode, gap, i);
        for (int j = 0; j < n; ++j) {
            const int32_t gap_s = row[j];
            uint8_t p;
            const int32_t sc = relax_cell(local, q[i], s[j], no_gap, gap_q, gap_s, same, diff, gap, &p);
            PF(i, j) = p;
            no_gap = gap_s; gap_q = sc; row[j] = sc;
        }
    }
    int i = r.pos_i, j = r.pos_j;
    uint8_t p = PF(i, j);
    while (p != PRED_NONE) {
        uint8_t sq = GAP_SYM, ss = GAP_SYM;
        const int out_pos = i + j + 1;
        if (p == PRED_NO_GAP || p == PRED_GAP_S) { sq = q[i]; --i; }
        if (p == PRED_NO_GAP || p == PRED_GAP_Q) { ss = s[j]; --j; }
        out_q[out_pos] = sq; out_s[out_pos] = ss;
        p = PF(i, j);
    }
    #undef PF
    if (start_out) { start_out[0] = i + 1; start_out[1] = j + 1; }      /* get_alignment_start */
    free(row); free(pred);
    return r.score;
}

/* Gotoh full-matrix traceback, BUILD-DEFINED like the other affine entry points
 * (*** parity unpinned vs the reference ***): 4 predecessor bits per cell and the
 * 3-state walk of the final pass of oracle_traceback_lintime_affine, over the whole
 * matrix, started in state H at the end cell of oracle_score_affine.  gi == 0
 * reproduces oracle_traceback_full. */
int32_t oracle_traceback_full_affine(int mode, const uint8_t* q, int m, const uint8_t* s, int n,
                                     int same, int diff, int gi, int ge, uint8_t* out_q, uint8_t* out_s,
                                     int32_t* start_out, int threads)
{
    const int glob = mode == ORACLE_GLOBAL, loc = mode == ORACLE_LOCAL;
    const int go = gi + ge;
    for (int k = 0; k < m + n; ++k) { out_q[k] = EMPTY_SYM; out_s[k] = EMPTY_SYM; }
    const oracle_result r = oracle_score_affine(mode, q, m, s, n, same, diff, gi, ge, threads, 0, 0);
    if (m <= 0 || n <= 0) { if (start_out) { start_out[0] = 0; start_out[1] = 0; } return r.score; }
    uint8_t* pred = (uint8_t*)malloc((size_t)m * (size_t)n);
    int32_t* Hrow = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    int32_t* Frow = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    for (int j = 0; j < n; ++j) { Hrow[j] = glob ? gi + (j + 1) * ge : 0; Frow[j] = NEG_INF; }
    int32_t diag0 = 0;
    for (int i = 0; i < m; ++i) {
        int32_t hleft = glob ? gi + (i + 1) * ge : 0;
        int32_t e = NEG_INF, diag = diag0;
        diag0 = hleft;
        for (int j = 0; j < n; ++j) {
            const int32_t up = Hrow[j];
            int eext = 0, fext = 0;
            int32_t eo = hleft + go;
            if (e + ge > eo) { eo = e + ge; eext = 1; }
            e = eo;
            int32_t fo = up + go;
            if (Frow[j] + ge > fo) { fo = Frow[j] + ge; fext = 1; }
            int32_t sc = diag + (q[i] == s[j] ? same : diff);
            int src = PRED_NO_GAP;
            if (e > sc) { sc = e; src = PRED_GAP_Q; }
            if (fo > sc) { sc = fo; src = PRED_GAP_S; }
            if (loc && 0 > sc) { sc = 0; src = PRED_NONE; }
            pred[(size_t)i * n + j] = (uint8_t)(src | (eext << 2) | (fext << 3));
            diag = up; hleft = sc; Hrow[j] = sc; Frow[j] = fo;
        }
    }
    int i = r.pos_i, j = r.pos_j, state = 0;
    for (;;) {
        if (i < 0 && j < 0) break;
        if (i < 0) { if (!glob) break; out_q[i + j + 1] = GAP_SYM; out_s[i + j + 1] = s[j]; --j; continue; }
        if (j < 0) { if (!glob) break; out_q[i + j + 1] = q[i]; out_s[i + j + 1] = GAP_SYM; --i; continue; }
        const uint8_t p = pred[(size_t)i * n + j];
        if (state == 0) {
            const int src = p & 3;
            if (src == PRED_NONE) break;
            if (src == PRED_NO_GAP) { out_q[i + j + 1] = q[i]; out_s[i + j + 1] = s[j]; --i; --j; }
            else state = (src == PRED_GAP_Q) ? 1 : 2;
        } else if (state == 1) {
            out_q[i + j + 1] = GAP_SYM; out_s[i + j + 1] = s[j];
            state = ((p >> 2) & 1) ? 1 : 0; --j;
        } else {
            out_q[i + j + 1] = q[i]; out_s[i + j + 1] = GAP_SYM;
            state = ((p >> 3) & 1) ? 2 : 0; --i;
        }
    }
    if (start_out) { start_out[0] = i + 1; start_out[1] = j + 1; }
    free(pred); free(Hrow); free(Frow);
    return r.score;
}

/* affine column score of an emitted alignment (test helper): gap runs cost gi + L*ge */
int64_t oracle_alignment_column_score_affine(const uint8_t* aq, const uint8_t* as, int len,
                                             int same, int diff, int gi, int ge)
{
    int64_t t = 0;
    int run = 0;                       /* 0 none, 1 gap in query row, 2 gap in subject row */
    for (int k = 0; k < len; ++k) {
        const uint8_t a = aq[k], b = as[k];
        if (a == EMPTY_SYM && b == EMPTY_SYM) continue;
        if (a == GAP_SYM) { t += ge + (run == 1 ? 0 : gi); run = 1; }
        else if (b == GAP_SYM) { t += ge + (run == 2 ? 0 : gi); run = 2; }
        else { t += (a == b) ? same : diff; run = 0; }
    }
    return t;
}

/* ========================================================================
 * Batches of independent pairs (BASELINE configs[3]): every pair goes through
 * oracle_score_affine / oracle_score_linear above with a team of ONE thread
 * (a 150 x 500 pair is a single 1024 x 1024 block anyway); the pairs are
 * spread over `threads` host threads.  scores[p] = score of pair p.
 * ====================================================================== */
void oracle_score_batch(int mode, const uint8_t* q, const int64_t* qoff, const uint8_t* s, const int64_t* soff,
                        int64_t npairs, int same, int diff, int gi, int ge, int threads, int32_t* scores)
{
    if (threads < 1) threads = 1;
    #pragma omp parallel for num_threads(threads) schedule(dynamic, 64)
    for (int64_t p = 0; p < npairs; ++p) {
        const int m = (int)(qoff[p + 1] - qoff[p]), n = (int)(soff[p + 1] - soff[p]);
        oracle_result r;
        if (m == 0 || n == 0) {
            /* no block ever runs (quirk Q12): global leaves the init value, semiglobal 0, local SCORE_MIN */
            const int L = m > n ? m : n;
            r.score = mode == ORACLE_GLOBAL ? (L > 0 ? gi + L * ge : 0) : (mode == ORACLE_SEMIGLOBAL ? 0 : ORACLE_SCORE_MIN);
        } else if (gi != 0) {
            r = oracle_score_affine(mode, q + qoff[p], m, s + soff[p], n, same, diff, gi, ge, 1, 0, 0);
        } else {
            r = oracle_score_linear(mode, q + qoff[p], m, s + soff[p], n, same, diff, ge, 1, 0, 0);
        }
        scores[p] = r.score;
    }
}
