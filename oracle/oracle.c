/*
 * oracle.c — CPU restatement of CONSEL's multiscale RELL bootstrap hot path.
 *
 * THIS FILE IS TEST INFRASTRUCTURE.  It is the checker for the CUDA path and is
 * never linked into, imported by or executed from the product (consel_b200/,
 * include/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function below
 * against (a) the committed golden vectors under tests/golden/ that were produced
 * by the unmodified reference binaries (oracle/gen_golden.py) and (b), when
 * oracle/_ref/ holds the compiled reference, against fresh runs of it.
 *
 * Each function cites the reference file:line (relative to /root/reference) whose
 * arithmetic it restates.  Operation ORDER is part of the contract: the replicate
 * sum is a sequential j=0..n-1 loop of a rounded product followed by a rounded
 * add and a final true division (src/makermt.c:65-70); build with
 * -ffp-contract=off so no FMA can merge the two roundings.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_HUGE 1.0e30 /* src/consel.c:155 (sentinel used as +-infinity) */

/* ------------------------------------------------------------------ MT19937 */
/* src/mt19937.c:57-60 (state), :63-76 (seeding), :108-127 (block regeneration),
 * :182-197 (tempering and the [0,1) conversion).                              */
typedef struct {
    uint32_t x[624];
    int pos; /* next word to temper; 624 = regenerate first */
} orc_mt;

void orc_mt_seed(orc_mt *g, uint64_t seed) {
    g->x[0] = (uint32_t)(seed & 0xffffffffu);
    for (int j = 1; j < 624; j++) {
        uint32_t p = g->x[j - 1];
        g->x[j] = 1812433253u * (p ^ (p >> 30)) + (uint32_t)j;
    }
    g->pos = 624;
}

static inline uint32_t orc_twist(uint32_t hi, uint32_t lo) {
    uint32_t y = (hi & 0x80000000u) | (lo & 0x7fffffffu);
    return (y >> 1) ^ ((lo & 1u) ? 0x9908b0dfu : 0u);
}

static void orc_mt_regen(orc_mt *g) {
    uint32_t *x = g->x;
    int k;
    for (k = 0; k < 624 - 397; k++) x[k] = x[k + 397] ^ orc_twist(x[k], x[k + 1]);
    for (; k < 623; k++) x[k] = x[k + 397 - 624] ^ orc_twist(x[k], x[k + 1]);
    x[623] = x[396] ^ orc_twist(x[623], x[0]);
    g->pos = 0;
}

uint32_t orc_mt_u32(orc_mt *g) {
    if (g->pos >= 624) orc_mt_regen(g);
    uint32_t y = g->x[g->pos++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* skip `skip` draws, then write `count` tempered 32-bit outputs */
void orc_mt_stream(uint64_t seed, uint64_t skip, uint64_t count, uint32_t *out) {
    orc_mt g;
    orc_mt_seed(&g, seed);
    for (uint64_t i = 0; i < skip; i++) (void)orc_mt_u32(&g);
    for (uint64_t i = 0; i < count; i++) out[i] = orc_mt_u32(&g);
}

/* raw state access so tests can check the device jump-ahead:
 * state after `blocks` block regenerations from the seed (pre-tempering words). */
void orc_mt_state_after_blocks(uint64_t seed, uint64_t blocks, uint32_t *state624) {
    orc_mt g;
    orc_mt_seed(&g, seed);
    for (uint64_t b = 0; b < blocks; b++) orc_mt_regen(&g);
    memcpy(state624, g.x, sizeof(g.x));
}

/* site index of one draw: src/mt19937.c:195 then src/makermt.c:55-60 */
static inline int orc_site(uint32_t y, int n) {
    double rv = (double)y * (1.0 / 4294967296.0);
    int k = (int)(rv * n);
    if (k < 0) k = 0;
    else if (k >= n) k = n - 1;
    return k;
}

/* ------------------------------------------------------------- scaleboot */
/* src/makermt.c:35-75.  L is K x n row-major (datmat[0] of new_mat).  rep is
 * K x bb row-major (tree-major, replicate fastest) — the byte order fwrite_bmat
 * emits.  counts (optional, may be NULL) receives the bb x n site-count matrix. */
void orc_scaleboot(const double *L, int K, int n, int bn, int bb, orc_mt *g,
                   double *rep, int32_t *counts) {
    double *wv = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    double r = (double)bn / (double)n;
    for (int b = 0; b < bb; b++) {
        for (int j = 0; j < n; j++) wv[j] = 0.0;
        for (int j = 0; j < bn; j++) wv[orc_site(orc_mt_u32(g), n)] += 1.0;
        if (counts)
            for (int j = 0; j < n; j++) counts[(size_t)b * n + j] = (int32_t)wv[j];
        for (int k = 0; k < K; k++) {
            const double *xp = L + (size_t)k * n;
            double x = 0.0;
            for (int j = 0; j < n; j++) x += wv[j] * xp[j];
            rep[(size_t)k * bb + b] = x / r;
        }
    }
    free(wv);
}

/* same contraction but driven by a caller-supplied count matrix (any RNG) */
void orc_contract_counts(const double *L, int K, int n, const int32_t *counts,
                         int bb, double r, double *rep) {
    for (int b = 0; b < bb; b++) {
        const int32_t *c = counts + (size_t)b * n;
        for (int k = 0; k < K; k++) {
            const double *xp = L + (size_t)k * n;
            double x = 0.0;
            for (int j = 0; j < n; j++) x += (double)c[j] * xp[j];
            rep[(size_t)k * bb + b] = x / r;
        }
    }
}

/* genrmt scale setup: src/makermt.c:159-168 */
void orc_scale_setup(const double *L, int K, int n, int S, const double *rr,
                     double *datvec, int *bn, double *rr1) {
    for (int i = 0; i < K; i++) {
        double x = 0;
        for (int j = 0; j < n; j++) x += L[(size_t)i * n + j];
        datvec[i] = x;
    }
    for (int s = 0; s < S; s++) {
        bn[s] = (int)(rr[s] * n);
        rr1[s] = (double)bn[s] / n;
    }
}

/* whole makermt resampling loop: src/makermt.c:201-211 with one global stream.
 * rep_all is the concatenation over scales of the K x bb[s] blocks.           */
void orc_makermt(const double *L, int K, int n, int S, const int *bn,
                 const int *bb, uint64_t seed, double *rep_all) {
    orc_mt g;
    orc_mt_seed(&g, seed);
    size_t off = 0;
    for (int s = 0; s < S; s++) {
        orc_scaleboot(L, K, n, bn[s], bb[s], &g, rep_all + off, NULL);
        off += (size_t)K * bb[s];
    }
}

/* ---------------------------------------------------- replicate statistics */
/* src/consel.c:1894-1916 calcmaxs: out[i] = max_{j != i} x[j] - x[i]          */
static void orc_calcmaxs(double *xx, int m, double *yy) {
    double x = -ORC_HUGE;
    for (int i = 0; i < m; i++) {
        yy[i] = x;
        if (xx[i] > x) x = xx[i];
    }
    x = -ORC_HUGE;
    for (int i = m - 1; i >= 0; i--) {
        if (x > yy[i]) yy[i] = x;
        if (xx[i] > x) x = xx[i];
    }
    for (int i = 0; i < m; i++) xx[i] = yy[i] - xx[i];
}

/* src/consel.c:1939-1959 repminmaxs (+ calcassmins :1923-1937).  Associations in
 * CSR form: item e owns trees ass_idx[ass_ptr[e] .. ass_ptr[e+1]).
 * rep: K x B row-major; stat: cm x B row-major.                               */
void orc_repminmaxs(const double *rep, int K, int B, int cm, const int *ass_ptr,
                    const int *ass_idx, double *stat) {
    double *buf = (double *)malloc(sizeof(double) * (size_t)K * 2);
    double *yy = buf + K;
    for (int b = 0; b < B; b++) {
        for (int j = 0; j < K; j++) buf[j] = rep[(size_t)j * B + b];
        orc_calcmaxs(buf, K, yy);
        for (int e = 0; e < cm; e++) {
            double y = ORC_HUGE;
            for (int q = ass_ptr[e]; q < ass_ptr[e + 1]; q++) {
                double x = buf[ass_idx[q]];
                if (x < y) y = x;
            }
            stat[(size_t)e * B + b] = y;
        }
    }
    free(buf);
}

static int orc_cmp_double(const void *a, const void *b) {
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}
/* src/misc.c:819-826 sort_vec: ascending sort of values (result is unique) */
void orc_sort_rows(double *mat, int rows, int B) {
    for (int i = 0; i < rows; i++) qsort(mat + (size_t)i * B, B, sizeof(double), orc_cmp_double);
}

/* src/consel.c:1796-1801 cntdist1, :1806-1821 cntdist2, :1842-1860 cntdist3 */
static int orc_cnt_le_sorted(const double *v, int bb, double t) {
    if (t < v[0]) return 0;
    if (v[bb - 1] <= t) return bb;
    int i0 = 0, i1 = bb - 1;
    while (i1 - i0 > 1) {
        int i = (i0 + i1) / 2;
        if (v[i] <= t) i0 = i; else i1 = i;
    }
    return i1;
}
double orc_cntdist(const double *v, int bb, double t, int mode) {
    if (mode == 1) {
        int c = 0;
        for (int i = 0; i < bb; i++) if (v[i] <= t) c++;
        return (double)c;
    }
    if (mode == 2) return (double)orc_cnt_le_sorted(v, bb, t);
    int i = orc_cnt_le_sorted(v, bb, t) - 1;
    double n = (double)bb, p;
    if (i < 0) {
        if (v[1] > v[0]) p = 0.5 + (t - v[0]) / (v[1] - v[0]);
        else p = 0.0;
    } else if (i < bb - 1) {
        if (v[i + 1] > v[i]) p = 0.5 + (double)i + (t - v[i]) / (v[i + 1] - v[i]);
        else p = 0.5 + (double)i;
    } else {
        if (v[bb - 1] - v[bb - 2] > 0) p = n - 0.5 + (t - v[bb - 1]) / (v[bb - 1] - v[bb - 2]);
        else p = n;
    }
    if (p > n) p = n; else if (p < 0.0) p = 0.0;
    return p;
}
/* src/consel.c:1876-1884 pvaldist with smoothing (sorts a private copy) */
static double orc_pvaldist(double *v, int bb, double t) {
    qsort(v, bb, sizeof(double), orc_cmp_double);
    return 1.0 - orc_cntdist(v, bb, t, 3) / bb;
}

/* src/consel.c:2010-2033 calcmaxass */
static void orc_calcmaxass(const double *xx, const double *wt, int m, const int *ass,
                           const int *cass, int n, double *out) {
    if (wt) {
        for (int i = 0; i < n; i++) {
            int ia = ass[i];
            double x0 = xx[ia], x = -ORC_HUGE;
            const double *wp = wt + (size_t)ia * m;
            for (int j = 0; j < m - n; j++) {
                int ja = cass[j];
                double y = wp[ja] * (xx[ja] - x0);
                if (y > x) x = y;
            }
            out[i] = x;
        }
    } else {
        double x = -ORC_HUGE;
        for (int j = 0; j < m - n; j++) {
            double y = xx[cass[j]];
            if (y > x) x = y;
        }
        for (int i = 0; i < n; i++) out[i] = x - xx[ass[i]];
    }
}

static void orc_complement(const int *ass, int n, int m, int *cass) {
    char *mark = (char *)calloc((size_t)m, 1);
    for (int i = 0; i < n; i++) mark[ass[i]] = 1;
    int j = 0;
    for (int i = 0; i < m; i++) if (!mark[i]) cass[j++] = i;
    free(mark);
}

/* observed statistic per association and the item order: src/consel.c:499-517
 * (calcmaxass on datvec, ascending sort of each vector, then ordering by
 * compdvec :1986-2000).  Writes obs_sorted_min[e] = min statistic of item e in
 * INPUT order; ordering itself is done by the caller (ties are undefined in the
 * reference's unstable quicksort and are avoided in the fixtures).             */
void orc_obs_stat(const double *datvec, int K, int cm, const int *ass_ptr,
                  const int *ass_idx, double *obs) {
    int *cass = (int *)malloc(sizeof(int) * (size_t)K);
    double *out = (double *)malloc(sizeof(double) * (size_t)K);
    for (int e = 0; e < cm; e++) {
        const int *ass = ass_idx + ass_ptr[e];
        int n = ass_ptr[e + 1] - ass_ptr[e];
        orc_complement(ass, n, K, cass);
        orc_calcmaxass(datvec, NULL, K, ass, cass, n, out);
        double y = out[0];
        for (int i = 1; i < n; i++) if (out[i] < y) y = out[i];
        obs[e] = y;
    }
    free(cass); free(out);
}

/* do_mctest: src/consel.c:1352-1439 with calckhpval :2111-2149 and calcmcpval
 * :2065-2102.  rep0 = K x B block of the scale nearest r=1 (row-major), NOT
 * modified.  Outputs cm-vectors kh, sh (mc), wkh, wsh (mcw).                   */
void orc_mctest(const double *rep0, const double *datvec, int K, int B, int cm,
                const int *ass_ptr, const int *ass_idx, double varadd,
                double *kh, double *sh, double *wkh, double *wsh) {
    size_t KB = (size_t)K * B;
    double *xc = (double *)malloc(sizeof(double) * KB);
    memcpy(xc, rep0, sizeof(double) * KB);
    for (int j = 0; j < K; j++) { /* centring :1378-1382 */
        double x = 0.0;
        for (int i = 0; i < B; i++) x += xc[(size_t)j * B + i];
        x = x / B;
        for (int i = 0; i < B; i++) xc[(size_t)j * B + i] -= x;
    }
    double *wt = (double *)calloc((size_t)K * K, sizeof(double));
    for (int i = 0; i < K; i++) { /* :1404-1415 */
        const double *xp = xc + (size_t)i * B;
        for (int j = 0; j < i; j++) {
            const double *yp = xc + (size_t)j * B;
            double x = 0.0;
            for (int ib = 0; ib < B; ib++) x += (xp[ib] - yp[ib]) * (xp[ib] - yp[ib]);
            wt[(size_t)i * K + j] = wt[(size_t)j * K + i] = 1.0 / sqrt((x + varadd) / (B - 1));
        }
    }
    int *cass = (int *)malloc(sizeof(int) * (size_t)K);
    double *buf = (double *)malloc(sizeof(double) * (size_t)K);
    double *stat = (double *)malloc(sizeof(double) * (size_t)K);
    double *repvec = (double *)malloc(sizeof(double) * (size_t)B);
    for (int pass = 0; pass < 2; pass++) {
        const double *w = pass ? wt : NULL;
        double *khout = pass ? wkh : kh, *mcout = pass ? wsh : sh;
        for (int e = 0; e < cm; e++) {
            const int *ass = ass_idx + ass_ptr[e];
            int alen = ass_ptr[e + 1] - ass_ptr[e];
            orc_complement(ass, alen, K, cass);
            /* ---- KH: pick the effective contrast :2121-2143 */
            int i0 = 0, j0 = 0, jj = 0;
            double x, y, x0;
            if (w) {
                x0 = ORC_HUGE;
                for (int i = 0; i < alen; i++) {
                    int ia = ass[i];
                    x = -ORC_HUGE;
                    for (int j = 0; j < K - alen; j++) {
                        int ja = cass[j];
                        y = w[(size_t)ia * K + ja] * (datvec[ja] - datvec[ia]);
                        if (y > x) { x = y; jj = j; }
                    }
                    if (x < x0) { x0 = x; i0 = i; j0 = jj; }
                }
            } else {
                x = -ORC_HUGE;
                for (int j = 0; j < K - alen; j++) {
                    y = datvec[cass[j]];
                    if (y > x) { x = y; j0 = j; }
                }
                x = ORC_HUGE;
                for (int i = 0; i < alen; i++) {
                    y = -datvec[ass[i]];
                    if (y < x) { x = y; i0 = i; }
                }
            }
            int ja = cass[j0], ia = ass[i0];
            x0 = datvec[ja] - datvec[ia];
            for (int i = 0; i < B; i++) repvec[i] = xc[(size_t)ja * B + i] - xc[(size_t)ia * B + i];
            khout[e] = orc_pvaldist(repvec, B, x0);
            /* ---- SH / wSH :2089-2101 */
            double *pmat = (double *)malloc(sizeof(double) * (size_t)alen * B);
            for (int i = 0; i < B; i++) {
                for (int j = 0; j < K; j++) buf[j] = xc[(size_t)j * B + i];
                orc_calcmaxass(buf, w, K, ass, cass, alen, stat);
                for (int j = 0; j < alen; j++) pmat[(size_t)j * B + i] = stat[j];
            }
            orc_calcmaxass(datvec, w, K, ass, cass, alen, stat);
            double best = 0.0;
            for (int i = 0; i < alen; i++) {
                double p = orc_pvaldist(pmat + (size_t)i * B, B, stat[i]);
                if (p > best) best = p;
            }
            mcout[e] = best;
            free(pmat);
        }
    }
    free(cass); free(buf); free(stat); free(repvec); free(wt); free(xc);
}

/* do_bptest: src/consel.c:1286-1303 on an already sorted stat row */
double orc_bp(const double *sorted_row, int B, double threshold) {
    return 1.0 - (1.0 - orc_cntdist(sorted_row, B, threshold, 3) / B);
}
