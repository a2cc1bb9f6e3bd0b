/*
 * c3poa_oracle.c -- CPU restatement (oracle) of C3POa's per-read consensus path.
 *
 * TEST INFRASTRUCTURE ONLY -- see c3poa_oracle.h for who may use this and for
 * the parity status of each stage (stage 2/3a pinned by the reference's own
 * Python; conk and abPOA 1.0.5 are "parity unpinned": neither dependency is
 * vendored in /root/reference, so they are restated from their call contract
 * and published algorithm).
 *
 * Reference lines followed (all under /root/reference):
 *   C3POa.py:110-165          analyze_reads (driver of the three stages)
 *   C3POa.py:106-108          rounding()  (banker's rounding to base 50)
 *   bin/savitzky_golay.py:7-38
 *   bin/call_peaks.py:8-16    (+ scipy.signal.find_peaks semantics)
 *   bin/determine_consensus.py:30-47   pyabpoa.msa_aligner(match=5).msa(...)
 */
#include "c3poa_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>

/* ===================================================================== */
/* stage 1: conk                                                         */
/* ===================================================================== */
static inline int base_code(char c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
    }
}

/* Call contract C3POa.py:123: scores = conk.conk(splint, seq, penalty); index d
 * of the result is a read offset (C3POa.py:127-130 adds len(splint)//2 and
 * discards >= len(seq)), so the profile has one entry per diagonal d=j-i>=0. */
int c3o_conk(const char *splint, int ls, const char *seq, int lr, int penalty,
             int32_t *out)
{
    if (ls <= 0 || lr <= 0) return -1;
    int32_t *prev = (int32_t *)calloc((size_t)lr + 1, sizeof(int32_t));
    int32_t *cur = (int32_t *)calloc((size_t)lr + 1, sizeof(int32_t));
    uint8_t *sq = (uint8_t *)malloc((size_t)lr);
    if (!prev || !cur || !sq) { free(prev); free(cur); free(sq); return -2; }
    for (int j = 0; j < lr; ++j) sq[j] = (uint8_t)base_code(seq[j]);
    memset(out, 0, (size_t)lr * sizeof(int32_t));
    for (int i = 0; i < ls; ++i) {
        int a = base_code(splint[i]);
        cur[0] = 0;
        for (int j = 0; j < lr; ++j) {
            int s = (a == sq[j] && a < 4) ? C3O_CONK_MATCH : C3O_CONK_MISMATCH;
            int h = prev[j] + s;                 /* H[i-1][j-1] + s */
            int u = prev[j + 1] - penalty;       /* H[i-1][j] - p   */
            int l = cur[j] - penalty;            /* H[i][j-1] - p   */
            if (u > h) h = u;
            if (l > h) h = l;
            if (h < 0) h = 0;
            cur[j + 1] = h;
            int d = j - i;
            if (d >= 0) out[d] += h;
        }
        int32_t *t = prev; prev = cur; cur = t;
    }
    free(prev); free(cur); free(sq);
    return 0;
}

/* ===================================================================== */
/* stage 2: savitzky_golay + call_peaks                                   */
/* ===================================================================== */
/* bin/savitzky_golay.py:33-36.  Products and sums are kept as separate IEEE
 * operations (volatile-free: the file is compiled with -ffp-contract=off). */
int c3o_savgol(const double *y, int n, const double *coef, int window, double *out)
{
    int half = (window - 1) / 2;
    if (n < half + 1 || window < 1 || (window & 1) == 0) return -1;
    double *yp = (double *)malloc(((size_t)n + 2 * (size_t)half) * sizeof(double));
    if (!yp) return -2;
    /* firstvals = y[0] - abs(y[1:half+1][::-1] - y[0]) */
    for (int k = 0; k < half; ++k) yp[k] = y[0] - fabs(y[half - k] - y[0]);
    memcpy(yp + half, y, (size_t)n * sizeof(double));
    /* lastvals = y[-1] + abs(y[-half-1:-1][::-1] - y[-1]) */
    for (int k = 0; k < half; ++k) yp[half + n + k] = y[n - 1] + fabs(y[n - 2 - k] - y[n - 1]);
    for (int t = 0; t < n; ++t) {
        double acc = 0.0;
        for (int k = 0; k < window; ++k) {
            double p = coef[k] * yp[t + k];
            acc = acc + p;
        }
        out[t] = acc;
    }
    free(yp);
    return 0;
}

static int cmp_double(const void *a, const void *b)
{
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

typedef struct { double pr; int32_t pos; int32_t idx; } cand_t;
static int cmp_cand(const void *a, const void *b)
{
    const cand_t *x = (const cand_t *)a, *y = (const cand_t *)b;
    if (x->pr < y->pr) return -1;
    if (x->pr > y->pr) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);   /* stable: ties keep index order */
}

int c3o_call_peaks(const int32_t *scores, int n, int min_dist, int iters,
                   const double *coef, int window, double *smoothed_out,
                   double *median_out, int32_t *peaks_out, int max_peaks)
{
    double *a = (double *)malloc((size_t)n * sizeof(double));
    double *b = (double *)malloc((size_t)n * sizeof(double));
    if (!a || !b) { free(a); free(b); return -2; }
    for (int i = 0; i < n; ++i) a[i] = (double)scores[i];
    for (int it = 0; it < iters; ++it) {
        int rc = c3o_savgol(a, n, coef, window, b);
        if (rc) { free(a); free(b); return rc; }
        double *t = a; a = b; b = t;
    }
    if (smoothed_out) memcpy(smoothed_out, a, (size_t)n * sizeof(double));
    /* med_score = np.median(scores) */
    memcpy(b, a, (size_t)n * sizeof(double));
    qsort(b, (size_t)n, sizeof(double), cmp_double);
    double med = (n & 1) ? b[n / 2] : (b[n / 2 - 1] + b[n / 2]) / 2.0;
    if (median_out) *median_out = med;
    double mx = a[0];
    for (int i = 1; i < n; ++i) if (a[i] > mx) mx = a[i];
    int np_ = 0;
    if (!(mx < 6 * med)) {
        /* scipy find_peaks(x, distance=min_dist, height=3*med):
         * _local_maxima_1d -> height (inclusive) -> _select_by_peak_distance */
        double hmin = med * 3;
        int cap = 64, nc = 0;
        cand_t *c = (cand_t *)malloc((size_t)cap * sizeof(cand_t));
        int i = 1, i_max = n - 1;
        while (i < i_max) {
            if (a[i - 1] < a[i]) {
                int ia = i + 1;
                while (ia < i_max && a[ia] == a[i]) ++ia;
                if (a[ia] < a[i]) {
                    int mid = (i + (ia - 1)) / 2;
                    if (hmin <= a[mid]) {
                        if (nc == cap) { cap *= 2; c = (cand_t *)realloc(c, (size_t)cap * sizeof(cand_t)); }
                        c[nc].pr = a[mid]; c[nc].pos = mid; c[nc].idx = nc; ++nc;
                    }
                    i = ia;
                }
            }
            ++i;
        }
        /* distance: ceil(distance); highest priority first; on equal priority the
         * later peak is visited first (stable ascending argsort walked from the top) */
        uint8_t *keep = (uint8_t *)malloc((size_t)(nc > 0 ? nc : 1));
        memset(keep, 1, (size_t)(nc > 0 ? nc : 1));
        cand_t *s = (cand_t *)malloc((size_t)(nc > 0 ? nc : 1) * sizeof(cand_t));
        memcpy(s, c, (size_t)nc * sizeof(cand_t));
        qsort(s, (size_t)nc, sizeof(cand_t), cmp_cand);
        for (int t = nc - 1; t >= 0; --t) {
            int j = s[t].idx;
            if (!keep[j]) continue;
            int k = j - 1;
            while (k >= 0 && c[j].pos - c[k].pos < min_dist) { keep[k] = 0; --k; }
            k = j + 1;
            while (k < nc && c[k].pos - c[j].pos < min_dist) { keep[k] = 0; ++k; }
        }
        for (int t = 0; t < nc; ++t)
            if (keep[t]) { if (np_ < max_peaks) peaks_out[np_] = c[t].pos; ++np_; }
        free(keep); free(s); free(c);
    }
    free(a); free(b);
    return np_ > max_peaks ? -3 : np_;
}

/* ===================================================================== */
/* stage 3a: split                                                        */
/* ===================================================================== */
static int cmp_int(const void *a, const void *b)
{
    int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}

int c3o_split(int32_t *peaks, int *n_peaks_io, int ls, int lr,
              int32_t *sub_bounds, int *n_sub, int32_t *dang_bounds, int *n_dang)
{
    int np_ = *n_peaks_io, k = 0;
    *n_sub = 0; *n_dang = 0;
    if (np_ == 0) return 1;                               /* C3POa.py:125-126 */
    for (int i = 0; i < np_; ++i) {                       /* C3POa.py:127-130 */
        int p = peaks[i] + ls / 2;
        if (p < lr) peaks[k++] = p;
    }
    *n_peaks_io = np_ = k;
    if (np_ == 0) return 1;                               /* C3POa.py:131-132 */
    if (np_ > 1) {
        int nl = np_ - 1;
        int *r = (int *)malloc((size_t)nl * sizeof(int));
        int *srt = (int *)malloc((size_t)nl * sizeof(int));
        for (int i = 0; i < nl; ++i) {
            double x = (double)(peaks[i + 1] - peaks[i]);
            r[i] = (int)(50 * rint(x / 50));              /* rounding(): Python round = half-to-even */
            srt[i] = r[i];
        }
        qsort(srt, (size_t)nl, sizeof(int), cmp_int);
        double med = (nl & 1) ? (double)srt[nl / 2]
                              : ((double)srt[nl / 2 - 1] + (double)srt[nl / 2]) / 2.0;
        double lo = med * 0.8, hi = med * 1.2;
        for (int i = 0; i < nl; ++i) {
            if (lo <= (double)r[i] && (double)r[i] <= hi) {
                sub_bounds[2 * *n_sub] = peaks[i];
                sub_bounds[2 * *n_sub + 1] = peaks[i + 1];
                ++*n_sub;
            }
        }
        if (peaks[0] > 100) { dang_bounds[0] = 0; dang_bounds[1] = peaks[0]; *n_dang = 1; }
        if (lr - peaks[np_ - 1] > 100) {
            dang_bounds[2 * *n_dang] = peaks[np_ - 1];
            dang_bounds[2 * *n_dang + 1] = lr;
            ++*n_dang;
        }
        free(r); free(srt);
    } else {
        dang_bounds[0] = 0; dang_bounds[1] = peaks[0];
        dang_bounds[2] = peaks[0]; dang_bounds[3] = lr;
        *n_dang = 2;
    }
    return 0;
}

/* ===================================================================== */
/* stage 3b: abPOA 1.0.5 restatement                                      */
/* ===================================================================== */
#define NEG_INF   (-(1 << 29))
#define NEG_HALF  (-(1 << 28))
#define SRC_ID  0
#define SINK_ID 1
#define ALPHA_M 5            /* abpt->m: A,C,G,T,N */

/* backtrack op masks (state machine of the convex-gap backtrack) */
#define OP_M   0x1
#define OP_E1  0x2
#define OP_E2  0x4
#define OP_E   0x6
#define OP_F1  0x8
#define OP_F2  0x10
#define OP_F   0x18
#define OP_ALL 0x1f

#define CG_MATCH 0
#define CG_INS   1
#define CG_DEL   2

void c3o_poa_default_para(c3o_poa_para_t *p)
{
    p->match = 5; p->mismatch = 4;
    p->gap_open1 = 4; p->gap_ext1 = 2; p->gap_open2 = 24; p->gap_ext2 = 1;
    p->wb = 10; p->wf = 0.01; p->simd_bits = 256;
    p->int8_lanes = 0; p->end_clamp = 0;
}

typedef struct {
    int in_n, in_m, *in_id;
    int out_n, out_m, *out_id, *out_w;
    uint64_t *out_reads;        /* read-id bitset per out edge (n_seq <= 64 tracked) */
    int aln_n, aln_id[ALPHA_M];
    int max_out_id;
    uint8_t base;
} pnode_t;

typedef struct {
    pnode_t *node; int node_n, node_m;
    int *index_to_node, *node_to_index, *max_pos_left, *max_pos_right, *max_remain, *msa_rank;
    int aux_m;
    int sorted;
} pgraph_t;

typedef struct { int op, node_id, qpos; } cg_t;     /* one op per base (INS runs expanded) */

static int g_add_node(pgraph_t *g, uint8_t base)
{
    if (g->node_n == g->node_m) {
        int m = g->node_m ? g->node_m * 2 : 1024;
        g->node = (pnode_t *)realloc(g->node, (size_t)m * sizeof(pnode_t));
        memset(g->node + g->node_m, 0, (size_t)(m - g->node_m) * sizeof(pnode_t));
        g->node_m = m;
    }
    pnode_t *nd = &g->node[g->node_n];
    nd->in_n = nd->out_n = nd->aln_n = 0; nd->base = base; nd->max_out_id = -1;
    return g->node_n++;
}

static void g_reset(pgraph_t *g)
{
    g->node_n = 0; g->sorted = 0;
    g_add_node(g, ALPHA_M); g_add_node(g, ALPHA_M);      /* src, sink */
}

static void g_free(pgraph_t *g)
{
    for (int i = 0; i < g->node_m; ++i) {
        free(g->node[i].in_id); free(g->node[i].out_id); free(g->node[i].out_w); free(g->node[i].out_reads);
    }
    free(g->node); free(g->index_to_node); free(g->node_to_index); free(g->max_pos_left);
    free(g->max_pos_right); free(g->max_remain); free(g->msa_rank);
    memset(g, 0, sizeof(*g));
}

/* abpoa_add_graph_edge: optional lookup of an existing edge (weight += w), else
 * append to the END of both the in- and the out-array (order is observable:
 * predecessor order drives backtrack ties, out order drives BFS/HB ties).   */
static void g_add_edge(pgraph_t *g, int from, int to, int check, int w, int read_id)
{
    pnode_t *f = &g->node[from], *t = &g->node[to];
    if (check) {
        for (int i = 0; i < f->out_n; ++i)
            if (f->out_id[i] == to) {
                f->out_w[i] += w;
                if (read_id < 64) f->out_reads[i] |= 1ull << read_id;
                return;
            }
    }
    if (t->in_n == t->in_m) {
        t->in_m = t->in_m ? t->in_m * 2 : 4;
        t->in_id = (int *)realloc(t->in_id, (size_t)t->in_m * sizeof(int));
    }
    t->in_id[t->in_n++] = from;
    if (f->out_n == f->out_m) {
        f->out_m = f->out_m ? f->out_m * 2 : 4;
        f->out_id = (int *)realloc(f->out_id, (size_t)f->out_m * sizeof(int));
        f->out_w = (int *)realloc(f->out_w, (size_t)f->out_m * sizeof(int));
        f->out_reads = (uint64_t *)realloc(f->out_reads, (size_t)f->out_m * sizeof(uint64_t));
    }
    f->out_id[f->out_n] = to; f->out_w[f->out_n] = w;
    f->out_reads[f->out_n] = read_id < 64 ? 1ull << read_id : 0;
    ++f->out_n;
}

static int g_get_aligned(pgraph_t *g, int node_id, uint8_t base)
{
    pnode_t *nd = &g->node[node_id];
    for (int i = 0; i < nd->aln_n; ++i)
        if (g->node[nd->aln_id[i]].base == base) return nd->aln_id[i];
    return -1;
}

static void g_add_aligned1(pnode_t *nd, int id) { if (nd->aln_n < ALPHA_M) nd->aln_id[nd->aln_n++] = id; }
static void g_add_aligned(pgraph_t *g, int node_id, int new_id)
{
    pnode_t *nd = &g->node[node_id];
    for (int i = 0; i < nd->aln_n; ++i) {
        g_add_aligned1(&g->node[nd->aln_id[i]], new_id);
        g_add_aligned1(&g->node[new_id], nd->aln_id[i]);
    }
    g_add_aligned1(&g->node[node_id], new_id);
    g_add_aligned1(&g->node[new_id], node_id);
}

static void g_aux_reserve(pgraph_t *g)
{
    if (g->aux_m >= g->node_n) return;
    int m = g->node_m;
    g->index_to_node = (int *)realloc(g->index_to_node, (size_t)m * sizeof(int));
    g->node_to_index = (int *)realloc(g->node_to_index, (size_t)m * sizeof(int));
    g->max_pos_left = (int *)realloc(g->max_pos_left, (size_t)m * sizeof(int));
    g->max_pos_right = (int *)realloc(g->max_pos_right, (size_t)m * sizeof(int));
    g->max_remain = (int *)realloc(g->max_remain, (size_t)m * sizeof(int));
    g->msa_rank = (int *)realloc(g->msa_rank, (size_t)m * sizeof(int));
    g->aux_m = m;
}

/* abpoa_BFS_set_node_index: Kahn BFS from the source; a node is released only
 * when it and every node aligned to it have in-degree 0; the group is queued
 * consecutively (node first, then its aligned ids in list order).           */
static int g_topo_sort(pgraph_t *g)
{
    int n = g->node_n;
    g_aux_reserve(g);
    int *indeg = (int *)malloc((size_t)n * sizeof(int));
    int *q = g->index_to_node;            /* the BFS queue IS the order */
    for (int i = 0; i < n; ++i) indeg[i] = g->node[i].in_n;
    int head = 0, tail = 0, ok = 0;
    q[tail++] = SRC_ID;
    while (head < tail) {
        int cur = q[head];
        g->node_to_index[cur] = head++;
        if (cur == SINK_ID) { ok = 1; break; }
        pnode_t *nd = &g->node[cur];
        for (int i = 0; i < nd->out_n; ++i) {
            int o = nd->out_id[i];
            if (--indeg[o] == 0) {
                pnode_t *on = &g->node[o];
                int ready = 1;
                for (int j = 0; j < on->aln_n; ++j)
                    if (indeg[on->aln_id[j]] != 0) { ready = 0; break; }
                if (!ready) continue;
                q[tail++] = o;
                for (int j = 0; j < on->aln_n; ++j) q[tail++] = on->aln_id[j];
            }
        }
    }
    free(indeg);
    if (!ok || head != n) return -1;
    /* band bookkeeping (abpoa_topological_sort, wb >= 0) */
    for (int i = 0; i < n; ++i) { g->max_pos_right[i] = 0; g->max_pos_left[i] = n; }
    /* abpoa_BFS_set_node_remain: reverse BFS from the sink; remain = 1 + remain of
     * the target of the HEAVIEST out edge (first maximal edge in out order).  */
    int *outdeg = (int *)malloc((size_t)n * sizeof(int));
    int *rq = (int *)malloc((size_t)n * sizeof(int));
    for (int i = 0; i < n; ++i) { outdeg[i] = g->node[i].out_n; g->max_remain[i] = 0; }
    head = tail = 0; rq[tail++] = SINK_ID; g->max_remain[SINK_ID] = -1;
    ok = 0;
    while (head < tail) {
        int cur = rq[head++];
        pnode_t *nd = &g->node[cur];
        if (cur != SINK_ID) {
            int max_w = -1, max_id = -1;
            for (int i = 0; i < nd->out_n; ++i)
                if (nd->out_w[i] > max_w) { max_w = nd->out_w[i]; max_id = nd->out_id[i]; }
            g->max_remain[cur] = g->max_remain[max_id] + 1;
        }
        if (cur == SRC_ID) { ok = 1; break; }
        for (int i = 0; i < nd->in_n; ++i) {
            int p = nd->in_id[i];
            if (--outdeg[p] == 0) rq[tail++] = p;
        }
    }
    free(outdeg); free(rq);
    if (!ok) return -2;
    g->sorted = 1;
    return 0;
}

/* per-thread alignment workspace */
typedef struct {
    int32_t *pool; size_t pool_m;          /* H,E1,E2,F1,F2 rows, banded */
    size_t *row_off; int *dp_beg, *dp_end, *dp_beg_sn, *dp_end_sn; int rows_m;
    int32_t *qp; size_t qp_m;              /* query profile [ALPHA_M][qlen+1] */
    int32_t *tm, *te1, *te2; int tmp_m;
    cg_t *cg; int cg_m;
} paln_ws_t;

static void ws_free(paln_ws_t *w)
{
    free(w->pool); free(w->row_off); free(w->dp_beg); free(w->dp_end); free(w->dp_beg_sn);
    free(w->dp_end_sn); free(w->qp); free(w->tm); free(w->te1); free(w->te2); free(w->cg);
    memset(w, 0, sizeof(*w));
}

static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

/* simd_abpoa_align_sequence_to_graph (convex gap, global, adaptive band).
 * Returns number of cigar ops (>=0) in ws->cg (forward order) or <0.         */
static int poa_align(pgraph_t *g, const c3o_poa_para_t *P, const uint8_t *query, int qlen,
                     paln_ws_t *ws, int *best_score_out, int64_t *cells_out)
{
    const int n = g->node_n;
    const int o1 = P->gap_open1, e1 = P->gap_ext1, o2 = P->gap_open2, e2 = P->gap_ext2;
    const int oe1 = o1 + e1, oe2 = o2 + e2;
    int mat[ALPHA_M][ALPHA_M];
    for (int a = 0; a < ALPHA_M; ++a)
        for (int b = 0; b < ALPHA_M; ++b)
            mat[a][b] = (a == ALPHA_M - 1 || b == ALPHA_M - 1) ? 0 : (a == b ? P->match : -P->mismatch);
    /* score width -> lanes per SIMD vector (band rounding granule) */
    int len = qlen > n ? qlen : n;
    int max_score = imax(qlen * ALPHA_M, len * e1 + o1);
    int pn = (max_score <= 32767 - P->mismatch - o1 - e1) ? P->simd_bits / 16 : P->simd_bits / 32;
    if (P->int8_lanes && max_score <= 127 - P->mismatch - o1 - e1) pn = P->simd_bits / 8;
    int w = P->wb < 0 ? qlen : P->wb + (int)(P->wf * qlen);

    if (ws->rows_m < n) {
        ws->rows_m = n + n / 2 + 16;
        ws->row_off = (size_t *)realloc(ws->row_off, (size_t)ws->rows_m * sizeof(size_t));
        ws->dp_beg = (int *)realloc(ws->dp_beg, (size_t)ws->rows_m * sizeof(int));
        ws->dp_end = (int *)realloc(ws->dp_end, (size_t)ws->rows_m * sizeof(int));
        ws->dp_beg_sn = (int *)realloc(ws->dp_beg_sn, (size_t)ws->rows_m * sizeof(int));
        ws->dp_end_sn = (int *)realloc(ws->dp_end_sn, (size_t)ws->rows_m * sizeof(int));
    }
    if (ws->qp_m < (size_t)ALPHA_M * (qlen + 1)) {
        ws->qp_m = (size_t)ALPHA_M * (qlen + 1);
        ws->qp = (int32_t *)realloc(ws->qp, ws->qp_m * sizeof(int32_t));
    }
    if (ws->tmp_m < qlen + 2) {
        ws->tmp_m = qlen + 2;
        ws->tm = (int32_t *)realloc(ws->tm, (size_t)ws->tmp_m * sizeof(int32_t));
        ws->te1 = (int32_t *)realloc(ws->te1, (size_t)ws->tmp_m * sizeof(int32_t));
        ws->te2 = (int32_t *)realloc(ws->te2, (size_t)ws->tmp_m * sizeof(int32_t));
    }
    for (int k = 0; k < ALPHA_M; ++k) {
        int32_t *qp = ws->qp + (size_t)k * (qlen + 1);
        qp[0] = 0;
        for (int j = 0; j < qlen; ++j) qp[j + 1] = mat[k][query[j]];
    }
    size_t used = 0;
    int64_t cells = 0;
#define ROW_H(i)  (ws->pool + ws->row_off[i])
#define ROW_W(i)  (ws->dp_end[i] - ws->dp_beg[i] + 1)
    /* ---- first row (source, index 0): simd_abpoa_cg_first_row/first_dp ---- */
    {
        g->max_pos_left[SRC_ID] = g->max_pos_right[SRC_ID] = 0;
        pnode_t *s = &g->node[SRC_ID];
        for (int i = 0; i < s->out_n; ++i)
            g->max_pos_left[s->out_id[i]] = g->max_pos_right[s->out_id[i]] = 1;
        int r = qlen - g->max_remain[SRC_ID];
        int beg = imax(0, imin(g->max_pos_left[SRC_ID], r) - w);
        int end = imin(qlen, imax(g->max_pos_right[SRC_ID], r) + w);
        ws->dp_beg_sn[0] = beg / pn; ws->dp_end_sn[0] = end / pn;
        ws->dp_beg[0] = ws->dp_beg_sn[0] * pn;
        ws->dp_end[0] = imin(qlen, (ws->dp_end_sn[0] + 1) * pn - 1);
    }
    /* pool sizing is incremental: ensure capacity per row */
#define ENSURE(extra) do { if (used + (extra) > ws->pool_m) { \
        ws->pool_m = (used + (extra)) * 2 + 4096; \
        ws->pool = (int32_t *)realloc(ws->pool, ws->pool_m * sizeof(int32_t)); } } while (0)
    {
        int wd = ROW_W(0);
        ENSURE((size_t)5 * wd);
        ws->row_off[0] = used; used += (size_t)5 * wd;
        int32_t *H = ROW_H(0), *E1 = H + wd, *E2 = E1 + wd, *F1 = E2 + wd, *F2 = F1 + wd;
        /* dp_beg[0] is 0 for the source (max_pos 0, w >= 0) */
        for (int j = 0; j < wd; ++j) { H[j] = E1[j] = E2[j] = F1[j] = F2[j] = NEG_INF; }
        if (ws->dp_beg[0] == 0) {
            H[0] = 0; E1[0] = -oe1; E2[0] = -oe2;
            for (int j = 1; j < wd; ++j) {
                F1[j] = -(o1 + e1 * j); F2[j] = -(o2 + e2 * j);
                H[j] = imax(F1[j], F2[j]);
            }
        }
    }
    /* ---- rows 1 .. n-2 ---- */
    for (int idx = 1; idx < n - 1; ++idx) {
        int node_id = g->index_to_node[idx];
        pnode_t *nd = &g->node[node_id];
        int r = qlen - g->max_remain[node_id];
        int beg = imax(0, imin(g->max_pos_left[node_id], r) - w);
        int end = imin(qlen, imax(g->max_pos_right[node_id], r) + w);
        int beg_sn = beg / pn, end_sn = end / pn;
        int min_pre_beg_sn = 0x7fffffff, max_pre_end_sn = -1;
        for (int k = 0; k < nd->in_n; ++k) {
            int pi = g->node_to_index[nd->in_id[k]];
            if (ws->dp_beg_sn[pi] < min_pre_beg_sn) min_pre_beg_sn = ws->dp_beg_sn[pi];
            if (ws->dp_end_sn[pi] > max_pre_end_sn) max_pre_end_sn = ws->dp_end_sn[pi];
        }
        if (beg_sn < min_pre_beg_sn) beg_sn = min_pre_beg_sn;
        if (P->end_clamp && end_sn > max_pre_end_sn + 1) end_sn = max_pre_end_sn + 1;
        if (end_sn < beg_sn) end_sn = beg_sn;             /* robustness guard (never seen) */
        ws->dp_beg_sn[idx] = beg_sn; ws->dp_end_sn[idx] = end_sn;
        beg = ws->dp_beg[idx] = beg_sn * pn;
        end = ws->dp_end[idx] = imin(qlen, (end_sn + 1) * pn - 1);
        int wd = end - beg + 1;
        if (wd <= 0) return -14;
        ENSURE((size_t)5 * wd);
        ws->row_off[idx] = used; used += (size_t)5 * wd;
        cells += wd;
        int32_t *H = ROW_H(idx), *E1 = H + wd, *E2 = E1 + wd, *F1 = E2 + wd, *F2 = F1 + wd;
        int32_t *tm = ws->tm, *te1 = ws->te1, *te2 = ws->te2;   /* indexed by j-beg */
        for (int j = 0; j < wd; ++j) { tm[j] = te1[j] = te2[j] = NEG_INF; }
        /* M from (pre, j-1): usable iff max(beg_pre, beg) <= j-1 <= end_pre;
         * E from (pre, j):   usable iff j inside both bands.                  */
        for (int k = 0; k < nd->in_n; ++k) {
            int pi = g->node_to_index[nd->in_id[k]];
            int pb = ws->dp_beg[pi], pe = ws->dp_end[pi], pw = pe - pb + 1;
            const int32_t *pH = ROW_H(pi), *pE1 = pH + pw, *pE2 = pE1 + pw;
            int lo = imax(pb, beg), hi = imin(pe, end - 1);
            for (int c = lo; c <= hi; ++c) {              /* c = j-1 */
                int v = pH[c - pb];
                if (v > tm[c + 1 - beg]) tm[c + 1 - beg] = v;
            }
            hi = imin(pe, end);
            for (int c = lo; c <= hi; ++c) {
                int v1 = pE1[c - pb], v2 = pE2[c - pb];
                if (v1 > te1[c - beg]) te1[c - beg] = v1;
                if (v2 > te2[c - beg]) te2[c - beg] = v2;
            }
        }
        const int32_t *qp = ws->qp + (size_t)nd->base * (qlen + 1);
        /* Hme = max(M+s, E1in, E2in); F by the scan definition; H; E for the next row */
        int32_t run1 = NEG_INF, run2 = NEG_INF;           /* max_{j'<j} Hme[j'] + e*j' */
        for (int j = beg; j <= end; ++j) {
            int c = j - beg;
            int m = tm[c] + qp[j];
            int hme = imax(m, imax(te1[c], te2[c]));
            int f1 = run1 - o1 - e1 * j;
            int f2 = run2 - o2 - e2 * j;
            int h = imax(hme, imax(f1, f2));
            H[c] = h; F1[c] = f1; F2[c] = f2;
            E1[c] = imax(h - oe1, te1[c] - e1);
            E2[c] = imax(h - oe2, te2[c] - e2);
            run1 = imax(run1, hme + e1 * j);
            run2 = imax(run2, hme + e2 * j);
        }
        /* simd_abpoa_ada_max_i: lane-wise running max (last vector seeds, earlier
         * vectors replace on strict >), then lanes scanned low->high on strict >. */
        {
            int best = NEG_INF, best_i = -1;
            for (int lane = 0; lane < pn; ++lane) {
                int lv = NEG_INF, li = -1;
                int jl = end_sn * pn + lane;
                if (jl <= qlen && jl >= beg) { lv = H[jl - beg]; li = jl; }
                for (int sn = beg_sn; sn < end_sn; ++sn) {
                    int j = sn * pn + lane;
                    if (H[j - beg] > lv) { lv = H[j - beg]; li = j; }
                }
                if (lv > best) { best = lv; best_i = li; }
            }
            if (best < NEG_HALF) best_i = -1;
            for (int i = 0; i < nd->out_n; ++i) {
                int o = nd->out_id[i];
                if (best_i + 1 > g->max_pos_right[o]) g->max_pos_right[o] = best_i + 1;
                if (best_i + 1 < g->max_pos_left[o]) g->max_pos_left[o] = best_i + 1;
            }
        }
    }
    *cells_out = cells;
    /* ---- simd_abpoa_global_get_max: best over the sink's predecessors ---- */
    int best_score = -0x7fffffff - 1, best_i = -1, best_j = -1;
    {
        pnode_t *sk = &g->node[SINK_ID];
        for (int k = 0; k < sk->in_n; ++k) {
            int pi = g->node_to_index[sk->in_id[k]];
            int e = imin(qlen, ws->dp_end[pi]);
            int v = ROW_H(pi)[e - ws->dp_beg[pi]];
            if (v > best_score) { best_score = v; best_i = pi; best_j = e; }
        }
    }
    *best_score_out = best_score;
    if (best_i < 0) return -10;
    /* ---- simd_abpoa_cg_backtrack ---- */
    if (ws->cg_m < qlen + n + 8) {
        ws->cg_m = qlen + n + 8;
        ws->cg = (cg_t *)realloc(ws->cg, (size_t)ws->cg_m * sizeof(cg_t));
    }
    cg_t *cg = ws->cg; int nc = 0;                        /* filled in reverse */
    int i = best_i, j = best_j;
    for (int t = qlen; t > best_j; --t) { cg[nc].op = CG_INS; cg[nc].node_id = -1; cg[nc].qpos = t - 1; ++nc; }
    int cur_op = OP_ALL;
    while (i > 0 && j > 0) {
        int id = g->index_to_node[i];
        pnode_t *nd = &g->node[id];
        int b = ws->dp_beg[i], wd = ROW_W(i);
        const int32_t *H = ROW_H(i), *E1 = H + wd, *E2 = E1 + wd, *F1 = E2 + wd, *F2 = F1 + wd;
        if (j < b || j > ws->dp_end[i]) return -11;
        int s = mat[nd->base][query[j - 1]];
        int hit = 0;
        if (cur_op & OP_M) {
            for (int k = 0; k < nd->in_n; ++k) {
                int pi = g->node_to_index[nd->in_id[k]];
                if (j - 1 < imax(ws->dp_beg[pi], b) || j - 1 > ws->dp_end[pi]) continue;
                if (ROW_H(pi)[j - 1 - ws->dp_beg[pi]] + s == H[j - b]) {
                    cg[nc].op = CG_MATCH; cg[nc].node_id = id; cg[nc].qpos = j - 1; ++nc;
                    i = pi; --j; hit = 1; cur_op = OP_ALL;
                    break;
                }
            }
        }
        if (!hit && (cur_op & OP_E)) {
            for (int k = 0; k < nd->in_n; ++k) {
                int pi = g->node_to_index[nd->in_id[k]];
                if (j < ws->dp_beg[pi] || j > ws->dp_end[pi]) continue;
                int pw = ROW_W(pi), pc = j - ws->dp_beg[pi];
                const int32_t *pH = ROW_H(pi), *pE1 = pH + pw, *pE2 = pE1 + pw;
                if (cur_op & OP_E1) {
                    if (cur_op & OP_M) {
                        if (H[j - b] == pE1[pc]) {
                            cur_op = (pH[pc] - oe1 == pE1[pc]) ? (OP_M | OP_F) : OP_E1;
                            hit = 1;
                        }
                    } else if (E1[j - b] == pE1[pc] - e1) {
                        cur_op = (pH[pc] - oe1 == pE1[pc]) ? (OP_M | OP_F) : OP_E1;
                        hit = 1;
                    }
                }
                if (!hit && (cur_op & OP_E2)) {
                    if (cur_op & OP_M) {
                        if (H[j - b] == pE2[pc]) {
                            cur_op = (pH[pc] - oe2 == pE2[pc]) ? (OP_M | OP_F) : OP_E2;
                            hit = 1;
                        }
                    } else if (E2[j - b] == pE2[pc] - e2) {
                        cur_op = (pH[pc] - oe2 == pE2[pc]) ? (OP_M | OP_F) : OP_E2;
                        hit = 1;
                    }
                }
                if (hit) {
                    cg[nc].op = CG_DEL; cg[nc].node_id = id; cg[nc].qpos = j - 1; ++nc;
                    i = pi;
                    break;
                }
            }
        }
        if (!hit && (cur_op & OP_F)) {
            if (j - 1 >= b) {
                if (cur_op & OP_F1) {
                    if (cur_op & OP_M) {
                        if (H[j - b] == F1[j - b]) {
                            if (H[j - 1 - b] - oe1 == F1[j - b]) { cur_op = OP_M | OP_E; hit = 1; }
                            else if (F1[j - 1 - b] - e1 == F1[j - b]) { cur_op = OP_F1; hit = 1; }
                        }
                    } else {
                        if (H[j - 1 - b] - oe1 == F1[j - b]) { cur_op = OP_M | OP_E; hit = 1; }
                        else if (F1[j - 1 - b] - e1 == F1[j - b]) { cur_op = OP_F1; hit = 1; }
                    }
                }
                if (!hit && (cur_op & OP_F2)) {
                    if (cur_op & OP_M) {
                        if (H[j - b] == F2[j - b]) {
                            if (H[j - 1 - b] - oe2 == F2[j - b]) { cur_op = OP_M | OP_E; hit = 1; }
                            else if (F2[j - 1 - b] - e2 == F2[j - b]) { cur_op = OP_F2; hit = 1; }
                        }
                    } else {
                        if (H[j - 1 - b] - oe2 == F2[j - b]) { cur_op = OP_M | OP_E; hit = 1; }
                        else if (F2[j - 1 - b] - e2 == F2[j - b]) { cur_op = OP_F2; hit = 1; }
                    }
                }
            }
            if (hit) { cg[nc].op = CG_INS; cg[nc].node_id = id; cg[nc].qpos = j - 1; ++nc; --j; }
        }
        if (!hit) return -12;
        if (nc >= ws->cg_m - 2) return -13;
    }
    for (; j > 0; --j) { cg[nc].op = CG_INS; cg[nc].node_id = -1; cg[nc].qpos = j - 1; ++nc; }
    /* reverse into forward order */
    for (int a = 0, z = nc - 1; a < z; ++a, --z) { cg_t t = cg[a]; cg[a] = cg[z]; cg[z] = t; }
    return nc;
#undef ROW_H
#undef ROW_W
#undef ENSURE
}

/* abpoa_add_graph_alignment */
static void poa_add_alignment(pgraph_t *g, const uint8_t *seq, int seq_l, const cg_t *cg, int nc,
                              int read_id, int first)
{
    int last_id = SRC_ID, last_new = 0;
    if (first) {                                          /* abpoa_add_graph_sequence */
        for (int i = 0; i < seq_l; ++i) {
            int id = g_add_node(g, seq[i]);
            g_add_edge(g, last_id, id, 0, 1, read_id);
            last_id = id;
        }
        g_add_edge(g, last_id, SINK_ID, 0, 1, read_id);
        g->sorted = 0;
        return;
    }
    for (int t = 0; t < nc; ++t) {
        if (cg[t].op == CG_MATCH) {
            int node_id = cg[t].node_id; uint8_t b = seq[cg[t].qpos];
            if (g->node[node_id].base != b) {
                int al = g_get_aligned(g, node_id, b);
                if (al != -1) {
                    g_add_edge(g, last_id, al, 1 - last_new, 1, read_id);
                    last_id = al; last_new = 0;
                } else {
                    int id = g_add_node(g, b);
                    g_add_edge(g, last_id, id, 0, 1, read_id);
                    last_id = id; last_new = 1;
                    g_add_aligned(g, node_id, id);
                }
            } else {
                g_add_edge(g, last_id, node_id, 1 - last_new, 1, read_id);
                last_id = node_id; last_new = 0;
            }
        } else if (cg[t].op == CG_INS) {
            int id = g_add_node(g, seq[cg[t].qpos]);
            g_add_edge(g, last_id, id, 0, 1, read_id);
            last_id = id; last_new = 1;
        }                                                 /* CG_DEL: nothing */
    }
    g_add_edge(g, last_id, SINK_ID, 1 - last_new, 1, read_id);
    g->sorted = 0;
}

/* abpoa_heaviest_bundling + consensus walk */
static int poa_consensus(pgraph_t *g, char *out, int cap)
{
    int n = g->node_n;
    int *outdeg = (int *)malloc((size_t)n * sizeof(int));
    int *score = (int *)calloc((size_t)n, sizeof(int));
    int *q = (int *)malloc((size_t)n * sizeof(int));
    for (int i = 0; i < n; ++i) outdeg[i] = g->node[i].out_n;
    int head = 0, tail = 0; q[tail++] = SINK_ID;
    while (head < tail) {
        int cur = q[head++];
        pnode_t *nd = &g->node[cur];
        if (cur == SINK_ID) { nd->max_out_id = -1; score[cur] = 0; }
        else if (cur == SRC_ID) {
            int max_id = -1, path_score = -1, path_w = -1;
            for (int i = 0; i < nd->out_n; ++i) {
                int o = nd->out_id[i], w = nd->out_w[i];
                if (w > path_w || (w == path_w && score[o] > path_score)) { max_id = o; path_score = score[o]; path_w = w; }
            }
            nd->max_out_id = max_id;
            break;
        } else {
            int max_w = -0x7fffffff - 1, max_id = -1;
            for (int i = 0; i < nd->out_n; ++i) {
                int o = nd->out_id[i], w = nd->out_w[i];
                if (max_w < w) { max_w = w; max_id = o; }
                else if (max_w == w && score[max_id] <= score[o]) max_id = o;
            }
            score[cur] = max_w + score[max_id];
            nd->max_out_id = max_id;
        }
        for (int i = 0; i < nd->in_n; ++i) {
            int p = nd->in_id[i];
            if (--outdeg[p] == 0) q[tail++] = p;
        }
    }
    free(outdeg); free(score); free(q);
    int len = 0, id = g->node[SRC_ID].max_out_id;
    while (id != SINK_ID) {
        if (id < 0 || len >= cap) return -1;
        out[len++] = "ACGTN"[g->node[id].base];
        id = g->node[id].max_out_id;
    }
    return len;
}

/* abpoa_generate_rc_msa (abpoa_DFS_set_msa_rank: LIFO traversal) */
static int poa_msa(pgraph_t *g, int n_seq, char *msa, int cap)
{
    int n = g->node_n;
    g_aux_reserve(g);
    int *indeg = (int *)malloc((size_t)n * sizeof(int));
    int *st = (int *)malloc((size_t)n * sizeof(int));
    int *rank = g->msa_rank;
    for (int i = 0; i < n; ++i) { indeg[i] = g->node[i].in_n; rank[i] = 0; }
    int top = 0, msa_rank = 0, ok = 0;
    st[top++] = SRC_ID; rank[SRC_ID] = -1;
    while (top > 0) {
        int cur = st[--top];
        pnode_t *nd = &g->node[cur];
        if (rank[cur] < 0) {
            rank[cur] = msa_rank;
            for (int i = 0; i < nd->aln_n; ++i) rank[nd->aln_id[i]] = msa_rank;
            ++msa_rank;
        }
        if (cur == SINK_ID) { ok = 1; break; }
        for (int i = 0; i < nd->out_n; ++i) {
            int o = nd->out_id[i];
            if (--indeg[o] == 0) {
                pnode_t *on = &g->node[o];
                int ready = 1;
                for (int j = 0; j < on->aln_n; ++j) if (indeg[on->aln_id[j]] != 0) { ready = 0; break; }
                if (!ready) continue;
                st[top++] = o; rank[o] = -1;
                for (int j = 0; j < on->aln_n; ++j) { st[top++] = on->aln_id[j]; rank[on->aln_id[j]] = -1; }
            }
        }
    }
    free(indeg); free(st);
    if (!ok) return -1;
    int msa_len = rank[SINK_ID] - 1;
    if (msa_len > cap) return -2;
    for (int r = 0; r < n_seq; ++r) memset(msa + (size_t)r * cap, '-', (size_t)msa_len);
    for (int i = 2; i < n; ++i) {
        pnode_t *nd = &g->node[i];
        int rk = rank[i];
        for (int k = 0; k < nd->aln_n; ++k) rk = imax(rk, rank[nd->aln_id[k]]);
        for (int j = 0; j < nd->out_n; ++j) {
            uint64_t b = nd->out_reads[j];
            for (int r = 0; r < n_seq && r < 64; ++r)
                if (b >> r & 1) msa[(size_t)r * cap + rk - 1] = "ACGTN"[nd->base];
        }
    }
    return msa_len;
}

typedef struct { pgraph_t g; paln_ws_t ws; uint8_t *bseq; int bseq_m; } poa_ctx_t;

static int poa_run(poa_ctx_t *ctx, const c3o_poa_para_t *P, int n_seq, const char *const *seqs,
                   const int32_t *lens, char *cons_out, int cons_cap, int *cons_len,
                   char *msa_out, int msa_cap, int *msa_len, c3o_poa_stats_t *st, int32_t *dbg)
{
    pgraph_t *g = &ctx->g;
    g_reset(g);
    memset(st, 0, sizeof(*st));
    for (int r = 0; r < n_seq; ++r) {
        int L = lens[r];
        if (ctx->bseq_m < L + 1) { ctx->bseq_m = L + 1024; ctx->bseq = (uint8_t *)realloc(ctx->bseq, (size_t)ctx->bseq_m); }
        for (int i = 0; i < L; ++i) ctx->bseq[i] = (uint8_t)base_code(seqs[r][i]);
        if (g->node_n <= 2) {
            poa_add_alignment(g, ctx->bseq, L, NULL, 0, r, 1);
            continue;
        }
        if (!g->sorted) { int rc = g_topo_sort(g); if (rc) { st->status = rc; return rc; } }
        int score = 0; int64_t cells = 0;
        int nc = poa_align(g, P, ctx->bseq, L, &ctx->ws, &score, &cells);
        if (nc < 0) { st->status = nc; return nc; }
        poa_add_alignment(g, ctx->bseq, L, ctx->ws.cg, nc, r, 0);
        if (dbg) { dbg[4 * st->n_aln] = score; dbg[4 * st->n_aln + 1] = nc; dbg[4 * st->n_aln + 2] = g->node_n; dbg[4 * st->n_aln + 3] = (int32_t)cells; }
        st->cells += cells; st->n_aln++; st->last_score = score;
    }
    st->node_n = g->node_n;
    if (cons_out) {
        int cl = g->node_n > 2 ? poa_consensus(g, cons_out, cons_cap) : 0;
        if (cl < 0) { st->status = -20; return -20; }
        *cons_len = cl;
    }
    if (msa_out) {
        int ml = g->node_n > 2 ? poa_msa(g, n_seq, msa_out, msa_cap) : 0;
        if (ml < 0) { st->status = -21; return -21; }
        *msa_len = ml;
    }
    return 0;
}

static void ctx_free(poa_ctx_t *c) { g_free(&c->g); ws_free(&c->ws); free(c->bseq); memset(c, 0, sizeof(*c)); }

int c3o_poa_msa(const c3o_poa_para_t *para, int n_seq, const char *const *seqs,
                const int32_t *seq_lens, char *cons_out, int cons_cap, int *cons_len,
                char *msa_out, int msa_cap, int *msa_len, c3o_poa_stats_t *stats, int32_t *dbg)
{
    poa_ctx_t ctx; memset(&ctx, 0, sizeof(ctx));
    c3o_poa_stats_t st;
    int rc = poa_run(&ctx, para, n_seq, seqs, seq_lens, cons_out, cons_cap, cons_len,
                     msa_out, msa_cap, msa_len, &st, dbg);
    if (stats) *stats = st;
    ctx_free(&ctx);
    return rc;
}

/* ===================================================================== */
/* whole per-read path, batch, pthreads (CPU baseline leg)                */
/* ===================================================================== */
typedef struct {
    int n_reads; const char *reads; const int64_t *read_off;
    const char *const *splints; const int32_t *splint_lens; const int32_t *splint_idx;
    int penalty, min_dist, iters; const double *coef; int window; const c3o_poa_para_t *para;
    int max_peaks; int32_t *peaks_out, *sub_bounds_out, *dang_bounds_out; int cons_cap; char *cons_out;
    c3o_read_result_t *results;
    int next; int err; pthread_mutex_t mu;
} batch_job_t;

static void batch_one(batch_job_t *J, int r, poa_ctx_t *ctx, int32_t **prof, int *prof_m,
                      const char **sp, int32_t *sl)
{
    const char *seq = J->reads + J->read_off[r];
    int lr = (int)(J->read_off[r + 1] - J->read_off[r]);
    c3o_read_result_t *res = &J->results[r];
    memset(res, 0, sizeof(*res));
    int si = J->splint_idx[r], ls = J->splint_lens[si], max_peaks = J->max_peaks;
    if (*prof_m < lr) { *prof_m = lr + 4096; *prof = (int32_t *)realloc(*prof, (size_t)*prof_m * sizeof(int32_t)); }
    if (c3o_conk(J->splints[si], ls, seq, lr, J->penalty, *prof)) { res->status = -1; J->err = 1; return; }
    res->conk_cells = (int64_t)ls * lr;
    int32_t *pk = J->peaks_out + (size_t)r * max_peaks;
    int np_ = c3o_call_peaks(*prof, lr, J->min_dist, J->iters, J->coef, J->window, NULL, NULL, pk, max_peaks);
    if (np_ < 0) { res->status = -2; J->err = 1; return; }
    int32_t *sb = J->sub_bounds_out + (size_t)r * 2 * max_peaks;
    int32_t *db = J->dang_bounds_out + (size_t)r * 4;
    int nsub = 0, ndang = 0;
    int skip = c3o_split(pk, &np_, ls, lr, sb, &nsub, db, &ndang);
    res->n_peaks = np_; res->n_sub = nsub; res->n_dang = ndang;
    if (skip) { res->status = 1; return; }
    char *co = J->cons_out + (size_t)r * J->cons_cap;
    if (nsub >= 3) {
        for (int k = 0; k < nsub; ++k) { sp[k] = seq + sb[2 * k]; sl[k] = sb[2 * k + 1] - sb[2 * k]; }
        c3o_poa_stats_t st; int cl = 0;
        int rc = poa_run(ctx, J->para, nsub, sp, sl, co, J->cons_cap, &cl, NULL, 0, NULL, &st, NULL);
        if (rc) { res->status = rc; J->err = 1; return; }
        res->cons_len = cl; res->poa_cells = st.cells; res->poa_nodes = st.node_n;
    } else if (nsub == 1) {                               /* determine_consensus.py:31-32 */
        int L = sb[1] - sb[0];
        if (L > J->cons_cap) { res->status = -3; J->err = 1; return; }
        memcpy(co, seq + sb[0], (size_t)L); res->cons_len = L;
    } else if (nsub == 2) {
        /* 2-repeat path (determine_consensus.py:33-41): msa(out_cons=False, out_msa=True); the two MSA
         * rows go to the consensus slot as [row0 | row1], cons_len = number of columns; the quality-aware
         * pairwise_consensus runs on the host (it needs the quality strings) */
        for (int k = 0; k < 2; ++k) { sp[k] = seq + sb[2 * k]; sl[k] = sb[2 * k + 1] - sb[2 * k]; }
        c3o_poa_stats_t st; int ml = 0, half = J->cons_cap / 2;
        int rc = poa_run(ctx, J->para, 2, sp, sl, NULL, 0, NULL, co, half, &ml, &st, NULL);
        if (rc) { res->status = rc == -21 ? -209 : rc; if (rc != -21) J->err = 1; return; }
        memmove(co + ml, co + half, (size_t)ml);
        res->status = 2; res->cons_len = ml; res->poa_cells = st.cells; res->poa_nodes = st.node_n;
    } else {
        res->status = 2;                                  /* 0-repeat path (needs mappy): bounds only */
    }
}

static void *batch_worker(void *arg)
{
    batch_job_t *J = (batch_job_t *)arg;
    poa_ctx_t ctx; memset(&ctx, 0, sizeof(ctx));
    int32_t *prof = NULL; int prof_m = 0;
    const char **sp = (const char **)malloc((size_t)(J->max_peaks + 1) * sizeof(char *));
    int32_t *sl = (int32_t *)malloc((size_t)(J->max_peaks + 1) * sizeof(int32_t));
    for (;;) {
        pthread_mutex_lock(&J->mu);
        int b = J->next; J->next += 4;
        pthread_mutex_unlock(&J->mu);
        if (b >= J->n_reads) break;
        int e = b + 4 < J->n_reads ? b + 4 : J->n_reads;
        for (int r = b; r < e; ++r) batch_one(J, r, &ctx, &prof, &prof_m, sp, sl);
    }
    ctx_free(&ctx); free(prof); free(sp); free(sl);
    return NULL;
}

int c3o_consensus_batch(int n_reads, const char *reads, const int64_t *read_off,
                        int n_splints, const char *const *splints, const int32_t *splint_lens,
                        const int32_t *splint_idx, int penalty, int min_dist, int iters,
                        const double *coef, int window, const c3o_poa_para_t *para,
                        int max_peaks, int32_t *peaks_out, int32_t *sub_bounds_out,
                        int32_t *dang_bounds_out, int cons_cap, char *cons_out,
                        c3o_read_result_t *results, int n_threads)
{
    (void)n_splints;
    batch_job_t J;
    J.n_reads = n_reads; J.reads = reads; J.read_off = read_off; J.splints = splints;
    J.splint_lens = splint_lens; J.splint_idx = splint_idx; J.penalty = penalty; J.min_dist = min_dist;
    J.iters = iters; J.coef = coef; J.window = window; J.para = para; J.max_peaks = max_peaks;
    J.peaks_out = peaks_out; J.sub_bounds_out = sub_bounds_out; J.dang_bounds_out = dang_bounds_out;
    J.cons_cap = cons_cap; J.cons_out = cons_out; J.results = results; J.next = 0; J.err = 0;
    pthread_mutex_init(&J.mu, NULL);
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 1024) n_threads = 1024;
    pthread_t *th = (pthread_t *)malloc((size_t)n_threads * sizeof(pthread_t));
    int started = 0;
    for (int t = 1; t < n_threads; ++t)
        if (pthread_create(&th[started], NULL, batch_worker, &J) == 0) ++started;
    batch_worker(&J);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    free(th);
    pthread_mutex_destroy(&J.mu);
    return J.err ? -1 : 0;
}
