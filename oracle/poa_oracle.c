/* poa_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never linked into, loaded by, or called from
 * the product (libabpoa_b200.so); only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may use it, and only as the checker.
 *
 * A plain scalar C restatement of abPOA's sequence-to-graph DP and backtrace
 * (reference src/abpoa_align_simd.c), written from the published recurrences rather than
 * from the SIMD macros: one int per cell, true "minus infinity" outside the band, the
 * adaptive band PUSHED to successors the way the reference does it.  Each block cites
 * the reference lines it follows.  Parity is PINNED: tests/test_oracle.py checks this file
 * against the unmodified reference built by oracle/Makefile (oracle/_ref/libabpoa_ref.so)
 * -- scores, every graph-CIGAR word, band of every row -- and against the golden vectors
 * in tests/golden/ generated from that same reference.
 *
 * It works on the abpoa.h structs (any library exporting that ABI can own the graph).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "abpoa.h"
#include "poa_oracle.h"

#define NINF (INT32_MIN / 4)
#define MAX2(a, b) ((a) > (b) ? (a) : (b))
#define MIN2(a, b) ((a) < (b) ? (a) : (b))

#define OP_M 0x1
#define OP_E1 0x2
#define OP_E2 0x4
#define OP_E 0x6
#define OP_F1 0x8
#define OP_F2 0x10
#define OP_F 0x18
#define OP_ALL 0x1f

typedef struct {
    int beg, end;          /* band of the row, inclusive                         */
    int *h, *e1, *e2, *f1, *f2;   /* planes, indexed by j - beg                  */
    /* banded linear-gap rows (global / extend): everything the reference's VECTOR row holds, cells
     * xbeg .. xend = whole vectors around the band (see lg_vector_row); h points into it */
    int xbeg, xend; int *hx;
    void *alloc;           /* what to free                                       */
} orow_t;

static void *xm(size_t n) { void *p = malloc(n ? n : 1); if (!p) { fprintf(stderr, "[poa_oracle] out of memory\n"); exit(1); } return p; }

static inline int cell(const int *plane, const orow_t *r, int j) {
    return (plane && j >= r->beg && j <= r->end) ? plane[j - r->beg] : NINF;
}
static inline int addinf(int v, int d) { return v <= NINF / 2 ? NINF : v + d; }   /* -inf stays -inf */
/* stored value of a vector-exact linear-gap row at column j (inside or outside its band) */
static inline int xcell(const orow_t *r, int j) { return (r->hx && j >= r->xbeg && j <= r->xend) ? r->hx[j - r->xbeg] : NINF; }

/* The reference's in-vector max-plus scan SIMD_SET_F (src/abpoa_align_simd.c:691-725), lane for lane.  Table semantics
 * (:480-495): PRE_MASK[c] keeps lanes 0..c, SUF_MIN[c] sets lanes > c to -inf, PRE_MIN[n] sets lanes < n to -inf.
 * set_num == pn: full log-step scan.  set_num < pn ("suffix MIN_INF"): at step k only lanes <= cov_k receive, cov_0 =
 * set_num, cov_k = cov_{k-1} + 2^k -- which is NOT a complete scan (set_num 0: lane 1 never receives, the last lane never
 * does), and that is part of what the reference computes at the right edge of a band. */
static void lg_set_f(int *F, int pn, int log_n, int set_num, int e1) {
    int G[64];
    int cov = set_num;
    for (int k = 0; k < log_n; ++k) {
        const int sh = 1 << k;
        if (set_num != pn && k > 0) cov += sh;
        for (int l = 0; l < pn; ++l) {
            int v = (l >= sh) ? addinf(F[l - sh], -e1 * sh) : NINF;
            if (set_num != pn && l > (cov < pn - 1 ? cov : pn - 1)) v = NINF;
            G[l] = v;
        }
        for (int l = 0; l < pn; ++l) F[l] = MAX2(F[l], G[l]);
    }
}

/* One banded linear-gap row in global / extend mode exactly as the reference's vector procedure leaves it
 * (simd_abpoa_lg_dp, src/abpoa_align_simd.c:727-815; pn = lanes of the AVX2 vector for the score width the reference
 * picks).  Differences from the textbook recurrence, all at band edges:
 *   - predecessor p is read in whole vectors _beg_sn .. min((pre_end+1)/pn, end_sn): cells of p outside its band but
 *     inside its stored vectors take part with whatever p left there;
 *   - after the scan the cells end+1 .. end of the last vector are NOT re-masked: they hold H[end] - k*E1 and are
 *     visible to successor rows;
 *   - vectors beyond the predecessors' last vector are scanned with set_num 1 / 0 (lg_set_f). */
static void lg_vector_row(orow_t *r, const orow_t *rows, const int *pre, int pre_n, const int *ps_of, const int *srow, const uint8_t *query,
                          int qlen, int pn, int log_n, int e1) {
    const int beg = r->beg, end = r->end, beg_sn = beg / pn, end_sn = end / pn, dp_sn = (qlen + 1 + pn - 1) / pn;
    r->xbeg = beg_sn * pn; r->xend = (end_sn + 1) * pn - 1;
    const int wd = r->xend - r->xbeg + 1;
    int *hx = (int *)xm((size_t)wd * sizeof(int));
    for (int x = 0; x < wd; ++x) hx[x] = NINF;
    r->hx = hx; r->alloc = hx; r->h = hx + (beg - r->xbeg);
    int max_pre_end_sn = -1;
    for (int k = 0; k < pre_n; ++k) { const int es = rows[pre[k]].end / pn; if (es > max_pre_end_sn) max_pre_end_sn = es; }
    for (int k = 0; k < pre_n; ++k) {
        const orow_t *p = &rows[pre[k]];
        const int ps = ps_of ? ps_of[k] : 0;
        const int pre_beg_sn = p->beg / pn;
        int _beg_sn, first;
        if (pre_beg_sn < beg_sn) { _beg_sn = beg_sn; first = xcell(p, beg_sn * pn - 1); }
        else { _beg_sn = pre_beg_sn; first = NINF; }
        int _end_sn = (p->end + 1) / pn;
        if (end_sn < _end_sn) _end_sn = end_sn;
        if (dp_sn - 1 < _end_sn) _end_sn = dp_sn - 1;
        for (int j = _beg_sn * pn; j < (_end_sn + 1) * pn; ++j) {
            const int pm1 = (j == _beg_sn * pn) ? first : xcell(p, j - 1);
            const int q = (j >= 1 && j <= qlen) ? srow[query[j - 1]] : 0;             /* query profile: qp[.][0] = 0, padding 0 */
            const int cand = MAX2(addinf(pm1, q + ps), addinf(xcell(p, j), ps - e1));
            int *c = &hx[j - r->xbeg];
            *c = (k == 0) ? cand : MAX2(cand, *c);
        }
    }
    for (int j = r->xbeg; j < beg; ++j) hx[j - r->xbeg] = NINF;
    for (int j = end + 1; j <= r->xend; ++j) hx[j - r->xbeg] = NINF;
    int first = hx[0];                                                                  /* lane 0 of the first vector */
    for (int sn = beg_sn; sn <= end_sn; ++sn) {
        int *F = hx + (sn * pn - r->xbeg);
        const int set_num = sn > max_pre_end_sn ? (sn == max_pre_end_sn + 1 ? 1 : 0) : pn;
        F[0] = MAX2(F[0], first);
        if (sn == end_sn) for (int j = end + 1; j < (end_sn + 1) * pn; ++j) F[j - sn * pn] = NINF;
        lg_set_f(F, pn, log_n, set_num, e1);
        first = addinf(F[pn - 1], -e1);
    }
}

/* -G path score of an in-edge: reference src/abpoa_graph.c:421-437 */
static int path_score(const abpoa_graph_t *g, int node_id, int k) {
    const abpoa_node_t *nd = &g->node[node_id], *pre = &g->node[nd->in_id[k]];
    int node_w = 0;
    for (int e = 0; e < pre->out_edge_n; ++e) node_w += pre->out_edge_weight[e];
    int edge_w = nd->in_edge_weight[k];
    if (node_w == 0 || edge_w == 0) return 0;
    int s = (int)round(log((double)edge_w / (double)node_w));
    return MAX2(s, -20);
}

typedef struct { int n, m; abpoa_cigar_t *a; } cig_t;
/* reference src/abpoa_align.h:54-73: consecutive insertions merge, nothing else does */
static void push(cig_t *c, int op, int len, int node_id, int qpos) {
    if (c->n > 0 && op == ABPOA_CINS && (int)(c->a[c->n - 1] & 0xf) == ABPOA_CINS) { c->a[c->n - 1] += (uint64_t)len << 4; return; }
    if (c->n == c->m) { c->m = c->m ? c->m << 1 : 4; c->a = (abpoa_cigar_t *)realloc(c->a, (size_t)c->m * sizeof(abpoa_cigar_t)); }
    if (op == ABPOA_CMATCH) c->a[c->n++] = ((uint64_t)node_id << 34) | ((uint64_t)qpos << 4) | op;
    else if (op == ABPOA_CINS) c->a[c->n++] = ((uint64_t)(uint32_t)qpos << 34) | ((uint64_t)len << 4) | op;
    else c->a[c->n++] = ((uint64_t)node_id << 34) | ((uint64_t)len << 4) | op;
}

/* first predecessor (in_id order) whose diagonal cell explains H[i][j] */
static int find_diag(const abpoa_graph_t *g, const abpoa_para_t *abpt, const orow_t *rows, const int *pre, const int *pre_k,
                     int n, int id, int j, int s, int hij) {
    for (int k = 0; k < n; ++k) {
        const orow_t *p = &rows[pre[k]];
        const int ps = abpt->inc_path_score ? path_score(g, id, pre_k[k]) : 0;
        if (j - 1 < p->beg || j - 1 > p->end) continue;
        if (cell(p->h, p, j - 1) + s + ps == hij) return k;
    }
    return -1;
}

int poa_oracle_score_bits(const abpoa_para_t *abpt, int qlen, int gn) {       /* reference :1293-1303 */
    int len = qlen > gn ? qlen : gn;
    int oe1 = abpt->gap_open1 + abpt->gap_ext1, oe2 = abpt->gap_open2 + abpt->gap_ext2;
    int max_score = MAX2(qlen * abpt->max_mat, len * abpt->gap_ext1 + abpt->gap_open1);
    return max_score <= INT16_MAX - abpt->min_mis - oe1 - oe2 ? 16 : 32;
}

int poa_oracle_align_sequence_to_subgraph(abpoa_t *ab, abpoa_para_t *abpt, int beg_node_id, int end_node_id,
                                          uint8_t *query, int qlen, abpoa_res_t *res, poa_oracle_info *info) {
    abpoa_graph_t *g = ab->abg;
    const int beg_index = g->node_id_to_index[beg_node_id], end_index = g->node_id_to_index[end_node_id];
    const int gn = end_index - beg_index + 1, m = abpt->m, *mat = abpt->mat;
    const int gap = abpt->gap_mode, mode = abpt->align_mode;
    const int e1 = abpt->gap_ext1, o1 = abpt->gap_open1, oe1 = o1 + e1;
    const int e2 = abpt->gap_ext2, o2 = abpt->gap_open2, oe2 = o2 + e2;
    const int banded = abpt->wb >= 0;
    const int w = banded ? abpt->wb + (int)(abpt->wf * qlen) : qlen;               /* :474 */
    const int pn = poa_oracle_score_bits(abpt, qlen, gn) == 16 ? 16 : 8;            /* AVX2 lanes */
    const int np_planes = gap == ABPOA_LINEAR_GAP ? 1 : (gap == ABPOA_AFFINE_GAP ? 3 : 5);
    /* banded linear gaps outside local mode: the specification IS the reference's vector procedure */
    const int lg_exact = gap == ABPOA_LINEAR_GAP && mode != ABPOA_LOCAL_MODE && banded;
    const int log_n = pn == 16 ? 4 : 3;

    /* rows reachable from the begin node: reference :1257-1269 */
    uint8_t *live = (uint8_t *)calloc((size_t)g->node_n, 1);
    live[beg_index] = live[end_index] = 1;
    for (int i = beg_index; i < end_index - 1; ++i) {
        if (!live[i]) continue;
        const abpoa_node_t *nd = &g->node[g->index_to_node_id[i]];
        for (int e = 0; e < nd->out_edge_n; ++e) live[g->node_id_to_index[nd->out_id[e]]] = 1;
    }
    /* predecessor rows in in_id order, filtered: reference :548-559 */
    int **pre = (int **)calloc((size_t)gn, sizeof(int *)), **pre_k = (int **)calloc((size_t)gn, sizeof(int *));
    int *pre_n = (int *)calloc((size_t)gn, sizeof(int));
    for (int r = 1; r < gn; ++r) {
        const int id = g->index_to_node_id[beg_index + r];
        const abpoa_node_t *nd = &g->node[id];
        pre[r] = (int *)xm((size_t)nd->in_edge_n * sizeof(int)); pre_k[r] = (int *)xm((size_t)nd->in_edge_n * sizeof(int));
        for (int k = 0; k < nd->in_edge_n; ++k) {
            int pi = g->node_id_to_index[nd->in_id[k]];
            if (live[pi]) { pre[r][pre_n[r]] = pi - beg_index; pre_k[r][pre_n[r]] = k; ++pre_n[r]; }
        }
    }
    orow_t *rows = (orow_t *)calloc((size_t)gn, sizeof(orow_t));
    int64_t cells = 0;
#define REMAIN(id) (g->node_id_to_max_remain[id] - g->node_id_to_max_remain[end_node_id] - 1)
#define BAND_BEG(id) MAX2(0, MIN2(g->node_id_to_max_pos_left[id], qlen - REMAIN(id)) - w)        /* abpoa_align.h:34 */
#define BAND_END(id) MIN2(qlen, MAX2(g->node_id_to_max_pos_right[id], qlen - REMAIN(id)) + w)     /* abpoa_align.h:35 */

    /* ---- first row: reference :582-688 ---- */
    {
        orow_t *r0 = &rows[0];
        r0->beg = 0; r0->end = qlen;
        if (banded) {
            g->node_id_to_max_pos_left[beg_node_id] = g->node_id_to_max_pos_right[beg_node_id] = 0;
            const abpoa_node_t *nd = &g->node[beg_node_id];
            for (int e = 0; e < nd->out_edge_n; ++e)
                if (live[g->node_id_to_index[nd->out_id[e]]])
                    g->node_id_to_max_pos_left[nd->out_id[e]] = g->node_id_to_max_pos_right[nd->out_id[e]] = 1;
            r0->end = BAND_END(beg_node_id);
        }
        const int wd = r0->end + 1;
        int *buf = (int *)xm((size_t)np_planes * wd * sizeof(int));
        r0->h = buf;
        if (np_planes >= 3) { r0->e1 = buf + wd; r0->f1 = buf + 2 * wd; }
        if (np_planes == 5) { r0->e2 = buf + 2 * wd; r0->f1 = buf + 3 * wd; r0->f2 = buf + 4 * wd; }
        for (int j = 0; j < wd; ++j) {
            if (mode == ABPOA_LOCAL_MODE) { for (int p = 0; p < np_planes; ++p) buf[p * wd + j] = 0; continue; }
            if (gap == ABPOA_LINEAR_GAP) r0->h[j] = -e1 * j;
            else if (j == 0) { r0->h[0] = 0; r0->e1[0] = -oe1; r0->f1[0] = NINF; if (r0->e2) { r0->e2[0] = -oe2; r0->f2[0] = NINF; } }
            else {
                r0->f1[j] = -o1 - e1 * j; r0->e1[j] = NINF; r0->h[j] = r0->f1[j];
                if (r0->e2) { r0->f2[j] = -o2 - e2 * j; r0->e2[j] = NINF; r0->h[j] = MAX2(r0->f1[j], r0->f2[j]); }
            }
        }
        cells += wd;
        if (lg_exact) { r0->hx = r0->h; r0->xbeg = 0; r0->xend = r0->end; }    /* beyond end0 the reference's first row holds -inf (:641-647) */
        if (info && info->row_cb) info->row_cb(info->row_user, 0, r0->beg, r0->end, r0->h, r0->e1, r0->e2, r0->f1, r0->f2);
    }

    int best_score = NINF, best_i = 0, best_j = 0, best_id = 0;
    /* ---- rows in topological order: reference drivers :1134-1231 ---- */
    for (int i = 1; i < gn - 1; ++i) {
        if (!live[beg_index + i]) continue;
        const int id = g->index_to_node_id[beg_index + i];
        const abpoa_node_t *nd = &g->node[id];
        orow_t *r = &rows[i];
        int beg = 0, end = qlen;
        if (banded) {                                                            /* band: e.g. :823-839 */
            beg = BAND_BEG(id); end = BAND_END(id);
            int min_pre_beg = INT32_MAX;
            for (int k = 0; k < pre_n[i]; ++k) if (rows[pre[i][k]].beg < min_pre_beg) min_pre_beg = rows[pre[i][k]].beg;
            if (pre_n[i] > 0 && beg / pn < min_pre_beg / pn) beg = min_pre_beg;
        }
        r->beg = beg; r->end = end;
        const int wd = end >= beg ? end - beg + 1 : 0;
        if (lg_exact && wd > 0) {
            int *ps_of = NULL;
            if (abpt->inc_path_score) { ps_of = (int *)xm((size_t)MAX2(pre_n[i], 1) * sizeof(int)); for (int k = 0; k < pre_n[i]; ++k) ps_of[k] = path_score(g, id, pre_k[i][k]); }
            lg_vector_row(r, rows, pre[i], pre_n[i], ps_of, mat + m * nd->base, query, qlen, pn, log_n, e1);
            free(ps_of);
            cells += wd;
            goto row_done;
        }
        int *buf = (int *)xm((size_t)np_planes * MAX2(wd, 1) * sizeof(int));
        r->h = buf;
        if (np_planes >= 3) { r->e1 = buf + wd; r->f1 = buf + 2 * wd; }
        if (np_planes == 5) { r->e2 = buf + 2 * wd; r->f1 = buf + 3 * wd; r->f2 = buf + 4 * wd; }
        cells += wd;
        const int *srow = mat + m * nd->base;
        int f1 = NINF, f2 = NINF, prevT = NINF, prevH = NINF;                    /* left neighbour state */
        for (int j = beg; j <= end; ++j) {
            int M = NINF, E1in = NINF, E2in = NINF, V = NINF;
            for (int k = 0; k < pre_n[i]; ++k) {
                const orow_t *p = &rows[pre[i][k]];
                const int ps = abpt->inc_path_score ? path_score(g, id, pre_k[i][k]) : 0;
                int hd = (j >= 1) ? cell(p->h, p, j - 1) : NINF;
                if (mode == ABPOA_LOCAL_MODE && j == 0) hd = 0;                  /* H[p][-1] = 0 (local) */
                M = MAX2(M, addinf(hd, ps));
                if (gap == ABPOA_LINEAR_GAP) V = MAX2(V, addinf(cell(p->h, p, j), -e1 + ps));
                else {
                    E1in = MAX2(E1in, addinf(cell(p->e1, p, j), ps));
                    if (gap == ABPOA_CONVEX_GAP) E2in = MAX2(E2in, addinf(cell(p->e2, p, j), ps));
                }
            }
            const int s = (j == 0) ? 0 : srow[query[j - 1]];                     /* query profile, qp[.][0] = 0 (:533-539) */
            const int Hm = addinf(M, s);
            const int x = j - beg;
            if (gap == ABPOA_LINEAR_GAP) {                                       /* :727-815 (textbook form) */
                int h = MAX2(Hm, V);
                h = MAX2(h, addinf(prevH, -e1));
                prevH = h;                                                       /* scan runs on the un-floored value */
                r->h[x] = (mode == ABPOA_LOCAL_MODE) ? MAX2(h, 0) : h;
            } else if (gap == ABPOA_AFFINE_GAP) {                                /* :898-931 */
                f1 = MAX2(addinf(prevT, -oe1), addinf(f1, -e1));                 /* opens from the M-only value */
                if (j == beg) f1 = NINF;
                const int T = MAX2(Hm, E1in);
                int h = MAX2(T, f1);
                if (mode == ABPOA_LOCAL_MODE) h = MAX2(h, 0);
                r->f1[x] = f1; r->h[x] = h;
                r->e1[x] = (h == T) ? MAX2(addinf(E1in, -e1), addinf(h, -oe1)) : (mode == ABPOA_LOCAL_MODE ? 0 : NINF);
                prevT = Hm;
            } else {                                                             /* :1032-1072 */
                const int T = MAX2(Hm, MAX2(E1in, E2in));
                f1 = MAX2(addinf(prevT, -oe1), addinf(f1, -e1));
                f2 = MAX2(addinf(prevT, -oe2), addinf(f2, -e2));
                if (j == beg) f1 = f2 = NINF;
                int h = MAX2(T, MAX2(f1, f2));
                if (mode == ABPOA_LOCAL_MODE) h = MAX2(h, 0);
                int x1 = MAX2(addinf(E1in, -e1), addinf(h, -oe1)), x2 = MAX2(addinf(E2in, -e2), addinf(h, -oe2));
                if (mode == ABPOA_LOCAL_MODE) { x1 = MAX2(x1, 0); x2 = MAX2(x2, 0); }
                r->f1[x] = f1; r->f2[x] = f2; r->h[x] = h; r->e1[x] = x1; r->e2[x] = x2;
                prevT = T;
            }
        }
row_done:
        if (info && info->row_cb) info->row_cb(info->row_user, i, beg, end, r->h, r->e1, r->e2, r->f1, r->f2);
        /* row maximum, first / last arg-max: reference :1107-1119 */
        int mx = NINF, left = -1, right = -1;
        if (banded || mode != ABPOA_GLOBAL_MODE) {
            for (int j = beg; j <= end; ++j) {
                const int v = r->h[j - beg];
                if (v > mx) { mx = v; left = right = j; } else if (v == mx && left >= 0) right = j;
            }
        }
        if (mode == ABPOA_LOCAL_MODE) { if (mx > best_score) { best_score = mx; best_i = i; best_j = left; } }
        else if (mode == ABPOA_EXTEND_MODE) {                                    /* :1082-1090 */
            if (mx > best_score) { best_score = mx; best_i = i; best_j = right; best_id = id; }
            else if (abpt->zdrop > 0) {
                int delta = g->node_id_to_max_remain[best_id] - g->node_id_to_max_remain[id];
                if (best_score - mx > abpt->zdrop + e1 * abs(delta - (right - best_j))) break;
            }
        }
        if (banded)                                                              /* band hints: :1121-1130 */
            for (int e = 0; e < nd->out_edge_n; ++e) {
                const int o = nd->out_id[e];
                if (right + 1 > g->node_id_to_max_pos_right[o]) g->node_id_to_max_pos_right[o] = right + 1;
                if (left + 1 < g->node_id_to_max_pos_left[o]) g->node_id_to_max_pos_left[o] = left + 1;
            }
    }
    /* ---- global: end cell among the sink's predecessors: reference :1092-1105 ---- */
    if (mode == ABPOA_GLOBAL_MODE)
        for (int k = 0; k < pre_n[gn - 1]; ++k) {
            const int pi = pre[gn - 1][k]; const orow_t *p = &rows[pi];
            const int endc = qlen > p->end ? p->end : qlen;
            const int v = cell(p->h, p, endc);
            if (v > best_score) { best_score = v; best_i = pi; best_j = endc; }
        }
    res->best_score = best_score;

    /* ---- backtrace: reference :116-458 ---- */
    if (abpt->ret_cigar) {
        cig_t cg = { 0, 0, NULL };
        int i = best_i, j = best_j, start_i = best_i, start_j = best_j, cur = OP_ALL;
        int gap_at_end = abpt->put_gap_at_end; const int gap_on_right = abpt->put_gap_on_right;
        if (best_j < qlen) push(&cg, ABPOA_CINS, qlen - best_j, -1, qlen - 1);
        while (i > 0 && j > 0) {
            const orow_t *r = &rows[i];
            const int hij = cell(r->h, r, j);
            if (mode == ABPOA_LOCAL_MODE && hij == 0) break;
            start_i = i; start_j = j;
            const int id = g->index_to_node_id[beg_index + i];
            const int s = mat[m * g->node[id].base + query[j - 1]];
            const int is_match = g->node[id].base == query[j - 1];
            int hit = 0;
            /* (1) diagonal move, unless gaps are to be preferred first (-R / -J) */
            if (!gap_on_right && !gap_at_end && (gap == ABPOA_LINEAR_GAP || (cur & OP_M))) {
                const int k = find_diag(g, abpt, rows, pre[i], pre_k[i], pre_n[i], id, j, s, hij);
                if (k >= 0) { push(&cg, ABPOA_CMATCH, 1, id, j - 1); i = pre[i][k]; --j; hit = 1; cur = OP_ALL;
                              ++res->n_aln_bases; res->n_matched_bases += is_match; }
            }
            /* (2) deletion: come from (p, j) */
            if (!hit && (gap == ABPOA_LINEAR_GAP || (cur & OP_E)))
                for (int k = 0; k < pre_n[i] && !hit; ++k) {
                    const orow_t *p = &rows[pre[i][k]];
                    const int ps = abpt->inc_path_score ? path_score(g, id, pre_k[i][k]) : 0;
                    if (j < p->beg || j > p->end) continue;
                    if (gap == ABPOA_LINEAR_GAP) {
                        if (cell(p->h, p, j) - e1 + ps == hij) hit = 1;
                    } else {
                        if (cur & OP_E1) {
                            const int ok = (cur & OP_M) ? (hij == cell(p->e1, p, j) + ps) : (cell(r->e1, r, j) == cell(p->e1, p, j) - e1 + ps);
                            if (ok) { cur = (cell(p->h, p, j) - oe1 == cell(p->e1, p, j)) ? (OP_M | OP_F) : OP_E1; hit = 1; }
                        }
                        if (!hit && gap == ABPOA_CONVEX_GAP && (cur & OP_E2)) {
                            const int ok = (cur & OP_M) ? (hij == cell(p->e2, p, j) + ps) : (cell(r->e2, r, j) == cell(p->e2, p, j) - e2 + ps);
                            if (ok) { cur = (cell(p->h, p, j) - oe2 == cell(p->e2, p, j)) ? (OP_M | OP_F) : OP_E2; hit = 1; }
                        }
                    }
                    if (hit) { push(&cg, ABPOA_CDEL, 1, id, j - 1); i = pre[i][k]; gap_at_end = 0; }
                }
            /* (3) insertion: come from (i, j-1) */
            if (!hit && (gap == ABPOA_LINEAR_GAP || (cur & OP_F))) {
                if (gap == ABPOA_LINEAR_GAP) { if (cell(r->h, r, j - 1) - e1 == hij) hit = 1; }
                else {
                    if (gap == ABPOA_AFFINE_GAP || (cur & OP_F1)) {
                        const int fij = cell(r->f1, r, j);
                        if (!(cur & OP_M) || hij == fij) {
                            if (cell(r->h, r, j - 1) - oe1 == fij) { cur = OP_M | OP_E; hit = 1; }
                            else if (cell(r->f1, r, j - 1) - e1 == fij) { cur = OP_F1; hit = 1; }
                        }
                    }
                    if (!hit && gap == ABPOA_CONVEX_GAP && (cur & OP_F2)) {
                        const int fij = cell(r->f2, r, j);
                        if (!(cur & OP_M) || hij == fij) {
                            if (cell(r->h, r, j - 1) - oe2 == fij) { cur = OP_M | OP_E; hit = 1; }
                            else if (cell(r->f2, r, j - 1) - e2 == fij) { cur = OP_F2; hit = 1; }
                        }
                    }
                }
                if (hit) { push(&cg, ABPOA_CINS, 1, id, j - 1); --j; gap_at_end = 0; ++res->n_aln_bases; }
            }
            /* (4) diagonal move as the fallback */
            if (!hit && (gap == ABPOA_LINEAR_GAP || (cur & OP_M))) {
                const int k = find_diag(g, abpt, rows, pre[i], pre_k[i], pre_n[i], id, j, s, hij);
                if (k >= 0) { push(&cg, ABPOA_CMATCH, 1, id, j - 1); i = pre[i][k]; --j; hit = 1; cur = OP_ALL; gap_at_end = 0;
                              ++res->n_aln_bases; res->n_matched_bases += is_match; }
            }
            if (!hit) { fprintf(stderr, "[poa_oracle] Error in backtrack (row %d, j %d, state 0x%x).\n", i, j, cur); exit(1); }
        }
        if (j > 0) push(&cg, ABPOA_CINS, j, -1, j - 1);
        if (!abpt->rev_cigar)
            for (int a = 0, b = cg.n - 1; a < b; ++a, --b) { abpoa_cigar_t t = cg.a[a]; cg.a[a] = cg.a[b]; cg.a[b] = t; }
        res->graph_cigar = cg.a; res->n_cigar = cg.n; res->m_cigar = cg.m;
        res->node_e = g->index_to_node_id[beg_index + best_i]; res->query_e = best_j - 1;
        res->node_s = g->index_to_node_id[beg_index + start_i]; res->query_s = start_j - 1;
    }
    if (info) {
        info->cells = cells; info->n_rows = gn; info->best_i = best_i; info->best_j = best_j;
        if (info->dp_beg && info->dp_end)
            for (int i = 0; i < gn - 1 && i < info->band_cap; ++i) { info->dp_beg[i] = rows[i].beg; info->dp_end[i] = rows[i].end; }
    }
    for (int i = 0; i < gn; ++i) { free(rows[i].alloc ? rows[i].alloc : (void *)rows[i].h); free(pre[i]); free(pre_k[i]); }
    free(rows); free(pre); free(pre_k); free(pre_n); free(live);
    return 0;
}
