/* poa_flat.c -- flatten the host graph into the device job blob.
 *
 * Mirrors the set-up half of the reference's DP entry
 * (src/abpoa_align_simd.c:1257-1269 index_map, :463-560 query profile / pre_index,
 *  :1293-1303 score width) -- but produces one contiguous, 16 B-aligned byte blob that
 * is copied to HBM with a single transfer (layout: PoaJobHeader in poa_device.cuh).
 */
#include "poa_internal.h"
#include "poa_device.cuh"

static inline size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }

/* w = wb + (int)(wf * qlen) with wf a C float, exactly as the reference evaluates it */
int poa_band_halfwidth(const abpoa_para_t *abpt, int qlen) {
    return abpt->wb < 0 ? -1 : abpt->wb + (int)(abpt->wf * qlen);
}

int poa_score_bits(const abpoa_para_t *abpt, int qlen, int n_rows) {
    const int len = qlen > n_rows ? qlen : n_rows;
    const int oe1 = abpt->gap_open1 + abpt->gap_ext1, oe2 = abpt->gap_open2 + abpt->gap_ext2;
    const int max_score = POA_MAX(qlen * abpt->max_mat, len * abpt->gap_ext1 + abpt->gap_open1);
    return max_score <= INT16_MAX - abpt->min_mis - oe1 - oe2 ? 16 : 32;
}

void poa_blob_plan_make(poa_blob_plan *pl, const abpoa_graph_t *abg, const abpoa_para_t *abpt,
                        int beg_node_id, int end_node_id, int qlen) {
    const int beg_index = abg->node_id_to_index[beg_node_id], end_index = abg->node_id_to_index[end_node_id];
    pl->beg_index = beg_index;
    pl->n_rows = end_index - beg_index + 1;
    pl->qlen = qlen;
    pl->whole_graph = (beg_node_id == ABPOA_SRC_NODE_ID && end_node_id == ABPOA_SINK_NODE_ID);
    pl->w = poa_band_halfwidth(abpt, qlen);
    pl->with_remain = (abpt->wb >= 0 || abpt->zdrop > 0);
    pl->with_score = abpt->inc_path_score ? 1 : 0;
    int n_pred = 0;
    if (pl->whole_graph) n_pred = (int)poa_graph_edge_count(abg);          /* every in-edge exactly once */
    else { const int *cin = poa_graph_in_degrees(abg); for (int r = 1; r < pl->n_rows; ++r) n_pred += cin[abg->index_to_node_id[beg_index + r]]; }
    pl->n_pred_max = n_pred;
    const size_t nr = (size_t)pl->n_rows;
    size_t b = al16(sizeof(struct PoaJobHeader));
    b += al16((nr + 1) * 8);                         /* rowmeta   */
    b += al16((size_t)n_pred * 4 + 4);               /* pred      */
    if (pl->with_score) b += al16((size_t)n_pred * 4 + 4);
    if (!pl->whole_graph) b += al16(nr);             /* live mask */
    b += al16((size_t)qlen + 1) + 16;                /* shifted query + one spare vector */
    pl->bytes = b;
}

void poa_blob_fill(uint8_t *dst, const poa_blob_plan *pl, const abpoa_graph_t *abg, const abpoa_para_t *abpt,
                   int beg_node_id, int end_node_id, const uint8_t *query) {
    struct PoaJobHeader *h = (struct PoaJobHeader *)dst;
    const int n_rows = pl->n_rows, beg_index = pl->beg_index, qlen = pl->qlen;
    const size_t nr = (size_t)n_rows;
    size_t off = al16(sizeof *h);
    memset(h, 0, sizeof *h);
    h->n_rows = n_rows; h->qlen = qlen; h->w = pl->w; h->node_n = abg->node_n;
    h->pn = poa_score_bits(abpt, qlen, n_rows) == 16 ? 16 : 8;
    h->off_rowmeta = (int32_t)off; off += al16((nr + 1) * 8);
    h->off_pred = (int32_t)off; off += al16((size_t)pl->n_pred_max * 4 + 4);
    h->off_predscore = -1;
    if (pl->with_score) { h->off_predscore = (int32_t)off; off += al16((size_t)pl->n_pred_max * 4 + 4); }
    h->off_live = -1;
    if (!pl->whole_graph) { h->off_live = (int32_t)off; off += al16(nr); }
    h->off_qs = (int32_t)off; off += al16((size_t)qlen + 1) + 16;
    h->blob_bytes = (int32_t)off;
    if (off != pl->bytes) poa_die(__func__, "blob size mismatch (%zu vs %zu)", off, pl->bytes);

    int32_t *rowmeta = (int32_t *)(dst + h->off_rowmeta), *pred = (int32_t *)(dst + h->off_pred);
    int32_t *pscore = pl->with_score ? (int32_t *)(dst + h->off_predscore) : NULL;
    uint8_t *live = pl->whole_graph ? NULL : dst + h->off_live;
    uint8_t *qs = dst + h->off_qs;

    /* rows reachable from the begin node inside [beg_index, end_index] */
    if (live) {
        memset(live, 0, nr);
        live[0] = live[n_rows - 1] = 1;
        for (int r = 0; r < n_rows - 2; ++r) {
            if (!live[r]) continue;
            const abpoa_node_t *nd = &abg->node[abg->index_to_node_id[beg_index + r]];
            for (int e = 0; e < nd->out_edge_n; ++e) {
                const int t = abg->node_id_to_index[nd->out_id[e]] - beg_index;
                if (t >= 0 && t < n_rows) live[t] = 1;
            }
        }
    }
    const int with_remain = pl->with_remain;
    const int end_remain = with_remain ? abg->node_id_to_max_remain[end_node_id] : 0;
    const uint8_t *cbase = poa_graph_bases(abg); const int *cin = poa_graph_in_degrees(abg);
    const int *id_of_index = abg->index_to_node_id + beg_index, *index_of_id = abg->node_id_to_index, *max_remain = abg->node_id_to_max_remain;
    int np = 0;
    for (int r = 0; r < n_rows; ++r) {
        const int id = id_of_index[r];
        if (r + 16 < n_rows) {                       /* rows ahead: their degree, residue, band centre and in-edge row */
            const int ia = id_of_index[r + 16];
            __builtin_prefetch(cin + ia, 0); __builtin_prefetch(cbase + ia, 0);
            if (with_remain) __builtin_prefetch(max_remain + ia, 0);
            __builtin_prefetch(poa_graph_in_ids_inline(abg, ia), 0);
        }
        if (r + 8 < n_rows) {                        /* second level: the row numbers of the predecessors of a nearer row */
            const int ib = id_of_index[r + 8];
            const int nb = cin[ib] < 4 ? cin[ib] : 4; const int *ids = poa_graph_in_ids_inline(abg, ib);
            for (int e = 0; e < nb; ++e) __builtin_prefetch(index_of_id + ids[e], 0);
        }
        const int rem = with_remain ? max_remain[id] - end_remain - 1 : 0;
        rowmeta[2 * r] = np;
        rowmeta[2 * r + 1] = (int32_t)((uint32_t)rem << 8) | cbase[id];
        if (r > 0) {
            const int ni = cin[id]; const int *iid = poa_graph_in_ids(abg, id);
            for (int e = 0; e < ni; ++e) {
                const int pr = index_of_id[iid[e]] - beg_index;
                if (pr < 0 || pr >= n_rows) continue;
                if (live && !live[pr]) continue;
                /* -G on a sub-graph: the reference indexes abpoa_get_incre_path_score with the position in the
                 * FILTERED predecessor list (pre_index k, src/abpoa_align_simd.c:134, :215), not with the in-edge
                 * index; for whole-graph alignments the two coincide.  Followed literally for bit-exactness. */
                if (pscore) pscore[np] = poa_edge_path_score(abg, id, np - rowmeta[2 * r]);
                pred[np++] = pr;
            }
        }
    }
    rowmeta[2 * n_rows] = np; rowmeta[2 * n_rows + 1] = 0;
    qs[0] = 0;
    memcpy(qs + 1, query, (size_t)qlen);
    memset(qs + 1 + qlen, 0, (size_t)h->blob_bytes - h->off_qs - 1 - (size_t)qlen);
}

/* May this alignment run on the packed int16x2 kernel?  Upper bounds are static; the lower side
 * is additionally watched at run time (POA_ST_RANGE -> the job is redone in 32 bits). */
int poa_p16_ok(const abpoa_para_t *abpt, int qlen, int n_rows) {
    const int oe1 = abpt->gap_open1 + abpt->gap_ext1, oe2 = abpt->gap_open2 + abpt->gap_ext2;
    const int emax = POA_MAX(abpt->gap_ext1, abpt->gap_ext2), oemax = POA_MAX(oe1, oe2);
    if (abpt->max_mat <= 0 || (int64_t)qlen * abpt->max_mat > 28000) return 0;
    if (abpt->min_mis > 1000 || abpt->max_mat > 1000 || oemax > 1000 || emax > 100) return 0;
    if (abpt->wb < 0) {
        if (abpt->align_mode != ABPOA_LOCAL_MODE) {          /* unbanded global / extend: column 0 sinks by e per row */
            const int64_t len = qlen > n_rows ? qlen : n_rows;
            if (len * abpt->gap_ext1 + abpt->gap_open1 > 26000) return 0;
        }
    } else {
        const int64_t w = poa_band_halfwidth(abpt, qlen);
        if (2 * w * emax + oemax > 12000) return 0;          /* worst in-band cell relative to the row maximum */
    }
    return 1;
}

/* debugging aid (tests): the job blob poa_blob_fill builds for aligning `query` to the whole graph of `ab`,
 * copied into `out` (capacity `cap`); returns its size or -1 */
int poa_debug_blob(abpoa_t *ab, abpoa_para_t *abpt, const uint8_t *query, int qlen, uint8_t *out, int cap) {
    abpoa_graph_t *abg = ab->abg;
    if (abg->node_n <= 2) return -1;
    if (!abg->is_topological_sorted) abpoa_topological_sort(abg, abpt);
    poa_blob_plan pl;
    poa_blob_plan_make(&pl, abg, abpt, ABPOA_SRC_NODE_ID, ABPOA_SINK_NODE_ID, qlen);
    if ((size_t)cap < pl.bytes) return -1;
    poa_blob_fill(out, &pl, abg, abpt, ABPOA_SRC_NODE_ID, ABPOA_SINK_NODE_ID, query);
    return (int)pl.bytes;
}
