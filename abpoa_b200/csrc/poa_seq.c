/* poa_seq.c -- per-handle read-name / strand bookkeeping (abpoa_seq_t).
 * Mirrors the container the reference keeps in src/abpoa_seq.c:100-172; only the parts
 * the MSA driver and the writers need (names, is_rc flags) are populated here -- file
 * parsing is outside the hot-path scope. */
#include "poa_internal.h"

#define POA_SEQ_CHUNK 1024

static void seq_zero_range(abpoa_seq_t *abs, int from, int to) {
    for (int i = from; i < to; ++i) {
        abs->seq[i].l = abs->seq[i].m = 0; abs->seq[i].s = NULL;
        abs->name[i].l = abs->name[i].m = 0; abs->name[i].s = NULL;
        abs->comment[i].l = abs->comment[i].m = 0; abs->comment[i].s = NULL;
        abs->qual[i].l = abs->qual[i].m = 0; abs->qual[i].s = NULL;
        abs->is_rc[i] = 0;
    }
}

abpoa_seq_t *poa_seq_new(void) {
    abpoa_seq_t *abs = (abpoa_seq_t *)poa_xmalloc(sizeof(abpoa_seq_t));
    abs->n_seq = 0; abs->m_seq = POA_SEQ_CHUNK;
    abs->seq = (abpoa_str_t *)poa_xmalloc(abs->m_seq * sizeof(abpoa_str_t));
    abs->name = (abpoa_str_t *)poa_xmalloc(abs->m_seq * sizeof(abpoa_str_t));
    abs->comment = (abpoa_str_t *)poa_xmalloc(abs->m_seq * sizeof(abpoa_str_t));
    abs->qual = (abpoa_str_t *)poa_xmalloc(abs->m_seq * sizeof(abpoa_str_t));
    abs->is_rc = (uint8_t *)poa_xmalloc(abs->m_seq);
    seq_zero_range(abs, 0, abs->m_seq);
    return abs;
}

void poa_seq_free(abpoa_seq_t *abs) {
    if (!abs) return;
    for (int i = 0; i < abs->m_seq; ++i) {
        if (abs->seq[i].m > 0) free(abs->seq[i].s);
        if (abs->name[i].m > 0) free(abs->name[i].s);
        if (abs->comment[i].m > 0) free(abs->comment[i].s);
        if (abs->qual[i].m > 0) free(abs->qual[i].s);
    }
    free(abs->seq); free(abs->name); free(abs->comment); free(abs->qual); free(abs->is_rc);
    free(abs);
}

void poa_seq_reserve(abpoa_seq_t *abs) {
    if (abs->n_seq < abs->m_seq) return;
    int m = POA_MAX(abs->n_seq, abs->m_seq << 1);
    abs->seq = (abpoa_str_t *)poa_xrealloc(abs->seq, m * sizeof(abpoa_str_t));
    abs->name = (abpoa_str_t *)poa_xrealloc(abs->name, m * sizeof(abpoa_str_t));
    abs->comment = (abpoa_str_t *)poa_xrealloc(abs->comment, m * sizeof(abpoa_str_t));
    abs->qual = (abpoa_str_t *)poa_xrealloc(abs->qual, m * sizeof(abpoa_str_t));
    abs->is_rc = (uint8_t *)poa_xrealloc(abs->is_rc, m);
    seq_zero_range(abs, abs->m_seq, m);
    abs->m_seq = m;
}

void poa_str_assign(abpoa_str_t *dst, const char *s, int l) {
    if (l <= 0) return;
    dst->s = (char *)(dst->m ? poa_xrealloc(dst->s, l + 1) : poa_xmalloc(l + 1));
    memcpy(dst->s, s, l); dst->s[l] = 0;
    dst->l = l; dst->m = l + 1;
}
