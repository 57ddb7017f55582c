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
    if (l <= 0) { dst->l = 0; if (dst->m > 0) dst->s[0] = 0; return; }
    dst->s = (char *)(dst->m ? poa_xrealloc(dst->s, l + 1) : poa_xmalloc(l + 1));
    memcpy(dst->s, s, l); dst->s[l] = 0;
    dst->l = l; dst->m = l + 1;
}

/* ------------------------------------------------------------------ FASTA / FASTQ input
 * Reads every record of a (possibly gzip-compressed) FASTA/FASTQ file and appends it to `abs` (name, comment,
 * bases, qualities), with the record grammar of the reference's reader (kseq, used by abpoa_read_seq,
 * reference src/abpoa_seq.c:173-193): a header line starts with '>' or '@', the name ends at the first
 * white space, the rest of the line is the comment; sequence lines run until a line that starts with '>', '@'
 * or '+'; after '+' quality characters are read (across lines) until as many as bases.  Returns the number
 * of records, -1 if the file cannot be opened. */
#include <zlib.h>

static void str_append(abpoa_str_t *d, const char *s, int l) {
    if (l <= 0) return;
    if (d->l + l + 1 > d->m) {
        int m = d->m ? d->m : 64;
        while (m < d->l + l + 1) m <<= 1;
        d->s = (char *)(d->m ? poa_xrealloc(d->s, (size_t)m) : poa_xmalloc((size_t)m));
        d->m = m;
    }
    memcpy(d->s + d->l, s, (size_t)l); d->l += l; d->s[d->l] = 0;
}

int poa_read_fastx(const char *fn, abpoa_seq_t *abs) {
    gzFile fp = (fn == NULL || strcmp(fn, "-") == 0) ? gzdopen(0, "r") : gzopen(fn, "r");
    if (!fp) return -1;
    size_t cap = 1 << 20, len = 0;
    char *buf = (char *)poa_xmalloc(cap);
    for (;;) {
        if (len + (1 << 16) > cap) { cap <<= 1; buf = (char *)poa_xrealloc(buf, cap); }
        const int got = gzread(fp, buf + len, 1 << 16);
        if (got <= 0) break;
        len += (size_t)got;
    }
    gzclose(fp);
    size_t p = 0; int n = 0;
    while (p < len && buf[p] != '>' && buf[p] != '@') ++p;            /* jump to the first header */
    while (p < len) {
        ++p;                                                          /* the '>' / '@' */
        abs->n_seq += 1; poa_seq_reserve(abs);
        const int i = abs->n_seq - 1;
        abs->name[i].l = abs->comment[i].l = abs->seq[i].l = abs->qual[i].l = 0; abs->is_rc[i] = 0;
        size_t q = p;
        while (q < len && buf[q] != '\n' && buf[q] != ' ' && buf[q] != '\t' && buf[q] != '\r') ++q;
        str_append(&abs->name[i], buf + p, (int)(q - p));
        if (q < len && buf[q] != '\n') {                              /* comment: the rest of the header line */
            size_t c0 = q + 1, c1 = c0;
            while (c1 < len && buf[c1] != '\n') ++c1;
            size_t ce = c1; while (ce > c0 && buf[ce - 1] == '\r') --ce;
            str_append(&abs->comment[i], buf + c0, (int)(ce - c0));
            q = c1;
        }
        p = q < len ? q + 1 : len;
        /* sequence lines */
        while (p < len && buf[p] != '>' && buf[p] != '@' && buf[p] != '+') {
            size_t e = p; while (e < len && buf[e] != '\n') ++e;
            size_t t = e; while (t > p && buf[t - 1] == '\r') --t;
            str_append(&abs->seq[i], buf + p, (int)(t - p));
            p = e < len ? e + 1 : len;
        }
        if (p < len && buf[p] == '+') {                               /* FASTQ: skip the '+' line, then the qualities */
            while (p < len && buf[p] != '\n') ++p;
            if (p < len) ++p;
            while (p < len && abs->qual[i].l < abs->seq[i].l) {
                size_t e = p; while (e < len && buf[e] != '\n') ++e;
                size_t t = e; while (t > p && buf[t - 1] == '\r') --t;
                str_append(&abs->qual[i], buf + p, (int)(t - p));
                p = e < len ? e + 1 : len;
            }
            while (p < len && buf[p] != '>' && buf[p] != '@') ++p;    /* to the next header */
        }
        ++n;
    }
    free(buf);
    return n;
}

/* residue letters -> codes with the alphabet table abpoa_post_set_para selected (reference: ab_char26_table lookups in
 * abpoa_msa1, src/abpoa_align.c:499).  A function rather than the table itself: the library is linked -Bsymbolic, so an
 * executable that referenced the table would get its own, never initialised copy (copy relocation). */
extern char ab_char26_table[256];
void poa_encode_residues(const char *s, int l, uint8_t *out) {
    for (int j = 0; j < l; ++j) out[j] = (uint8_t)ab_char26_table[(int)(unsigned char)s[j]];
}
