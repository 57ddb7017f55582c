/* poa_para.c -- parameter object, scoring matrices and alphabet tables.
 *
 * Host-side mirror of the reference's parameter handling so that callers configure
 * the GPU engine exactly as they configure abPOA:
 *   abpoa_init_para        reference src/abpoa_align.c:100-157 (same defaults)
 *   abpoa_post_set_para    reference src/abpoa_align.c:159-184 (derived fields)
 *   abpoa_set_mat_from_file reference src/abpoa_align.c:61-85  (matrix text format)
 *   alphabet tables        reference src/abpoa_seq.c:15-98     (values are ABI data)
 */
#include <ctype.h>
#include <stdarg.h>
#include "poa_internal.h"

/* ---------------------------------------------------------------- fatal helpers */
void poa_die(const char *where, const char *fmt, ...) {
    va_list ap;
    fflush(stdout);
    fprintf(stderr, "[%s] ", where);
    va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap);
    fputc('\n', stderr);
    exit(EXIT_FAILURE);
}
void *poa_xmalloc(size_t n) {
    void *p = malloc(n ? n : 1);
    if (!p) poa_die("poa_xmalloc", "out of memory (%zu bytes)", n);
    return p;
}
void *poa_xcalloc(size_t n, size_t sz) {
    void *p = calloc(n ? n : 1, sz ? sz : 1);
    if (!p) poa_die("poa_xcalloc", "out of memory (%zu x %zu bytes)", n, sz);
    return p;
}
void *poa_xrealloc(void *p, size_t n) {
    void *q = realloc(p, n ? n : 1);
    if (!q) poa_die("poa_xrealloc", "out of memory (%zu bytes)", n);
    return q;
}

/* ---------------------------------------------------------------- alphabet tables
 * Exported under the reference's names because bindings read them directly
 * (python/cabpoa.pxd).  The mapping itself is data fixed by the interface:
 *   nucleotides  A C G T(U) N -> 0..4, everything else 4
 *   amino acids  the 26 letters in the order "ACGTNBDEFHIJKLMOPQRSUVWXYZ" -> 0..25,
 *                everything else 26
 * and the inverse tables print code m (the MSA gap) as '-'. */
unsigned char ab_nt4_table[256];
char ab_nt256_table[256];
unsigned char ab_aa26_table[256];
char ab_aa256_table[256];
char ab_char26_table[256];
char ab_char256_table[256];

static const char POA_AA_ORDER[] = "ACGTNBDEFHIJKLMOPQRSUVWXYZ";

static void __attribute__((constructor)) poa_build_alphabets(void) {
    int c, k;
    /* the batch engine drives up to 32 CUDA streams: give each its own hardware work queue
     * (read by the driver when the context is created; a user setting wins) */
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    /* Load every kernel when the module loads: with lazy loading the FIRST launch of a kernel has to
     * synchronise the context, which can never complete while the resident kernel is running. */
    setenv("CUDA_MODULE_LOADING", "EAGER", 0);
    for (c = 0; c < 256; ++c) {
        ab_nt4_table[c] = 4; ab_nt256_table[c] = 'N';
        ab_aa26_table[c] = 26; ab_aa256_table[c] = '*';
    }
    /* already-encoded input passes through unchanged */
    for (c = 0; c < 4; ++c) ab_nt4_table[c] = (unsigned char)c;
    for (c = 0; c < 26; ++c) ab_aa26_table[c] = (unsigned char)c;
    /* letters -> codes */
    for (k = 0; k < 4; ++k) {
        ab_nt4_table[(int)"ACGT"[k]] = (unsigned char)k;
        ab_nt4_table[tolower("ACGT"[k])] = (unsigned char)k;
    }
    ab_nt4_table['U'] = ab_nt4_table['u'] = 3;
    for (k = 0; k < 26; ++k) {
        ab_aa26_table[(int)POA_AA_ORDER[k]] = (unsigned char)k;
        ab_aa26_table[tolower(POA_AA_ORDER[k])] = (unsigned char)k;
    }
    /* codes -> letters */
    for (k = 0; k < 4; ++k) ab_nt256_table[k] = "ACGT"[k];
    ab_nt256_table[5] = '-'; ab_nt256_table[27] = '-';
    for (k = 0; k < 26; ++k) ab_aa256_table[k] = POA_AA_ORDER[k];
    ab_aa256_table[27] = '-';
    /* letters -> canonical letters */
    for (k = 0; k < 4; ++k) {
        ab_nt256_table[(int)"ACGT"[k]] = "ACGT"[k];
        ab_nt256_table[tolower("ACGT"[k])] = "ACGT"[k];
    }
    ab_nt256_table['U'] = ab_nt256_table['u'] = 'T';
    for (c = 'A'; c <= 'Z'; ++c) {
        ab_aa256_table[c] = (char)c; ab_aa256_table[tolower(c)] = (char)c;
    }
    /* until abpoa_post_set_para() picks an alphabet, behave as nucleotide */
    memcpy(ab_char26_table, ab_nt4_table, 256);
    memcpy(ab_char256_table, ab_nt256_table, 256);
}

/* ---------------------------------------------------------------- scoring */
/* match/mismatch matrix with a neutral last residue (N / '*'): reference
 * src/abpoa_align.c:12-25 */
static void poa_fill_simple_matrix(abpoa_para_t *abpt) {
    const int m = abpt->m;
    const int ma = abs(abpt->match), mi = -abs(abpt->mismatch);
    for (int r = 0; r < m; ++r)
        for (int c = 0; c < m; ++c)
            abpt->mat[r * m + c] = (r == m - 1 || c == m - 1) ? 0 : (r == c ? ma : mi);
    abpt->max_mat = ma;
    abpt->min_mis = -mi;
}

/* Matrix text format: '#' comments; first data line lists the column residues; each
 * following line is "<row residue> <score> <score> ...".  Scores land at
 * mat[row * m + column_code]. */
void abpoa_set_mat_from_file(abpoa_para_t *abpt, char *mat_fn) {
    FILE *fp = fopen(mat_fn, "r");
    if (!fp) poa_die(__func__, "Unable to open scoring matrix file: \"%s\"", mat_fn);
    const int m = abpt->m;
    int *col_code = (int *)poa_xmalloc(256 * sizeof(int)), n_col = 0, have_header = 0;
    char line[4096];
    while (fgets(line, sizeof line, fp)) {
        if (line[0] == '#') continue;
        if (!have_header) {
            for (char *p = line; *p; ++p)
                if (!isspace((unsigned char)*p) && n_col < 256) col_code[n_col++] = ab_char26_table[(unsigned char)*p];
            have_header = 1;
            continue;
        }
        char *p = line; int row = -1, n = 0;
        while (*p) {
            if (!isalpha((unsigned char)*p) && !isdigit((unsigned char)*p) && *p != '+' && *p != '-') { ++p; continue; }
            if (row < 0) {
                row = ab_char26_table[(unsigned char)*p];
                if (row >= m) poa_die(__func__, "Unknown base: \"%c\" (%d).", *p, row);
                ++p;
            } else {
                if (n == m || n >= n_col) poa_die(__func__, "Too many scores in matrix.");
                char *end; long s = strtol(p, &end, 10);
                if (end == p) { ++p; continue; }
                if (col_code[n] < m) abpt->mat[row * m + col_code[n]] = (int)s;
                ++n; p = end;
            }
        }
    }
    abpt->min_mis = 0; abpt->max_mat = 0;
    for (int i = 0; i < m * m; ++i) {
        if (abpt->mat[i] > abpt->max_mat) abpt->max_mat = abpt->mat[i];
        if (-abpt->mat[i] > abpt->min_mis) abpt->min_mis = -abpt->mat[i];
    }
    free(col_code); fclose(fp);
}

/* gap model from the open penalties: O1 == 0 linear, O2 == 0 affine, else convex
 * (reference src/abpoa_align.c:87-98) */
static void poa_pick_gap_mode(abpoa_para_t *abpt) {
    if (abpt->match < 0 || abpt->mismatch < 0 || abpt->gap_open1 < 0 || abpt->gap_open2 < 0 ||
        abpt->gap_ext1 < 0 || abpt->gap_ext2 < 0)
        poa_die("abpoa_set_gap_mode", "Invalid negative scoring parameters: match=%d, mismatch=%d, gap_open1=%d, gap_open2=%d, gap_ext1=%d, gap_ext2=%d.",
                abpt->match, abpt->mismatch, abpt->gap_open1, abpt->gap_open2, abpt->gap_ext1, abpt->gap_ext2);
    if (abpt->gap_ext1 == 0 && abpt->gap_ext2 == 0)
        poa_die("abpoa_set_gap_mode", "Invalid gap extension parameters, expect at least one postive: gap_ext1=%d, gap_ext2=%d.",
                abpt->gap_ext1, abpt->gap_ext2);
    if (abpt->gap_open1 == 0) abpt->gap_mode = ABPOA_LINEAR_GAP;
    else if (abpt->gap_open2 == 0) abpt->gap_mode = ABPOA_AFFINE_GAP;
    else abpt->gap_mode = ABPOA_CONVEX_GAP;
}

abpoa_para_t *abpoa_init_para(void) {
    abpoa_para_t *p = (abpoa_para_t *)poa_xcalloc(1, sizeof(abpoa_para_t));
    /* alignment */
    p->align_mode = ABPOA_GLOBAL_MODE;
    p->gap_mode = ABPOA_CONVEX_GAP;
    p->match = 2; p->mismatch = 4;
    p->gap_open1 = 4; p->gap_ext1 = 2;
    p->gap_open2 = 24; p->gap_ext2 = 1;
    p->wb = ABPOA_EXTRA_B; p->wf = ABPOA_EXTRA_F;
    p->zdrop = -1; p->end_bonus = -1;
    p->ret_cigar = 1;
    /* alphabet / matrix: 5 nucleotide codes until the caller says otherwise */
    p->m = 5;
    p->mat = (int *)poa_xmalloc((size_t)p->m * p->m * sizeof(int));
    /* output */
    p->out_cons = 1;
    p->cons_algrm = ABPOA_HB;
    p->max_n_cons = 1;
    p->min_freq = 0.25;
    /* seeding is present in the struct for ABI reasons only; off by default */
    p->disable_seeding = 1;
    p->k = 19; p->w = 10; p->min_w = 500;
    p->verbose = ABPOA_NONE_VERBOSE;
    return p;
}

void abpoa_post_set_para(abpoa_para_t *abpt) {
    poa_pick_gap_mode(abpt);
    if (abpt->out_msa || abpt->out_gfa || abpt->max_n_cons > 1 || abpt->cons_algrm == ABPOA_MF) {
        abpt->use_read_ids = 1;
        if (abpt->out_msa || abpt->out_gfa || abpt->max_n_cons > 1) poa_set_65536_table();
        if (abpt->max_n_cons > 1 || abpt->cons_algrm == ABPOA_MF) poa_set_bit_table16();
    }
    if (abpt->align_mode == ABPOA_LOCAL_MODE) abpt->wb = -1;   /* local alignment is never banded */
    if (abpt->m > 5) {
        memcpy(ab_char26_table, ab_aa26_table, 256);
        memcpy(ab_char256_table, ab_aa256_table, 256);
        if (abpt->k > 11) { abpt->k = 7; abpt->w = 4; }
    } else {
        memcpy(ab_char26_table, ab_nt4_table, 256);
        memcpy(ab_char256_table, ab_nt256_table, 256);
    }
    if (abpt->use_score_matrix == 0) poa_fill_simple_matrix(abpt);
    else abpoa_set_mat_from_file(abpt, abpt->mat_fn);
}

void abpoa_free_para(abpoa_para_t *abpt) {
    if (!abpt) return;
    free(abpt->mat); free(abpt->mat_fn); free(abpt->out_pog); free(abpt->incr_fn);
    free(abpt);
}
