/* poa_oracle.h -- interface of the scalar CPU restatement (TEST INFRASTRUCTURE ONLY). */
#ifndef POA_ORACLE_H
#define POA_ORACLE_H
#include <stdint.h>
#include "abpoa.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct {
    int64_t cells;            /* sum over DP rows 0..gn-2 of end - beg + 1               */
    int n_rows, best_i, best_j;
    int *dp_beg, *dp_end;     /* optional caller buffers [band_cap]: band of every row  */
    int band_cap;
    /* optional: called once per computed DP row with its planes (NULL for absent planes) */
    void (*row_cb)(void *user, int row, int beg, int end, const int *h, const int *e1, const int *e2, const int *f1, const int *f2);
    void *row_user;
} poa_oracle_info;

/* same contract as the reference seam simd_abpoa_align_sequence_to_subgraph
 * (src/abpoa_align_simd.h:12): graph must be topologically sorted; mutates the graph's
 * max_pos_left/right scratch; res->graph_cigar is malloc'ed. */
int poa_oracle_align_sequence_to_subgraph(abpoa_t *ab, abpoa_para_t *abpt, int beg_node_id, int end_node_id,
                                          uint8_t *query, int qlen, abpoa_res_t *res, poa_oracle_info *info);
int poa_oracle_score_bits(const abpoa_para_t *abpt, int qlen, int gn);
#ifdef __cplusplus
}
#endif
#endif
