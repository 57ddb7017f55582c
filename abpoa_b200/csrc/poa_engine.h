/* poa_engine.h -- private interface between the abpoa.h / abpoa_gpu.h front ends and the
 * CUDA stream contexts of poa_cuda.cu. */
#ifndef POA_ENGINE_H
#define POA_ENGINE_H
#include "poa_internal.h"

#ifdef __cplusplus
extern "C" {
#endif

/* one sequence-to-graph alignment handed to a stream context */
typedef struct {
    /* in */
    const abpoa_graph_t *abg; int beg_node_id, end_node_id;
    const uint8_t *query;
    poa_blob_plan plan;
    int want_bands;                 /* copy the per-row (beg,end,left,right) back            */
    void *tag;                      /* caller's cookie                                        */
    /* out (valid inside the sink callback only) */
    int status, bits, ref_bits;     /* bits: kernel variant used (15 packed int16, 16, 32); ref_bits: the reference's width */
    int best_score, best_i, best_j, start_i, start_j, n_aln_bases, n_matched_bases, max_band;
    int64_t cells; uint64_t plane_units;
    int n_ops; const uint64_t *ops; /* graph-CIGAR words in backtrack (reversed) order        */
    const int32_t *bands;           /* [n_rows][4] when want_bands                            */
} poa_job;

typedef void (*poa_job_sink)(void *user, poa_job *job);

/* a captured job: the exact bytes that were sent to the device, plus what came back */
typedef struct {
    uint8_t *blob; size_t bytes;
    int n_rows, qlen, w, n_pred, bits;
    int best_score, n_ops; int64_t cells;
    uint64_t plane_units;           /* 8-cell units of plane storage the job really used */
} poa_captured_job;
typedef void (*poa_capture_fn)(void *user, const poa_captured_job *cj);

typedef struct {
    double kernel_ms;               /* CUDA-event time of the alignment kernels               */
    int64_t cells, alignments, launches, retries;
    uint64_t h2d_bytes, d2h_bytes;
    int64_t fwd_clk, bt_clk;            /* SM cycles in the forward DP / backtrace, summed over alignments */
    int64_t prof[6];                    /* -DPOA_KPROF builds: per-phase cycles */
    int64_t diag[4];                    /* -DPOA_KPROF builds: rows on the straight-line path / generic because of np, ring distance, band width */
    double fill_ms, wait_ms, copy_ms;   /* host-side phases of a launch: blob fill, kernel wait, result copies */
} poa_engine_stats;

/* One big HBM region shared by the stream contexts of a batch engine: score planes live
 * only between a launch and its result copy, so streams borrow and return slices. */
typedef struct poa_arena poa_arena;
poa_arena *poa_arena_new(int dev, size_t bytes);
void poa_arena_destroy(poa_arena *a);
size_t poa_arena_capacity(const poa_arena *a);
poa_dev_ctx *poa_dev_ctx_new_on(int dev);
void poa_dev_ctx_use_arena(poa_dev_ctx *c, poa_arena *a);
void poa_dev_ctx_set_capture(poa_dev_ctx *c, poa_capture_fn fn, void *user);
/* called (on the launching thread) before a context waits for plane memory: the owner drains its other launches */
typedef void (*poa_pressure_fn)(void *user);
void poa_dev_ctx_set_pressure_cb(poa_dev_ctx *c, poa_pressure_fn fn, void *user);
/* replay support: run pre-uploaded blobs (device pointers) as one launch on the context's stream */
typedef struct { const uint8_t *d_blob; int n_rows, qlen, w; uint64_t plane_units; } poa_replay_job;
double poa_dev_ctx_replay_launch(poa_dev_ctx *c, const abpoa_para_t *abpt, const poa_replay_job *jobs, int n, int bits,
                                 int32_t *out_score, int32_t *out_nops, int64_t *out_cells);
void poa_dev_ctx_reserve(poa_dev_ctx *c, int jobs, int rows_hint, int qlen_hint);

struct PoaResultDev; struct PoaParamsDev;
uint8_t *poa_arena_borrow(poa_arena *a, size_t bytes);
uint8_t *poa_arena_try_borrow(poa_arena *a, size_t bytes);
void poa_arena_return(poa_arena *a, uint8_t *p, size_t bytes);
void poa_fill_params(struct PoaParamsDev *p, const abpoa_para_t *abpt, int bits);

void poa_engine_run(poa_dev_ctx *c, const abpoa_para_t *abpt, poa_job *jobs, int n, poa_job_sink sink, void *user);
int poa_engine_submit(poa_dev_ctx *c, const abpoa_para_t *abpt, poa_job *jobs, int n);   /* 1 launched, 0 not one launch, -1 arena short */
void poa_engine_collect(poa_dev_ctx *c, poa_job_sink sink, void *user);
void poa_job_to_res(const poa_job *j, const abpoa_para_t *abpt, abpoa_res_t *res);
void poa_dev_ctx_set_planes_limit(poa_dev_ctx *c, size_t bytes);
const poa_engine_stats *poa_dev_ctx_stats(const poa_dev_ctx *c);
void poa_dev_ctx_reset_stats(poa_dev_ctx *c);
int poa_dev_ctx_device(const poa_dev_ctx *c);

#ifdef __cplusplus
}
#endif
#endif
