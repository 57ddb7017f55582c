/* poa_chain_host.h -- host interface of the device-resident chain engine (poa_chain.cu); private. */
#ifndef POA_CHAIN_HOST_H
#define POA_CHAIN_HOST_H
#include <vector>
#include "abpoa_gpu.h"
#include "poa_engine.h"

typedef struct {
    double device_ms;                   /* CUDA-event time from "inputs resident in HBM" to "last fuse kernel done", summed over waves */
    int64_t cells, alignments, launches;
    uint64_t h2d_bytes, d2h_bytes;
    int groups_done, groups_failed;     /* finished on the device / handed to the launch-per-round engine */
    double dp_ms, fuse_ms;              /* per-launch CUDA-event times of the two kernels, summed over rounds AND cohort streams
                                           (cohorts run concurrently: dp_ms + fuse_ms ~ n_cohorts x device_ms) */
    int64_t dp_launches, fuse_launches;
    int64_t fwd_clk, bt_clk;            /* SM cycles inside the forward DP / the backtrace, summed over alignments */
    double wait_ms;                     /* free-running mode: time the alignment warps waited for their fuse tasks, summed over groups
                                           (then dp_ms / fuse_ms are per-group sums of time inside the alignments / inside chain_fuse) */
    int free_running;
} PoaChainStats;

/* may this parameter set run on the device chain at all? */
int poa_chain_eligible(const abpoa_para_t *abpt);
/* run groups[todo[*]]; groups the device could not finish are appended to `fallback` */
int poa_chain_run(int dev, poa_arena *arena, abpoa_para_t *abpt, int n_workers, const abpoa_gpu_group_t *groups,
                  abpoa_gpu_group_result_t *results, const std::vector<int> &todo, int flags, std::vector<int> &fallback, PoaChainStats *stats,
                  struct PoaEmit *emit);
/* consensus / MSA of a finished group copied into the caller's record (poa_batch.cu).  emit != NULL
 * (abpoa_gpu_msa_batch_write): also keep, for group `gidx`, the text abpoa_output() prints. */
struct PoaEmit;
void poa_finish_group_result(abpoa_t *ab, abpoa_para_t *abpt, abpoa_gpu_group_result_t *o, struct PoaEmit *emit, int gidx);
#endif
