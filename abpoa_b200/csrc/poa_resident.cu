/* poa_resident.cu -- host side of the resident alignment kernel (poa_resident_kernel_p16).
 *
 * A batch of read groups is ~1000 independent, strictly sequential chains
 *     flatten graph -> align read r on the GPU -> fuse the graph-CIGAR -> flatten ... (r+1)
 * (reference: one abpoa_msa() per group, src/abpoa_align.c:401-471).  Driving those chains with
 * launches and stream copies couples them into rounds and runs into the 32 hardware work queues.
 * Here the device side is ONE kernel that stays resident for the whole batch call: CTA s owns slot
 * s (a private HBM workspace) and waits on a mailbox in mapped pinned host memory.  The host
 * thread that owns a group writes the job blob into the slot's pinned staging buffer and bumps the
 * mailbox; the warp pulls the blob into HBM, aligns, pushes the CIGAR back and stamps the result.
 * Steady state needs no CUDA API call at all.
 *
 * Safety: the kernel exits when ctl->quit is raised (end of the batch call) and, as a backstop, when
 * its time budget is used up; a dying host process tears the context (and the kernel) down.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "poa_internal.h"
#include "poa_engine.h"
#include "poa_device.cuh"

extern "C" cudaError_t poa_launch_resident_p16(int gap_mode, int align_mode, int query_only, int *max_ctas_per_sm, const PoaSlotDev *slots,
                                               const PoaParamsDev *prm, int n_slots, int ring_rows, int ring_cells,
                                               const PoaResidentCtl *ctl, uint64_t budget_ns, cudaStream_t st);
extern "C" void poa_pick_ring(int gap_mode, int bits, int band_cells, size_t smem_budget, int *ring_rows, int *ring_cells);

#define CKR(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) poa_die("libabpoa_b200/resident", "%s failed: %s", #call, cudaGetErrorString(e_)); } while (0)

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

struct poa_resident {
    int dev = 0; cudaStream_t st = NULL; poa_arena *arena = NULL;
    /* geometry of the current (or last) start() */
    int n_slots = 0, stage_rows = 0, qlen_cap = 0, m = 0, gap_mode = 0, align_mode = 0;
    size_t blob_cap = 0, cigar_words = 0;
    bool running = false;
    /* mapped pinned host memory */
    uint8_t *h_blobs = NULL; size_t h_blobs_cap = 0;
    uint64_t *h_cigars = NULL; size_t h_cigars_cap = 0;
    PoaResultDev *h_results = NULL; PoaMailbox *h_mail = NULL; size_t h_slots_cap = 0;
    PoaResidentCtl *h_ctl = NULL;
    /* device memory */
    PoaSlotDev *d_slots = NULL; size_t d_slots_cap = 0;
    PoaParamsDev *d_prm = NULL;
    /* per slot (host): sequence number and the arena slice of the job in flight */
    uint32_t *seq = NULL; uint8_t **job_mem = NULL; size_t *job_bytes = NULL;
    /* every slot owns a fixed region sized for the usual graph growth (no allocator on the hot path);
     * only outliers (full-rectangle retries, unusually bushy graphs) borrow from the arena per job */
    uint8_t *slab = NULL; size_t slot_bytes = 0;
    int64_t launches = 0;
};

extern "C" poa_resident *poa_resident_new(int dev, poa_arena *arena) {
    poa_resident *r = new poa_resident();
    r->dev = dev; r->arena = arena;
    CKR(cudaSetDevice(dev));
    CKR(cudaStreamCreateWithFlags(&r->st, cudaStreamNonBlocking));
    CKR(cudaHostAlloc((void **)&r->h_ctl, sizeof(PoaResidentCtl), cudaHostAllocMapped | cudaHostAllocPortable));
    memset(r->h_ctl, 0, sizeof(PoaResidentCtl));
    CKR(cudaMalloc((void **)&r->d_prm, sizeof(PoaParamsDev)));
    return r;
}

extern "C" void poa_resident_stop(poa_resident *r);
extern "C" void poa_resident_free(poa_resident *r) {
    if (!r) return;
    if (r->running) poa_resident_stop(r);
    cudaSetDevice(r->dev);
    if (r->h_blobs) cudaFreeHost(r->h_blobs);
    if (r->h_cigars) cudaFreeHost(r->h_cigars);
    if (r->h_results) cudaFreeHost(r->h_results);
    if (r->h_mail) cudaFreeHost(r->h_mail);
    if (r->h_ctl) cudaFreeHost(r->h_ctl);
    if (r->d_slots) cudaFree(r->d_slots);
    if (r->d_prm) cudaFree(r->d_prm);
    free(r->seq); free(r->job_mem); free(r->job_bytes);
    if (r->st) cudaStreamDestroy(r->st);
    delete r;
}

static void *dev_view(void *host_ptr) {
    void *d = NULL;
    CKR(cudaHostGetDevicePointer(&d, host_ptr, 0));
    return d;
}

/* HBM workspace of one job: blob | rowinfo | rowoff | cigar | query profile | planes */
struct JobLayout { size_t o_blob, o_info, o_off, o_cig, o_qp, o_planes, total; uint64_t units; size_t cigar_cap; };
static JobLayout job_layout(const abpoa_para_t *abpt, int n_rows, int qlen, int w, size_t blob_bytes, int generous) {
    JobLayout L;
    const int P = abpt->gap_mode == ABPOA_LINEAR_GAP ? 1 : (abpt->gap_mode == ABPOA_AFFINE_GAP ? 3 : 5);
    uint64_t per_row = (uint64_t)((qlen + 1 + 7) / 8 + 1);
    if (!generous && w >= 0) { const uint64_t est = (uint64_t)((2 * w + 1 + 32 + 7) / 8 + 2); if (est < per_row) per_row = est; }
    L.units = per_row * (uint64_t)P * (uint64_t)n_rows;
    const size_t qstride = (((size_t)qlen + 1 + 7) & ~(size_t)7) + 8;
    L.cigar_cap = (size_t)qlen + n_rows + 8;
    size_t off = 0;
    L.o_blob = off; off += al256(blob_bytes);
    L.o_info = off; off += al256((size_t)n_rows * sizeof(PoaRowInfo));
    L.o_off = off; off += al256((size_t)n_rows * 4);
    L.o_cig = off; off += al256(L.cigar_cap * 8);
    L.o_qp = off; off += al256((size_t)abpt->m * qstride * 2);
    L.o_planes = off; off += al256((size_t)L.units * POA_GROUP * 2);
    L.total = off;
    return L;
}

/* Start the resident kernel for one batch call.  want_slots = number of groups that should be in
 * flight together; qmax = longest read of the batch.  Returns the number of slots (0: this
 * configuration is not served by the resident kernel -- the caller uses the launch-per-round path). */
extern "C" int poa_resident_start(poa_resident *r, const abpoa_para_t *abpt, int want_slots, int qmax) {
    if (r->running) poa_die("libabpoa_b200/resident", "resident kernel already running");
    if (!r->arena) return 0;
    CKR(cudaSetDevice(r->dev));
    /* staging (pinned host) is sized for graphs of up to 4 nodes per read base; a group whose graph
     * outgrows it is finished by the launch path after the resident kernel has stopped */
    const int stage_rows = 4 * qmax + 1024;
    if (!poa_p16_ok(abpt, qmax, 2 * qmax) || abpt->m > POA_MAX_M) return 0;
    const int w = poa_band_halfwidth(abpt, qmax);
    /* shared-memory ring as the launch path picks it; occupancy from the driver */
    const int band_cells = w >= 0 ? (2 * w + 1 + 40 + 7) / 8 * 8 : (qmax + 1 + 7) / 8 * 8 + 8;
    static const size_t smem_budget = [] { const char *e = getenv("ABPOA_GPU_SMEM_KB"); return (size_t)(e && *e ? atoi(e) : 28) * 1024; }();
    int ring_rows = 2, ring_cells = 64;
    poa_pick_ring(abpt->gap_mode, 16, band_cells, smem_budget, &ring_rows, &ring_cells);
    int per_sm = 0, n_sm = 0;
    CKR(poa_launch_resident_p16(abpt->gap_mode, abpt->align_mode, 1, &per_sm, NULL, NULL, 0, ring_rows, ring_cells, NULL, 0, r->st));
    CKR(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, r->dev));
    /* every slot's CTA must be resident at the same time; keep one CTA slot per SM free */
    const int max_slots = (per_sm - 1) * n_sm;
    if (max_slots < n_sm) return 0;
    const int n_slots = want_slots < max_slots ? want_slots : max_slots;
    if (n_slots < 1) return 0;

    const size_t n_pred_cap = (size_t)3 * stage_rows;
    const size_t blob_cap = al256(sizeof(PoaJobHeader) + 64 + ((size_t)stage_rows + 1) * 8 + 32 + n_pred_cap * 4 * (abpt->inc_path_score ? 2 : 1) + 64 + (size_t)qmax + 64);
    const size_t cigar_words = (size_t)qmax + stage_rows + 8;

    /* (re)allocate: grow-only, reused across batch calls */
    const size_t need_blobs = (size_t)n_slots * blob_cap, need_cig = (size_t)n_slots * cigar_words * 8;
    if (need_blobs > r->h_blobs_cap) { if (r->h_blobs) cudaFreeHost(r->h_blobs); CKR(cudaHostAlloc((void **)&r->h_blobs, need_blobs, cudaHostAllocMapped | cudaHostAllocPortable)); r->h_blobs_cap = need_blobs; }
    if (need_cig > r->h_cigars_cap) { if (r->h_cigars) cudaFreeHost(r->h_cigars); CKR(cudaHostAlloc((void **)&r->h_cigars, need_cig, cudaHostAllocMapped | cudaHostAllocPortable)); r->h_cigars_cap = need_cig; }
    if ((size_t)n_slots > r->h_slots_cap) {
        if (r->h_results) cudaFreeHost(r->h_results);
        if (r->h_mail) cudaFreeHost(r->h_mail);
        CKR(cudaHostAlloc((void **)&r->h_results, (size_t)n_slots * sizeof(PoaResultDev), cudaHostAllocMapped | cudaHostAllocPortable));
        CKR(cudaHostAlloc((void **)&r->h_mail, (size_t)n_slots * sizeof(PoaMailbox), cudaHostAllocMapped | cudaHostAllocPortable));
        r->seq = (uint32_t *)poa_xrealloc(r->seq, (size_t)n_slots * sizeof(uint32_t));
        r->job_mem = (uint8_t **)poa_xrealloc(r->job_mem, (size_t)n_slots * sizeof(uint8_t *));
        r->job_bytes = (size_t *)poa_xrealloc(r->job_bytes, (size_t)n_slots * sizeof(size_t));
        r->h_slots_cap = (size_t)n_slots;
    }
    if ((size_t)n_slots > r->d_slots_cap) { if (r->d_slots) cudaFree(r->d_slots); CKR(cudaMalloc((void **)&r->d_slots, (size_t)n_slots * sizeof(PoaSlotDev))); r->d_slots_cap = (size_t)n_slots; }

    /* fixed per-slot workspace: graphs of 5 % error reads end near 2.5 rows per read base */
    {
        int slot_rows = 3 * qmax + 512;
        const size_t lim = poa_arena_capacity(r->arena) / 10 * 7;
        for (;;) {
            const size_t bb = al256(sizeof(PoaJobHeader) + 64 + ((size_t)slot_rows + 1) * 8 + 32 + (size_t)2 * slot_rows * 4 * (abpt->inc_path_score ? 2 : 1) + 64 + (size_t)qmax + 64);
            r->slot_bytes = job_layout(abpt, slot_rows, qmax, w, bb, 0).total;
            if ((size_t)n_slots * r->slot_bytes <= lim || slot_rows <= qmax + 512) break;
            slot_rows = slot_rows / 10 * 9;
        }
        if ((size_t)n_slots * r->slot_bytes > lim) return 0;
        r->slab = poa_arena_borrow(r->arena, (size_t)n_slots * r->slot_bytes);
    }
    r->n_slots = n_slots; r->stage_rows = stage_rows; r->qlen_cap = qmax; r->m = abpt->m; r->gap_mode = abpt->gap_mode; r->align_mode = abpt->align_mode;
    r->blob_cap = blob_cap; r->cigar_words = cigar_words;

    /* slot table */
    PoaSlotDev *tab = (PoaSlotDev *)poa_xmalloc((size_t)n_slots * sizeof(PoaSlotDev));
    uint8_t *d_blobs_view = (uint8_t *)dev_view(r->h_blobs);
    uint64_t *d_cig_view = (uint64_t *)dev_view(r->h_cigars);
    PoaResultDev *d_res_view = (PoaResultDev *)dev_view(r->h_results);
    PoaMailbox *d_mail_view = (PoaMailbox *)dev_view(r->h_mail);
    for (int s = 0; s < n_slots; ++s) {
        PoaSlotDev &t = tab[s];
        t.host_blob = d_blobs_view + (size_t)s * blob_cap;
        t.result = d_res_view + s;
        t.host_cigar = d_cig_view + (size_t)s * cigar_words;
        t.mail = d_mail_view + s;
        r->seq[s] = 0; r->job_mem[s] = NULL; r->job_bytes[s] = 0;
    }
    memset(r->h_mail, 0, (size_t)n_slots * sizeof(PoaMailbox));
    memset(r->h_results, 0, (size_t)n_slots * sizeof(PoaResultDev));
    r->h_ctl->quit = 0;
    __sync_synchronize();
    PoaParamsDev prm; poa_fill_params(&prm, abpt, 15);
    CKR(cudaMemcpyAsync(r->d_prm, &prm, sizeof prm, cudaMemcpyHostToDevice, r->st));
    CKR(cudaMemcpyAsync(r->d_slots, tab, (size_t)n_slots * sizeof(PoaSlotDev), cudaMemcpyHostToDevice, r->st));
    CKR(cudaStreamSynchronize(r->st));
    free(tab);
    static const uint64_t budget_ns = [] { const char *e = getenv("ABPOA_GPU_RESIDENT_BUDGET_S"); return (uint64_t)(e && *e ? atoll(e) : 900) * 1000000000ull; }();
    /* Ask for so much shared memory per CTA that exactly ceil(n_slots / SMs) CTAs fit one SM: whatever order
     * the block scheduler fills SMs in, the slots end up spread evenly over the whole chip. */
    static const bool spread = [] { const char *e = getenv("ABPOA_GPU_RESIDENT_SPREAD"); return e && *e == '1'; }();
    int smem_ask = 0;
    if (spread) {
        const int per = (n_slots + n_sm - 1) / n_sm;
        smem_ask = (int)((size_t)227 * 1024 / (size_t)per) - 1024 - 256;      /* 1 KB static + allocation granularity */
        smem_ask &= ~255;
    }
    CKR(poa_launch_resident_p16(abpt->gap_mode, abpt->align_mode, 0, spread ? &smem_ask : NULL, r->d_slots, r->d_prm, n_slots, ring_rows, ring_cells,
                                (const PoaResidentCtl *)dev_view(r->h_ctl), budget_ns, r->st));
    r->running = true; r->launches += 1;
    poa_hold_frees(1);
    return n_slots;
}

extern "C" void poa_resident_stop(poa_resident *r) {
    if (!r->running) return;
    CKR(cudaSetDevice(r->dev));
    __atomic_store_n(&r->h_ctl->quit, 1u, __ATOMIC_RELEASE);
    CKR(cudaStreamSynchronize(r->st));
    poa_hold_frees(0);
    for (int s = 0; s < r->n_slots; ++s) if (r->job_mem[s]) { poa_arena_return(r->arena, r->job_mem[s], r->job_bytes[s]); r->job_mem[s] = NULL; }
    if (r->slab) { poa_arena_return(r->arena, r->slab, (size_t)r->n_slots * r->slot_bytes); r->slab = NULL; }
    r->running = false;
}

extern "C" int poa_resident_slots(const poa_resident *r) { return r->running ? r->n_slots : 0; }
extern "C" int64_t poa_resident_launches(const poa_resident *r) { return r->launches; }

/* may this job go through a slot? */
extern "C" int poa_resident_fits(const poa_resident *r, const abpoa_para_t *abpt, const poa_blob_plan *pl) {
    return pl->whole_graph && pl->n_rows <= r->stage_rows && pl->qlen <= r->qlen_cap && pl->bytes <= r->blob_cap &&
           poa_p16_ok(abpt, pl->qlen, pl->n_rows);
}
extern "C" uint8_t *poa_resident_stage(poa_resident *r, int slot) { return r->h_blobs + (size_t)slot * r->blob_cap; }

/* The blob is in poa_resident_stage(slot): carve the job's workspace out of the arena and hand the job
 * to the slot's warp.  generous: planes for the full rectangle (retry of a PLANE_OVF job).  Returns 0
 * when the arena has no room right now (nothing submitted; the caller retries on a later sweep). */
extern "C" int poa_resident_submit(poa_resident *r, int slot, const abpoa_para_t *abpt, const poa_blob_plan *pl, int generous) {
    const JobLayout L = job_layout(abpt, pl->n_rows, pl->qlen, pl->w, pl->bytes, generous);
    uint8_t *mem;
    if (L.total <= r->slot_bytes) mem = r->slab + (size_t)slot * r->slot_bytes;          /* the slot's own region */
    else {
        mem = poa_arena_try_borrow(r->arena, L.total);
        if (!mem) return 0;
        r->job_mem[slot] = mem; r->job_bytes[slot] = L.total;
    }
    PoaMailbox *mb = r->h_mail + slot;
    *(volatile uint64_t *)&r->h_results[slot].t_end_ns = 0;
    mb->blob_bytes = (uint32_t)pl->bytes; mb->cigar_cap = (uint32_t)L.cigar_cap; mb->pad0 = 0; mb->plane_cap_units = L.units;
    mb->blob = mem + L.o_blob; mb->planes = mem + L.o_planes; mb->rowinfo = (PoaRowInfo *)(mem + L.o_info); mb->rowoff = (uint32_t *)(mem + L.o_off);
    mb->cigar = (uint64_t *)(mem + L.o_cig); mb->qprof = (int16_t *)(mem + L.o_qp);
    __atomic_store_n(&mb->seq, ++r->seq[slot], __ATOMIC_RELEASE);
    return 1;
}
extern "C" void poa_resident_release(poa_resident *r, int slot) {
    if (r->job_mem[slot]) { poa_arena_return(r->arena, r->job_mem[slot], r->job_bytes[slot]); r->job_mem[slot] = NULL; r->job_bytes[slot] = 0; }
}
/* NULL while the slot's job is running */
extern "C" const PoaResultDev *poa_resident_poll(const poa_resident *r, int slot) {
    if (*(volatile const uint64_t *)&r->h_results[slot].t_end_ns == 0) return NULL;
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    return r->h_results + slot;
}
extern "C" const uint64_t *poa_resident_cigar(const poa_resident *r, int slot) { return r->h_cigars + (size_t)slot * r->cigar_words; }
/* has the kernel ended on its own (budget) or died?  cudaSuccess = it is gone */
extern "C" int poa_resident_alive(poa_resident *r) {
    if (!r->running) return 0;
    cudaError_t e = cudaStreamQuery(r->st);
    if (e == cudaErrorNotReady) return 1;
    if (e != cudaSuccess) poa_die("libabpoa_b200/resident", "resident kernel failed: %s", cudaGetErrorString(e));
    return 0;
}
