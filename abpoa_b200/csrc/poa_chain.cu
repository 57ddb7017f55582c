/* poa_chain.cu -- the device-resident progressive POA ("chain engine") behind abpoa_gpu_msa_batch.
 *
 * Reads of a group are strictly sequential (read i+1 is aligned to the graph that already contains
 * read i; reference src/abpoa_align.c:312-352).  The launch-per-round engine of poa_batch.cu pays a
 * host round trip per read (flatten -> H2D -> kernel -> D2H -> host fusion).  Here the graph lives in
 * HBM: per round and cohort of groups the stream carries exactly two kernels,
 *
 *     poa_chain_align_kernel_p16   one warp per group: DP + backtrace of read r (poa_kernels.cu)
 *     poa_chain_fuse_kernel        one CTA per group: fuse the graph-CIGAR, splice the order, edge order,
 *                                  max_remain, flatten graph + read r+1 into the next job (poa_chain.cuh)
 *
 * and the host does nothing until the group is finished: reads go up once, the final graph comes
 * back once (compact export, rebuilt by poa_graph_import) and the host runs consensus on it.
 * Groups the device cannot finish (capacity, int16 window, band wider than estimated) are reported
 * back and completed by the launch-per-round engine -- same results either way.
 *
 * Scope of the chain: global alignment, banded (wb >= 0), packed-int16 admissible scores, consensus
 * output (no per-edge read sets), unit base weights.  Everything else takes the other engine.
 */
#include <cuda_runtime.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <thread>
#include <vector>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "abpoa_gpu.h"
#include "poa_internal.h"
#include "poa_engine.h"
#include "poa_device.cuh"
#include "poa_chain.cuh"
#include "poa_chain_host.h"

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    poa_die("libabpoa_b200/chain", "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); } while (0)

extern "C" cudaError_t poa_launch_chain_dp_worker(int gap_mode, const int *gaps, PoaChainSlot *slots, PoaChainSync *sync, int n_groups,
                                                  const PoaParamsDev *prm, int ring_rows, int ring_cells, cudaStream_t st);
extern "C" cudaError_t poa_launch_chain_align_p16(int gap_mode, const int *gaps, const PoaChainSlot *slots, const int32_t *idx, int n_jobs, int round,
                                                  const PoaParamsDev *prm, int ring_rows, int ring_cells, cudaStream_t st);
extern "C" void poa_pick_ring(int gap_mode, int bits, int band_cells, size_t smem_budget, int *ring_rows, int *ring_cells);

static_assert(offsetof(PoaChainSync, q_tail) == 128 && offsetof(PoaChainSync, total) == 256 && offsetof(PoaChainSync, abort) == 384,
              "every polled / bumped word of PoaChainSync sits in its own 128-byte line");

/* ------------------------------------------------------------------ kernels */
__global__ void __launch_bounds__(POA_CHAIN_T) poa_chain_seed_kernel(PoaChainSlot *slots, const PoaChainParams *cp, int n) {
    if ((int)blockIdx.x >= n) return;
    chain_seed(&slots[blockIdx.x], cp);
}

__global__ void __launch_bounds__(POA_CHAIN_T) poa_chain_fuse_kernel(PoaChainSlot *slots, const int32_t *idx, const PoaChainParams *cp, int n, int round) {
    if ((int)blockIdx.x >= n) return;
    chain_fuse(&slots[idx[blockIdx.x]], cp, round);
}

/* Free-running chain, fuse side: persistent CTAs draw tickets from PoaChainSync; ticket t is served when tasks[t] holds a group.
 * (The alignment side is poa_chain_dp_worker_kernel in poa_kernels.cu.) */
__device__ __forceinline__ int sync_ld(const int32_t *p) { int v; asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void sync_st(int32_t *p, int v) { asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned long long sync_now_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

__global__ void __launch_bounds__(POA_CHAIN_T) poa_chain_fuse_worker_kernel(PoaChainSlot *slots, PoaChainSync *sync, const PoaChainParams *cp) {
    __shared__ int task_s;
    for (;;) {
        if (threadIdx.x < 32) {                                /* warp 0 waits, warp-uniformly (see poa_chain_dp_worker_kernel) */
            unsigned ticket = 0;
            if (threadIdx.x == 0) ticket = atomicAdd(&sync->q_head, 1u);
            ticket = __shfl_sync(0xffffffffu, ticket, 0);
            int g = -2; unsigned ns = 250, polls = 0;
            const unsigned long long t0 = sync_now_ns(), limit = sync->watchdog_ns;
            const int32_t *tasks = sync->tasks;
            for (;;) {
                /* the ticket's own task word is what is polled; the shared words (total, abort) and the clock only every 32nd time */
                if ((polls++ & 31u) == 0) {
                    const int tot = __shfl_sync(0xffffffffu, sync_ld(&sync->total), 0), ab = __shfl_sync(0xffffffffu, sync_ld(&sync->abort), 0);
                    if ((long long)ticket >= (long long)tot || ab) { g = -2; break; }
                    /* nobody appended a task for this long: the alignment kernel is not running next to this one (a tool that
                     * serialises kernels, a device shared with a long-running grid): give up, the launch engine finishes the groups */
                    const unsigned late = __shfl_sync(0xffffffffu, (unsigned)(polls > 1 && sync_now_ns() - t0 > limit), 0);
                    if (late) { if (threadIdx.x == 0) sync_st(&sync->abort, 1); g = -2; break; }
                }
                g = __shfl_sync(0xffffffffu, sync_ld(&tasks[ticket]), 0);
                if (g >= 0) break;
                __nanosleep(ns); if (ns < 4000) ns <<= 1;
            }
            if (threadIdx.x == 0) task_s = g;
        }
        __syncthreads();
        const int g = task_s;
        __syncthreads();
        if (g < 0) return;
        __threadfence();                                       /* acquire: graph arrays / CIGAR of this group may have been written from another SM */
        PoaChainSlot *s = &slots[g];
        const unsigned long long t0 = sync_now_ns();
        chain_fuse(s, cp, 0);
        __syncthreads();
        __threadfence();                                       /* release */
        __syncthreads();
        if (threadIdx.x == 0) { s->fuse_ns += sync_now_ns() - t0; __threadfence(); sync_st(&s->turn, 0); }
    }
}

/* Compact export of the final graphs (layout: poa_graph_import in poa_graph.c).  ex_off[g] = first int32 word of
 * group g's record in `ex`; a record is written only if it fits ex_cap[g] words (else word 0 = -1). */
__global__ void __launch_bounds__(POA_CHAIN_T) poa_chain_export_kernel(PoaChainSlot *slots, const PoaChainParams *cp, int n,
                                                                       int32_t *ex, const int64_t *ex_off, const int32_t *ex_cap) {
    if ((int)blockIdx.x >= n) return;
    PoaChainSlot *s = &slots[blockIdx.x];
    int32_t *o = ex + ex_off[blockIdx.x];
    const int K = cp->K, A = cp->A, nn = s->n_nodes;
    if (s->failed) { if (threadIdx.x == 0) o[0] = -1; return; }
    int32_t *ci = s->scr[0], *ca = s->scr[1];
    POA_PAR_FOR(v, nn) { ci[v] = s->in_cnt[v]; ca[v] = s->aln_cnt[v]; }
    __syncthreads();
    const int n_in = cta_excl_scan(ci, nn);
    __syncthreads();
    const int n_aln = cta_excl_scan(ca, nn);
    __syncthreads();
    int32_t *co = s->scr[2];
    POA_PAR_FOR(v, nn) co[v] = s->out_cnt[v];
    __syncthreads();
    const int n_out = cta_excl_scan(co, nn);
    __syncthreads();
    const int64_t words = 4 + 5ll * nn + 4ll * n_in + n_aln;
    if (words > ex_cap[blockIdx.x] || n_out != n_in) { if (threadIdx.x == 0) o[0] = -1; return; }
    int32_t *base = o + 4, *n_read = base + nn, *in_cnt = n_read + nn, *out_cnt = in_cnt + nn, *aln_cnt = out_cnt + nn;
    int32_t *in_id = aln_cnt + nn, *in_w = in_id + n_in, *out_id = in_w + n_in, *out_w = out_id + n_in, *aln = out_w + n_in;
    if (threadIdx.x == 0) { o[0] = nn; o[1] = n_in; o[2] = n_aln; o[3] = s->fused; }
    POA_PAR_FOR(v, nn) {
        base[v] = s->base[v]; n_read[v] = s->n_read[v]; in_cnt[v] = s->in_cnt[v]; out_cnt[v] = s->out_cnt[v]; aln_cnt[v] = s->aln_cnt[v];
        for (int e = 0; e < s->in_cnt[v]; ++e) { in_id[ci[v] + e] = s->in_id[(size_t)v * K + e]; in_w[ci[v] + e] = s->in_w[(size_t)v * K + e]; }
        for (int e = 0; e < s->out_cnt[v]; ++e) { out_id[co[v] + e] = s->out_id[(size_t)v * K + e]; out_w[co[v] + e] = s->out_w[(size_t)v * K + e]; }
        for (int a = 0; a < s->aln_cnt[v]; ++a) aln[ca[v] + a] = s->aln_id[(size_t)v * A + a];
    }
}

/* consensus of every finished group (chain_consensus: heaviest bundling on one thread per group); records are packed
 * back to back into `out` through an atomic cursor: rec_off[g] = first word of group g's record, -1 if none */
__global__ void __launch_bounds__(32) poa_chain_consensus_kernel(PoaChainSlot *slots, const PoaChainParams *cp, int n,
                                                                 int32_t *out, unsigned long long *cursor, unsigned long long out_words, int64_t *rec_off) {
    if ((int)blockIdx.x >= n || threadIdx.x != 0) return;
    PoaChainSlot *s = &slots[blockIdx.x];
    int32_t *tmp = s->scr[2];                               /* [n_cap]: the consensus is never longer than the graph */
    chain_consensus(s, cp, tmp, s->n_cap);
    const int len = tmp[0];
    if (len < 0) { rec_off[blockIdx.x] = -1; return; }
    const unsigned long long at = atomicAdd(cursor, (unsigned long long)(len + 1));
    if (at + (unsigned long long)(len + 1) > out_words) { rec_off[blockIdx.x] = -1; return; }
    for (int k = 0; k <= len; ++k) out[at + k] = tmp[k];
    rec_off[blockIdx.x] = (int64_t)at;
}

/* ------------------------------------------------------------------ host side */
static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

int poa_chain_eligible(const abpoa_para_t *abpt) {
    const char *off = getenv("ABPOA_GPU_NO_CHAIN");
    if (off && *off == '1') return 0;
    { const char *np = getenv("ABPOA_GPU_NO_P16"); if (np && *np == '1') return 0; }      /* the chain only has the packed int16 kernel */
    if (abpt->align_mode != ABPOA_GLOBAL_MODE || abpt->wb < 0) return 0;
    if (abpt->gap_mode == ABPOA_LINEAR_GAP) return 0;                      /* banded linear gaps: generic kernel only (lane-exact band edges) */
    if (abpt->use_read_ids || abpt->out_msa || abpt->out_gfa || abpt->max_n_cons > 1 || abpt->cons_algrm != ABPOA_HB) return 0;
    if (abpt->use_qv || abpt->amb_strand || abpt->inc_path_score || abpt->zdrop > 0 || abpt->rev_cigar || !abpt->ret_cigar) return 0;
    if (abpt->put_gap_on_right || abpt->put_gap_at_end) return 0;         /* handled by the kernels, but keep the chain on the common configuration */
    if (abpt->m > POA_MAX_M) return 0;
    if (!(abpt->disable_seeding && abpt->progressive_poa == 0)) return 0;
    return 1;
}

namespace {

struct GroupPlan {
    int g;                  /* index into the caller's groups */
    int n_reads, qmax; int64_t bases;
    int n_cap; size_t static_bytes; double pool_units_est;
};

struct Cohort {
    std::vector<int> members;               /* indices into the wave's plan list */
    cudaStream_t st = NULL;
    cudaEvent_t ev_begin = NULL, ev_end = NULL;
    std::vector<cudaEvent_t> marks;         /* per round: before DP, between DP and fuse, after fuse */
};

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

/* Pinned staging buffers are kept for the life of the process (grow-only, one per purpose): cudaHostAlloc / cudaFreeHost of
 * half a gigabyte per batch call cost ~0.2 s, more than the copies they serve. */
struct PinnedSlot { void *p = NULL; size_t cap = 0; bool busy = false; };
std::mutex g_pin_mu; PinnedSlot g_pin[4];
void *pinned_get(int which, size_t bytes) {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    PinnedSlot &ps = g_pin[which];
    if (ps.busy) { void *q = NULL; CK(cudaHostAlloc(&q, bytes ? bytes : 1, cudaHostAllocPortable)); return q; }      /* concurrent caller: private buffer */
    if (bytes > ps.cap) {
        if (ps.p) CK(cudaFreeHost(ps.p));
        ps.cap = bytes + bytes / 4 + 4096;
        CK(cudaHostAlloc(&ps.p, ps.cap, cudaHostAllocPortable));
    }
    ps.busy = true;
    return ps.p;
}
void pinned_put(int which, void *q) {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    if (q == g_pin[which].p) g_pin[which].busy = false; else CK(cudaFreeHost(q));
}

}  // namespace

/* Run the groups listed in `todo` (eligible ones) through the device chain.  Groups that could not be
 * finished are appended to `fallback`.  Returns 0. */
int poa_chain_run(int dev, poa_arena *arena, abpoa_para_t *abpt, int n_workers, const abpoa_gpu_group_t *groups,
                  abpoa_gpu_group_result_t *results, const std::vector<int> &todo, int flags, std::vector<int> &fallback, PoaChainStats *stats,
                  struct PoaEmit *emit) {
    CK(cudaSetDevice(dev));
    const bool record = (flags & ABPOA_GPU_RECORD_READS) != 0;
    const bool verbose = getenv("ABPOA_GPU_PROFILE") != NULL;
    const int m = abpt->m;
    const int K = [&] { const char *e = getenv("ABPOA_GPU_CHAIN_K"); return e && *e ? atoi(e) : (m > 5 ? 24 : 12); }();
    const int A = m - 1 > 1 ? m - 1 : 1;
    const int P = abpt->gap_mode == ABPOA_LINEAR_GAP ? 1 : (abpt->gap_mode == ABPOA_AFFINE_GAP ? 3 : 5);
    int n_cohorts = [&] { const char *e = getenv("ABPOA_GPU_CHAIN_COHORTS"); return e && *e ? atoi(e) : 4; }();
    if (n_cohorts < 1) n_cohorts = 1;
    if (n_cohorts > 16) n_cohorts = 16;
    /* free-running (default): every group advances at its own pace (PoaChainSync); ABPOA_GPU_CHAIN_ROUNDS=1: lock-step rounds,
     * two kernels per round and cohort */
    /* two kernels that wait for each other need to run CONCURRENTLY: under a tool that injects into the CUDA driver and
     * serialises kernel launches (ncu, compute-sanitizer), or with blocking launches, use the round schedule.  Looked up once. */
    static const bool serialised_launches = [] {
        { const char *lb = getenv("CUDA_LAUNCH_BLOCKING"); if (lb && *lb == '1') return true; }      /* the second kernel would never be launched */
        extern char **environ;
        for (char **v = environ; v && *v; ++v)
            if (!strncmp(*v, "CUDA_INJECTION64_PATH=", 22) || !strncmp(*v, "NV_NSIGHT_INJECTION", 19) || !strncmp(*v, "NV_COMPUTE_PROFILER", 19) ||
                !strncmp(*v, "NV_SANITIZER_INJECTION", 22) || !strncmp(*v, "NV_TPS_LAUNCH_", 14)) return true;
        /* ... or whose injection library is already mapped into this process */
        if (FILE *mp = fopen("/proc/self/maps", "r")) {
            char line[512]; bool hit = false;
            while (!hit && fgets(line, sizeof line, mp))
                hit = strstr(line, "InjectionTarget") || strstr(line, "cuda-injection") || strstr(line, "libsanitizer-collection") || strstr(line, "TreeLauncherTarget");
            fclose(mp);
            if (hit) return true;
        }
        return false;
    }();
    const bool free_run = [] {
        const char *e = getenv("ABPOA_GPU_CHAIN_ROUNDS");        /* per call: the tests switch schedules inside one process */
        if (e && *e) return *e != '1';
        return !serialised_launches;
    }();
    const double pool_margin = 1.15;

    /* ---- per-group sizes ---- */
    std::vector<GroupPlan> plans;
    for (int g : todo) {
        const abpoa_gpu_group_t &in = groups[g];
        GroupPlan p; p.g = g; p.n_reads = in.n_seq; p.qmax = 0; p.bases = 0;
        bool ok = in.n_seq >= 2;
        for (int i = 0; i < in.n_seq; ++i) {
            const int l = in.seq_lens[i];
            if (l < 1) ok = false;
            if (l > p.qmax) p.qmax = l;
            p.bases += l;
            if (ok && !poa_p16_ok(abpt, l, 3 * l)) ok = false;
        }
        if (!ok || p.qmax > (1 << 24)) { fallback.push_back(g); continue; }
        const int64_t grow = (int64_t)((double)p.qmax * (1.0 + 0.10 * (p.n_reads - 1))) + 256;
        p.n_cap = (int)std::min<int64_t>(2 + p.bases, grow);
        const size_t nc = (size_t)p.n_cap, scr_n = std::max<size_t>((size_t)p.qmax + 2, nc);
        size_t b = 0;
        b += al256(nc);                                            /* base            */
        b += 4 * al256(nc * 4);                                    /* counts, n_read  */
        b += 4 * al256(nc * K * 4);                                /* edge lists      */
        b += al256(nc * A * 4);                                    /* aligned sets    */
        b += 4 * al256(nc * 4);                                    /* order x2, node_row, rem_row */
        b += 6 * al256(scr_n * 4);                                 /* scratch         */
        b += al256((size_t)p.bases) + al256(((size_t)p.n_reads + 1) * 4) + al256((size_t)p.n_reads * 4);   /* reads, offsets, w */
        const size_t pred_cap = nc * 3;
        b += al256(256 + (nc + 1) * 8 + pred_cap * 4 + 4 + (size_t)p.qmax + 64);    /* job blob        */
        b += al256(nc * sizeof(PoaRowInfo)) + al256(nc * sizeof(PoaRowOff));       /* rowinfo, rowoff */
        b += al256(((size_t)p.qmax + nc + 8) * 8);                 /* graph-CIGAR     */
        b += al256((size_t)m * ((((size_t)p.qmax + 1 + 7) & ~(size_t)7) + 8) * 2);  /* query profile   */
        b += al256(sizeof(PoaResultDev)) + al256(nc * sizeof(PoaBtRec));
        if (record) b += 2 * al256((size_t)p.n_reads * 4) + al256((size_t)p.n_reads * 8);
        p.static_bytes = b;
        const int wmax = poa_band_halfwidth(abpt, p.qmax);
        const double rows_final = std::min<double>(2.0 + (double)p.bases, (double)p.qmax * (1.0 + 0.045 * (p.n_reads - 1)) + 64);
        p.pool_units_est = rows_final * (double)((2 * wmax + 1 + 32 + 7) / 8 + 2) * P;
        plans.push_back(p);
    }
    if (plans.empty()) return 0;

    /* ---- waves: as many groups as the arena holds (static regions + plane pool) ---- */
    const size_t arena_cap = poa_arena_capacity(arena);
    size_t pos = 0;
    std::vector<abpoa_t *> handles;
    /* how many groups fit one wave; when several waves are needed they get equal shares (a small tail wave would leave the
     * device nearly empty for as long as a full one takes: a wave lasts as long as one group's chain) */
    auto fit_from = [&](size_t from, size_t limit, size_t *need_static_out, double *need_pool_out) {
        size_t end = from, need_static = 0; double need_pool = 0;
        while (end < plans.size() && end - from < limit) {
            const size_t s2 = need_static + plans[end].static_bytes + sizeof(PoaChainSlot) + 4096;
            const double p2 = need_pool + plans[end].pool_units_est * 16.0 * pool_margin;
            if (end > from && (double)s2 + p2 + (64 << 20) > (double)arena_cap) break;
            need_static = s2; need_pool = p2; ++end;
        }
        if (need_static_out) *need_static_out = need_static;
        if (need_pool_out) *need_pool_out = need_pool;
        return end;
    };
    size_t n_waves = 0;
    for (size_t q = 0; q < plans.size(); ++n_waves) q = fit_from(q, plans.size(), NULL, NULL);
    const size_t per_wave = (plans.size() + n_waves - 1) / std::max<size_t>(n_waves, 1);
    while (pos < plans.size()) {
        size_t need_static = 0; double need_pool = 0;
        const size_t end = fit_from(pos, per_wave, &need_static, &need_pool);
        if ((double)need_static + need_pool * 0.5 + (64 << 20) > (double)arena_cap) {       /* a single group that does not fit */
            fallback.push_back(plans[pos].g); ++pos; continue;
        }
        const int nw = (int)(end - pos);
        const double t_wave0 = now_ms();
        /* ---- carve the arena: everything of the wave in one borrow ---- */
        const size_t total = arena_cap;
        uint8_t *d_base = poa_arena_borrow(arena, total);
        size_t doff = 0;
        auto dtake = [&](size_t b) { uint8_t *q = d_base + doff; doff += al256(b); return q; };
        PoaChainSlot *d_slots = (PoaChainSlot *)dtake((size_t)nw * sizeof(PoaChainSlot));
        PoaChainParams *d_cp = (PoaChainParams *)dtake(sizeof(PoaChainParams));
        PoaParamsDev *d_prm = (PoaParamsDev *)dtake(sizeof(PoaParamsDev));
        unsigned long long *d_cursors = (unsigned long long *)dtake((size_t)32 * sizeof(unsigned long long));       /* 2 per cohort, <= 16 cohorts */
        PoaChainSync *d_sync = (PoaChainSync *)dtake(sizeof(PoaChainSync));
        int64_t n_tasks = 0;
        for (int t = 0; t < nw; ++t) n_tasks += plans[pos + t].n_reads - 1;
        int32_t *d_tasks = (int32_t *)dtake((size_t)std::max<int64_t>(n_tasks, 1) * 4);

        /* pinned staging: slots | params | reads + offsets + w of every group | round index lists */
        std::vector<PoaChainSlot> hs((size_t)nw);
        size_t reads_bytes = 0;
        for (int t = 0; t < nw; ++t) reads_bytes += al256((size_t)plans[pos + t].bases) + al256(((size_t)plans[pos + t].n_reads + 1) * 4) + al256((size_t)plans[pos + t].n_reads * 4);
        uint8_t *h_reads = (uint8_t *)pinned_get(0, reads_bytes + 256);
        uint8_t *d_reads = dtake(reads_bytes);
        size_t roff = 0;
        int max_reads = 0, band_cells = 64;
        struct ReadCopy { uint8_t *dst; int g; };
        std::vector<ReadCopy> read_copies((size_t)nw);
        for (int t = 0; t < nw; ++t) {
            const GroupPlan &p = plans[pos + t];
            const abpoa_gpu_group_t &in = groups[p.g];
            PoaChainSlot &s = hs[t]; memset(&s, 0, sizeof s);
            const size_t nc = (size_t)p.n_cap, scr_n = std::max<size_t>((size_t)p.qmax + 2, nc);
            s.n_cap = p.n_cap; s.pred_cap = (int32_t)(nc * 3); s.n_reads = p.n_reads;
            s.base = dtake(nc);
            s.in_cnt = (int32_t *)dtake(nc * 4); s.out_cnt = (int32_t *)dtake(nc * 4); s.aln_cnt = (int32_t *)dtake(nc * 4); s.n_read = (int32_t *)dtake(nc * 4);
            s.in_id = (int32_t *)dtake(nc * K * 4); s.in_w = (int32_t *)dtake(nc * K * 4); s.out_id = (int32_t *)dtake(nc * K * 4); s.out_w = (int32_t *)dtake(nc * K * 4);
            s.aln_id = (int32_t *)dtake(nc * A * 4);
            s.order[0] = (int32_t *)dtake(nc * 4); s.order[1] = (int32_t *)dtake(nc * 4); s.node_row = (int32_t *)dtake(nc * 4); s.rem_row = (int32_t *)dtake(nc * 4);
            for (int k = 0; k < 6; ++k) s.scr[k] = (int32_t *)dtake(scr_n * 4);
            /* reads */
            uint8_t *hr = h_reads + roff; const size_t rb = al256((size_t)p.bases), ob = al256(((size_t)p.n_reads + 1) * 4);
            int32_t *hoff = (int32_t *)(hr + rb), *hw = (int32_t *)(hr + rb + ob);
            int acc = 0;
            read_copies[t] = { hr, p.g };
            for (int i = 0; i < p.n_reads; ++i) {
                hoff[i] = acc; acc += in.seq_lens[i];
                hw[i] = poa_band_halfwidth(abpt, in.seq_lens[i]);
                const int bc = (2 * hw[i] + 1 + 104 + 7) / 8 * 8;
                if (bc > band_cells) band_cells = bc;
            }
            hoff[p.n_reads] = acc;
            s.reads = d_reads + roff; s.read_off = (const int32_t *)(d_reads + roff + rb); s.read_w = (const int32_t *)(d_reads + roff + rb + ob);
            roff += rb + ob + al256((size_t)p.n_reads * 4);
            /* job */
            s.blob_cap = (int32_t)(256 + (nc + 1) * 8 + (size_t)s.pred_cap * 4 + 4 + (size_t)p.qmax + 64);
            s.jd.blob = dtake((size_t)s.blob_cap);
            s.jd.rowinfo = (PoaRowInfo *)dtake(nc * sizeof(PoaRowInfo)); s.jd.rowoff = (PoaRowOff *)dtake(nc * sizeof(PoaRowOff));
            s.jd.cigar_cap = (int32_t)(p.qmax + p.n_cap + 8);
            s.jd.cigar = (uint64_t *)dtake((size_t)s.jd.cigar_cap * 8);
            s.jd.qprof = (int16_t *)dtake((size_t)m * ((((size_t)p.qmax + 1 + 7) & ~(size_t)7) + 8) * 2);
            s.jd.result = (PoaResultDev *)dtake(sizeof(PoaResultDev));
            s.jd.btrec = (PoaBtRec *)dtake(nc * sizeof(PoaBtRec));
            if (record) { s.rec_score = (int32_t *)dtake((size_t)p.n_reads * 4); s.rec_nops = (int32_t *)dtake((size_t)p.n_reads * 4); s.rec_hash = (uint64_t *)dtake((size_t)p.n_reads * 8); }
            if (p.n_reads > max_reads) max_reads = p.n_reads;
        }
        /* the read bytes themselves: half a gigabyte at BASELINE size, copied into the pinned buffer by all workers */
        {
            std::atomic<int> nx(0);
            auto copy_reads = [&]() {
                for (int t; (t = nx.fetch_add(1)) < nw;) {
                    const abpoa_gpu_group_t &in = groups[read_copies[t].g];
                    uint8_t *q = read_copies[t].dst;
                    for (int i = 0; i < in.n_seq; ++i) { memcpy(q, in.seqs[i], (size_t)in.seq_lens[i]); q += in.seq_lens[i]; }
                }
            };
            const int nth = reads_bytes < (8u << 20) ? 1 : std::max(1, std::min(n_workers, 16));
            std::vector<std::thread> th;
            for (int w = 1; w < nth; ++w) th.emplace_back(copy_reads);
            copy_reads();
            for (auto &x : th) x.join();
        }
        const double t_staged = now_ms();
        /* round index lists per cohort: wave-local slot indices of the groups that still have a read r */
        /* cohorts: each cohort's alignment kernel is one CTA (warp) per group; with at most one CTA per SM per cohort the
         * concurrently running kernels of all cohorts load every SM alike (a 250-CTA grid next to three more would put
         * twice as many warps on the first 102 SMs as on the rest, and a round ends when its slowest warp does) */
        int sm_count = 148; if (cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sm_count < 1) sm_count = 148;
        const int auto_cohorts = std::max(1, std::min(16, (nw + sm_count - 1) / sm_count));
        const int use_cohorts = free_run ? 1 : (getenv("ABPOA_GPU_CHAIN_COHORTS") ? n_cohorts : auto_cohorts);
        std::vector<Cohort> coh((size_t)std::min(use_cohorts, nw));
        for (int t = 0; t < nw; ++t) coh[(size_t)t % coh.size()].members.push_back(t);
        std::vector<int32_t> h_idx; std::vector<std::vector<std::pair<size_t, int>>> round_of(coh.size());   /* (offset into h_idx, count) per round */
        /* a group that has to re-run an alignment (band wider than its plane slab) falls one round behind the schedule:
         * every group stays listed EXTRA rounds beyond its last read (slots with nothing to do return at once) */
        const int EXTRA = 2, n_rounds = max_reads - 1 + EXTRA;
        for (size_t c = 0; c < coh.size(); ++c)
            for (int r = 1; r <= n_rounds; ++r) {
                const size_t o = h_idx.size(); int cnt = 0;
                for (int t : coh[c].members) if (plans[pos + t].n_reads + EXTRA > r) { h_idx.push_back(t); ++cnt; }
                round_of[c].push_back({o, cnt});
            }
        int32_t *d_idx = (int32_t *)dtake(std::max<size_t>(h_idx.size(), 1) * 4);
        /* export buffers */
        std::vector<int64_t> h_exoff((size_t)nw); std::vector<int32_t> h_excap((size_t)nw); int64_t ex_words = 0;
        for (int t = 0; t < nw; ++t) {
            const GroupPlan &p = plans[pos + t];
            const int64_t cap = 4 + 5ll * p.n_cap + 4ll * 3 * p.n_cap + (int64_t)p.n_cap * 2;
            h_exoff[t] = ex_words; h_excap[t] = (int32_t)std::min<int64_t>(cap, INT32_MAX); ex_words += (cap + 63) & ~63ll;
        }
        int64_t *d_exoff = (int64_t *)dtake((size_t)nw * 8); int32_t *d_excap = (int32_t *)dtake((size_t)nw * 4);
        /* the export buffer and the plane pool share what is left: planes are dead when the export runs */
        doff = al256(doff);
        if (doff > total) poa_die("libabpoa_b200/chain", "wave layout (%zu bytes) exceeds the arena (%zu bytes)", doff, total);
        if (doff + (size_t)ex_words * 4 + (32 << 20) > total) {       /* cannot happen with the wave sizing above; be safe */
            poa_arena_return(arena, d_base, total); pinned_put(0, h_reads);
            for (int t = 0; t < nw; ++t) fallback.push_back(plans[pos + t].g);
            pos = end; continue;
        }
        uint8_t *d_pool = d_base + doff; const size_t pool_bytes = total - doff;
        int32_t *d_ex = (int32_t *)d_pool;
        /* pool shares per cohort, proportional to the estimates */
        std::vector<double> est(coh.size(), 0.0); double est_tot = 0;
        for (size_t c = 0; c < coh.size(); ++c) { for (int t : coh[c].members) est[c] += plans[pos + t].pool_units_est; est_tot += est[c]; }
        if (free_run) {
            size_t o = 0;
            for (int t = 0; t < nw; ++t) {
                size_t share = (size_t)((double)pool_bytes * plans[pos + t].pool_units_est / est_tot) & ~(size_t)255;
                { static const double slab_x = [] { const char *e = getenv("ABPOA_GPU_CHAIN_SLAB_X"); return e && *e ? atof(e) : 0.0; }();      /* experiment: compact slabs */
                  if (slab_x > 0) share = std::min(share, (size_t)(plans[pos + t].pool_units_est * 16.0 * slab_x) & ~(size_t)255); }
                hs[t].pool_base = d_pool + o; hs[t].pool_units = share / 16; hs[t].pool_cursor = NULL;
                o += share;
            }
        } else {
            size_t o = 0;
            for (size_t c = 0; c < coh.size(); ++c) {
                const size_t share = c + 1 == coh.size() ? pool_bytes - o : (size_t)((double)pool_bytes * est[c] / est_tot) & ~(size_t)255;
                for (int t : coh[c].members) { hs[t].pool_base = d_pool + o; hs[t].pool_units = share / 16; hs[t].pool_cursor = d_cursors + 2 * c; }
                o += share;
            }
        }
        PoaChainParams hcp; memset(&hcp, 0, sizeof hcp);
        hcp.K = K; hcp.A = A; hcp.m = m; hcp.max_mat = abpt->max_mat; hcp.min_mis = abpt->min_mis; hcp.o1 = abpt->gap_open1; hcp.e1 = abpt->gap_ext1;
        hcp.oe1 = abpt->gap_open1 + abpt->gap_ext1; hcp.oe2 = abpt->gap_open2 + abpt->gap_ext2; hcp.record = record ? 1 : 0; hcp.P = P;
        PoaParamsDev hprm; poa_fill_params(&hprm, abpt, 15);

        /* ---- upload (stream 0 of the wave), then fork the cohort streams ---- */
        for (Cohort &c : coh) { CK(cudaStreamCreateWithFlags(&c.st, cudaStreamNonBlocking)); CK(cudaEventCreate(&c.ev_begin)); CK(cudaEventCreate(&c.ev_end)); }
        cudaStream_t s0 = coh[0].st;
        cudaEvent_t ev_up, ev_t0, ev_t1; CK(cudaEventCreateWithFlags(&ev_up, cudaEventDisableTiming)); CK(cudaEventCreate(&ev_t0)); CK(cudaEventCreate(&ev_t1));
        CK(cudaMemcpyAsync(d_reads, h_reads, reads_bytes, cudaMemcpyHostToDevice, s0));
        CK(cudaMemcpyAsync(d_slots, hs.data(), (size_t)nw * sizeof(PoaChainSlot), cudaMemcpyHostToDevice, s0));
        CK(cudaMemcpyAsync(d_cp, &hcp, sizeof hcp, cudaMemcpyHostToDevice, s0));
        CK(cudaMemcpyAsync(d_prm, &hprm, sizeof hprm, cudaMemcpyHostToDevice, s0));
        CK(cudaMemcpyAsync(d_idx, h_idx.data(), h_idx.size() * 4, cudaMemcpyHostToDevice, s0));
        CK(cudaMemcpyAsync(d_exoff, h_exoff.data(), (size_t)nw * 8, cudaMemcpyHostToDevice, s0));
        CK(cudaMemcpyAsync(d_excap, h_excap.data(), (size_t)nw * 4, cudaMemcpyHostToDevice, s0));
        CK(cudaMemsetAsync(d_cursors, 0, (size_t)32 * sizeof(unsigned long long), s0));
        const uint64_t h2d = reads_bytes + (uint64_t)nw * sizeof(PoaChainSlot) + sizeof hcp + sizeof hprm + h_idx.size() * 4 + (uint64_t)nw * 12;
        /* timed region of the device work: inputs are resident when ev_t0 fires */
        CK(cudaEventRecord(ev_t0, s0));
        poa_chain_seed_kernel<<<nw, POA_CHAIN_T, 0, s0>>>(d_slots, d_cp, nw);
        CK(cudaGetLastError());
        CK(cudaEventRecord(ev_up, s0));
        static const size_t smem_budget = [] { const char *e = getenv("ABPOA_GPU_SMEM_KB"); return (size_t)(e && *e ? atoi(e) : 28) * 1024; }();
        int ring_rows = 2, ring_cells = 64;
        poa_pick_ring(abpt->gap_mode, 16, band_cells, smem_budget, &ring_rows, &ring_cells);
        int64_t launches = 1;
        const int gaps[4] = { abpt->gap_ext1, abpt->gap_open1 + abpt->gap_ext1, abpt->gap_ext2, abpt->gap_open2 + abpt->gap_ext2 };
        const double t_uploaded = now_ms();
        for (size_t c = 1; c < coh.size(); ++c) CK(cudaStreamWaitEvent(coh[c].st, ev_up, 0));
        cudaStream_t st_dp = NULL;
        if (free_run) {
            /* two persistent kernels: the fuse workers first (one CTA per SM; they must be resident while alignment warps
             * wait for them -- 10 alignment CTAs, the most one SM takes, leave registers and shared memory for one), then one
             * alignment warp per group */
            static const double watchdog_s = [] { const char *e = getenv("ABPOA_GPU_CHAIN_WATCHDOG_S"); return e && *e ? atof(e) : 30.0; }();
            static const int fuse_per_sm = [] { const char *e = getenv("ABPOA_GPU_CHAIN_FUSE_PER_SM"); return e && *e ? std::max(1, atoi(e)) : 1; }();
            PoaChainSync hsync; memset(&hsync, 0, sizeof hsync);
            hsync.total = (int32_t)n_tasks; hsync.watchdog_ns = (unsigned long long)(watchdog_s * 1e9); hsync.tasks = d_tasks;
            CK(cudaMemcpyAsync(d_sync, &hsync, sizeof hsync, cudaMemcpyHostToDevice, s0));
            CK(cudaMemsetAsync(d_tasks, 0xff, (size_t)std::max<int64_t>(n_tasks, 1) * 4, s0));
            CK(cudaStreamCreateWithFlags(&st_dp, cudaStreamNonBlocking));
            cudaEvent_t ev_sync; CK(cudaEventCreateWithFlags(&ev_sync, cudaEventDisableTiming));
            CK(cudaEventRecord(ev_sync, s0));
            static const int fuse_cap = [] { const char *e = getenv("ABPOA_GPU_CHAIN_FUSE_WORKERS"); return e && *e ? std::max(1, atoi(e)) : (1 << 30); }();
            const int n_fuse = std::max(1, std::min(std::min(nw, sm_count * fuse_per_sm), fuse_cap));
            /* Both persistent kernels must ask for the SAME shared-memory configuration of the SM: a resident CTA pins the
             * SM's L1/shared split, and the other kernel's CTAs are only dispatched to SMs whose split matches its launch --
             * with one fuse CTA on every SM (default split of a 5 KB kernel: 64 KB) the alignment grid (split: maximum) never
             * started, and neither kernel ever ends by itself (measured: the fuse workers' watchdog fired, then the alignment
             * grid ran). */
            CK(cudaFuncSetAttribute(poa_chain_fuse_worker_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            static const bool dp_first = [] { const char *e = getenv("ABPOA_GPU_CHAIN_DP_FIRST"); return e && *e == '1'; }();     /* experiment */
            CK(cudaStreamWaitEvent(st_dp, ev_sync, 0));
            if (dp_first) CK(poa_launch_chain_dp_worker(abpt->gap_mode, gaps, d_slots, d_sync, nw, d_prm, ring_rows, ring_cells, st_dp));
            poa_chain_fuse_worker_kernel<<<n_fuse, POA_CHAIN_T, 0, s0>>>(d_slots, d_sync, d_cp);
            CK(cudaGetLastError());
            if (!dp_first) CK(poa_launch_chain_dp_worker(abpt->gap_mode, gaps, d_slots, d_sync, nw, d_prm, ring_rows, ring_cells, st_dp));
            cudaEvent_t ev_dp; CK(cudaEventCreateWithFlags(&ev_dp, cudaEventDisableTiming));
            CK(cudaEventRecord(ev_dp, st_dp));
            CK(cudaStreamWaitEvent(s0, ev_dp, 0));
            CK(cudaEventDestroy(ev_sync)); CK(cudaEventDestroy(ev_dp));
            launches += 2;
        }
        /* round-major enqueue: every cohort's stream gets its first kernels at once */
        for (int r = 1; r <= (free_run ? 0 : n_rounds); ++r)
            for (size_t c = 0; c < coh.size(); ++c) {
                cudaStream_t st = coh[c].st;
                const std::pair<size_t, int> &ro = round_of[c][(size_t)r - 1];
                if (ro.second == 0) continue;
                cudaEvent_t e0, e1, e2; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
                coh[c].marks.push_back(e0); coh[c].marks.push_back(e1); coh[c].marks.push_back(e2);
                CK(cudaEventRecord(e0, st));
                CK(poa_launch_chain_align_p16(abpt->gap_mode, gaps, d_slots, d_idx + ro.first, ro.second, r, d_prm, ring_rows, ring_cells, st));
                CK(cudaEventRecord(e1, st));
                poa_chain_fuse_kernel<<<ro.second, POA_CHAIN_T, 0, st>>>(d_slots, d_idx + ro.first, d_cp, ro.second, r);
                CK(cudaGetLastError());
                CK(cudaEventRecord(e2, st));
                launches += 2;
            }
        for (size_t c = 0; c < coh.size(); ++c) CK(cudaEventRecord(coh[c].ev_end, coh[c].st));
        const double t_enqueued = now_ms();
        for (size_t c = 1; c < coh.size(); ++c) CK(cudaStreamWaitEvent(s0, coh[c].ev_end, 0));
        CK(cudaEventRecord(ev_t1, s0));
        /* ---- results.  Default: heaviest-bundling consensus on the device, only consensus bytes come back.
         *      ABPOA_GPU_CHAIN_EXPORT_GRAPH=1: the whole graph comes back (compact export) and the host layer computes the
         *      consensus on it -- the cross-check of the device graph against the host code. ---- */
        const bool export_graph = [] { const char *e = getenv("ABPOA_GPU_CHAIN_EXPORT_GRAPH"); return e && *e == '1'; }();
        unsigned long long *d_ccur = d_cursors;               /* the pool cursors are idle now: reuse the first as the record cursor */
        int64_t *d_recoff = d_exoff;                          /* and the export offsets as record offsets */
        if (export_graph) poa_chain_export_kernel<<<nw, POA_CHAIN_T, 0, s0>>>(d_slots, d_cp, nw, d_ex, d_exoff, d_excap);
        else {
            CK(cudaMemsetAsync(d_ccur, 0, sizeof(unsigned long long), s0));
            poa_chain_consensus_kernel<<<nw, 32, 0, s0>>>(d_slots, d_cp, nw, d_ex, d_ccur, (unsigned long long)(pool_bytes / 4), d_recoff);
        }
        CK(cudaGetLastError());
        ++launches;
        std::vector<PoaChainSlot> fin((size_t)nw);
        PoaChainSlot *h_fin = (PoaChainSlot *)pinned_get(1, (size_t)nw * sizeof(PoaChainSlot));
        CK(cudaMemcpyAsync(h_fin, d_slots, (size_t)nw * sizeof(PoaChainSlot), cudaMemcpyDeviceToHost, s0));
        CK(cudaStreamSynchronize(s0));
        float dev_ms = 0.f; CK(cudaEventElapsedTime(&dev_ms, ev_t0, ev_t1));
        double dp_ms = 0, fuse_ms = 0; int64_t n_marks = 0;
        for (Cohort &c : coh) {
            for (size_t k = 0; k + 2 < c.marks.size(); k += 3) {
                float a = 0.f, b = 0.f;
                CK(cudaEventElapsedTime(&a, c.marks[k], c.marks[k + 1])); CK(cudaEventElapsedTime(&b, c.marks[k + 1], c.marks[k + 2]));
                dp_ms += a; fuse_ms += b; ++n_marks;
            }
            for (cudaEvent_t ev : c.marks) cudaEventDestroy(ev);
            c.marks.clear();
        }
        memcpy(fin.data(), h_fin, (size_t)nw * sizeof(PoaChainSlot));
        pinned_put(1, h_fin);
        double wait_ms = 0;
        if (free_run) {                   /* no per-launch marks: time inside the alignments (SM cycles) and inside chain_fuse, summed over groups */
            int khz = 0; if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev) != cudaSuccess || khz <= 0) khz = 1965000;
            for (int t = 0; t < nw; ++t) {
                dp_ms += (double)(fin[t].fwd_clk + fin[t].bt_clk) / (double)khz; fuse_ms += (double)fin[t].fuse_ns * 1e-6; wait_ms += (double)fin[t].wait_ns * 1e-6;
                n_marks += fin[t].fused > 0 ? fin[t].fused - 1 : 0;
            }
            CK(cudaStreamDestroy(st_dp));
            PoaChainSync hs2; CK(cudaMemcpy(&hs2, d_sync, sizeof hs2, cudaMemcpyDeviceToHost));
            if (hs2.abort) fprintf(stderr, "[libabpoa_b200/chain] watchdog: the alignment and the fuse kernel did not make progress together within ABPOA_GPU_CHAIN_WATCHDOG_S; "
                                           "unfinished groups go to the launch engine (ABPOA_GPU_CHAIN_ROUNDS=1 selects the round schedule)\n");
        }
        const double t_dev_done = now_ms();
        /* ---- device consensus: record offsets, then one copy of all records ---- */
        std::vector<int64_t> recoff((size_t)nw, -1); int32_t *h_cons = NULL; unsigned long long cons_words = 0;
        if (!export_graph) {
            int64_t *h_ro = NULL; CK(cudaHostAlloc((void **)&h_ro, (size_t)nw * 8 + 8, cudaHostAllocDefault));
            CK(cudaMemcpyAsync(h_ro, d_recoff, (size_t)nw * 8, cudaMemcpyDeviceToHost, s0));
            CK(cudaMemcpyAsync(h_ro + nw, d_ccur, 8, cudaMemcpyDeviceToHost, s0));
            CK(cudaStreamSynchronize(s0));
            memcpy(recoff.data(), h_ro, (size_t)nw * 8); cons_words = (unsigned long long)h_ro[nw];
            CK(cudaFreeHost(h_ro));
            if (cons_words > pool_bytes / 4) cons_words = pool_bytes / 4;
            h_cons = (int32_t *)pinned_get(2, (size_t)std::max<unsigned long long>(cons_words, 1) * 4);
            if (cons_words) CK(cudaMemcpyAsync(h_cons, d_ex, (size_t)cons_words * 4, cudaMemcpyDeviceToHost, s0));
        }
        /* word counts: header words 0..3 of every record */
        std::vector<int32_t> hdr4((size_t)nw * 4);
        if (export_graph) {
            int32_t *h_hdr = NULL; CK(cudaHostAlloc((void **)&h_hdr, (size_t)nw * 16, cudaHostAllocDefault));
            for (int t = 0; t < nw; ++t) CK(cudaMemcpyAsync(h_hdr + 4 * t, d_ex + h_exoff[t], 16, cudaMemcpyDeviceToHost, s0));
            CK(cudaStreamSynchronize(s0));
            memcpy(hdr4.data(), h_hdr, (size_t)nw * 16);
            CK(cudaFreeHost(h_hdr));
        }
        std::vector<int64_t> words((size_t)nw, 0), hoff2((size_t)nw, 0); int64_t tot_words = 0;
        for (int t = 0; t < nw; ++t) {
            if (!export_graph) { words[t] = (!fin[t].failed && recoff[t] >= 0) ? 1 : 0; continue; }
            if (fin[t].failed || hdr4[4 * t] < 2) continue;
            words[t] = 4 + 5ll * hdr4[4 * t] + 4ll * hdr4[4 * t + 1] + hdr4[4 * t + 2];
            hoff2[t] = tot_words; tot_words += words[t];
        }
        int32_t *h_ex = (int32_t *)pinned_get(3, (size_t)std::max<int64_t>(tot_words, 1) * 4);
        if (export_graph) for (int t = 0; t < nw; ++t) if (words[t]) CK(cudaMemcpyAsync(h_ex + hoff2[t], d_ex + h_exoff[t], (size_t)words[t] * 4, cudaMemcpyDeviceToHost, s0));
        /* per-read records */
        std::vector<std::vector<int32_t>> rs((size_t)nw), rn((size_t)nw); std::vector<std::vector<uint64_t>> rh((size_t)nw);
        if (record) for (int t = 0; t < nw; ++t) {
            const int nr = plans[pos + t].n_reads;
            rs[t].resize(nr); rn[t].resize(nr); rh[t].resize(nr);
            CK(cudaMemcpyAsync(rs[t].data(), fin[t].rec_score, (size_t)nr * 4, cudaMemcpyDeviceToHost, s0));
            CK(cudaMemcpyAsync(rn[t].data(), fin[t].rec_nops, (size_t)nr * 4, cudaMemcpyDeviceToHost, s0));
            CK(cudaMemcpyAsync(rh[t].data(), fin[t].rec_hash, (size_t)nr * 8, cudaMemcpyDeviceToHost, s0));
        }
        CK(cudaStreamSynchronize(s0));
        const uint64_t d2h = (uint64_t)tot_words * 4 + (uint64_t)cons_words * 4 + (uint64_t)nw * (sizeof(PoaChainSlot) + 16);
        poa_arena_return(arena, d_base, total);
        const double t_copied = now_ms();

        /* ---- host: rebuild each graph, consensus ---- */
        std::atomic<int> next(0); std::atomic<int> n_failed(0);
        std::vector<int> failed_groups; std::mutex fmu;
        int64_t cells = 0, alns = 0, fwd_clk = 0, bt_clk = 0;
        for (int t = 0; t < nw; ++t) if (!fin[t].failed && words[t] && fin[t].fused == plans[pos + t].n_reads) { cells += fin[t].cells; alns += plans[pos + t].n_reads - 1; fwd_clk += fin[t].fwd_clk; bt_clk += fin[t].bt_clk; }
        const int nth = std::max(1, std::min(n_workers, nw));
        std::vector<std::thread> th;
        for (int w = 0; w < nth; ++w)
            th.emplace_back([&, w]() {
                (void)w;
                abpoa_t *ab = abpoa_init();
                for (;;) {
                    const int t = next.fetch_add(1);
                    if (t >= nw) break;
                    const GroupPlan &p = plans[pos + t];
                    if (fin[t].failed || !words[t] || fin[t].fused != p.n_reads) {
                        if (verbose) fprintf(stderr, "[chain] group %d left the device chain after %d reads (flags 0x%x)\n", p.g, fin[t].fused, fin[t].failed);
                        std::lock_guard<std::mutex> lk(fmu); failed_groups.push_back(p.g); n_failed += 1; continue;
                    }
                    abpoa_gpu_group_result_t *o = &results[p.g];
                    memset(o, 0, sizeof *o);
                    abpoa_reset(ab, abpt, p.qmax);
                    abpoa_seq_t *abs = ab->abs;
                    abs->n_seq = p.n_reads; poa_seq_reserve(abs);
                    for (int i = 0; i < p.n_reads; ++i) { abs->is_rc[i] = 0; abs->name[i].l = 0; }
                    if (export_graph) poa_graph_import(ab, abpt, h_ex + hoff2[t]);
                    else {                                     /* the device's consensus: base | coverage << 8 per position */
                        const int32_t *rec = h_cons + recoff[t];
                        const int len = rec[0];
                        std::vector<uint8_t> cb((size_t)(len > 0 ? len : 1)); std::vector<int> cc((size_t)(len > 0 ? len : 1));
                        for (int j = 0; j < len; ++j) { cb[j] = (uint8_t)(rec[1 + j] & 0xff); cc[j] = rec[1 + j] >> 8; }
                        poa_cons_install(ab, p.n_reads, len, cb.data(), cc.data());
                    }
                    poa_finish_group_result(ab, abpt, o, emit, p.g);
                    o->dp_cells = fin[t].cells; o->n_aligned = p.n_reads - 1;
                    if (record) {
                        const int nr = p.n_reads;
                        o->read_best_score = (int32_t *)poa_xcalloc((size_t)nr, sizeof(int32_t));
                        o->read_n_cigar = (int32_t *)poa_xcalloc((size_t)nr, sizeof(int32_t));
                        o->read_cigar_hash = (uint64_t *)poa_xcalloc((size_t)nr, sizeof(uint64_t));
                        o->read_cigar_hash[0] = 1469598103934665603ull;               /* FNV-1a of an empty CIGAR, as the other engine records it */
                        for (int i = 1; i < nr; ++i) { o->read_best_score[i] = rs[t][i]; o->read_n_cigar[i] = rn[t][i]; o->read_cigar_hash[i] = rh[t][i]; }
                    }
                }
                abpoa_free(ab);
            });
        for (auto &x : th) x.join();
        for (int g : failed_groups) fallback.push_back(g);
        pinned_put(3, h_ex); pinned_put(0, h_reads); if (h_cons) pinned_put(2, h_cons);
        for (Cohort &c : coh) { cudaEventDestroy(c.ev_begin); cudaEventDestroy(c.ev_end); cudaStreamDestroy(c.st); }
        cudaEventDestroy(ev_up); cudaEventDestroy(ev_t0); cudaEventDestroy(ev_t1);
        if (stats) {
            stats->device_ms += dev_ms; stats->cells += cells; stats->alignments += alns; stats->launches += launches;
            stats->h2d_bytes += h2d; stats->d2h_bytes += d2h; stats->groups_done += nw - n_failed.load(); stats->groups_failed += n_failed.load();
            stats->fwd_clk += fwd_clk; stats->bt_clk += bt_clk;
            stats->dp_ms += dp_ms; stats->fuse_ms += fuse_ms; stats->dp_launches += n_marks; stats->fuse_launches += n_marks;
            stats->wait_ms += wait_ms; stats->free_running = free_run ? 1 : 0;
        }
#ifdef POA_KPROF
        if (verbose) {
            double pf[6] = {0, 0, 0, 0, 0, 0}; double na = 0;
            for (int t = 0; t < nw; ++t) { for (int z = 0; z < 6; ++z) pf[z] += (double)fin[t].prof[z]; na += fin[t].fused > 0 ? fin[t].fused - 1 : 0; }
            double bd[4] = {0, 0, 0, 0};
            for (int t = 0; t < nw; ++t) for (int z = 0; z < 4; ++z) bd[z] += (double)fin[t].btdiag[z];
            if (na > 0) fprintf(stderr, "[chain, backtrace per alignment] %.0f steps in %.0f speculative rounds, %.0f general steps costing %.0f k-cycles\n", bd[0] / na, bd[1] / na, bd[2] / na, bd[3] / na);
            if (na > 0) fprintf(stderr, "[chain, k-cycles/alignment] -DPOA_KPROF phases: setup %.0f pred %.0f compute %.0f store %.0f rowmax %.0f tail+prefetch %.0f\n",
                                pf[0] / na / 1e3, pf[1] / na / 1e3, pf[2] / na / 1e3, pf[3] / na / 1e3, pf[4] / na / 1e3, pf[5] / na / 1e3);
        }
#endif
        if (verbose)
            free_run ? fprintf(stderr, "[chain] free-running: per group on average %.1f ms inside alignments + %.1f ms inside fuse (waited %.1f ms for fuse workers incl. the fuse itself), %lld alignments\n",
                               dp_ms / nw, fuse_ms / nw, wait_ms / nw, (long long)n_marks)
                     : fprintf(stderr, "[chain] DP kernels %.1f ms + fuse kernels %.1f ms summed over %zu concurrent cohort streams (%lld rounds)\n", dp_ms, fuse_ms, coh.size(), (long long)n_marks),
            fprintf(stderr, "[chain] wave of %d groups (%zu cohorts, K=%d): stage %.0f + upload-enqueue %.0f + launch-enqueue %.0f ms, stage+launch+device %.0f ms (device %.1f ms), export copy %.0f ms, import+consensus %.0f ms; "
                            "static %.2f GB, pool %.2f GB, export %.1f MB; %d groups handed to the launch engine\n",
                    nw, coh.size(), K, t_staged - t_wave0, t_uploaded - t_staged, t_enqueued - t_uploaded, t_dev_done - t_wave0, dev_ms, t_copied - t_dev_done, now_ms() - t_copied,
                    (double)doff / 1e9, (double)pool_bytes / 1e9, (double)tot_words * 4 / 1e6, n_failed.load());
        pos = end;
    }
    return 0;
}
