/* chain_emul.cpp -- TEST INFRASTRUCTURE: the device-side graph code of abpoa_b200/csrc/poa_chain.cuh compiled
 * for the host (-DPOA_CHAIN_EMUL: PAR_FOR = plain loop, barrier = nothing) behind a tiny C interface, so that
 * the CPU suite can drive it read by read next to the host graph layer (poa_graph.c / poa_flat.c) and compare
 * every array.  Nothing in the product links this file. */
#define POA_CHAIN_EMUL 1
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "poa_chain.cuh"

struct Emul {
    PoaChainSlot s; PoaChainParams cp;
    std::vector<uint8_t> mem, blob, reads; std::vector<int32_t> read_off, read_w;
    std::vector<uint64_t> cigar; PoaResultDev res;
    std::vector<int32_t> rec_score, rec_nops; std::vector<uint64_t> rec_hash;
};

extern "C" Emul *chain_emul_new(int n_reads, const int32_t *lens, const uint8_t *const *seqs, const int32_t *w, int n_cap, int K, int A,
                                int m, int max_mat, int min_mis, int o1, int e1, int oe1, int oe2) {
    Emul *e = new Emul();
    memset(&e->s, 0, sizeof e->s); memset(&e->cp, 0, sizeof e->cp); memset(&e->res, 0, sizeof e->res);
    e->cp.K = K; e->cp.A = A; e->cp.m = m; e->cp.max_mat = max_mat; e->cp.min_mis = min_mis; e->cp.o1 = o1; e->cp.e1 = e1; e->cp.oe1 = oe1; e->cp.oe2 = oe2; e->cp.record = 1;
    int qmax = 1; e->read_off.push_back(0);
    for (int i = 0; i < n_reads; ++i) { e->reads.insert(e->reads.end(), seqs[i], seqs[i] + lens[i]); e->read_off.push_back((int32_t)e->reads.size()); e->read_w.push_back(w[i]); if (lens[i] > qmax) qmax = lens[i]; }
    PoaChainSlot &s = e->s;
    s.n_cap = n_cap; s.pred_cap = n_cap * 4; s.n_reads = n_reads;
    const size_t scr_n = (size_t)(qmax + 2 > n_cap ? qmax + 2 : n_cap);
    size_t bytes = (size_t)n_cap * (1 + 4 * 4 + 4 * K * 4 + A * 4 + 4 * 4) + 6 * scr_n * 4 + 4096;
    e->mem.assign(bytes, 0xcd);                            /* poison: nothing may rely on zeroed memory */
    uint8_t *p = e->mem.data();
    auto take = [&](size_t b) { uint8_t *q = p; p += (b + 15) & ~(size_t)15; return q; };
    s.base = take(n_cap);
    s.in_cnt = (int32_t *)take((size_t)n_cap * 4); s.out_cnt = (int32_t *)take((size_t)n_cap * 4); s.aln_cnt = (int32_t *)take((size_t)n_cap * 4); s.n_read = (int32_t *)take((size_t)n_cap * 4);
    s.in_id = (int32_t *)take((size_t)n_cap * K * 4); s.in_w = (int32_t *)take((size_t)n_cap * K * 4);
    s.out_id = (int32_t *)take((size_t)n_cap * K * 4); s.out_w = (int32_t *)take((size_t)n_cap * K * 4);
    s.aln_id = (int32_t *)take((size_t)n_cap * A * 4);
    s.order[0] = (int32_t *)take((size_t)n_cap * 4); s.order[1] = (int32_t *)take((size_t)n_cap * 4);
    s.node_row = (int32_t *)take((size_t)n_cap * 4); s.rem_row = (int32_t *)take((size_t)n_cap * 4);
    for (int k = 0; k < 6; ++k) s.scr[k] = (int32_t *)take(scr_n * 4);
    s.blob_cap = (int32_t)(256 + ((size_t)n_cap + 1) * 8 + (size_t)s.pred_cap * 4 + qmax + 64);
    e->blob.assign((size_t)s.blob_cap, 0xcd);
    e->cigar.assign((size_t)qmax + n_cap + 8, 0);
    s.reads = e->reads.data(); s.read_off = e->read_off.data(); s.read_w = e->read_w.data();
    s.jd.blob = e->blob.data(); s.jd.cigar = e->cigar.data(); s.jd.result = &e->res;
    e->rec_score.assign(n_reads, 0); e->rec_nops.assign(n_reads, 0); e->rec_hash.assign(n_reads, 0);
    s.rec_score = e->rec_score.data(); s.rec_nops = e->rec_nops.data(); s.rec_hash = e->rec_hash.data();
    return e;
}
extern "C" void chain_emul_free(Emul *e) { delete e; }
extern "C" void chain_emul_seed(Emul *e) { chain_seed(&e->s, &e->cp); }
/* ops: graph-CIGAR words in BACKTRACK order with DP rows (what the alignment kernel leaves in jd.cigar) */
extern "C" int chain_emul_fuse(Emul *e, const uint64_t *ops, int n_ops, int best_score, int64_t cells) {
    memcpy(e->cigar.data(), ops, (size_t)n_ops * 8);
    e->res.status = POA_ST_OK; e->res.n_ops = n_ops; e->res.best_score = best_score; e->res.cells = cells;
    chain_fuse(&e->s, &e->cp, e->s.fused);
    return e->s.failed;
}
extern "C" const PoaChainSlot *chain_emul_slot(Emul *e) { return &e->s; }
extern "C" int chain_emul_n_nodes(Emul *e) { return e->s.n_nodes; }
extern "C" int chain_emul_failed(Emul *e) { return e->s.failed; }
extern "C" const int32_t *chain_emul_array(Emul *e, int which) {
    PoaChainSlot &s = e->s;
    switch (which) {
    case 0: return s.in_cnt; case 1: return s.out_cnt; case 2: return s.aln_cnt; case 3: return s.n_read;
    case 4: return s.in_id; case 5: return s.in_w; case 6: return s.out_id; case 7: return s.out_w; case 8: return s.aln_id;
    case 9: return s.order[s.cur]; case 10: return s.node_row; case 11: return s.rem_row;
    case 12: return s.rec_score; case 13: return s.rec_nops;
    }
    return NULL;
}
extern "C" const uint8_t *chain_emul_bases(Emul *e) { return e->s.base; }
extern "C" const uint8_t *chain_emul_blob(Emul *e) { return e->blob.data(); }
extern "C" const uint64_t *chain_emul_hashes(Emul *e) { return e->rec_hash.data(); }
extern "C" int64_t chain_emul_cells(Emul *e) { return e->s.cells; }
extern "C" int chain_emul_consensus(Emul *e, int32_t *out, int cap) { chain_consensus(&e->s, &e->cp, out, cap); return out[0]; }
