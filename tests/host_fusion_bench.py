#!/usr/bin/env python
"""Dev tool (not a test): time the product's HOST graph code -- graph fusion + topological sort --
without a GPU.  Alignments (graph-CIGARs) are recorded once from the unmodified reference
(oracle/_ref) and replayed into the product's graph code, many groups interleaved round by round the
way a batch worker cycles through its chunk (so the cache behaviour is comparable).

    python tests/host_fusion_bench.py [workload] [n_groups] [reps] [fast_order=1]
"""
from __future__ import annotations

import ctypes as C
import pickle
import sys
import time
from multiprocessing import Pool
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from abpoa_b200 import capi, synth                      # noqa: E402
from abpoa_b200.aligner import PoaSession, make_para    # noqa: E402
from abpoa_b200.capi import abpoa_res_t, c_int_p, c_u8_p  # noqa: E402


def _record(args):
    name, gi = args
    w = synth.WORKLOADS[name]
    reads = w.groups(gi + 1)[gi]
    with PoaSession(w.cfg, capi.load_library(Path(__file__).resolve().parents[1] / "oracle" / "_ref" / "libabpoa_ref.so")) as s:
        alns = s.run_reads(reads, count_cells=False)
    return [a.cigar for a in alns]


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "convex_10k"
    G = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    fast = int(sys.argv[4]) if len(sys.argv) > 4 else 1      # spliced topological order (what the batch engine uses)
    cache = Path(f"/tmp/hfb_{name}_{G}.pkl")
    w = synth.WORKLOADS[name]
    groups = w.groups(G)
    if cache.exists():
        cigars = pickle.loads(cache.read_bytes())
    else:
        t0 = time.time()
        with Pool(8) as p:
            cigars = p.map(_record, [(name, g) for g in range(G)])
        cache.write_bytes(pickle.dumps(cigars))
        print(f"recorded {G} groups with the reference in {time.time() - t0:.1f}s", flush=True)

    lib = capi.product()
    abpt = make_para(lib, w.cfg)
    dll = lib.dll
    dll.poa_prof_snapshot.argtypes = [C.POINTER(C.c_double), C.c_int]
    dll.poa_add_alignment_nosync.restype = C.c_int
    prof = (C.c_double * 8)()

    class BlobPlan(C.Structure):
        _fields_ = [("n_rows", C.c_int), ("n_pred_max", C.c_int), ("qlen", C.c_int), ("beg_index", C.c_int), ("whole_graph", C.c_int),
                    ("w", C.c_int), ("with_remain", C.c_int), ("with_score", C.c_int), ("bytes", C.c_size_t)]
    plan = BlobPlan()
    blob = np.zeros(1 << 20, dtype=np.uint8)
    n_reads = max(len(g) for g in groups)
    for rep in range(reps):
        abs_ = [lib.abpoa_init() for _ in range(G)]
        for g in range(G):
            lib.abpoa_reset(abs_[g], abpt, max(len(r) for r in groups[g]))
            dll.poa_graph_set_fast_order(abs_[g].contents.abg, fast)
        dll.poa_prof_snapshot(prof, 1)
        t_sort = t_fuse = t_flat = 0.0
        for r in range(n_reads):
            for g in range(G):
                if r >= len(groups[g]):
                    continue
                ab = abs_[g]
                abg = ab.contents.abg
                t0 = time.perf_counter()
                if abg.contents.node_n > 2 and not abg.contents.is_topological_sorted:
                    lib.abpoa_topological_sort(abg, abpt)
                t1 = time.perf_counter()
                res = abpoa_res_t()
                cg = cigars[g][r]
                res.n_cigar = len(cg)
                res.graph_cigar = cg.ctypes.data_as(C.POINTER(C.c_uint64)) if len(cg) else None
                seq = np.ascontiguousarray(groups[g][r], dtype=np.uint8)
                dll.poa_add_alignment_nosync(ab, abpt, 0, 1, seq.ctypes.data_as(c_u8_p), None, len(seq), None, res, r, len(groups[g]), 1)
                t2 = time.perf_counter()
                # flattening of the fused graph for the NEXT read (what the batch worker does before a launch)
                if r + 1 < len(groups[g]):
                    nxt = np.ascontiguousarray(groups[g][r + 1], dtype=np.uint8)
                    dll.poa_blob_plan_make(C.byref(plan), abg, abpt, 0, 1, len(nxt))
                    if plan.bytes > len(blob):
                        blob = np.zeros(int(plan.bytes) * 2, dtype=np.uint8)
                    dll.poa_blob_fill(blob.ctypes.data_as(c_u8_p), C.byref(plan), abg, abpt, 0, 1, nxt.ctypes.data_as(c_u8_p))
                t3 = time.perf_counter()
                t_sort += t1 - t0
                t_fuse += t2 - t1
                t_flat += t3 - t2
        dll.poa_prof_snapshot(prof, 0)
        n_f = sum(len(g) for g in groups)
        print(f"{name} G={G}: per fusion: sort {t_sort / n_f * 1e3:.3f} ms (bfs {prof[0] / n_f:.3f} order {prof[1] / n_f:.3f} remain {prof[2] / n_f:.3f}) "
              f"fuse {t_fuse / n_f * 1e3:.3f} ms (thread_cigar {prof[3] / n_f:.3f}) flatten {t_flat / n_f * 1e3:.3f} ms", flush=True)
        sig = 0
        for g in range(G):
            sig = (sig * 1000003 + abs_[g].contents.abg.contents.node_n) & 0xFFFFFFFF
            lib.abpoa_free(abs_[g])
        print("  node-count signature", sig)


if __name__ == "__main__":
    main()
