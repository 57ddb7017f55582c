#!/usr/bin/env python
"""Quick experiment: e2e throughput of the batch engine on a workload subset."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from abpoa_b200 import synth
from abpoa_b200.aligner import make_para
from abpoa_b200.batch import BatchEngine, PackedGroups
from abpoa_b200 import capi

name = sys.argv[1]; n_groups = int(sys.argv[2]); workers = int(sys.argv[3]); gpl = int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
w = synth.WORKLOADS[name]
t0 = time.time(); groups = w.groups(n_groups); print(f"gen {time.time()-t0:.1f}s", flush=True)
packed = PackedGroups(groups)
lib = capi.product()
abpt = make_para(lib, w.cfg)
with BatchEngine(n_workers=workers, groups_per_launch=gpl) as eng:
    for rep in range(reps):
        eng.reset_stats()
        t0 = time.time()
        res = eng.run_packed(abpt, packed, keep_results=False)
        dt = time.time() - t0
        st = eng.stats()
        cells = sum(r[0] for r in res)
        print(f"{name} groups={n_groups} workers={workers} gpl={gpl}: wall {dt:.2f}s cells {cells/1e9:.2f}G -> {cells/dt/1e9:.2f} GCUPS e2e, "
              f"reads/s {packed.total_reads/dt:.0f}; kernel_ms(sum over streams) {st['kernel_ms']:.0f} launches {st['launches']} retries {st['retries']} "
              f"h2d {st['h2d_bytes']/1e9:.2f}GB d2h {st['d2h_bytes']/1e9:.2f}GB; per-aln fwd {st['fwd_clk']/max(st['alignments'],1)/1e3:.0f}k clk bt {st['bt_clk']/max(st['alignments'],1)/1e3:.0f}k clk", flush=True)
