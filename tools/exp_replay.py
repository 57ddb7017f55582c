#!/usr/bin/env python
"""Quick experiment: device-resident replay throughput (the bench's `value`) on a subset of a workload.
    python tools/exp_replay.py convex_10k 200 [workers]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from abpoa_b200 import synth, capi
from abpoa_b200.aligner import make_para
from abpoa_b200.batch import BatchEngine, PackedGroups

name = sys.argv[1]; n_groups = int(sys.argv[2]); workers = int(sys.argv[3]) if len(sys.argv) > 3 else 32
gpl = int(sys.argv[4]) if len(sys.argv) > 4 else 0
w = synth.WORKLOADS[name]
packed = PackedGroups(w.groups(n_groups))
lib = capi.product()
abpt = make_para(lib, w.cfg)
with BatchEngine(n_workers=workers, groups_per_launch=gpl) as eng:
    t0 = time.time()
    eng.run_packed(abpt, packed, keep_results=False, capture=True)
    print(f"capture pass {time.time()-t0:.2f}s", flush=True)
    import torch                      # NVTX range so that `ncu --nvtx --nvtx-include replay/` profiles only the replay launches
    torch.cuda.nvtx.range_push("replay")
    r = eng.replay(abpt, warmup=1, repeats=2)
    torch.cuda.nvtx.range_pop()
    print(f"{name} n_groups={n_groups}: replay jobs {r['n_jobs']} cells {r['cells']/1e9:.2f}G kernel_ms {r['kernel_ms']:.1f} (min {r['kernel_ms_min']:.1f}) -> "
          f"{r['cells']/r['kernel_ms_min']/1e6:.1f} GCUPS; launches {r['launches']} mismatches {r['mismatches']}", flush=True)
