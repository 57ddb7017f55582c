"""CPU suite, part 3: the N>1 path on the gloo backend (world_size 2).

Groups are sharded over ranks, each rank computes its shard with no communication, rank 0 gathers.
On the CPU box the per-rank runner is the oracle-driven harness (the product's host layer + scalar
oracle alignments); on a GPU box the same function runs the batch engine per rank."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

from abpoa_b200.parallel import lpt_assignment, shard_bounds

ROOT = Path(__file__).resolve().parent.parent

WORKER = r'''
import os, sys, pickle
sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
import numpy as np
import torch.distributed as dist
from abpoa_b200 import capi, synth
from abpoa_b200.aligner import PoaConfig
from abpoa_b200.parallel import distributed_msa
from helpers import run_group

def runner(cfg, groups):
    lib = capi.product()
    return [run_group(lib, cfg, g, want_msa=False, use_oracle=True)["cons"] for g in groups]

dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
cfg = PoaConfig()
groups = [synth.make_group(300 + g, 4 + g %% 3, 150 + 40 * (g %% 4), 0.05) for g in range(7)] if dist.get_rank() == 0 else None
out = distributed_msa(groups, cfg, runner=runner)
if dist.get_rank() == 0:
    single = runner(cfg, groups)
    ok = len(out) == len(single) and all(len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b)) for a, b in zip(out, single))
    print("PARITY_OK" if ok else "PARITY_FAIL")
dist.barrier()
dist.destroy_process_group()
'''


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_lpt_assignment_is_a_partition_and_balanced():
    rng = np.random.default_rng(0)
    costs = rng.integers(1, 100, size=41).tolist()
    parts = lpt_assignment(costs, 4)
    assert sorted(i for p in parts for i in p) == list(range(41))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= max(costs)


def test_two_ranks_gloo_match_single_process(tmp_path):
    port = 29500 + os.getpid() % 2000
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": str(ROOT), "tests": str(ROOT / "tests"), "port": port})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "PARITY_OK" in outs[0], outs[0]
