"""Multi-GPU: shard independent read groups over ranks (one process per GPU).

The path shards naturally: groups share nothing, a single alignment never spans GPUs,
so there is NO collective on the data path.  torch.distributed (NCCL on GPUs, gloo on CPU for the
tests) is used only to scatter the encoded reads from rank 0 and to gather the per-group results.
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np


def shard_bounds(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [lo, hi) of `n_items` for `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def lpt_assignment(costs: Sequence[float], world: int) -> list[list[int]]:
    """Longest-processing-time-first assignment of groups to ranks for unequal groups
    (cost ~ sum of read_len^2); returns the group indices of every rank, each in input order."""
    load = [0.0] * world
    out: list[list[int]] = [[] for _ in range(world)]
    for g in sorted(range(len(costs)), key=lambda i: -costs[i]):
        r = min(range(world), key=lambda k: load[k])
        out[r].append(g)
        load[r] += costs[g]
    return [sorted(x) for x in out]


def group_cost(reads: Sequence[np.ndarray]) -> float:
    return float(sum(len(r) for r in reads)) * (max((len(r) for r in reads), default=0) + 1)


def distributed_msa(groups, cfg, runner: Callable | None = None, balance: bool = True):
    """Run the MSA of every group on the ranks of the default process group.

    groups : list of read groups on rank 0 (ignored elsewhere).
    runner : callable(cfg, list_of_groups) -> list of per-group results; defaults to the B200
             batch engine on the rank's current CUDA device.
    Returns the list of per-group results in input order on rank 0, None on the other ranks.
    """
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    if runner is None:
        from .batch import BatchEngine

        def runner(c, gs):
            with BatchEngine() as eng:
                return eng.run(c, gs)

    # ---- scatter: rank 0 decides the assignment and ships every rank its groups ----
    if rank == 0:
        n = len(groups)
        if balance:
            assign = lpt_assignment([group_cost(g) for g in groups], world)
        else:
            assign = [list(range(*shard_bounds(n, r, world))) for r in range(world)]
        payload = [(idx, [groups[i] for i in idx]) for idx in assign]
    else:
        payload = [None] * world
    mine = [None]
    dist.scatter_object_list(mine, payload, src=0)
    idx, my_groups = mine[0]

    # ---- compute: no communication ----
    results = runner(cfg, my_groups) if my_groups else []

    # ---- gather ----
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((idx, results), gathered, dst=0)
    if rank != 0:
        return None
    out = [None] * sum(len(i) for i, _ in gathered)
    for i, res in gathered:
        for g, r in zip(i, res):
            out[g] = r
    return out
