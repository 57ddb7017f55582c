"""Synthetic read groups of the shapes BASELINE.json names.

Per group g (seed = base_seed + g): a template of length L drawn uniformly over the
alphabet; every read is the template sent through an i.i.d. per-base channel with total
error e: substitution 0.4 e (uniform over the other letters), deletion 0.3 e,
insertion-after 0.3 e (uniform letter).  Reads stay in generation order.  numpy's PCG64
is used (instead of the survey's pure-python generator) so that the 10-kbp x 50 x 1000
set is produced in seconds; both arms of every comparison consume the same arrays.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .aligner import AA_ORDER, PoaConfig
from .capi import ABPOA_GLOBAL_MODE, ABPOA_LOCAL_MODE

AA20 = "ACDEFGHIKLMNPQRSTVWY"


def _alphabet_codes(m: int) -> np.ndarray:
    if m == 5:
        return np.arange(4, dtype=np.uint8)
    return np.array([AA_ORDER.index(c) for c in AA20], dtype=np.uint8)


def make_group(seed: int, n_reads: int, length: int, err: float, m: int = 5) -> list[np.ndarray]:
    rng = np.random.default_rng(seed)
    alpha = _alphabet_codes(m)
    k = len(alpha)
    template = rng.integers(0, k, size=length)
    reads = []
    for _ in range(n_reads):
        u = rng.random(length)
        keep = u >= 0.3 * err                              # deletion below
        sub = (u >= 0.3 * err) & (u < 0.7 * err)           # substitution
        base = template.copy()
        shift = rng.integers(1, k, size=length)
        base[sub] = (base[sub] + shift[sub]) % k
        ins = rng.random(length) < 0.3 * err               # insertion after position i
        ins_base = rng.integers(0, k, size=length)
        # interleave: position i contributes [base_i if keep_i] + [ins_i if ins]
        cnt = keep.astype(np.int64) + ins.astype(np.int64)
        off = np.concatenate(([0], np.cumsum(cnt)))
        out = np.empty(off[-1], dtype=np.int64)
        out[off[:-1][keep]] = base[keep]
        out[(off[:-1] + keep)[ins]] = ins_base[ins]
        reads.append(alpha[out])
    return reads


@dataclass(frozen=True)
class Workload:
    """One BASELINE.json config."""
    name: str
    n_groups: int
    n_reads: int
    length: int
    err: float
    cfg: PoaConfig

    def groups(self, n_groups: int | None = None, base_seed: int = 1000, first: int = 0):
        n = self.n_groups if n_groups is None else n_groups
        return [make_group(base_seed + first + g, self.n_reads, self.length, self.err, self.cfg.m) for g in range(n)]


def _blosum_path() -> str:
    from .capi import REPO_ROOT
    return str(REPO_ROOT / "abpoa_b200" / "data" / "BLOSUM62.mtx")


WORKLOADS = {
    # configs[1]: 1000 groups x 20 reads x 1 kbp, global, affine (-O 4 -E 2)
    "affine_1k": Workload("affine_1k", 1000, 20, 1000, 0.05,
                          PoaConfig(align_mode=ABPOA_GLOBAL_MODE, gap_open1=4, gap_ext1=2, gap_open2=0, gap_ext2=0)),
    # configs[2]: 1000 groups x 50 reads x 10 kbp, global, convex (-O 4,24 -E 2,1): the headline
    "convex_10k": Workload("convex_10k", 1000, 50, 10000, 0.05,
                           PoaConfig(align_mode=ABPOA_GLOBAL_MODE, gap_open1=4, gap_ext1=2, gap_open2=24, gap_ext2=1)),
    # not a BASELINE config: the headline shape with AFFINE gaps -- the shape north_star's "60 % HBM roofline on the
    # affine inner kernel" is measured on (widest practical rows, many groups)
    "affine_10k": Workload("affine_10k", 1000, 50, 10000, 0.05,
                           PoaConfig(align_mode=ABPOA_GLOBAL_MODE, gap_open1=4, gap_ext1=2, gap_open2=0, gap_ext2=0)),
    # configs[3]: 500 groups x 100 reads x 5 kbp, local, linear (-m1 -O 0 -E 2)
    "local_linear_5k": Workload("local_linear_5k", 500, 100, 5000, 0.05,
                                PoaConfig(align_mode=ABPOA_LOCAL_MODE, gap_open1=0, gap_ext1=2, gap_open2=0, gap_ext2=0)),
    # configs[4]: 200 groups x 30 seqs x 2 kaa, BLOSUM62, global, affine
    "aa_blosum62_2k": Workload("aa_blosum62_2k", 200, 30, 2000, 0.10,
                               PoaConfig(align_mode=ABPOA_GLOBAL_MODE, m=27, score_matrix=_blosum_path(),
                                         gap_open1=4, gap_ext1=2, gap_open2=0, gap_ext2=0)),
}
