"""Named parity cases shared by the golden-vector generator and the tests.
Each case = PoaConfig keyword arguments + either a fixture file or a synthetic shape."""
from __future__ import annotations

from abpoa_b200 import synth
from abpoa_b200.capi import ABPOA_EXTEND_MODE, ABPOA_LOCAL_MODE
from helpers import INPUTS, read_fasta

AFFINE = dict(gap_open1=4, gap_ext1=2, gap_open2=0, gap_ext2=0)
LINEAR = dict(gap_open1=0, gap_ext1=2, gap_open2=0, gap_ext2=0)
BLOSUM = synth.WORKLOADS["aa_blosum62_2k"].cfg.score_matrix

CASES = {
    # the reference's own inputs (config 1 of BASELINE.json and friends)
    "seq_affine": dict(cfg=dict(**AFFINE), file="seq.fa"),
    "seq_convex": dict(cfg=dict(), file="seq.fa"),
    "test_convex": dict(cfg=dict(), file="test.fa"),
    "heter_convex": dict(cfg=dict(), file="heter.fa"),
    "3alleles_affine": dict(cfg=dict(**AFFINE), file="3alleles.fa"),
    # synthetic shapes (small versions of configs 2-5)
    "syn_affine_1k": dict(cfg=dict(**AFFINE), synth=(101, 20, 1000, 0.05, 5)),
    "syn_convex_2k": dict(cfg=dict(), synth=(102, 10, 2000, 0.05, 5)),
    "syn_convex_int32": dict(cfg=dict(match=20, mismatch=40, gap_open1=40, gap_ext1=20, gap_open2=240, gap_ext2=10), synth=(103, 8, 1500, 0.05, 5)),
    "syn_linear_banded": dict(cfg=dict(**LINEAR), synth=(104, 10, 800, 0.05, 5)),
    "syn_local_linear": dict(cfg=dict(align_mode=ABPOA_LOCAL_MODE, **LINEAR), synth=(105, 8, 600, 0.05, 5)),
    "syn_local_affine": dict(cfg=dict(align_mode=ABPOA_LOCAL_MODE, **AFFINE), synth=(106, 8, 600, 0.05, 5)),
    "syn_local_convex": dict(cfg=dict(align_mode=ABPOA_LOCAL_MODE), synth=(107, 8, 600, 0.05, 5)),
    "syn_extend_affine": dict(cfg=dict(align_mode=ABPOA_EXTEND_MODE, **AFFINE), synth=(108, 8, 700, 0.05, 5)),
    "syn_extend_convex_zdrop": dict(cfg=dict(align_mode=ABPOA_EXTEND_MODE, zdrop=100), synth=(109, 8, 700, 0.05, 5)),
    "syn_aa_blosum62": dict(cfg=dict(m=27, score_matrix=BLOSUM, **AFFINE), synth=(110, 10, 1000, 0.10, 27)),
    "syn_unbanded_affine": dict(cfg=dict(wb=-1, **AFFINE), synth=(111, 6, 400, 0.08, 5)),
    "syn_path_score": dict(cfg=dict(inc_path_score=True), synth=(112, 8, 600, 0.08, 5)),
    "syn_gap_on_right": dict(cfg=dict(put_gap_on_right=True), synth=(113, 8, 600, 0.08, 5)),
    "syn_gap_at_end": dict(cfg=dict(put_gap_at_end=True), synth=(114, 8, 600, 0.08, 5)),
    "syn_ragged": dict(cfg=dict(), synth_ragged=(115, [900, 40, 1200, 7, 600, 1, 1000])),
    "syn_high_error": dict(cfg=dict(), synth=(116, 8, 800, 0.25, 5)),
    # -Q: per-base quality weights become edge weights (predecessor order, consensus); the DP itself is unchanged
    "syn_qv_weights": dict(cfg=dict(use_qv=True), synth=(117, 8, 600, 0.08, 5), weights=117),
}


def case_weights(case, reads):
    """Deterministic quality-like weights (1..40 per base) for cases that ask for them, else None."""
    if "weights" not in case:
        return None
    import numpy as np
    rng = np.random.default_rng(case["weights"])
    return [rng.integers(1, 41, size=len(r)).astype(np.int32) for r in reads]


def case_reads(case):
    if "file" in case:
        return read_fasta(INPUTS / case["file"], case["cfg"].get("m", 5))
    if "synth_ragged" in case:
        seed, lens = case["synth_ragged"]
        import numpy as np
        base = synth.make_group(seed, len(lens), max(lens), 0.05)
        return [np.ascontiguousarray(r[:n]) for r, n in zip(base, lens)]
    seed, n, length, err, m = case["synth"]
    return synth.make_group(seed, n, length, err, m)
