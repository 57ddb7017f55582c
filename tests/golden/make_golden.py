#!/usr/bin/env python
"""Generate tests/golden/golden.json from the UNMODIFIED reference (oracle/_ref/libabpoa_ref.so).

Run in the build container (where /root/reference exists and `make -C oracle ref` works):
    python tests/golden/make_golden.py
For every case the reference is stepped read by read and we record, per alignment, the best
score, the number of DP cells, the CIGAR length and the sha1 of the raw 64-bit graph-CIGAR
words, plus the final consensus, its coverage, and the sha1 of every RC-MSA row.
The CLI md5 vectors come from oracle/_ref/abpoa_ref (they equal the values recorded in SURVEY.md 8c).
"""
import hashlib
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from abpoa_b200 import capi, synth  # noqa: E402
from abpoa_b200.aligner import PoaConfig, decode  # noqa: E402
from cases import CASES, case_reads, case_weights  # noqa: E402
from helpers import run_group  # noqa: E402


def sha(a) -> str:
    return hashlib.sha1(a.tobytes()).hexdigest()


def main():
    ref = capi.load_library(Path(__file__).resolve().parents[2] / "oracle" / "_ref" / "libabpoa_ref.so")
    out = {"reference": "abPOA v1.5.6 (make avx2=1 flags), built by oracle/Makefile", "cases": {}, "cli_md5": {}}
    for name, case in CASES.items():
        cfg = PoaConfig(**case["cfg"])
        reads = case_reads(case)
        r = run_group(ref, cfg, reads, want_msa=True, weights=case_weights(case, reads))
        out["cases"][name] = {
            "alns": [
                {"aligned": a.aligned, "score": a.best_score, "cells": a.cells, "n_cigar": int(len(a.cigar)), "cigar_sha1": sha(a.cigar),
                 "ends": [a.node_s, a.node_e, a.query_s, a.query_e]} for a in r["alns"]],
            "cons": [decode(c, cfg.m) for c in r["cons"]],
            "cov_sha1": [sha(c) for c in r["cov"]],
            "msa_sha1": [sha(m) for m in r["msa"]],
            "msa_len": int(len(r["msa"][0])) if r["msa"] else 0,
        }
        print(name, "ok", len(reads), "reads")
    cli = ROOT / "oracle" / "_ref" / "abpoa_ref"
    inputs = ROOT / "tests" / "golden" / "inputs"
    for tag, args in {
        "seq.fa -O 4 -E 2": ["-O", "4", "-E", "2", str(inputs / "seq.fa")],
        "seq.fa -O 4 -E 2 -r1": ["-O", "4", "-E", "2", "-r1", str(inputs / "seq.fa")],
        "seq.fa -O 4 -E 2 -r2": ["-O", "4", "-E", "2", "-r2", str(inputs / "seq.fa")],
        "seq.fa": [str(inputs / "seq.fa")],
        "test.fa": [str(inputs / "test.fa")],
        "heter.fa -r2": ["-r2", str(inputs / "heter.fa")],
    }.items():
        p = subprocess.run([str(cli)] + args, capture_output=True, check=True)
        out["cli_md5"][tag] = hashlib.md5(p.stdout).hexdigest()
    (ROOT / "tests" / "golden" / "golden.json").write_text(json.dumps(out, indent=1))
    print("written")


if __name__ == "__main__":
    main()
