#!/usr/bin/env python
"""Copy the artefacts of the record GPU pass (gpurun_out/r02f_*) into profiles/ (tracked) and print the README table."""
import json, shutil, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
G, P = ROOT / "gpurun_out", ROOT / "profiles"


def last_json(path):
    lines = [l for l in path.read_text().splitlines() if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


def main():
    rows = []
    for wl in ("convex_10k", "affine_10k", "affine_1k", "aa_blosum62_2k", "local_linear_5k"):
        src = G / f"r02f_bench_{wl}.json"
        if not src.exists():
            continue
        d = last_json(src)
        if d is None:
            continue
        (P / f"r02_bench_{wl}.json").write_text(json.dumps(d) + "\n")
        cb = d.get("cpu_baseline") or {}
        rf = d.get("roofline") or {}
        rows.append((wl, d["config"].get("workload", wl), d["value"], d["e2e"]["value"], d.get("reads_per_s"), rf.get("kernel_alone_gcups"), rf.get("frac"),
                     rf.get("dram_frac_measured"), cb.get("value"), (cb.get("single_socket") or {}).get("value"), (d.get("parity_sample") or {}).get("consensus_identical")))
    ref = G / "r02f_bench_reference_convex_10k.json"
    if ref.exists() and last_json(ref):
        (P / "r02_bench_reference_convex_10k.json").write_text(json.dumps(last_json(ref)) + "\n")
    for name, dst in (("r02f_pytest.log", "r02_pytest_gpu.log"), ("r02f_launches_chain_convex.csv", "r02_ncu_launches_chain_convex.csv"),
                      ("r02f_kprof_convex_64.log", "r02_kprof_convex_64groups.log"), ("r02f_gpu.txt", "r02_gpu.txt")):
        if (G / name).exists():
            if name.endswith(".log") and "kprof" in name:
                keep = [l for l in (G / name).read_text().splitlines() if l.startswith("[chain") or "GCUPS" in l]
                (P / dst).write_text("\n".join(keep) + "\n")
            else:
                shutil.copy(G / name, P / dst)
    for rep, txt, js, note in (("r02f_chain_align_convex_full", "r02_ncu_full_chain_align_convex.txt", "r02_ncu_traffic_convex.json",
                                "poa_chain_align_kernel_p16<CG>: one round of the round schedule, 1000 groups x 10 kbp (ABPOA_GPU_CHAIN_COHORTS=1), ncu --set full --clock-control none"),
                               ("r02f_chain_align_affine_full", "r02_ncu_full_chain_align_affine.txt", "r02_ncu_traffic_affine.json",
                                "poa_chain_align_kernel_p16<AG>: one round of the round schedule, 1000 groups x 10 kbp (ABPOA_GPU_CHAIN_COHORTS=1), ncu --set full --clock-control none"),
                               ("r02f_chain_fuse_full", "r02_ncu_full_chain_fuse.txt", None, "poa_chain_fuse_kernel: one round, 1000 groups x 10 kbp, ncu --set full --clock-control none")):
        if (G / f"{rep}.ncu-rep").exists():
            cmd = [sys.executable, str(ROOT / "tools" / "ncu_summary.py"), str(G / f"{rep}.ncu-rep"), note]
            if js:
                cmd += ["--json", str(P / js)]
            out = subprocess.run(cmd, capture_output=True, text=True).stdout
            (P / txt).write_text(out)
            if js and (P / js).exists():
                t = json.loads((P / js).read_text()); t["source"] = f"profiles/{txt}"; (P / js).write_text(json.dumps(t, indent=1) + "\n")
    print("| workload | value (device-resident) | e2e (host buffers) | reads/s e2e | DP kernel alone | roofline frac (algorithmic / measured DRAM) | reference, 64 cores (32 cores) | e2e ÷ reference | consensus sample |")
    print("|---|---|---|---|---|---|---|---|---|")
    for wl, desc, v, e, rps, ka, fr, dfr, cb, cb1, par in rows:
        f = lambda x, n=1: "-" if x is None else f"{x:.{n}f}"
        print(f"| `{wl}` | {f(v)} GCUPS | {f(e)} GCUPS | {f(rps, 0)} | {f(ka)} GCUPS | {f(fr, 3)} / {f(dfr, 3)} | {f(cb, 3)} ({f(cb1, 3)}) GCUPS | {f(e / cb if cb else None, 0)}x | {par}/64 |")


if __name__ == "__main__":
    main()
