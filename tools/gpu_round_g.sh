#!/bin/bash
# round-2 GPU pass G: free-running chain (two persistent kernels) vs lock-step rounds
set -u
O=gpurun_out; mkdir -p $O
export CUDA_MODULE_LOADING=EAGER
ABPOA_GPU_CHAIN_WATCHDOG_S=5 ABPOA_GPU_PROFILE=1 timeout 120 python tools/exp_batch.py affine_1k 64 0 0 2 > $O/r02g_quick.log 2>&1; echo "quick rc=$?"; grep -E "GCUPS|free-running|watchdog|handed" $O/r02g_quick.log | tail -5
ABPOA_GPU_CHAIN_WATCHDOG_S=10 timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_fullshape.py -q -m gpu -p no:cacheprovider 2>&1 | tail -15 > $O/r02g_pytest_chain.log; echo "pytest chain rc=${PIPESTATUS[0]}"; tail -5 $O/r02g_pytest_chain.log
for mode in 0 1; do
  ABPOA_GPU_CHAIN_ROUNDS=$mode ABPOA_GPU_PROFILE=1 timeout 300 python tools/exp_batch.py convex_10k 1000 0 0 3 > $O/r02g_e2e_rounds$mode.log 2>&1; echo "exp convex rounds=$mode rc=$?"
  grep -E "GCUPS|free-running|wave of" $O/r02g_e2e_rounds$mode.log | tail -4
done
ABPOA_GPU_PROFILE=1 timeout 900 python bench.py --steps 2 --warmup 3 > $O/r02g_bench_convex_10k.json 2> $O/r02g_bench_convex_10k.err; echo "bench convex rc=$?"
ABPOA_GPU_PROFILE=1 timeout 900 python bench.py --workload affine_10k --groups 1000 --steps 2 --warmup 3 > $O/r02g_bench_affine_10k.json 2> $O/r02g_bench_affine_10k.err; echo "bench affine_10k rc=$?"
timeout 600 python bench.py --workload affine_1k --steps 2 --warmup 3 > $O/r02g_bench_affine_1k.json 2> $O/r02g_bench_affine_1k.err; echo "bench affine_1k rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02g_bench_*.json")):
    try: d=json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as e: print(f, "unreadable", e); continue
    print(f.split("r02g_bench_")[1], "value %.2f e2e %.2f ms/step %.0f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), "kernel_alone", (d.get("roofline") or {}).get("kernel_alone_gcups"), "parity", (d.get("parity_sample") or {}).get("consensus_identical"), "chain", d.get("chain"))
PY
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --deselect tests/test_gpu_chain.py --deselect tests/test_gpu_fullshape.py 2>&1 | tail -8 > $O/r02g_pytest_rest.log; echo "pytest rest rc=${PIPESTATUS[0]}"; tail -4 $O/r02g_pytest_rest.log
