#!/bin/bash
# round-2 GPU pass K: free-running chain with matching L1/shared splits; A/B against the round schedule, tests, bench
set -u
O=gpurun_out; mkdir -p $O
export CUDA_MODULE_LOADING=EAGER ABPOA_GPU_CHAIN_WATCHDOG_S=8 ABPOA_GPU_PROFILE=1
run() { tag=$1; shift; env "$@" timeout 200 python tools/exp_batch.py $WL $NG 0 0 $REPS > $O/r02k_$tag.log 2>&1; echo "== $tag rc=$?"; grep -E "GCUPS|k-cycles|watchdog|free-running|wave of" $O/r02k_$tag.log | sed 's/.*GCUPS e2e, reads.s/reads.s/' | cut -c1-300 | tail -6; }
WL=convex_10k NG=1000 REPS=4
run c1000_free ABPOA_GPU_CHAIN_ROUNDS=0
if grep -q watchdog $O/r02k_c1000_free.log; then echo "free-running still stalls: stop here"; exit 0; fi
REPS=2
run c1000_rounds ABPOA_GPU_CHAIN_ROUNDS=1
WL=affine_10k
run a10k_free ABPOA_GPU_CHAIN_ROUNDS=0
WL=affine_1k NG=1000 REPS=3
run a1k_free ABPOA_GPU_CHAIN_ROUNDS=0
run a1k_rounds ABPOA_GPU_CHAIN_ROUNDS=1
WL=aa_blosum62_2k NG=200
run aa_free ABPOA_GPU_CHAIN_ROUNDS=0
run aa_rounds ABPOA_GPU_CHAIN_ROUNDS=1
ABPOA_GPU_CHAIN_WATCHDOG_S=10 ABPOA_GPU_PROFILE= timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_fullshape.py -q -m gpu -p no:cacheprovider 2>&1 | tail -8 > $O/r02k_pytest_chain.log; echo "pytest chain rc=${PIPESTATUS[0]}"; tail -4 $O/r02k_pytest_chain.log
unset ABPOA_GPU_CHAIN_WATCHDOG_S
timeout 900 python bench.py --steps 2 --warmup 3 > $O/r02k_bench_convex_10k.json 2> $O/r02k_bench_convex_10k.err; echo "bench convex rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02k_bench_convex_10k.json") if l.startswith("{")][-1])
print("convex: value %.2f e2e %.2f ms/step %.0f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), "kernel_alone", (d.get("roofline") or {}).get("kernel_alone_gcups"), "parity", (d.get("parity_sample") or {}).get("consensus_identical"), "chain", d.get("chain"))
PY
