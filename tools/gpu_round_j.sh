#!/bin/bash
# round-2 GPU pass J: free-running chain with warp-uniform wait loops
set -u
O=gpurun_out; mkdir -p $O
export CUDA_MODULE_LOADING=EAGER ABPOA_GPU_CHAIN_WATCHDOG_S=8 ABPOA_GPU_PROFILE=1
K=$PWD/abpoa_b200/lib/libabpoa_b200_kprof.so
run() { tag=$1; shift; env "$@" timeout 200 python tools/exp_batch.py $WL $NG 0 0 $REPS > $O/r02j_$tag.log 2>&1; echo "== $tag rc=$?"; grep -E "GCUPS|k-cycles|watchdog|free-running|wave of" $O/r02j_$tag.log | sed 's/.*GCUPS e2e, reads.s/reads.s/' | cut -c1-300 | tail -4; }
WL=convex_10k NG=64 REPS=1
run c64_kprof_free ABPOA_B200_LIB=$K ABPOA_GPU_CHAIN_ROUNDS=0
run c64_free ABPOA_GPU_CHAIN_ROUNDS=0
fwd=$(grep -oE "per-aln fwd [0-9]+k" $O/r02j_c64_free.log | tail -1 | grep -oE "[0-9]+")
echo "fwd k-clk per alignment: $fwd"
if [ "${fwd:-999999}" -gt 70000 ]; then echo "alignments inside the worker kernel are still slow: stop here"; exit 0; fi
WL=convex_10k NG=1000 REPS=3
run c1000_free ABPOA_GPU_CHAIN_ROUNDS=0
if grep -q watchdog $O/r02j_c1000_free.log; then echo "free-running still stalls: stop here"; exit 0; fi
run c1000_rounds ABPOA_GPU_CHAIN_ROUNDS=1
WL=affine_1k NG=1000 REPS=3
run a1k_free ABPOA_GPU_CHAIN_ROUNDS=0
run a1k_rounds ABPOA_GPU_CHAIN_ROUNDS=1
ABPOA_GPU_CHAIN_WATCHDOG_S=10 ABPOA_GPU_PROFILE= timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_fullshape.py -q -m gpu -p no:cacheprovider 2>&1 | tail -8 > $O/r02j_pytest_chain.log; echo "pytest chain rc=${PIPESTATUS[0]}"; tail -4 $O/r02j_pytest_chain.log
