#!/bin/bash
# round-2 GPU pass N: backtrace with speculative cell prefetch (PoaBtRec carries the plane offset): split, speed, parity
set -u
O=gpurun_out; mkdir -p $O
export CUDA_MODULE_LOADING=EAGER ABPOA_GPU_CHAIN_WATCHDOG_S=8 ABPOA_GPU_PROFILE=1
K=$PWD/abpoa_b200/lib/libabpoa_b200_kprof.so
run() { tag=$1; shift; env "$@" timeout 200 python tools/exp_batch.py $WL $NG 0 0 $REPS > $O/r02n_$tag.log 2>&1; echo "== $tag rc=$?"; grep -E "GCUPS|k-cycles|backtrace per|watchdog|free-running|wave of" $O/r02n_$tag.log | sed 's/.*GCUPS e2e, reads.s/reads.s/' | cut -c1-300 | tail -5; }
WL=convex_10k NG=64 REPS=1
run c64_kprof ABPOA_B200_LIB=$K
WL=convex_10k NG=1000 REPS=2
run c1000_free
WL=affine_10k
run a10k_free
ABPOA_GPU_PROFILE= timeout 1200 python -m pytest tests/test_gpu_chain.py tests/test_gpu_cases.py tests/test_gpu_fullshape.py -q -m gpu -p no:cacheprovider 2>&1 | tail -8 > $O/r02n_pytest.log; echo "pytest rc=${PIPESTATUS[0]}"; tail -4 $O/r02n_pytest.log
