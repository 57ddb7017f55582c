#!/bin/bash
# round-2 GPU pass I: why is the alignment 4x slower inside the persistent worker kernel?  64 groups, one variable at a time
set -u
O=gpurun_out; mkdir -p $O
export CUDA_MODULE_LOADING=EAGER ABPOA_GPU_CHAIN_WATCHDOG_S=8 ABPOA_GPU_PROFILE=1
K=$PWD/abpoa_b200/lib/libabpoa_b200_kprof.so
run() { tag=$1; shift; env "$@" timeout 120 python tools/exp_batch.py convex_10k 64 0 0 1 > $O/r02i_$tag.log 2>&1; echo "== $tag rc=$?"; grep -E "GCUPS|k-cycles|watchdog" $O/r02i_$tag.log | sed 's/.*GCUPS e2e, reads.s [0-9]*; //' | cut -c1-260 | tail -3; }
run kprof_rounds ABPOA_B200_LIB=$K ABPOA_GPU_CHAIN_ROUNDS=1
run kprof_free ABPOA_B200_LIB=$K ABPOA_GPU_CHAIN_ROUNDS=0
run free_slab ABPOA_GPU_CHAIN_ROUNDS=0 ABPOA_GPU_CHAIN_SLAB_X=1.2
run free_nofence ABPOA_GPU_CHAIN_ROUNDS=0 ABPOA_GPU_CHAIN_DBG=1
run free_nopad ABPOA_GPU_CHAIN_ROUNDS=0 ABPOA_GPU_CHAIN_DBG=2
run free_carve30 ABPOA_GPU_CHAIN_ROUNDS=0 ABPOA_GPU_CHAIN_CARVEOUT=30
run free_dpfirst ABPOA_GPU_CHAIN_ROUNDS=0 ABPOA_GPU_CHAIN_DP_FIRST=1
run free_1worker ABPOA_GPU_CHAIN_ROUNDS=0 ABPOA_GPU_CHAIN_FUSE_WORKERS=1
