#!/bin/bash
# round-2 GPU pass B: chain engine + LEAN kernel: parity suite, benches, phase profile, ncu
set -u
O=gpurun_out; mkdir -p $O
export CUDA_MODULE_LOADING=EAGER
timeout 2400 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider -x > $O/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02b_pytest.log
tail -15 $O/r02b_pytest.log
ABPOA_GPU_PROFILE=1 timeout 900 python bench.py --workload convex_10k --steps 2 --warmup 3 > $O/r02b_bench_convex_10k.json 2> $O/r02b_bench_convex_10k.err; echo "bench convex rc=$?"
tail -c 1500 $O/r02b_bench_convex_10k.json; grep "chain\]" $O/r02b_bench_convex_10k.err | tail -4
# kernel A/B: the same replay with the straight-line row path off
ABPOA_GPU_NO_LEAN=1 timeout 600 python tools/exp_replay.py convex_10k 200 32 0 > $O/r02b_replay_nolean.log 2>&1; tail -1 $O/r02b_replay_nolean.log
timeout 600 python tools/exp_replay.py convex_10k 200 32 0 > $O/r02b_replay_lean.log 2>&1; tail -1 $O/r02b_replay_lean.log
# where the cycles of a row go (-DPOA_KPROF build, launch engine)
ABPOA_B200_LIB=$PWD/abpoa_b200/lib/libabpoa_b200_kprof.so ABPOA_GPU_PROFILE=1 timeout 600 python tools/exp_batch.py convex_10k 256 32 0 1 > $O/r02b_kprof_lean.log 2>&1; grep "kernel, k-cycles" $O/r02b_kprof_lean.log | head -2
ABPOA_GPU_NO_LEAN=1 ABPOA_B200_LIB=$PWD/abpoa_b200/lib/libabpoa_b200_kprof.so ABPOA_GPU_PROFILE=1 timeout 600 python tools/exp_batch.py convex_10k 256 32 0 1 > $O/r02b_kprof_nolean.log 2>&1; grep "kernel, k-cycles" $O/r02b_kprof_nolean.log | head -2
for wl in "affine_1k 0" "aa_blosum62_2k 0" "affine_10k 500"; do
  set -- $wl
  ABPOA_GPU_PROFILE=1 timeout 900 python bench.py --workload $1 --groups $2 --steps 2 --warmup 3 > $O/r02b_bench_$1.json 2> $O/r02b_bench_$1.err; echo "bench $1 rc=$?"
  tail -c 400 $O/r02b_bench_$1.json
done
# ncu: launch list of a short chain run; full captures of the chain's two kernels (affine: the north_star roofline kernel; convex)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02b_launches_chain_convex.csv \
   python tools/exp_batch.py convex_10k 148 8 0 1 > $O/r02b_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:poa_chain_align_kernel_p16 -s 30 -c 1 -o $O/r02b_chain_align_convex_full -f \
   env ABPOA_GPU_CHAIN_COHORTS=1 python tools/exp_batch.py convex_10k 1000 8 0 1 > $O/r02b_ncu_full_convex.log 2>&1; echo "ncu full convex rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:poa_chain_align_kernel_p16 -s 30 -c 1 -o $O/r02b_chain_align_affine_full -f \
   env ABPOA_GPU_CHAIN_COHORTS=1 python tools/exp_batch.py affine_10k 1000 8 0 1 > $O/r02b_ncu_full_affine.log 2>&1; echo "ncu full affine rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:poa_chain_fuse_kernel -s 30 -c 1 -o $O/r02b_chain_fuse_full -f \
   env ABPOA_GPU_CHAIN_COHORTS=1 python tools/exp_batch.py convex_10k 1000 8 0 1 > $O/r02b_ncu_full_fuse.log 2>&1; echo "ncu full fuse rc=$?"
ls -la $O | grep r02b
