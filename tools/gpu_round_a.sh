#!/bin/bash
# round-2 GPU pass A: full GPU test suite, the four BASELINE workloads + affine 10 kbp, ncu of the affine kernel
set -u
O=gpurun_out; mkdir -p $O
export CUDA_MODULE_LOADING=EAGER
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $O/r02a_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > $O/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02a_pytest.log
tail -5 $O/r02a_pytest.log
for wl in "convex_10k 0" "affine_1k 0" "aa_blosum62_2k 0" "local_linear_5k 100" "affine_10k 500"; do
  set -- $wl
  ABPOA_GPU_PROFILE=1 timeout 900 python bench.py --workload $1 --groups $2 --steps 2 --warmup 3 > $O/r02a_bench_$1.json 2> $O/r02a_bench_$1.err; echo "bench $1 rc=$?"
  tail -c 600 $O/r02a_bench_$1.json
done
# ncu: launch list of the affine bench (short), full capture of one replay launch of the affine kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02a_launches_affine_10k.csv \
   python bench.py --workload affine_10k --groups 200 --steps 1 --warmup 3 --no-cpu-baseline > $O/r02a_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none --nvtx --nvtx-include "replay/" -c 1 -o $O/r02a_affine_replay_full -f \
   python tools/exp_replay.py affine_10k 60 8 8 > $O/r02a_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $O | tail -20
