#!/bin/bash
# round-2 GPU pass C: full parity suite on the current tree, benches of all workloads, TMA A/B, ncu captures
set -u
O=gpurun_out; mkdir -p $O
export CUDA_MODULE_LOADING=EAGER
timeout 2700 python -m pytest tests -m gpu -q --timeout=1200 -p no:cacheprovider > $O/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02c_pytest.log
tail -25 $O/r02c_pytest.log
ABPOA_GPU_PROFILE=1 timeout 900 python bench.py --workload convex_10k --steps 2 --warmup 3 > $O/r02c_bench_convex_10k.json 2> $O/r02c_bench_convex_10k.err; echo "bench convex rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02c_bench_convex_10k.json"))
print("convex: value %.1f e2e %.1f kernel_alone %.1f ms/step %.0f chain %s" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_alone_gcups"], d["ms_per_step"], d["chain"]))
PY
grep "chain\]" $O/r02c_bench_convex_10k.err | tail -3
# bulk-store (TMA) variant against the default, same 1000-group e2e call
timeout 600 python tools/exp_batch.py convex_10k 1000 32 0 3 > $O/r02c_e2e_default.log 2>&1; tail -2 $O/r02c_e2e_default.log
ABPOA_GPU_TMA=1 timeout 600 python tools/exp_batch.py convex_10k 1000 32 0 3 > $O/r02c_e2e_tma.log 2>&1; tail -2 $O/r02c_e2e_tma.log
ABPOA_GPU_TMA=1 timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider > $O/r02c_pytest_tma.log 2>&1; tail -3 $O/r02c_pytest_tma.log
# where rows go and where the cycles of a row go (-DPOA_KPROF build, launch engine so that the worker prints the per-kernel phases)
ABPOA_GPU_NO_CHAIN=1 ABPOA_B200_LIB=$PWD/abpoa_b200/lib/libabpoa_b200_kprof.so ABPOA_GPU_PROFILE=1 timeout 600 python tools/exp_batch.py convex_10k 256 32 0 1 > $O/r02c_kprof.log 2>&1; grep "kernel rows\|kernel, k-cycles" $O/r02c_kprof.log | head -4
for wl in "affine_1k 0" "aa_blosum62_2k 0" "affine_10k 1000" "local_linear_5k 100"; do
  set -- $wl
  ABPOA_GPU_PROFILE=1 timeout 900 python bench.py --workload $1 --groups $2 --steps 2 --warmup 3 > $O/r02c_bench_$1.json 2> $O/r02c_bench_$1.err; echo "bench $1 rc=$?"
  python - "$1" <<'PY'
import json,sys
d=json.load(open("gpurun_out/r02c_bench_%s.json" % sys.argv[1]))
print("%s: value %.1f e2e %.1f kernel_alone %.1f frac %.3f ms/step %.0f ref %.3f parity %s" % (sys.argv[1], d["value"], d["e2e"]["value"], d["roofline"]["kernel_alone_gcups"], d["roofline"]["frac"], d["ms_per_step"], (d.get("cpu_baseline") or {}).get("value",0), d.get("parity_sample")))
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r02c_launches_chain_convex.csv \
   python tools/exp_batch.py convex_10k 296 8 0 1 > $O/r02c_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:poa_chain_align_kernel_p16 -s 30 -c 1 -o $O/r02c_chain_align_convex_full -f \
   env ABPOA_GPU_CHAIN_COHORTS=1 python tools/exp_batch.py convex_10k 1000 8 0 1 > $O/r02c_ncu_full_convex.log 2>&1; echo "ncu full convex rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:poa_chain_align_kernel_p16 -s 30 -c 1 -o $O/r02c_chain_align_affine_full -f \
   env ABPOA_GPU_CHAIN_COHORTS=1 python tools/exp_batch.py affine_10k 1000 8 0 1 > $O/r02c_ncu_full_affine.log 2>&1; echo "ncu full affine rc=$?"
ls -la $O | grep r02c
