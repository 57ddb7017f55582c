#!/bin/bash
# round-2 GPU pass F: the record run -- all GPU tests, bench lines of every workload, reference arm, ncu launch list + full captures
set -u
O=gpurun_out; mkdir -p $O
export CUDA_MODULE_LOADING=EAGER
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem --format=csv > $O/r02f_gpu.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -12 > $O/r02f_pytest.log; echo "pytest rc=${PIPESTATUS[0]}"; tail -3 $O/r02f_pytest.log
ABPOA_GPU_PROFILE=1 timeout 900 python bench.py --steps 2 --warmup 3 > $O/r02f_bench_convex_10k.json 2> $O/r02f_bench_convex_10k.err; echo "bench convex rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 3 > $O/r02f_bench_reference_convex_10k.json 2> $O/r02f_bench_reference.err; echo "reference arm rc=$?"
for wl in "affine_1k 0" "aa_blosum62_2k 0" "affine_10k 1000" "local_linear_5k 100"; do
  set -- $wl
  ABPOA_GPU_PROFILE=1 timeout 1200 python bench.py --workload $1 --groups $2 --steps 2 --warmup 3 > $O/r02f_bench_$1.json 2> $O/r02f_bench_$1.err; echo "bench $1 rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02f_bench_*.json")):
    try: d=json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as e: print(f, "unreadable", e); continue
    print(f.split("r02f_bench_")[1], "value %.2f e2e %.2f ms/step %.0f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), "kernel_alone", (d.get("roofline") or {}).get("kernel_alone_gcups"), "frac", (d.get("roofline") or {}).get("frac"), "parity", (d.get("parity_sample") or {}).get("consensus_identical"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
PY
# where the cycles go (profiling build), free-running schedule, 64 groups
ABPOA_B200_LIB=$PWD/abpoa_b200/lib/libabpoa_b200_kprof.so ABPOA_GPU_PROFILE=1 timeout 200 python tools/exp_batch.py convex_10k 64 0 0 1 > $O/r02f_kprof_convex_64.log 2>&1; grep -E "k-cycles|backtrace per" $O/r02f_kprof_convex_64.log
# profilers serialise kernels: the round schedule is what they see (same job function)
export ABPOA_GPU_CHAIN_ROUNDS=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r02f_launches_chain_convex.csv \
   python tools/exp_batch.py convex_10k 296 8 0 1 > $O/r02f_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:poa_chain_align_kernel_p16 -s 30 -c 1 -o $O/r02f_chain_align_convex_full -f \
   env ABPOA_GPU_CHAIN_COHORTS=1 python tools/exp_batch.py convex_10k 1000 8 0 1 > $O/r02f_ncu_full_convex.log 2>&1; echo "ncu full convex rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:poa_chain_align_kernel_p16 -s 30 -c 1 -o $O/r02f_chain_align_affine_full -f \
   env ABPOA_GPU_CHAIN_COHORTS=1 python tools/exp_batch.py affine_10k 1000 8 0 1 > $O/r02f_ncu_full_affine.log 2>&1; echo "ncu full affine rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:poa_chain_fuse_kernel -s 30 -c 1 -o $O/r02f_chain_fuse_full -f \
   env ABPOA_GPU_CHAIN_COHORTS=1 python tools/exp_batch.py convex_10k 1000 8 0 1 > $O/r02f_ncu_full_fuse.log 2>&1; echo "ncu full fuse rc=$?"
ls -la $O | grep r02f
