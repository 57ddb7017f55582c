#!/bin/bash
# round-2 GPU pass D: parity suite + benches after the backtrace shortcut
set -u
O=gpurun_out; mkdir -p $O
export CUDA_MODULE_LOADING=EAGER
timeout 2700 python -m pytest tests -m gpu -q --timeout=1200 -p no:cacheprovider > $O/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02d_pytest.log
tail -12 $O/r02d_pytest.log
timeout 600 python tools/exp_batch.py convex_10k 1000 32 0 3 > $O/r02d_e2e_default.log 2>&1; tail -2 $O/r02d_e2e_default.log
ABPOA_GPU_PROFILE=1 timeout 900 python bench.py --workload convex_10k --steps 2 --warmup 3 > $O/r02d_bench_convex_10k.json 2> $O/r02d_bench_convex_10k.err; echo "bench convex rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02d_bench_convex_10k.json"))
print("convex: value %.1f e2e %.1f kernel_alone %.1f ms/step %.0f chain %s" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_alone_gcups"], d["ms_per_step"], d["chain"]))
PY
for wl in "affine_1k 0" "aa_blosum62_2k 0"; do
  set -- $wl
  timeout 900 python bench.py --workload $1 --groups $2 --steps 2 --warmup 3 --no-cpu-baseline > $O/r02d_bench_$1.json 2> $O/r02d_bench_$1.err; echo "bench $1 rc=$?"
  python - "$1" <<'PY'
import json,sys
d=json.load(open("gpurun_out/r02d_bench_%s.json" % sys.argv[1]))
print("%s: value %.1f e2e %.1f kernel_alone %.1f ms/step %.0f" % (sys.argv[1], d["value"], d["e2e"]["value"], d["roofline"]["kernel_alone_gcups"], d["ms_per_step"]))
PY
done
ls -la $O | grep r02d
