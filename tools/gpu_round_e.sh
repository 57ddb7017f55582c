#!/bin/bash
# round-2 GPU pass E: N-GPU weak scaling of bench.py (launched the way the driver launches it)
set -u
N=${1:-8}
O=gpurun_out; mkdir -p $O
export CUDA_MODULE_LOADING=EAGER
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 2 --warmup 3 > $O/r02e_bench_n$N.json 2> $O/r02e_bench_n$N.err; echo "bench N=$N rc=$?"
python - "$N" <<'PY'
import json,sys
n=sys.argv[1]
txt=[l for l in open(f"gpurun_out/r02e_bench_n{n}.json") if l.startswith("{")]
d=json.loads(txt[-1])
print("N=%s: value %.1f e2e %.1f (per GPU %.1f) ms/step %.0f dist_check %s chain %s" % (n, d["value"], d["e2e"]["value"], d["e2e"]["per_gpu"], d["ms_per_step"], d.get("distributed_check"), d.get("chain")))
PY
tail -5 $O/r02e_bench_n$N.err
