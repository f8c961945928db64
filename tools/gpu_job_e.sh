#!/bin/bash
# round 2, job E (N GPUs): multi-GPU tests + gather-inclusive bench lines
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/e_test_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/e_test_n$N.log
  tail -6 gpurun_out/e_test_n$N.log | cut -c1-300
fi
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N "$@" > gpurun_out/e_bench_${name}_n$N.json 2> gpurun_out/e_bench_${name}_n$N.err; echo "rc=$?"; tail -4 gpurun_out/e_bench_${name}_n$N.err | cut -c1-300; }
run c2 --steps 20 --warmup 5
run c5fast --config c5 --solver fast --steps 5 --warmup 2
if [ "$N" = "8" ]; then run c5lbfgsb --config c5 --steps 2 --warmup 1; else T2FIT_BENCH_SCALE=0.5 run c5lbfgsb_half --config c5 --steps 2 --warmup 1; fi
run c4 --config c4 --steps 3 --warmup 3
run c3 --config c3 --steps 2 --warmup 3 --no-secondary
for f in gpurun_out/e_bench_*_n$N.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus")}, "ms/pass", d["config"].get("ms_per_pass"), d["config"].get("ms_per_volume"))
    if d.get("sharded"): print("  sharded", {k:v for k,v in d["sharded"].items() if k not in ("op","note","voxels_per_rank")})
    if d.get("replicas"): print("  replicas", d["replicas"]["value"], d["replicas"]["ms_per_pass"])
    if d.get("e2e"): print("  e2e", d["e2e"].get("value"), (d["e2e"].get("strong_single_volume") or {}))
except Exception as e: print("bad", e)
PY
done
