#!/bin/bash
# round 2, job G (N = 4 or 8 GPUs): gather-inclusive bench lines
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-4}
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N "$@" > gpurun_out/g_bench_${name}_n$N.json 2> gpurun_out/g_bench_${name}_n$N.err; echo "rc=$?"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/g_bench_${name}_n$N.err | tail -6 | cut -c1-400; }
run c2 --steps 20 --warmup 5
run c5fast --config c5 --solver fast --steps 5 --warmup 2
if [ "$N" = "8" ]; then
  run c5lbfgsb --config c5 --steps 2 --warmup 1
  run c4 --config c4 --steps 3 --warmup 3
else
  T2FIT_BENCH_SCALE=0.5 run c5lbfgsb_half --config c5 --steps 2 --warmup 1
fi
for f in gpurun_out/g_bench_*_n$N.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus")}, "ms/pass", d["config"].get("ms_per_pass"), d["config"].get("ms_per_volume"))
    for k in ("sharded","sharded_nccl"):
        if d.get(k): print(" ",k, {a:b for a,b in d[k].items() if a not in ("op","note","what")})
    if d.get("replicas"): print("  replicas", d["replicas"]["value"], d["replicas"]["ms_per_pass"])
    if d.get("e2e"): print("  e2e", d["e2e"].get("value"), (d["e2e"].get("strong_single_volume") or {}).get("value"))
except Exception as e: print("bad", e)
PY
done
