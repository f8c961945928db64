#!/bin/bash
# round 2, job F (N GPUs): multi-GPU tests (N = 2 only) + the c2 / c3 sharded bench lines with the fused all-gather
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/f_test_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/f_test_n$N.log
  tail -12 gpurun_out/f_test_n$N.log | cut -c1-400
fi
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N "$@" > gpurun_out/f_bench_${name}_n$N.json 2> gpurun_out/f_bench_${name}_n$N.err; echo "rc=$?"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/f_bench_${name}_n$N.err | tail -6 | cut -c1-400; }
run c2 --steps 20 --warmup 5
run c3 --config c3 --steps 2 --warmup 3 --no-secondary
for f in gpurun_out/f_bench_*_n$N.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus")}, "ms/pass", d["config"].get("ms_per_pass"))
    for k in ("sharded","sharded_nccl"):
        if d.get(k): print(" ",k, {a:b for a,b in d[k].items() if a not in ("op","note","what")})
    if d.get("replicas"): print("  replicas", d["replicas"]["value"], d["replicas"]["ms_per_pass"])
    if d.get("e2e"): print("  e2e", d["e2e"].get("value"), (d["e2e"].get("strong_single_volume") or {}).get("value"))
except Exception as e: print("bad", e)
PY
done
