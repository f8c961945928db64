#!/bin/bash
# round 2, job D (2 GPUs): two-GPU tests + gather-inclusive bench lines at N = 2
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_abi4.py "tests/test_gpu_parity.py::test_fused_zero_fill_ragged_volumes" "tests/test_gpu_parity.py::test_floor_model_reaches_a_bounded_minimum" -q -x > gpurun_out/d_test.log 2>&1; echo "rc=$?" >> gpurun_out/d_test.log
tail -8 gpurun_out/d_test.log
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N "$@" > gpurun_out/d_bench_${name}_n$N.json 2> gpurun_out/d_bench_${name}_n$N.err; echo "rc=$?"; tail -4 gpurun_out/d_bench_${name}_n$N.err | cut -c1-300; }
run c2 --steps 20 --warmup 5
run c5fast --config c5 --solver fast --steps 5 --warmup 2
T2FIT_BENCH_SCALE=0.5 run c5lbfgsb_half --config c5 --steps 2 --warmup 1
run c4 --config c4 --steps 3 --warmup 3
timeout 600 python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/d_bench_c4_n1.json 2> gpurun_out/d_bench_c4_n1.err; echo "rc=$?"
for f in gpurun_out/d_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus")}, (d.get("sharded") or {}), (d.get("replicas") or {}).get("value"))
except Exception as e: print("bad", e)
PY
done
