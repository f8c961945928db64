#!/bin/bash
# round 2, job C: full GPU test suite + bench lines of all five configurations on one GPU
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/c_test_all.log 2>&1; echo "rc=$?" >> gpurun_out/c_test_all.log
tail -15 gpurun_out/c_test_all.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c_bench_c2.json 2> gpurun_out/c_bench_c2.err; echo "rc=$?"; tail -3 gpurun_out/c_bench_c2.err
T2FIT_REF_BUDGET_S=40 timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/c_bench_c2_ref.json 2>&1
timeout 600 python bench.py --config c1 --steps 20 --warmup 3 > gpurun_out/c_bench_c1.json 2> gpurun_out/c_bench_c1.err; echo "rc=$?"; tail -3 gpurun_out/c_bench_c1.err
timeout 900 python bench.py --config c3 --steps 5 --warmup 3 > gpurun_out/c_bench_c3.json 2> gpurun_out/c_bench_c3.err; echo "rc=$?"; tail -3 gpurun_out/c_bench_c3.err
timeout 900 python bench.py --config c4 --steps 5 --warmup 3 > gpurun_out/c_bench_c4.json 2> gpurun_out/c_bench_c4.err; echo "rc=$?"; tail -3 gpurun_out/c_bench_c4.err
T2FIT_BENCH_SCALE=0.5 timeout 900 python bench.py --config c5 --steps 3 --warmup 1 > gpurun_out/c_bench_c5_lbfgsb_half.json 2> gpurun_out/c_bench_c5.err; echo "rc=$?"; tail -3 gpurun_out/c_bench_c5.err
timeout 900 python bench.py --config c5 --solver fast --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/c_bench_c5_fast.json 2> gpurun_out/c_bench_c5_fast.err; echo "rc=$?"; tail -3 gpurun_out/c_bench_c5_fast.err
for f in gpurun_out/c_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus")}, (d.get("config") or {}).get("passes_per_step"), (d.get("roofline") or {}).get("frac"), (d.get("e2e") or {}).get("value"))
except Exception as e: print("bad", e)
PY
done
