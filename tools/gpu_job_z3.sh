#!/bin/bash
# last single-GPU check of the tree as committed: default bench line + smoke
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== bench (default)"; timeout 200 python bench.py > gpurun_out/z3_bench_c2.json 2> gpurun_out/z3_bench_c2.err; python - <<'P'
import json
d = json.loads(open("gpurun_out/z3_bench_c2.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "gpu_launches", "dtype")}, d["e2e"]["value"], d["roofline"]["frac"], d["clocks"])
P
echo "== smoke"; timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
} 2>&1 | tee gpurun_out/z3_job.log
