#!/bin/bash
# re-check of the library rebuilt from a clean checkout (same SASS as the committed evidence): default bench line, smoke, GPU suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== bench (default)"; timeout 300 python bench.py > gpurun_out/z_bench_c2.json 2> gpurun_out/z_bench_c2.err; python - <<'P'
import json
d = json.loads(open("gpurun_out/z_bench_c2.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "gpu_launches", "dtype")}, d["e2e"]["value"], d["roofline"]["frac"], d["clocks"])
P
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== pytest -m gpu"; timeout 420 python -m pytest tests -q -x -m gpu 2>&1 | tail -3
} 2>&1 | tee gpurun_out/z_job.log
