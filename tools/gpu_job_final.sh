#!/bin/bash
# last check of the shipped library: full GPU suite, smoke, the default bench line and the reference arm as the driver runs them
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -2
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench (default)"; timeout 900 python bench.py > gpurun_out/final_bench_c2.json 2> gpurun_out/final_bench_c2.err; python - <<'P'
import json
d = json.loads(open("gpurun_out/final_bench_c2.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "gpu_launches", "dtype")}, d["e2e"]["value"], d["roofline"]["frac"], d["clocks"])
P
echo "== bench c3 / c5"; timeout 900 python bench.py --config c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench_c3.json 2> gpurun_out/final_bench_c3.err
timeout 900 python bench.py --config c5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench_c5.json 2> gpurun_out/final_bench_c5.err
python - <<'P'
import json
for c in ("c3", "c5"):
    d = json.loads(open(f"gpurun_out/final_bench_{c}.json").read().strip().splitlines()[-1])
    print(c, "value %.4g ms/step %.2f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
P
} 2>&1 | tee gpurun_out/final_job.log
