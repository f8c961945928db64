#!/bin/bash
# bulk asynchronous zero-fill (T2_FILL_BULK variant library) against the shipped fused fill, plus the thread L-BFGS-B kernel at 8 blocks/SM
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B=$PWD/fetal_t2mapping_b200/csrc/libt2fit_bulk.so
{
echo "== fused-fill tests, bulk variant"
T2FIT_LIB=$B timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "fused_zero_fill or dense or volume" 2>&1 | tail -5
for rep in 1 2; do
echo "== bench c2, shipped library (rep $rep)"
timeout 600 python bench.py --config c2 --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/k_c2_main_$rep.json 2> gpurun_out/k_c2_main_$rep.err; python - <<P
import json; d=json.loads(open("gpurun_out/k_c2_main_$rep.json").read().strip().splitlines()[-1]); print(d["value"], d["config"]["ms_per_pass"], d["e2e"]["value"], d["roofline"]["frac"])
P
echo "== bench c2, bulk variant (rep $rep)"
T2FIT_LIB=$B timeout 600 python bench.py --config c2 --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/k_c2_bulk_$rep.json 2> gpurun_out/k_c2_bulk_$rep.err; python - <<P
import json; d=json.loads(open("gpurun_out/k_c2_bulk_$rep.json").read().strip().splitlines()[-1]); print(d["value"], d["config"]["ms_per_pass"], d["e2e"]["value"], d["roofline"]["frac"])
P
done
echo "== thread L-BFGS-B kernel, shipped library (8 blocks/SM)"
timeout 600 python tools/lb_bench.py c2 c3 c5 --kernels thread 2>&1 | grep -v "^$"
} 2>&1 | tee gpurun_out/k_bulk_fill.log
