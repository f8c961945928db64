#!/bin/bash
# round 2, job H (1 GPU): full GPU tests + smoke + bench c2 after the warp-cooperative row fetch; ncu evidence of the c2 step
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/h_test_all.log 2>&1; echo "rc=$?" >> gpurun_out/h_test_all.log
tail -8 gpurun_out/h_test_all.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/h_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/h_smoke.log; tail -3 gpurun_out/h_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/h_bench_c2.json 2> gpurun_out/h_bench_c2.err; echo "rc=$?"; tail -3 gpurun_out/h_bench_c2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/h_bench_c2.json').read().strip().splitlines()[-1])
print("value", d["value"], "ms/pass", d["config"]["ms_per_pass"], "roof", d["roofline"]["frac"], "fp32", d["roofline_fp32"]["kernel_ms"], d["roofline_fp32"]["frac"])
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "all", d["e2e"]["all_outputs_int64_indices"]["value"], "pageable", d["e2e"]["pageable_input"]["value"])
PY
T2FIT_BENCH_MIN_S=0.002 timeout 900 bash tools/profile_step.sh r02 > gpurun_out/h_profile_step.log 2>&1; tail -5 gpurun_out/h_profile_step.log
python tools/ncu_summary.py gpurun_out/r02_step_raw.csv > gpurun_out/r02_step_summary.txt 2>&1
rm -f gpurun_out/r02_step.ncu-rep
gzip -f gpurun_out/r02_step_source.csv
ls -la gpurun_out | grep r02
