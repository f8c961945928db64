#!/bin/bash
# after the bulk zero-fill + PDL + L-BFGS-B code-shape changes: full GPU suite, bench lines, launch list, one ncu capture of the step kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -4
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench c2 (default command)"
timeout 900 python bench.py > gpurun_out/m_bench_c2.json 2> gpurun_out/m_bench_c2.err; tail -c 600 gpurun_out/m_bench_c2.json | head -c 300; echo
python - <<'P'
import json
d=json.loads(open("gpurun_out/m_bench_c2.json").read().strip().splitlines()[-1])
print("c2 value %.4g ms/pass %.5f e2e %.4g frac %.3f launches %s timed %.3f" % (d["value"], d["config"]["ms_per_pass"], d["e2e"]["value"], d["roofline"]["frac"], d["gpu_launches"], d["config"]["timed_region_s"]))
print("cpu_baseline", d.get("cpu_baseline")); print("clocks", d.get("clocks"))
P
echo "== bench c3"
timeout 900 python bench.py --config c3 --steps 2 --warmup 3 > gpurun_out/m_bench_c3.json 2> gpurun_out/m_bench_c3.err
echo "== bench c5 (half)"
T2FIT_BENCH_SCALE=0.5 timeout 900 python bench.py --config c5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/m_bench_c5_lbfgsb_half.json 2> gpurun_out/m_bench_c5.err
python - <<'P'
import json
for n in ("c3","c5_lbfgsb_half"):
    try:
        d=json.loads(open(f"gpurun_out/m_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, "value %.4g ms/step %.1f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), "parity", json.dumps(d.get("parity"))[:400])
    except Exception as e: print(n, "bad", e)
P
echo "== launch list of the default bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/m_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/m_ncu_bench.log 2>&1
tail -3 gpurun_out/m_launches.csv | cut -c1-200
echo "== ncu --set full, step kernel (fit + bulk zero-fill)"
MB_ONLY=step timeout 900 ncu --set full --clock-control none --import-source on -k regex:fit_kernel -s 2 -c 1 -f -o /tmp/m_step python tools/microbench.py > gpurun_out/m_ncu_step.log 2>&1
ncu -i /tmp/m_step.ncu-rep --page raw --csv > gpurun_out/m_step_raw.csv
ncu -i /tmp/m_step.ncu-rep --page source --csv > /tmp/m_step_source.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/m_step_raw.csv > gpurun_out/m_step_summary.txt; head -70 gpurun_out/m_step_summary.txt
python tools/ncu_source_hist.py /tmp/m_step_source.csv > gpurun_out/m_step_source_hist.txt 2>&1; head -50 gpurun_out/m_step_source_hist.txt
} 2>&1 | tee gpurun_out/m_job.log
