#!/bin/bash
# round-2 closing run on one GPU: full GPU suite, smoke, one bench line per config, launch list, ncu capture of the step kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
show() { python - "$@" <<'P'
import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); c=d["config"]
        print(f.split("/")[-1], "value %.4g ms/step %.2f ms/pass %s ms/vol %s e2e %.4g frac %.3f launches %s" % (d["value"], d["ms_per_step"], c.get("ms_per_pass"), c.get("ms_per_volume"), d["e2e"]["value"], d["roofline"]["frac"], d.get("gpu_launches")))
        if d.get("parity"): print("   parity", json.dumps(d["parity"])[:700])
        if d.get("cpu_baseline"): print("   cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
    except Exception as e: print(f, "bad", e)
P
}
{
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -3
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
echo "== bench lines"
timeout 900 python bench.py > gpurun_out/p_bench_c2.json 2> gpurun_out/p_bench_c2.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/p_bench_c2_reference.json 2> gpurun_out/p_bench_c2_reference.err
timeout 900 python bench.py --config c1 > gpurun_out/p_bench_c1.json 2> gpurun_out/p_bench_c1.err
timeout 900 python bench.py --config c3 --steps 2 --warmup 3 > gpurun_out/p_bench_c3.json 2> gpurun_out/p_bench_c3.err
timeout 900 python bench.py --config c4 --steps 3 --warmup 3 > gpurun_out/p_bench_c4.json 2> gpurun_out/p_bench_c4.err
T2FIT_BENCH_SCALE=0.5 timeout 900 python bench.py --config c5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/p_bench_c5_lbfgsb_half.json 2> gpurun_out/p_bench_c5h.err
timeout 900 python bench.py --config c5 --solver fast --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/p_bench_c5_fast.json 2> gpurun_out/p_bench_c5f.err
show gpurun_out/p_bench_c2.json gpurun_out/p_bench_c1.json gpurun_out/p_bench_c3.json gpurun_out/p_bench_c4.json gpurun_out/p_bench_c5_lbfgsb_half.json gpurun_out/p_bench_c5_fast.json
tail -c 400 gpurun_out/p_bench_c2_reference.json
echo "== launch list of the default bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/p_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/p_ncu_bench.log 2>&1
tail -2 gpurun_out/p_launches.csv | cut -c1-200
echo "== ncu --set full, step kernel"
MB_ONLY=step timeout 900 ncu --set full --clock-control none --import-source on -k regex:fit_kernel -s 2 -c 1 -f -o /tmp/p_step python tools/microbench.py > gpurun_out/p_ncu_step.log 2>&1
ncu -i /tmp/p_step.ncu-rep --page raw --csv > gpurun_out/p_step_raw.csv
ncu -i /tmp/p_step.ncu-rep --page source --csv > /tmp/p_step_source.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/p_step_raw.csv > gpurun_out/p_step_summary.txt; head -12 gpurun_out/p_step_summary.txt
python tools/ncu_source_hist.py /tmp/p_step_source.csv > gpurun_out/p_step_source_hist.txt 2>&1
echo "== thread L-BFGS-B kernel ncu summary (c3 slab)"
timeout 600 python tools/lb_bench.py c3 --kernels thread 2>&1 | grep -v "^$"
} 2>&1 | tee gpurun_out/p_job.log
