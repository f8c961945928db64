#!/bin/bash
# closing run on one GPU: full GPU suite, smoke, one bench line per config (c3 / c5 on the dense L-BFGS-B kernel, c5 at full
# size), reference arm, launch list of the default bench command, ncu --set full of the dense kernel on a c3 half volume
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
show() { python - "$@" <<'P'
import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); c=d["config"]
        print(f.split("/")[-1], "value %.4g ms/step %.2f ms/pass %s ms/vol %s e2e %.4g frac %.3g launches %s" % (d["value"], d["ms_per_step"], c.get("ms_per_pass"), c.get("ms_per_volume"), (d.get("e2e") or {}).get("value", float("nan")), d["roofline"]["frac"], d.get("gpu_launches")))
        if d.get("parity"): print("   parity", json.dumps(d["parity"])[:1000])
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
timeout 900 python bench.py > gpurun_out/w_bench_c2.json 2> gpurun_out/w_bench_c2.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/w_bench_c2_reference.json 2> gpurun_out/w_bench_c2_reference.err
timeout 900 python bench.py --config c1 > gpurun_out/w_bench_c1.json 2> gpurun_out/w_bench_c1.err
timeout 900 python bench.py --config c3 --steps 5 --warmup 3 > gpurun_out/w_bench_c3.json 2> gpurun_out/w_bench_c3.err
timeout 900 python bench.py --config c4 --steps 3 --warmup 3 > gpurun_out/w_bench_c4.json 2> gpurun_out/w_bench_c4.err
timeout 900 python bench.py --config c5 --steps 2 --warmup 3 > gpurun_out/w_bench_c5.json 2> gpurun_out/w_bench_c5.err
timeout 900 python bench.py --config c5 --solver fast --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/w_bench_c5_fast.json 2> gpurun_out/w_bench_c5f.err
show gpurun_out/w_bench_c2.json gpurun_out/w_bench_c1.json gpurun_out/w_bench_c3.json gpurun_out/w_bench_c4.json gpurun_out/w_bench_c5.json gpurun_out/w_bench_c5_fast.json
tail -c 400 gpurun_out/w_bench_c2_reference.json
echo "== launch list of the default bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/w_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/w_ncu_bench.log 2>&1
tail -2 gpurun_out/w_launches.csv | cut -c1-200
echo "== launch list of bench --config c3"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/w_launches_c3.csv python bench.py --config c3 --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/w_ncu_bench_c3.log 2>&1
grep -c lbfgsb_dense gpurun_out/w_launches_c3.csv; grep lbfgsb_dense gpurun_out/w_launches_c3.csv | tail -2 | cut -c1-260
echo "== ncu --set full, dense kernel, c3 x 0.5"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lbfgsb_dense_kernel -s 1 -c 1 -f -o /tmp/w_dense python tools/lb_profile.py c3 0.5 gaussian_rician lbfgsb_dense > gpurun_out/w_ncu_dense.log 2>&1
ncu -i /tmp/w_dense.ncu-rep --page raw --csv > gpurun_out/w_dense_raw.csv
ncu -i /tmp/w_dense.ncu-rep --page source --csv > gpurun_out/w_dense_source.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/w_dense_raw.csv > gpurun_out/w_dense_summary.txt; grep -E "time_duration|dram__bytes|issue_active|pipe_fp64|thread_inst_executed_per" gpurun_out/w_dense_summary.txt
} 2>&1 | tee gpurun_out/w_job.log
