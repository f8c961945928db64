#!/bin/bash
# re-entry run: dense L-BFGS-B kernel — GPU parity tests, A/B vs compact, bench lines c3 / c5, ncu --set full + source histogram
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
show() { python - "$@" <<'P'
import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); c=d["config"]
        print(f.split("/")[-1], "value %.4g ms/step %.2f ms/pass %s ms/vol %s e2e %.4g frac %.3g launches %s" % (d["value"], d["ms_per_step"], c.get("ms_per_pass"), c.get("ms_per_volume"), d["e2e"]["value"], d["roofline"]["frac"], d.get("gpu_launches")))
        if d.get("parity"): print("   parity", json.dumps(d["parity"])[:900])
    except Exception as e: print(f, "bad", e)
P
}
{
echo "== pytest dense + parity"
timeout 900 python -m pytest tests/test_gpu_lbfgsb_dense.py tests/test_gpu_parity.py -q -x 2>&1 | tail -3
echo "== lb_bench thread vs dense"
timeout 600 python tools/lb_bench.py c2 c3 c3r c5 --kernels thread,dense 2>&1 | grep -v "^$"
echo "== bench c3 / c5"
timeout 900 python bench.py --config c3 --steps 3 --warmup 3 > gpurun_out/r_bench_c3.json 2> gpurun_out/r_bench_c3.err
T2FIT_BENCH_SCALE=0.5 timeout 900 python bench.py --config c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r_bench_c5_half.json 2> gpurun_out/r_bench_c5h.err
show gpurun_out/r_bench_c3.json gpurun_out/r_bench_c5_half.json
echo "== ncu --set full, dense kernel, c3 x 0.25"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lbfgsb_dense_kernel -s 1 -c 1 -f -o /tmp/r_dense python tools/lb_profile.py c3 0.25 gaussian_rician lbfgsb_dense > gpurun_out/r_ncu_dense.log 2>&1
ncu -i /tmp/r_dense.ncu-rep --page raw --csv > gpurun_out/r_dense_raw.csv
ncu -i /tmp/r_dense.ncu-rep --page source --csv > gpurun_out/r_dense_source.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r_dense_raw.csv > gpurun_out/r_dense_summary.txt; head -70 gpurun_out/r_dense_summary.txt
python tools/ncu_source_hist.py gpurun_out/r_dense_source.csv > gpurun_out/r_dense_source_hist.txt 2>&1; head -60 gpurun_out/r_dense_source_hist.txt
} 2>&1 | tee gpurun_out/r_job.log
