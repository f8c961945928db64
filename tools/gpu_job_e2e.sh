#!/bin/bash
# same-box A/B of the c2 bench line (device step + e2e from host memory) between the shipped library and libt2fit_$v.so
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
line() { python - "$1" <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c=d["config"]
print("   value %.4g  us/pass %.2f  e2e %.4g (%.4f ms/call)  frac %.3f" % (d["value"], 1e3*c.get("ms_per_pass"), d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"]))
P
}
{
for rep in 1 2; do
echo "== shipped (rep $rep)"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/e2e_a$rep.json 2> gpurun_out/e2e_a$rep.err; line gpurun_out/e2e_a$rep.json
for v in $VARIANTS; do
echo "== $v (rep $rep)"; T2FIT_LIB=$PWD/fetal_t2mapping_b200/csrc/libt2fit_$v.so timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/e2e_${v}$rep.json 2> gpurun_out/e2e_${v}$rep.err; line gpurun_out/e2e_${v}$rep.json
done
done
} 2>&1 | tee gpurun_out/e2e_job.log
