#!/bin/bash
# finer sweep of the L2 prefetch-ahead distance of the c2 step (blocks), plus c4 per-volume device part and c1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
line() { python - "$1" <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c=d["config"]
print("   value %.4g  ms/pass %s ms/volume %s e2e %.4g  frac %.3f" % (d["value"], c.get("ms_per_pass"), c.get("ms_per_volume"), d["e2e"]["value"], d["roofline"]["frac"]))
P
}
{
for a in 0 148 296 444 592 740 888 1036 1184 0 740; do
echo "== bench c2 ahead $a"
T2FIT_PREFETCH_AHEAD=$a timeout 600 python bench.py --config c2 --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/o_c2_a${a}.json 2> gpurun_out/o_c2_a${a}.err; line gpurun_out/o_c2_a${a}.json
done
for a in 0 740; do
echo "== bench c2 scale 0.7 ahead $a"
T2FIT_BENCH_SCALE=0.7 T2FIT_PREFETCH_AHEAD=$a timeout 600 python bench.py --config c2 --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/o_c2s_a${a}.json 2> gpurun_out/o_c2s_a${a}.err; line gpurun_out/o_c2s_a${a}.json
done
} 2>&1 | tee gpurun_out/o_job.log
