#!/bin/bash
# programmatic dependent launch (T2FIT_PDL) x bulk asynchronous zero-fill (T2_FILL_BULK variant library): tests, c2 bench A/B, one ncu capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B=$PWD/fetal_t2mapping_b200/csrc/libt2fit_bulk.so
M=$PWD/fetal_t2mapping_b200/csrc/libt2fit.so
line() { python - "$1" <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print("   value %.4g  ms/pass %.5f  e2e %.4g  frac %.3f" % (d["value"], d["config"]["ms_per_pass"], d["e2e"]["value"], d["roofline"]["frac"]))
P
}
{
for lib in $M $B; do for pdl in 1 0; do
echo "== fused-fill / volume tests, $(basename $lib) PDL=$pdl"
T2FIT_PDL=$pdl T2FIT_LIB=$lib timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi4.py -q -x -m gpu 2>&1 | tail -3
done; done
for rep in 1 2; do for lib in $M $B; do for pdl in 0 1; do
n=$(basename $lib .so)_pdl${pdl}_$rep
echo "== bench c2 $n"
T2FIT_PDL=$pdl T2FIT_LIB=$lib timeout 600 python bench.py --config c2 --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/l_$n.json 2> gpurun_out/l_$n.err; line gpurun_out/l_$n.json
done; done; done
echo "== ncu, bulk variant, step kernel"
T2FIT_LIB=$B MB_ONLY=step timeout 600 python tools/microbench.py 2>&1 | tail -4
T2FIT_LIB=$B MB_ONLY=step timeout 900 ncu --set full --clock-control none --import-source on -k regex:fit_kernel -s 6 -c 1 -f -o /tmp/l_step python tools/microbench.py > gpurun_out/l_ncu_step.log 2>&1
ncu -i /tmp/l_step.ncu-rep --page raw --csv > gpurun_out/l_step_bulk_raw.csv
python tools/ncu_summary.py gpurun_out/l_step_bulk_raw.csv > gpurun_out/l_step_bulk_summary.txt; cat gpurun_out/l_step_bulk_summary.txt | head -60
} 2>&1 | tee gpurun_out/l_pdl_bulk.log
