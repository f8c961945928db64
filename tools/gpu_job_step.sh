#!/bin/bash
# ncu --set full of the c2 step kernel (fit + fused zero-fill) of the shipped library: summary, source histogram, traffic.json input
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
MB_ONLY=step timeout 600 python tools/microbench.py 2>&1 | tail -3
MB_ONLY=step timeout 900 ncu --set full --clock-control none --import-source on -k regex:fit_kernel -s 2 -c 1 -f -o /tmp/step python tools/microbench.py > gpurun_out/step_ncu.log 2>&1
ncu -i /tmp/step.ncu-rep --page raw --csv > gpurun_out/step_raw.csv
ncu -i /tmp/step.ncu-rep --page source --csv > gpurun_out/step_source.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/step_raw.csv > gpurun_out/step_summary.txt; head -14 gpurun_out/step_summary.txt
python tools/ncu_source_hist.py gpurun_out/step_source.csv > gpurun_out/step_source_hist.txt 2>&1; head -8 gpurun_out/step_source_hist.txt
} 2>&1 | tee gpurun_out/step_job.log
