#!/bin/bash
# ncu of the cooperative L-BFGS-B kernel, summaries exported on the box (the .ncu-rep files are too large to pull back)
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in c2 c3; do
  for k in coop8 coop32 thread; do
    kn=lbfgsb_coop; [ $k = thread ] && kn=lbfgsb_kernel
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kn -c 1 -f -o /tmp/rep_${cfg}_$k python tools/lb_bench.py $cfg --kernels $k --scale 0.25 > gpurun_out/b_ncu_${cfg}_$k.log 2>&1
    ncu -i /tmp/rep_${cfg}_$k.ncu-rep --page raw --csv > /tmp/raw_${cfg}_$k.csv 2>/dev/null
    python tools/ncu_summary.py /tmp/raw_${cfg}_$k.csv > gpurun_out/b_${cfg}_${k}_summary.txt 2>&1
    ncu -i /tmp/rep_${cfg}_$k.ncu-rep --page source --csv > /tmp/src_${cfg}_$k.csv 2>/dev/null
    python tools/ncu_func_hist.py /tmp/src_${cfg}_$k.csv fetal_t2mapping_b200/csrc/libt2fit.so $kn >> gpurun_out/b_${cfg}_${k}_summary.txt 2>&1
    gzip -c /tmp/src_${cfg}_$k.csv > gpurun_out/b_src_${cfg}_$k.csv.gz
  done
done
ls -la gpurun_out
