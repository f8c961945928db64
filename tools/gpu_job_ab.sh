#!/bin/bash
# A/B of library variants on one box: lb_bench (dense kernel) for the shipped library and every libt2fit_$v.so in $VARIANTS, twice
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for rep in 1 2; do
echo "== shipped (rep $rep)"
timeout 600 python tools/lb_bench.py c2 c3 c3r c5 --kernels dense 2>&1 | grep -v "^$"
for v in $VARIANTS; do
  echo "== variant $v (rep $rep)"
  T2FIT_LIB=$PWD/fetal_t2mapping_b200/csrc/libt2fit_$v.so timeout 600 python tools/lb_bench.py c2 c3 c3r c5 --kernels dense 2>&1 | grep -v "^$"
done
done
} 2>&1 | tee gpurun_out/ab_job.log
