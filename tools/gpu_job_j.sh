#!/bin/bash
# A/B of inner-loop unrolling in the thread-per-voxel L-BFGS-B kernel (variant libraries built with -DT2_INNER_UNROLL=n)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in _b7 _b8 _b10 _b12; do
  echo "== variant libt2fit$v.so"
  T2FIT_LIB=$PWD/fetal_t2mapping_b200/csrc/libt2fit$v.so timeout 600 python tools/lb_bench.py c2 c3 c5 --kernels thread 2>&1 | grep -v "^$"
done | tee gpurun_out/j_unroll_ab.log
