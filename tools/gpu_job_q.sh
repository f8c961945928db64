#!/bin/bash
# dense-matrix L-BFGS-B kernel: GPU parity tests, A/B against the compact thread kernel, occupancy variants, ncu capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest dense kernel"
timeout 900 python -m pytest tests/test_gpu_lbfgsb_dense.py -q -x -s 2>&1 | grep -v "^$" | tail -40
echo "== lb_bench thread vs dense"
timeout 900 python tools/lb_bench.py c2 c3 c3r c5 --kernels thread,dense 2>&1 | grep -v "^$"
for mb in 3 6 8; do
  if [ -f fetal_t2mapping_b200/csrc/libt2fit_lbd$mb.so ]; then
    echo "== variant T2_LBD_MIN_BLOCKS=$mb"
    T2FIT_LIB=$PWD/fetal_t2mapping_b200/csrc/libt2fit_lbd$mb.so timeout 600 python tools/lb_bench.py c2 c3 c5 --kernels dense 2>&1 | grep -v "^$"
  fi
done
echo "== ncu --set full, dense kernel, c3 x 0.25"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lbfgsb_dense_kernel -s 1 -c 1 -f -o /tmp/q_dense python tools/lb_profile.py c3 0.25 gaussian_rician lbfgsb_dense > gpurun_out/q_ncu_dense.log 2>&1
ncu -i /tmp/q_dense.ncu-rep --page raw --csv > gpurun_out/q_dense_raw.csv
ncu -i /tmp/q_dense.ncu-rep --page source --csv > gpurun_out/q_dense_source.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/q_dense_raw.csv > gpurun_out/q_dense_summary.txt; head -60 gpurun_out/q_dense_summary.txt
python tools/ncu_source_hist.py gpurun_out/q_dense_source.csv > gpurun_out/q_dense_source_hist.txt 2>&1; head -40 gpurun_out/q_dense_source_hist.txt
} 2>&1 | tee gpurun_out/q_job.log
