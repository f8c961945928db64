#!/bin/bash
# dense L-BFGS-B kernel with the fused echo loop + shared-memory sums: parity tests, A/B against the round-2 base library and
# occupancy variants, ncu --set full on a full-size slab (lanes are refilled from the queue)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-s}
{
echo "== pytest dense"
timeout 900 python -m pytest tests/test_gpu_lbfgsb_dense.py -q -x 2>&1 | tail -3
echo "== lb_bench dense (new)"
timeout 600 python tools/lb_bench.py c2 c3 c3r c5 --kernels dense 2>&1 | grep -v "^$"
for v in $VARIANTS; do
  if [ -f fetal_t2mapping_b200/csrc/libt2fit_$v.so ]; then
    echo "== variant $v"
    T2FIT_LIB=$PWD/fetal_t2mapping_b200/csrc/libt2fit_$v.so timeout 600 python tools/lb_bench.py c2 c3 c3r c5 --kernels dense 2>&1 | grep -v "^$"
  fi
done
echo "== ncu --set full, dense kernel, c3 x 0.5"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lbfgsb_dense_kernel -s 1 -c 1 -f -o /tmp/${tag}_dense python tools/lb_profile.py c3 0.5 gaussian_rician lbfgsb_dense > gpurun_out/${tag}_ncu_dense.log 2>&1
ncu -i /tmp/${tag}_dense.ncu-rep --page raw --csv > gpurun_out/${tag}_dense_raw.csv
ncu -i /tmp/${tag}_dense.ncu-rep --page source --csv > gpurun_out/${tag}_dense_source.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${tag}_dense_raw.csv > gpurun_out/${tag}_dense_summary.txt; grep -E "time_duration|dram__bytes|registers_per|warps_active|issue_active|stalled_(long|no_inst|wait|short|math)|inst_executed.sum|thread_inst_executed_per|pipe_fp64|pipe_fma|pipe_alu|pipe_xu|shared" gpurun_out/${tag}_dense_summary.txt
python tools/ncu_source_hist.py gpurun_out/${tag}_dense_source.csv > gpurun_out/${tag}_dense_source_hist.txt 2>&1; head -42 gpurun_out/${tag}_dense_source_hist.txt
} 2>&1 | tee gpurun_out/${tag}_job.log
