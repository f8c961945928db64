#!/bin/bash
# first GPU job of round 2: cooperative L-BFGS-B kernel -- correctness, A/B timing, racecheck, ncu
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt
timeout 600 python -m pytest tests/test_gpu_lbfgsb_coop.py -x -q > gpurun_out/a_test_coop.log 2>&1; echo "rc=$?" >> gpurun_out/a_test_coop.log
tail -5 gpurun_out/a_test_coop.log
timeout 900 python tools/lb_bench.py c2 c3 c3r c5 > gpurun_out/a_lb_bench.log 2>&1; echo "rc=$?" >> gpurun_out/a_lb_bench.log
cat gpurun_out/a_lb_bench.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/a_test_all.log 2>&1; echo "rc=$?" >> gpurun_out/a_test_all.log
tail -5 gpurun_out/a_test_all.log
cat > /tmp/race.py <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import fetal_t2mapping_b200 as t2
from tests.conftest import load_golden, fit_params_of
t2.init(0)
for name in ("c3_floor_noprior", "c2_gaussian_noprior"):
    g = load_golden(name); fp = fit_params_of(g)
    for k in ("coop8", "coop32"):
        os.environ["T2FIT_LB_KERNEL"] = k
        r = t2.fit_voxels_batch(torch.from_numpy(g["rows"][:96]).cuda(), None, g["te"], g["fit"], fp, prior=g["prior"], norm=g["norm"], solver="lbfgsb")
        torch.cuda.synchronize()
        print(name, k, float(r.nit.float().mean()))
PY
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python /tmp/race.py > gpurun_out/a_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/a_racecheck.log
tail -15 gpurun_out/a_racecheck.log
timeout 600 compute-sanitizer --tool memcheck python /tmp/race.py > gpurun_out/a_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/a_memcheck.log
tail -5 gpurun_out/a_memcheck.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lbfgsb_coop -c 1 -f -o gpurun_out/r02_coop8_c3 python tools/lb_bench.py c3 --kernels coop8 --scale 0.3 > gpurun_out/a_ncu_c3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lbfgsb_coop -c 1 -f -o gpurun_out/r02_coop8_c2 python tools/lb_bench.py c2 --kernels coop8 --scale 0.3 > gpurun_out/a_ncu_c2.log 2>&1
ls -la gpurun_out/
