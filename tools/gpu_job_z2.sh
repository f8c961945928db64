#!/bin/bash
# two-GPU re-check of the rebuilt library: the two-GPU tests and the N = 2 bench line (sharded step with the fused all-gather)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest two-GPU tests"; timeout 200 python -m pytest tests/test_gpu_multi.py -q -x -m gpu 2>&1 | tail -3
echo "== bench N=2"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/z2_bench_c2_n2.json 2> gpurun_out/z2_bench_c2_n2.err; python - <<'P'
import json
d = json.loads(open("gpurun_out/z2_bench_c2_n2.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "n_gpus", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["sharded"]["ms_per_pass"], d["sharded"]["equals_single_gpu_fit"], d["sharded"]["gathered_equals_owner"])
P
} 2>&1 | tee gpurun_out/z2_job.log
