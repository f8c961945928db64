#!/bin/bash
# round 2, job I (N GPUs): the c2 bench line only
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-4}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/i_bench_c2_n$N.json 2> gpurun_out/i_bench_c2_n$N.err; echo "rc=$?"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/i_bench_c2_n$N.err | tail -6 | cut -c1-400
python - gpurun_out/i_bench_c2_n$N.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","n_gpus")}, "ms/pass", d["config"].get("ms_per_pass"))
for k in ("sharded","sharded_t2_s0_only","sharded_nccl"):
    if d.get(k): print(" ",k, {a:b for a,b in d[k].items() if a not in ("op","note","what")})
PY
