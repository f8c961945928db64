#!/bin/bash
# multi-GPU bench lines of the configurations the dense L-BFGS-B kernel serves: N=$1, configs $2...
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}; shift
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N "$@" > gpurun_out/x_bench_${name}_n$N.json 2> gpurun_out/x_bench_${name}_n$N.err; echo "rc=$?"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/x_bench_${name}_n$N.err | tail -4 | cut -c1-400; python - gpurun_out/x_bench_${name}_n$N.json <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c=d["config"]
    print("   value %.4g ms/step %.2f scaling %s sharded %s" % (d["value"], d["ms_per_step"], d.get("scaling"), json.dumps(d.get("sharded"))[:400]))
except Exception as e: print("bad", e)
P
}
{
for cfg in "$@"; do
  case $cfg in
    c3) run c3 --config c3 --steps 5 --warmup 3 --no-cpu-baseline ;;
    c5) run c5 --config c5 --steps 3 --warmup 3 --no-cpu-baseline ;;
    c2) run c2 --steps 20 --warmup 5 --no-cpu-baseline ;;
    multi) timeout 900 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -3 ;;
  esac
done
} 2>&1 | tee gpurun_out/x_job_n$N.log
