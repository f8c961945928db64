#!/bin/bash
# NVSwitch multicast form of the fused all-gather: two-GPU tests (multicast="on" block), then bench lines with T2FIT_MULTICAST on / off
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
{
nvidia-smi topo -m 2>&1 | head -12
if [ "$N" = "2" ]; then
echo "== pytest two-GPU tests"; timeout 240 python -m pytest tests/test_gpu_multi.py -q -x -m gpu -s 2>&1 | grep -E "multicast|passed|failed|Error|error" | tail -12
fi
for mc in ${2-on off}; do
echo "== bench N=$N T2FIT_MULTICAST=$mc"; T2FIT_MULTICAST=$mc timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --no-cpu-baseline > gpurun_out/mc_${mc}_bench_c2_n$N.json 2> gpurun_out/mc_${mc}_bench_c2_n$N.err || tail -5 gpurun_out/mc_${mc}_bench_c2_n$N.err
python - <<P
import json
d = json.loads(open("gpurun_out/mc_${mc}_bench_c2_n$N.json").read().strip().splitlines()[-1])
s = d["sharded"]
print("value %.4g ms/pass %.4f fit %.4f multicast %s note %s GB/s %.0f" % (d["value"], s["ms_per_pass"], s["fit_only_ms"], s.get("multicast"), s.get("multicast_note"), s["gather_gbs_received_per_rank"]), d.get("sharded_t2_s0_only", {}).get("ms_per_pass"))
P
done
} 2>&1 | tee gpurun_out/mc_job_n$N.log
