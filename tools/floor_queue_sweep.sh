python -m pytest tests/test_gpu_parity.py -q -x -k "floor_queue or floor_model or fused_zero_fill or edge_cases or host_and_device" 2>&1 | tail -5
for k in oneshot queue; do for r in 4 8 16; do
  if [ $k = oneshot ] && [ $r != 8 ]; then continue; fi
  echo "== kernel $k refill $r"; T2FIT_FLOOR_KERNEL=$k T2FIT_QUEUE_REFILL=$r python tools/bench_configs.py c3 c5 2>&1 | grep "solver=fast"
done; done
