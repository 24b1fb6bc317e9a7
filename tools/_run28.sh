set -x
mkdir -p gpurun_out
for conn in default 32; do
  if [ "$conn" = "default" ]; then unset CUDA_DEVICE_MAX_CONNECTIONS; else export CUDA_DEVICE_MAX_CONNECTIONS=$conn; fi
  for extra in "--tail-priority" "--tail-priority --streams 2" "--tail-priority --streams 4"; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29592 bench.py --gpus 2 --steps 2000 --warmup 5 --e2e-steps 10 $extra > gpurun_out/r2_t.json 2> gpurun_out/r2_t.err || tail -5 gpurun_out/r2_t.err
    python - <<PY
import json
d=json.loads(open('gpurun_out/r2_t.json').read().strip().splitlines()[-1])
print('$conn', '$extra', round(d['value']), d['run']['per_rank_ms_per_step']['all'], round(d['host_issue_us_per_step'],1))
PY
  done
done
