set -x
mkdir -p gpurun_out
for conn in default 32 32 default; do
  if [ "$conn" = "default" ]; then unset CUDA_DEVICE_MAX_CONNECTIONS; else export CUDA_DEVICE_MAX_CONNECTIONS=$conn; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 2 --steps 2000 --warmup 5 --e2e-steps 10 > gpurun_out/r2_s_bench_n2_$conn.json 2> gpurun_out/r2_s_bench_n2.err || tail -5 gpurun_out/r2_s_bench_n2.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_s_bench_n2_$conn.json').read().strip().splitlines()[-1])
print('$conn', round(d['value']), d['run']['per_rank_ms_per_step']['all'])
PY
done
