set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 2000 --warmup 5 > gpurun_out/r2_r_bench_n8.json 2> gpurun_out/r2_r_bench_n8.err || tail -5 gpurun_out/r2_r_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29582 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_r_bench_n8_s20.json 2> gpurun_out/r2_r_bench_n8_s20.err || tail -5 gpurun_out/r2_r_bench_n8_s20.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29583 bench.py --gpus 8 --steps 2000 --warmup 5 --e2e-static > gpurun_out/r2_r_bench_n8_static.json 2> gpurun_out/r2_r_bench_n8_static.err || tail -5 gpurun_out/r2_r_bench_n8_static.err
python - <<'PY'
import json
for f in ('n8','n8_s20','n8_static'):
    try:
        d=json.loads(open('gpurun_out/r2_r_bench_%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step']*1e3,2), 'e2e', round(d['e2e']['value']), d['e2e'].get('steps_per_rank'), d['e2e'].get('split'), d['run']['per_rank_ms_per_step']['min'], d['run']['per_rank_ms_per_step']['max'], d.get('gather',{}).get('gather_ms'))
    except Exception as e: print(f,'ERR',e)
PY
