set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 8 --steps 2000 --warmup 5 > gpurun_out/r2_v_bench_n8.json 2> gpurun_out/r2_v_bench_n8.err || tail -5 gpurun_out/r2_v_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_v_bench_n8_s20.json 2> gpurun_out/r2_v_bench_n8_s20.err || tail -5 gpurun_out/r2_v_bench_n8_s20.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29603 bench.py --gpus 8 --steps 2000 --warmup 5 --dtype f16 > gpurun_out/r2_v_bench_n8_f16.json 2> gpurun_out/r2_v_bench_n8_f16.err || tail -5 gpurun_out/r2_v_bench_n8_f16.err
python - <<'PY'
import json
for f in ('n8','n8_s20','n8_f16'):
    try:
        d=json.loads(open('gpurun_out/r2_v_bench_%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step']*1e3,2), 'e2e', round(d['e2e']['value']), d['e2e'].get('steps_per_rank'), d['run']['per_rank_ms_per_step']['min'], d['run']['per_rank_ms_per_step']['max'], d.get('gather',{}).get('gather_ms'))
    except Exception as e: print(f,'ERR',e)
PY
