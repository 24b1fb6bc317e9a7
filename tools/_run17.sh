set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_n8.txt 2>&1
timeout 600 python -m pytest tests/test_baseline_sizes_gpu.py::test_second_device_runs_on_that_device tests/test_pipeline_gpu.py -m gpu -q -k "second_device" 2>&1 | tail -4 > gpurun_out/r2_n8_pytest_second_device.txt
cat gpurun_out/r2_n8_pytest_second_device.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 2000 --warmup 5 > gpurun_out/r2_n_bench_n8.json 2> gpurun_out/r2_n_bench_n8.err || tail -5 gpurun_out/r2_n_bench_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29572 tools/h2d_scaling_probe.py > gpurun_out/r2_h2d_probe_n8.json 2> gpurun_out/r2_h2d_probe_n8.err || tail -5 gpurun_out/r2_h2d_probe_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29573 tools/sweep_cfg5.py --frames 100000 > gpurun_out/r2_cfg5_n8.json 2> gpurun_out/r2_cfg5_n8.err || tail -5 gpurun_out/r2_cfg5_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29574 bench.py --gpus 8 --steps 2000 --warmup 5 --dtype f16 > gpurun_out/r2_n_bench_n8_f16.json 2> gpurun_out/r2_n_bench_n8_f16.err || tail -5 gpurun_out/r2_n_bench_n8_f16.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29575 bench.py --gpus 8 --steps 500 --warmup 5 --zero-copy-cms > gpurun_out/r2_n_bench_n8_zc.json 2> gpurun_out/r2_n_bench_n8_zc.err || tail -5 gpurun_out/r2_n_bench_n8_zc.err
python - <<'PY'
import json
for f in ('n8','n8_f16','n8_zc'):
    try:
        d=json.loads(open('gpurun_out/r2_n_bench_%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step']*1e3,2), 'e2e', round(d['e2e']['value']), d['run']['per_rank_ms_per_step']['min'], d['run']['per_rank_ms_per_step']['max'], d.get('gather',{}).get('gather_ms'), d['e2e'].get('host_cpus_rank0'))
    except Exception as e: print(f,'ERR',e)
try:
    d=json.loads(open('gpurun_out/r2_h2d_probe_n8.json').read().strip().splitlines()[-1])
    print({k:v for k,v in d.items() if 'GBps' in k})
    print(d['numa_nodes'][:400]); print(d['cpu_count'], d['affinity'])
except Exception as e: print('probe ERR',e)
print(open('gpurun_out/r2_cfg5_n8.json').read()[:1500])
PY
cat gpurun_out/r2_topo_n8.txt | head -14
