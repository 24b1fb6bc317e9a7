set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_n4.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 tools/h2d_scaling_probe.py > gpurun_out/r2_h2d_probe_n4.json 2> gpurun_out/r2_h2d_probe_n4.err || tail -5 gpurun_out/r2_h2d_probe_n4.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/h2d_scaling_probe.py > gpurun_out/r2_h2d_probe_n2.json 2> gpurun_out/r2_h2d_probe_n2.err || tail -5 gpurun_out/r2_h2d_probe_n2.err
for n in 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 2000 --warmup 5 > gpurun_out/r2_e_bench_n$n.json 2> gpurun_out/r2_e_bench_n$n.err || tail -5 gpurun_out/r2_e_bench_n$n.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29549 bench.py --gpus 4 --steps 2000 --warmup 5 --dtype f16 > gpurun_out/r2_e_bench_n4_f16.json 2> gpurun_out/r2_e_bench_n4_f16.err || tail -5 gpurun_out/r2_e_bench_n4_f16.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29550 bench.py --gpus 4 --steps 2000 --warmup 5 --graph > gpurun_out/r2_e_bench_n4_graph.json 2> gpurun_out/r2_e_bench_n4_graph.err || tail -5 gpurun_out/r2_e_bench_n4_graph.err
timeout 300 python bench.py --steps 2000 --warmup 5 --no-extras --no-cpu-baseline --graph > gpurun_out/r2_e_bench_n1_graph.json 2> gpurun_out/r2_e_bench_n1_graph.err || tail -5 gpurun_out/r2_e_bench_n1_graph.err
python - <<'PY'
import json
for f in ('n4','n2','n4_f16','n4_graph','n1_graph'):
    try:
        d=json.loads(open('gpurun_out/r2_e_bench_%s.json'%f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, round(d['value']), round(d['ms_per_step']*1e3,2), 'e2e', round(d['e2e']['value']), 'issue', round(d['host_issue_us_per_step'],1), d['run']['per_rank_ms_per_step'], d.get('gather'))
    except Exception as e: print(f,'ERR',e)
for n in (4,2):
    try:
        d=json.loads(open('gpurun_out/r2_h2d_probe_n%d.json'%n).read().strip().splitlines()[-1])
        print(n,{k:v for k,v in d.items() if 'GBps' in k})
    except Exception as e: print(n,'ERR',e)
PY
cat gpurun_out/r2_topo_n4.txt
