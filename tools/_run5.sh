set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_d_pytest.txt
cat gpurun_out/r2_d_pytest.txt
timeout 600 python tools/bench_kernels.py --iters 200 --only k1_cfg3_f32,k1_cfg3_f16,k1_cfg3_bf16,k2_cfg2,k2_cfg2_b1024,k2_cfg2_f16 > gpurun_out/r2_d_kernels.jsonl 2> gpurun_out/r2_d_kernels.err
python -c "
import json
for l in open('gpurun_out/r2_d_kernels.jsonl'):
    d=json.loads(l); print(d.get('bench'), round(d.get('avg_launch_ms',0)*1e3,2),'us', round(d.get('frac',0),3), d.get('error',''))
"
timeout 300 python bench.py --steps 2000 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_d_bench_f32.json 2> gpurun_out/r2_d_bench_f32.err || tail -5 gpurun_out/r2_d_bench_f32.err
timeout 300 python bench.py --steps 2000 --warmup 5 --no-extras --no-cpu-baseline --dtype f16 > gpurun_out/r2_d_bench_f16.json 2> gpurun_out/r2_d_bench_f16.err || tail -5 gpurun_out/r2_d_bench_f16.err
timeout 300 python bench.py --steps 2000 --warmup 5 --no-extras --no-cpu-baseline --graph > gpurun_out/r2_d_bench_f32_graph.json 2> gpurun_out/r2_d_bench_f32_graph.err || tail -5 gpurun_out/r2_d_bench_f32_graph.err
python -c "
import json
for f in ('f32','f16','f32_graph'):
    d=json.loads(open('gpurun_out/r2_d_bench_%s.json'%f).read().strip().splitlines()[-1]); r=d['roofline']
    print(f, round(d['value']), round(d['ms_per_step']*1e3,2), 'detect', round(r['avg_launch_ms']*1e3,2), round(r['frac'],3), 'insitu', round(r['in_situ_avg_launch_ms']*1e3,2), 'e2e', round(d['e2e']['value']), 'issue', round(d['host_issue_us_per_step'],1))
"
timeout 300 python tools/h2d_scaling_probe.py > gpurun_out/r2_h2d_probe_n1.json 2> gpurun_out/r2_h2d_probe_n1.err || tail -5 gpurun_out/r2_h2d_probe_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_h2d_probe_n1.json').read().strip().splitlines()[-1])
print({k:v for k,v in d.items() if k.endswith('_sum') or k.startswith('frames')})
"
