set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_c_pytest.txt
cat gpurun_out/r2_c_pytest.txt
timeout 600 python tools/bench_kernels.py --iters 200 --only k1_cfg3_f32,k1_cfg3_f16,k1_cfg3_bf16,k2_cfg2,k2_cfg2_b1024,k2_cfg2_f16,k7_cfg4,k7_cfg4_bf16,k7_cfg4_g1,k7_cfg4_g1_bf16,k8_cfg4,k8_cfg4_g8,k8_cfg4_bf16,k8_cfg4_g8_bf16 > gpurun_out/r2_c_kernels.jsonl 2> gpurun_out/r2_c_kernels.err
python -c "
import json
for l in open('gpurun_out/r2_c_kernels.jsonl'):
    d=json.loads(l); print(d.get('bench'), round(d.get('avg_launch_ms',0)*1e3,2),'us', round(d.get('frac',0),3), d.get('error',''))
"
timeout 300 python bench.py --steps 2000 --warmup 5 --no-extras > gpurun_out/r2_c_bench_f32.json 2> gpurun_out/r2_c_bench_f32.err || tail -5 gpurun_out/r2_c_bench_f32.err
timeout 300 python bench.py --steps 2000 --warmup 5 --no-extras --dtype f16 > gpurun_out/r2_c_bench_f16.json 2> gpurun_out/r2_c_bench_f16.err || tail -5 gpurun_out/r2_c_bench_f16.err
python -c "
import json
for f in ('f32','f16'):
    d=json.loads(open('gpurun_out/r2_c_bench_%s.json'%f).read().strip().splitlines()[-1]); r=d['roofline']
    print(f, round(d['value']), round(d['ms_per_step']*1e3,2), 'detect', round(r['avg_launch_ms']*1e3,2), round(r['frac'],3), 'e2e', round(d['e2e']['value']), d.get('parity'))
"
