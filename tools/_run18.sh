set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_final_pytest.txt
cat gpurun_out/r2_final_pytest.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.txt 2>&1; tail -4 gpurun_out/r2_final_smoke.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench_s20.json 2> gpurun_out/r2_final_bench_s20.err || tail -20 gpurun_out/r2_final_bench_s20.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_final_bench_ref.json 2> gpurun_out/r2_final_bench_ref.err || tail -5 gpurun_out/r2_final_bench_ref.err
timeout 300 python bench.py --steps 2000 --warmup 5 --no-extras > gpurun_out/r2_final_bench_f32.json 2> gpurun_out/r2_final_bench_f32.err || tail -5 gpurun_out/r2_final_bench_f32.err
timeout 300 python bench.py --steps 2000 --warmup 5 --no-extras --dtype f16 > gpurun_out/r2_final_bench_f16.json 2> gpurun_out/r2_final_bench_f16.err || tail -5 gpurun_out/r2_final_bench_f16.err
timeout 600 python tools/latency_small_batch.py > gpurun_out/r2_final_latency.jsonl 2> gpurun_out/r2_final_latency.err
python - <<'PY'
import json
for f in ('s20','f32','f16'):
    try:
        d=json.loads(open('gpurun_out/r2_final_bench_%s.json'%f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, round(d['value']), round(d['ms_per_step']*1e3,2), 'detect', round(r['avg_launch_ms']*1e3,2), round(r['frac'],3), 'e2e', round(d['e2e']['value']), d.get('parity'), (d.get('cpu_baseline') or {}).get('value'))
        for k,v in (d.get('extra') or {}).items(): print('   ',k, {kk:(round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk!='kernel'})
    except Exception as e: print(f,'ERR',e)
d=json.loads(open('gpurun_out/r2_final_bench_ref.json').read().strip().splitlines()[-1]); print('ref', d['value'], d['cpu_baseline']['kind'], d['cpu_baseline']['one_thread'])
for l in open('gpurun_out/r2_final_latency.jsonl'):
    d=json.loads(l); print(d['config'], {k:(round(v['chain_us'],1), round(v['tail_us'],1)) for k,v in d.items() if isinstance(v,dict) and 'chain_us' in v}, d.get('cpu_reference',{}).get('ms_per_call'))
PY
