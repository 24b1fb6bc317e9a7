set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_p_pytest.txt
cat gpurun_out/r2_p_pytest.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_p_bench_s20.json 2> gpurun_out/r2_p_bench_s20.err || tail -20 gpurun_out/r2_p_bench_s20.err
timeout 300 python bench.py --steps 2000 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_p_bench_f32.json 2> gpurun_out/r2_p_bench_f32.err || tail -5 gpurun_out/r2_p_bench_f32.err
python - <<'PY'
import json
for f in ('s20','f32'):
    try:
        d=json.loads(open('gpurun_out/r2_p_bench_%s.json'%f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, round(d['value']), round(d['ms_per_step']*1e3,2), 'detect', round(r['avg_launch_ms']*1e3,2), round(r['frac'],3), 'traffic', r['traffic'], (r['traffic_record'] or {}).get('source_matches'), 'e2e', round(d['e2e']['value']), 'issue', round(d['host_issue_us_per_step'],1))
        for k,v in (d.get('extra') or {}).items():
            if 'k7' in k or 'targets' in k: print('   ',k, {kk:(round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk!='kernel'})
    except Exception as e: print(f,'ERR',e)
PY
