set -x
mkdir -p gpurun_out
timeout 300 python bench.py --steps 2000 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_q_bench_f32.json 2> gpurun_out/r2_q_bench_f32.err || tail -5 gpurun_out/r2_q_bench_f32.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_q_bench_s20.json 2> gpurun_out/r2_q_bench_s20.err || tail -5 gpurun_out/r2_q_bench_s20.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_q_bench_s20b.json 2> gpurun_out/r2_q_bench_s20b.err || tail -5 gpurun_out/r2_q_bench_s20b.err
timeout 300 python bench.py --steps 2000 --warmup 5 --no-extras --no-cpu-baseline --dtype f16 > gpurun_out/r2_q_bench_f16.json 2> gpurun_out/r2_q_bench_f16.err || tail -5 gpurun_out/r2_q_bench_f16.err
timeout 300 python bench.py --steps 2000 --warmup 5 --no-extras --no-cpu-baseline --streams 4 > gpurun_out/r2_q_bench_f32_s4.json 2> gpurun_out/r2_q_bench_f32_s4.err || tail -5 gpurun_out/r2_q_bench_f32_s4.err
timeout 300 python bench.py --steps 2000 --warmup 5 --no-extras --no-cpu-baseline --streams 2 > gpurun_out/r2_q_bench_f32_s2.json 2> gpurun_out/r2_q_bench_f32_s2.err || tail -5 gpurun_out/r2_q_bench_f32_s2.err
python - <<'PY'
import json
for f in ('f32','s20','s20b','f16','f32_s4','f32_s2'):
    try:
        d=json.loads(open('gpurun_out/r2_q_bench_%s.json'%f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, round(d['value']), round(d['ms_per_step']*1e3,2), 'detect', round(r['avg_launch_ms']*1e3,2), round(r['frac'],3), 'e2e', round(d['e2e']['value']), 'issue', round(d['host_issue_us_per_step'],1))
    except Exception as e: print(f,'ERR',e)
PY
