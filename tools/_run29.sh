set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_sharding.py tests/test_baseline_sizes_gpu.py -m gpu -x -q 2>&1 | tail -3
for extra in "" "--no-tail-priority" "--streams 2" "--dtype f16"; do
  timeout 300 python bench.py --steps 2000 --warmup 5 --no-extras --no-cpu-baseline $extra > gpurun_out/r2_u.json 2> gpurun_out/r2_u.err || tail -5 gpurun_out/r2_u.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_u.json').read().strip().splitlines()[-1]); r=d['roofline']
print('2000 [$extra]', round(d['value']), round(d['ms_per_step']*1e3,2), 'detect', round(r['avg_launch_ms']*1e3,2), round(r['frac'],3), 'e2e', round(d['e2e']['value']), 'issue', round(d['host_issue_us_per_step'],1))
PY
done
for extra in "" "--no-tail-priority"; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline $extra > gpurun_out/r2_u.json 2> gpurun_out/r2_u.err || tail -5 gpurun_out/r2_u.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_u.json').read().strip().splitlines()[-1]); r=d['roofline']
print('20 [$extra]', round(d['value']), round(d['ms_per_step']*1e3,2))
PY
done
timeout 300 python tools/bench_kernels.py --iters 60 --only chain_cfg4 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l); print(d['bench'], round(d['avg_launch_ms']*1e3,1), round(d['frac'],3))
    except Exception: pass
"
