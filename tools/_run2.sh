set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_a_pytest.txt
cat gpurun_out/r2_a_pytest.txt
timeout 300 python bench.py --steps 2000 --warmup 5 --no-cpu-baseline > gpurun_out/r2_a_bench_f32.json 2> gpurun_out/r2_a_bench_f32.err || tail -5 gpurun_out/r2_a_bench_f32.err
timeout 300 python bench.py --steps 2000 --warmup 5 --no-cpu-baseline --dtype f16 > gpurun_out/r2_a_bench_f16.json 2> gpurun_out/r2_a_bench_f16.err || tail -5 gpurun_out/r2_a_bench_f16.err
timeout 300 python bench.py --steps 2000 --warmup 5 --no-cpu-baseline --dtype bf16 > gpurun_out/r2_a_bench_bf16.json 2> gpurun_out/r2_a_bench_bf16.err || tail -5 gpurun_out/r2_a_bench_bf16.err
timeout 300 python bench.py --steps 500 --warmup 5 --no-cpu-baseline --zero-copy-cms > gpurun_out/r2_a_bench_f32_zc.json 2> gpurun_out/r2_a_bench_f32_zc.err || tail -5 gpurun_out/r2_a_bench_f32_zc.err
timeout 300 python bench.py --steps 500 --warmup 5 --no-cpu-baseline --zero-copy-cms --dtype f16 > gpurun_out/r2_a_bench_f16_zc.json 2> gpurun_out/r2_a_bench_f16_zc.err || tail -5 gpurun_out/r2_a_bench_f16_zc.err
cat gpurun_out/r2_a_bench_*.json
