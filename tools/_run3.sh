set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_b_pytest.txt
cat gpurun_out/r2_b_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_b_bench_f32_s20.json 2> gpurun_out/r2_b_bench_f32_s20.err || tail -20 gpurun_out/r2_b_bench_f32_s20.err
timeout 300 python bench.py --steps 2000 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_b_bench_f32.json 2> gpurun_out/r2_b_bench_f32.err || tail -5 gpurun_out/r2_b_bench_f32.err
timeout 300 python bench.py --steps 2000 --warmup 5 --no-extras --dtype f16 > gpurun_out/r2_b_bench_f16.json 2> gpurun_out/r2_b_bench_f16.err || tail -5 gpurun_out/r2_b_bench_f16.err
timeout 300 python bench.py --steps 2000 --warmup 5 --no-extras --no-cpu-baseline --dtype bf16 > gpurun_out/r2_b_bench_bf16.json 2> gpurun_out/r2_b_bench_bf16.err || tail -5 gpurun_out/r2_b_bench_bf16.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_b_bench_ref.json 2> gpurun_out/r2_b_bench_ref.err || tail -5 gpurun_out/r2_b_bench_ref.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:local_peaks_detect -s 12 -c 3 -o gpurun_out/r2_b_detect_f32 -f python tools/bench_kernels.py --iters 10 --only k1_cfg3_f32 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:local_peaks_detect -s 12 -c 3 -o gpurun_out/r2_b_detect_f16 -f python tools/bench_kernels.py --iters 10 --only k1_cfg3_f16 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_b_bench_launches.csv python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -20
