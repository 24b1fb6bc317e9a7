set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:local_peaks_detect -s 12 -c 3 -o gpurun_out/r2_detect_f32 -f python tools/bench_kernels.py --iters 10 --only k1_cfg3_f32 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:local_peaks_detect -s 12 -c 3 -o gpurun_out/r2_detect_f16 -f python tools/bench_kernels.py --iters 10 --only k1_cfg3_f16 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:local_peaks_detect -s 12 -c 3 -o gpurun_out/r2_detect_bf16 -f python tools/bench_kernels.py --iters 10 --only k1_cfg3_bf16 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bottomup_tail|local_peaks_detect" -s 8 -c 4 -o gpurun_out/r2_bench_detect_tail -f python bench.py --steps 6 --warmup 3 --no-extras --no-cpu-baseline --streams 1 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_kernels_launches.csv python tools/bench_kernels.py --iters 3 --only k2_cfg2,k2_cfg2_f16,k7_cfg4,k7_cfg4_bf16,k8_cfg4_g8,targets_cfg4_fused,chain_cfg4 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/*launches.csv
