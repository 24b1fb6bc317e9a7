set -x
python tools/sweep_cfg5.py > gpurun_out/r1d_cfg5_n1.json 2> gpurun_out/r1d_cfg5_n1.err; cat gpurun_out/r1d_cfg5_n1.json | cut -c1-600
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r1d_bench_ref.json 2>/dev/null
python bench.py > gpurun_out/r1d_bench_n1.json 2> gpurun_out/r1d_bench_n1.err; cut -c1-400 gpurun_out/r1d_bench_n1.json
python tools/bench_kernels.py > gpurun_out/r1d_bench_kernels.jsonl 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1d_bench_launches.csv python bench.py --steps 4 --warmup 3 --e2e-steps 2 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/r1d_kernels_launches.csv python tools/bench_kernels.py --iters 3 --only k2_cfg2,k7_cfg4,k8_cfg4_g8 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"confmaps_rows2" -s 12 -c 1 -o gpurun_out/r1d_k7 python tools/bench_kernels.py --iters 3 --only k7_cfg4 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"local_peaks_detect|bottomup_tail" -s 8 -c 2 -o gpurun_out/r1d_cfg3 python tools/chain_once.py 8 > /dev/null 2>&1
ls -la gpurun_out | tail -12
