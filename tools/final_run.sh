#!/bin/bash
# One gpurun call that regenerates the round's evidence:  gpurun --timeout 2400 -- 'bash tools/final_run.sh'
# Bench numbers are never taken under ncu; the ncu passes below only feed profiles/ (run tools/ncu_traffic.py and
# tools/ncu_summary.py on the results afterwards, WITHOUT editing csrc/peaks.cu in between: traffic.json keys on its hash).
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_final_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.txt 2>&1
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_final_bench_ref.json 2> /dev/null
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench_s20.json 2> gpurun_out/r2_final_bench_s20.err
python bench.py --steps 2000 --warmup 20 --no-extras > gpurun_out/r2_final_bench_f32.json 2> /dev/null
python bench.py --steps 2000 --warmup 20 --no-extras --dtype f16 > gpurun_out/r2_final_bench_f16.json 2> /dev/null
python bench.py --steps 2000 --warmup 20 --no-extras --graph > gpurun_out/r2_final_bench_graph.json 2> /dev/null
for dt in f32 f16 bf16; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:local_peaks_detect -s 12 -c 3 -o gpurun_out/r2_detect_$dt -f python tools/bench_kernels.py --iters 10 --only k1_cfg3_$dt > /dev/null 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bottomup_tail|local_peaks_detect" -s 8 -c 4 -o gpurun_out/r2_bench_detect_tail -f python bench.py --steps 6 --warmup 3 --no-extras --no-cpu-baseline --streams 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:confmaps_sep -s 12 -c 1 -o gpurun_out/r2_k7_sep_bf16 -f python tools/bench_kernels.py --iters 4 --only k7_cfg4_bf16 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1
# small-launch product numbers, tail phases (profiling builds: see the tools' docstrings for how the .so files are built)
SLEAPNN_B200_LIB=$PWD/sleap_nn_b200/lib/libsleapnn_b200_ab.so timeout 600 python tools/sweep_small_launch.py ceil prod > gpurun_out/r2_sweep_small_final.jsonl 2> /dev/null
SLEAPNN_B200_LIB=$PWD/sleap_nn_b200/lib/libsleapnn_b200_timing.so timeout 300 python tools/tail_phases.py cfg4 > gpurun_out/r2_tail_phases_cfg4_final.txt 2>&1
SLEAPNN_B200_LIB=$PWD/sleap_nn_b200/lib/libsleapnn_b200_timing.so timeout 300 python tools/tail_phases.py cfg3 > gpurun_out/r2_tail_phases_cfg3_final.txt 2>&1
timeout 600 python tools/latency_small_batch.py > gpurun_out/r2_latency_small_batch_final.jsonl 2> /dev/null
ls -la gpurun_out | tail -14
