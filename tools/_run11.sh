set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:confmaps_rows2 -s 12 -c 1 -o gpurun_out/r2_j_k7_g1 -f python tools/bench_kernels.py --iters 4 --only k7_cfg4_g1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:confmaps_sep -s 12 -c 1 -o gpurun_out/r2_j_k7_sep -f python tools/bench_kernels.py --iters 4 --only k7_cfg4_bf16 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 600 python tools/latency_small_batch.py > gpurun_out/r2_j_latency.jsonl 2> gpurun_out/r2_j_latency.err || tail -5 gpurun_out/r2_j_latency.err
cat gpurun_out/r2_j_latency.jsonl
SNB_LIB_NAME=libsleapnn_b200_timing.so SNB_NVCC_EXTRA=-DSNB_TAIL_TIMING bash sleap_nn_b200/csrc/build.sh > /dev/null 2>&1
SLEAPNN_B200_LIB=sleap_nn_b200/lib/libsleapnn_b200_timing.so python tools/tail_phases.py cfg4 > gpurun_out/r2_j_tail_phases_cfg4.txt 2>&1
cat gpurun_out/r2_j_tail_phases_cfg4.txt
