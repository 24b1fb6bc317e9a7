mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pipeline_gpu.py tests/test_paf_gpu.py tests/test_baseline_sizes_gpu.py -m gpu -q 2>&1 | tail -4
SLEAPNN_B200_LIB=$PWD/sleap_nn_b200/lib/libsleapnn_b200_timing.so timeout 200 python tools/tail_phases.py cfg3 > gpurun_out/r2_tail_phases_cfg3_v7.txt 2>&1
SLEAPNN_B200_LIB=$PWD/sleap_nn_b200/lib/libsleapnn_b200_timing.so timeout 200 python tools/tail_phases.py cfg4 > gpurun_out/r2_tail_phases_cfg4_v7.txt 2>&1
grep "wall\|greedy\|total" gpurun_out/r2_tail_phases_cfg3_v7.txt gpurun_out/r2_tail_phases_cfg4_v7.txt
