mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
SLEAPNN_B200_LIB=$PWD/sleap_nn_b200/lib/libsleapnn_b200_timing.so timeout 300 python tools/tail_phases.py cfg4 > gpurun_out/r2_tail_phases_cfg4_v5.txt 2>&1
SLEAPNN_B200_LIB=$PWD/sleap_nn_b200/lib/libsleapnn_b200_timing.so timeout 300 python tools/tail_phases.py cfg3 > gpurun_out/r2_tail_phases_cfg3_v5.txt 2>&1
cat gpurun_out/r2_tail_phases_cfg4_v5.txt gpurun_out/r2_tail_phases_cfg3_v5.txt
