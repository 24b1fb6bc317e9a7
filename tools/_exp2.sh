set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "target or paf or confmap or global or baseline or peaks" 2>&1 | tail -3
export SLEAPNN_B200_LIB=$PWD/sleap_nn_b200/lib/libsleapnn_b200_ab.so
SNB_K2_ROT=7 timeout 600 python -m pytest tests/test_peaks_gpu.py tests/test_baseline_sizes_gpu.py tests/test_layers_gpu.py -m gpu -x -q -k "global or cfg2 or topdown or centroid or single" 2>&1 | tail -3
timeout 900 python tools/sweep_small_launch.py k2rot k2u k7g k8g > gpurun_out/r2_sweep_small2.jsonl 2> gpurun_out/r2_sweep_small2.err
tail -5 gpurun_out/r2_sweep_small2.err
cat gpurun_out/r2_sweep_small2.jsonl
