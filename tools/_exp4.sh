set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
SLEAPNN_B200_LIB=$PWD/sleap_nn_b200/lib/libsleapnn_b200_timing.so timeout 300 python tools/tail_phases.py cfg4 > gpurun_out/r2_tail_phases_cfg4_v2.txt 2>&1
SLEAPNN_B200_LIB=$PWD/sleap_nn_b200/lib/libsleapnn_b200_timing.so timeout 300 python tools/tail_phases.py cfg3 > gpurun_out/r2_tail_phases_cfg3_v2.txt 2>&1
cat gpurun_out/r2_tail_phases_cfg4_v2.txt gpurun_out/r2_tail_phases_cfg3_v2.txt
timeout 300 python tools/latency_small_batch.py > gpurun_out/r2_latency_small_batch_v2.jsonl 2>/dev/null
cut -c1-420 gpurun_out/r2_latency_small_batch_v2.jsonl
SLEAPNN_B200_LIB=$PWD/sleap_nn_b200/lib/libsleapnn_b200_ab.so timeout 900 python tools/sweep_small_launch.py prod > gpurun_out/r2_sweep_small4.jsonl 2> gpurun_out/r2_sweep_small4.err
cat gpurun_out/r2_sweep_small4.jsonl
