set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_targets_gpu.py tests/test_batched_targets_gpu.py tests/test_baseline_sizes_gpu.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_k_pytest.txt
cat gpurun_out/r2_k_pytest.txt
timeout 600 python tools/bench_kernels.py --iters 200 --only targets_cfg4_fused,targets_cfg4_fused_bf16,k7_cfg4,k7_cfg4_bf16,k7_cfg4_g1_bf16 > gpurun_out/r2_k_kernels.jsonl 2> gpurun_out/r2_k_kernels.err
python -c "
import json
for l in open('gpurun_out/r2_k_kernels.jsonl'):
    d=json.loads(l); print(d.get('bench'), round(d.get('avg_launch_ms',0)*1e3,2),'us', round(d.get('frac',0),3), d.get('error',''))
"
SNB_LIB_NAME=libsleapnn_b200_timing.so SNB_NVCC_EXTRA=-DSNB_TAIL_TIMING bash sleap_nn_b200/csrc/build.sh > /dev/null 2>&1
SLEAPNN_B200_LIB=sleap_nn_b200/lib/libsleapnn_b200_timing.so python tools/tail_phases.py cfg4 > gpurun_out/r2_k_tail_phases_cfg4.txt 2>&1
cat gpurun_out/r2_k_tail_phases_cfg4.txt
SLEAPNN_B200_LIB=sleap_nn_b200/lib/libsleapnn_b200_timing.so python tools/tail_phases.py cfg3 > gpurun_out/r2_k_tail_phases_cfg3.txt 2>&1
cat gpurun_out/r2_k_tail_phases_cfg3.txt
