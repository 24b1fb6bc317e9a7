set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_targets_gpu.py tests/test_batched_targets_gpu.py tests/test_baseline_sizes_gpu.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_g_pytest.txt
cat gpurun_out/r2_g_pytest.txt
timeout 600 python tools/bench_kernels.py --iters 200 --only targets_cfg4_fused,targets_cfg4_fused_bf16,targets_cfg4_fused_g8,k7_cfg4,k7_cfg4_bf16,k7_cfg4_g1,k7_cfg4_g1_bf16,k8_cfg4,k8_cfg4_g8,k8_cfg4_bf16,k8_cfg4_g8_bf16 > gpurun_out/r2_g_kernels.jsonl 2> gpurun_out/r2_g_kernels.err
python -c "
import json
for l in open('gpurun_out/r2_g_kernels.jsonl'):
    d=json.loads(l); print(d.get('bench'), round(d.get('avg_launch_ms',0)*1e3,2),'us', round(d.get('frac',0),3), d.get('error',''))
"
tail -3 gpurun_out/r2_g_kernels.err
