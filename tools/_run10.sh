set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_targets_gpu.py tests/test_batched_targets_gpu.py tests/test_pipeline_gpu.py tests/test_baseline_sizes_gpu.py tests/test_paf_gpu.py -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2_i_pytest.txt
cat gpurun_out/r2_i_pytest.txt
timeout 600 python tools/bench_kernels.py --iters 200 --only targets_cfg4_fused,targets_cfg4_fused_bf16,targets_cfg4_fused_g8,k7_cfg4,k7_cfg4_bf16,k7_cfg4_g1,k7_cfg4_g1_bf16,k8_cfg4,k8_cfg4_g8,k8_cfg4_bf16,k8_cfg4_g8_bf16,k2_cfg2,k2_cfg2_f16,k1_cfg3_f32,k1_cfg3_f16 > gpurun_out/r2_i_kernels.jsonl 2> gpurun_out/r2_i_kernels.err
python -c "
import json
for l in open('gpurun_out/r2_i_kernels.jsonl'):
    d=json.loads(l); print(d.get('bench'), round(d.get('avg_launch_ms',0)*1e3,2),'us', round(d.get('frac',0),3), 'eager', round(d.get('eager_loop_ms',0)*1e3,2), 'graph', round((d.get('graph_replay_ms') or 0)*1e3,2), d.get('error',''))
"
tail -3 gpurun_out/r2_i_kernels.err
