set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_targets_gpu.py tests/test_batched_targets_gpu.py tests/test_baseline_sizes_gpu.py tests/test_pipeline_gpu.py tests/test_paf_gpu.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_o_pytest.txt
cat gpurun_out/r2_o_pytest.txt
timeout 600 python tools/bench_kernels.py --iters 200 --only targets_cfg4_fused_bf16,k7_cfg4_bf16,k7_cfg4_g1_bf16 > gpurun_out/r2_o_kernels.jsonl 2> gpurun_out/r2_o_kernels.err
python -c "
import json
for l in open('gpurun_out/r2_o_kernels.jsonl'):
    d=json.loads(l); print(d.get('bench'), round(d.get('avg_launch_ms',0)*1e3,2),'us', round(d.get('frac',0),3), d.get('error',''))
"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:confmaps_sep -s 12 -c 1 -o gpurun_out/r2_o_k7_sep -f python tools/bench_kernels.py --iters 4 --only k7_cfg4_bf16 > /dev/null 2>&1
