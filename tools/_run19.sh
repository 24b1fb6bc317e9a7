set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -4
# sanitizer on the kernels touched this round (small cases only: the tools slow kernels down 10-100x)
SAN="tests/test_pipeline_gpu.py::test_cluster_tail_equals_single_cta_tail_and_oracle tests/test_pipeline_gpu.py::test_fused_pipeline_matches_reference_golden tests/test_baseline_sizes_gpu.py::test_half_maps_every_detect_path tests/test_batched_targets_gpu.py::test_batched_targets_golden tests/test_peaks_gpu.py"
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 5 python -m pytest $SAN -m gpu -x -q -k "not 1024 and not full_size and not batch256" > gpurun_out/r2_sanitizer_$tool.txt 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/r2_sanitizer_$tool.txt | tail -4
done
