set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
SLEAPNN_B200_LIB=$PWD/sleap_nn_b200/lib/libsleapnn_b200_timing.so timeout 300 python tools/tail_phases.py cfg4 > gpurun_out/r2_tail_phases_cfg4_v4.txt 2>&1
SLEAPNN_B200_LIB=$PWD/sleap_nn_b200/lib/libsleapnn_b200_timing.so timeout 300 python tools/tail_phases.py cfg3 > gpurun_out/r2_tail_phases_cfg3_v4.txt 2>&1
cat gpurun_out/r2_tail_phases_cfg4_v4.txt gpurun_out/r2_tail_phases_cfg3_v4.txt
SNB_LATENCY_NO_CPU=1 timeout 300 python tools/latency_small_batch.py > gpurun_out/r2_latency_small_batch_v4.jsonl 2>/dev/null
cut -c1-330 gpurun_out/r2_latency_small_batch_v4.jsonl
python bench.py --steps 2000 --warmup 20 --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench2000', d['value'], d['ms_per_step'], d['extra'].get('chain_cfg4'))"
