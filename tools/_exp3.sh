set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "target or global or baseline or peaks or layers" 2>&1 | tail -3
export SLEAPNN_B200_LIB=$PWD/sleap_nn_b200/lib/libsleapnn_b200_ab.so
timeout 900 python tools/sweep_small_launch.py k2u prod > gpurun_out/r2_sweep_small3.jsonl 2> gpurun_out/r2_sweep_small3.err
tail -5 gpurun_out/r2_sweep_small3.err
cat gpurun_out/r2_sweep_small3.jsonl
unset SLEAPNN_B200_LIB
for r in 1 2; do
python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('plain   ', d['value'], d['ms_per_step'])"
python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --e2e-steps 2 --fill-stagger 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stagger ', d['value'], d['ms_per_step'])"
done
python bench.py --steps 2000 --warmup 20 --no-extras --no-cpu-baseline --e2e-steps 2 --fill-stagger 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stagger2000 ', d['value'], d['ms_per_step'])"
