set -x
mkdir -p gpurun_out
SNB_LIB_NAME=libsleapnn_b200_ab.so SNB_NVCC_EXTRA=-DSNB_AB_VARIANTS bash sleap_nn_b200/csrc/build.sh > /dev/null 2>&1
export SLEAPNN_B200_LIB=sleap_nn_b200/lib/libsleapnn_b200_ab.so
: > gpurun_out/r2_detect_tma.jsonl
timeout 120 python tools/detect_variants.py f32 >> gpurun_out/r2_detect_tma.jsonl
for c in 3 2 1; do SNB_DETECT_TMA=1 SNB_DETECT_TMA_CTAS=$c timeout 120 python tools/detect_variants.py f32 >> gpurun_out/r2_detect_tma.jsonl || echo "tma f32 ctas $c rc=$?"; done
SNB_DETECT_BULK=1 timeout 120 python tools/detect_variants.py f32 >> gpurun_out/r2_detect_tma.jsonl
timeout 120 python tools/detect_variants.py f16 >> gpurun_out/r2_detect_tma.jsonl
for c in 3 2; do SNB_DETECT_TMA=1 SNB_DETECT_TMA_CTAS=$c timeout 120 python tools/detect_variants.py f16 >> gpurun_out/r2_detect_tma.jsonl || echo "tma f16 rc=$?"; done
cat gpurun_out/r2_detect_tma.jsonl
