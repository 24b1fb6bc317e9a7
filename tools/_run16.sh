set -x
mkdir -p gpurun_out
SNB_LIB_NAME=libsleapnn_b200_ab.so SNB_NVCC_EXTRA=-DSNB_AB_VARIANTS bash sleap_nn_b200/csrc/build.sh > /dev/null 2>&1
export SLEAPNN_B200_LIB=sleap_nn_b200/lib/libsleapnn_b200_ab.so
: > gpurun_out/r2_detect_ab.jsonl
python tools/detect_variants.py f32 >> gpurun_out/r2_detect_ab.jsonl
SNB_DETECT_BULK=1 python tools/detect_variants.py f32 >> gpurun_out/r2_detect_ab.jsonl
for v in 1 4 8; do SNB_DETECT_VARIANT=$v python tools/detect_variants.py f32 >> gpurun_out/r2_detect_ab.jsonl; done
python tools/detect_variants.py f16 >> gpurun_out/r2_detect_ab.jsonl
for v in 7 8 0; do SNB_DETECT_VARIANT=$v python tools/detect_variants.py f16 >> gpurun_out/r2_detect_ab.jsonl; done
cat gpurun_out/r2_detect_ab.jsonl
