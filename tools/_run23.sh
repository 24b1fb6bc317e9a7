set -x
mkdir -p gpurun_out
for cfg in "8192 8" "4096 16" "32768 2"; do
  set -- $cfg
  SNB_LIB_NAME=libsleapnn_b200_ab.so SNB_NVCC_EXTRA="-DSNB_AB_VARIANTS -DSNB_TMA_STAGE_BYTES=$1 -DSNB_TMA_STAGES=$2" bash sleap_nn_b200/csrc/build.sh > /dev/null 2>&1
  echo "stage_bytes=$1 stages=$2" >> gpurun_out/r2_detect_tma.jsonl
  SLEAPNN_B200_LIB=sleap_nn_b200/lib/libsleapnn_b200_ab.so SNB_DETECT_TMA=1 SNB_DETECT_TMA_CTAS=3 timeout 120 python tools/detect_variants.py f32 >> gpurun_out/r2_detect_tma.jsonl || echo rc=$?
done
tail -6 gpurun_out/r2_detect_tma.jsonl
