#!/usr/bin/env bash
# Build libsleapnn_b200.so IN-TREE for sm_100a (the built .so travels to the GPU box with the repo snapshot).
# No fast-math: denormals, IEEE division and sqrt are part of the parity contract.
# Every .cu is compiled to an object in parallel, then linked.
#   SNB_NVCC_EXTRA   extra nvcc flags (e.g. -DSNB_AB_VARIANTS: the A/B kernel variants and their getenv switches,
#                    -DSNB_TAIL_TIMING: clock64 stamps in the fused tail) - profiling builds only (tools/)
#   SNB_LIB_NAME     output name of such a build (default libsleapnn_b200.so, the product)
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../lib"
mkdir -p "${OUT}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17
       -ftz=false -prec-div=true -prec-sqrt=true -fmad=true
       -Xcompiler -fPIC -Xcompiler -fvisibility=default ${SNB_NVCC_EXTRA:-})
LIBNAME="${SNB_LIB_NAME:-libsleapnn_b200.so}"
OBJ="${OUT}/obj_${LIBNAME%.so}"
mkdir -p "${OBJ}"
pids=()
for src in "${HERE}"/*.cu; do
  "${NVCC}" "${FLAGS[@]}" -c -o "${OBJ}/$(basename "${src}" .cu).o" "${src}" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "${p}"; done
"${NVCC}" -gencode arch=compute_100a,code=sm_100a -shared -o "${OUT}/${LIBNAME}" "${OBJ}"/*.o -lcudart
echo "built ${OUT}/${LIBNAME}"
