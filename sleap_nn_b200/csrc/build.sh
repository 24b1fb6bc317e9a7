#!/usr/bin/env bash
# Build libsleapnn_b200.so IN-TREE for sm_100a (the built .so travels to the GPU box with the repo snapshot).
# No fast-math: denormals, IEEE division and sqrt are part of the parity contract.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../lib"
mkdir -p "${OUT}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17
       -ftz=false -prec-div=true -prec-sqrt=true -fmad=true
       -Xcompiler -fPIC -Xcompiler -fvisibility=default ${SNB_NVCC_EXTRA:-})
SRCS=("${HERE}"/*.cu)
LIBNAME="${SNB_LIB_NAME:-libsleapnn_b200.so}"  # SNB_LIB_NAME / SNB_NVCC_EXTRA: profiling variants (tools/)
"${NVCC}" "${FLAGS[@]}" -shared -o "${OUT}/${LIBNAME}" "${SRCS[@]}" -lcudart
echo "built ${OUT}/${LIBNAME}"
