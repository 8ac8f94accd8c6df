#!/bin/bash
# Builds a kernel-variant library for tools/kernel_sweep.py:  tools/build_variant.sh NAME -DD3Q_...=..
# -> gpurun_out/variants/libd3q19b200_NAME.so  (travels to the GPU box? no: gpurun_out/ is scratch and
# is NOT sent; variants are therefore written to build/variants/, which is git-ignored but sent).
set -e
here=$(cd "$(dirname "$0")/.." && pwd)
name=$1; shift
mkdir -p "$here/build/variants"
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared "$@" \
    -o "$here/build/variants/libd3q19b200_$name.so" "$here/d3q19-single-phase_b200/csrc/d3q19_api.cu" -ldl
echo "$here/build/variants/libd3q19b200_$name.so"
