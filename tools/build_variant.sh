#!/bin/sh
# Build a variant of the CUDA library for A/B timing: tools/build_variant.sh NAME [-DFLAG ...]
# -> gpurun_out/variants/libgadfly_b200_NAME.so  (gpurun_out/ does not travel; copy to variants/)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" \
    -o variants/libgadfly_b200_$name.so gadfly_b200/csrc/*.cu -ldl
echo variants/libgadfly_b200_$name.so
