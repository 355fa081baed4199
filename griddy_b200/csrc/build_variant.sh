#!/bin/bash
# build_variant.sh NAME [nvcc flags...] — an extra build of the library for kernel experiments:
# griddy_b200/csrc/variants/libgriddy_cuda_NAME.so, selected at run time with GRIDDY_CUDA_LIB.
set -e
here="$(cd "$(dirname "$0")" && pwd)"
name="$1"; shift
src="${GRIDDY_VARIANT_SRC:-$here}"
mkdir -p "$here/variants"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false \
  -I "$src/../../include" -shared -Xcompiler -fPIC "$@" -o "$here/variants/libgriddy_cuda_$name.so" "$src/griddy_cuda.cu"
echo "built variants/libgriddy_cuda_$name.so"
