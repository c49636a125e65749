#!/bin/bash
# builds a variant of the library into build_variants/<tag>.so:  scripts/build_variant.sh tag [-DMACRO ...]
tag=$1; shift
cd "$(dirname "$0")/../datok_b200/csrc"
mkdir -p ../../build_variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
  -shared -o ../../build_variants/$tag.so kernels.cu format_kernels.cu api.cu model.cpp format.cpp multi.cpp -lz -lcudart_static -ldl -lrt -lpthread 2>&1 | grep -E "error" | head
ls -la ../../build_variants/$tag.so | awk '{print $5, $9}'
