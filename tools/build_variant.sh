#!/bin/bash
# Development aid: tools/build_variant.sh NAME [nvcc -D flags...] -> cilqr_b200/lib/variants/libcilqr_b200_NAME.so
set -e
cd "$(dirname "$0")/../cilqr_b200/csrc"
name=$1; shift
mkdir -p ../lib/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared --cudart static \
  -Xptxas -v "$@" -o ../lib/variants/libcilqr_b200_$name.so cilqr_capi.cu 2>&1 | grep -E "Used|spill" | grep -E "Used|[1-9][0-9]* bytes spill" | head -5
