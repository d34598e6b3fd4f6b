#!/bin/bash
# Build libltx_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC \
     ${NVCC_EXTRA:-} -o ../libltx_b200.so ltx_b200.cu
