#!/bin/bash
# Builds tokenmonster_b200/libtmgpu.so (sm_100a only).  nvcc cross-compiles without a GPU.
set -e
cd "$(dirname "$0")/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo \
  -Xcompiler -fPIC,-O2,-Wall -shared -cudart shared \
  -o ../libtmgpu.so abi.cu abi_tokenize.cu abi_decode.cu abi_wire.cu kernels.cu vocab_tables.cpp unicode_tables.cpp "$@"
echo "built $(cd .. && pwd)/libtmgpu.so"
