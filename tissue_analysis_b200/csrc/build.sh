#!/bin/bash
# Build libtissue_b200.so for sm_100a (in-tree; the .so travels to the GPU box, it is git-ignored).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -Xcompiler -fPIC -Xcompiler -O2 -shared -cudart shared \
  ${TA_NVCC_EXTRA} -o ${TA_OUT:-../libtissue_b200.so} ta_api.cu
