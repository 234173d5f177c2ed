#!/bin/bash
# usage: scripts/build_variant.sh NAME [-DFLAG=V ...]  ->  build/variants/NAME.so  (A/B kernels; select with PB200_LIB_PATH)
set -e
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true \
     -ftz=false -shared -Xcompiler -fPIC,-O2,-Wall -cudart static -Xptxas -v "$@" \
     -o build/variants/$name.so proteus_b200/csrc/pb200_api.cu 2>&1 | grep -A2 "dswx_fused_stream_kernelILb1" | grep -v "^--" | tail -2
