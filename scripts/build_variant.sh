#!/bin/bash
# build an experiment variant of the library next to the shipped one: bash scripts/build_variant.sh _old -DAGX_PERSISTENT=0
# (scripts/ab_variants.sh then A/Bs the variants inside one gpurun call)
set -e
suffix=$1; shift
cd "$(dirname "$0")/../agilex-ntt_b200/csrc"
mkdir -p ../lib
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -shared \
     -o ../lib/libagxntt$suffix.so agx_api.cu agx_tables.cpp
echo built ../lib/libagxntt$suffix.so "$@"
