#!/bin/bash
# build_variant.sh <name> [-D flags...]  -> gpurun_variants/lib_<name>.so (experiments; not the product build)
NAME=$1; shift
mkdir -p gpurun_variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared "$@" -Xptxas -v \
  -o gpurun_variants/lib_${NAME}.so mgnet_b200/csrc/mgvs_api.cu 2>&1 | grep -E "error|fwd_kernelILb1|bwd_kernelILb1|bwd_stash_kernelILb1ELb0" -A2 | grep -E "error|registers|spill" | tr '\n' ' '
echo " <- $NAME"
