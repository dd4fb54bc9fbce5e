#!/bin/bash
# build.sh <name>: nvcc scripts/microbench/<name>.cu -> scripts/microbench/<name> (sm_100a), then run it
set -e
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o "$1" "$1.cu"
./"$1"
