#!/bin/bash
TAG=r02e
mkdir -p gpurun_out
for rep in 1 2; do
MGVS_FORWARD_MODE=exact MGVS_LIB_PATH=gpurun_variants/lib_base.so timeout 300 python scripts/time_kernels.py c2 c4 2>&1 | tail -1
MGVS_FORWARD_MODE=gated MGVS_LIB_PATH=gpurun_variants/lib_base.so timeout 300 python scripts/time_kernels.py c2 c4 2>&1 | tail -1
MGVS_FORWARD_MODE=gated MGVS_LIB_PATH=gpurun_variants/lib_abl16.so timeout 300 python scripts/time_kernels.py c2 c4 2>&1 | tail -1
MGVS_BACKWARD=recompute MGVS_FORWARD_MODE=exact MGVS_LIB_PATH=gpurun_variants/lib_base.so timeout 300 python scripts/time_kernels.py c2 c4 2>&1 | tail -1
MGVS_BACKWARD=recompute MGVS_FORWARD_MODE=gated MGVS_LIB_PATH=gpurun_variants/lib_abl16.so timeout 300 python scripts/time_kernels.py c2 c4 2>&1 | tail -1
done | tee gpurun_out/${TAG}_modes.txt
