#!/bin/bash
TAG=r02v
for rep in 1 2; do for v in base min1; do MGVS_LIB_PATH=gpurun_variants/lib_$v.so timeout 300 python scripts/time_kernels.py c2 c4 2>&1 | tail -1; done; done | tee gpurun_out/${TAG}_min1.txt
