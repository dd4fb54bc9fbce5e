#!/bin/bash
TAG=r02u
for rep in 1 2 3; do for v in base rollid; do MGVS_LIB_PATH=gpurun_variants/lib_$v.so timeout 300 python scripts/time_kernels.py c2 c4 2>&1 | tail -1; done; done | tee gpurun_out/${TAG}_rollid.txt
MGVS_LIB_PATH=gpurun_variants/lib_rollid.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "golden or against_oracle" 2>&1 | tail -2 | tee -a gpurun_out/${TAG}_rollid.txt
