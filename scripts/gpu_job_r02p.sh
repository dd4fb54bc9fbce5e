#!/bin/bash
TAG=r02p
for rep in 1 2; do for v in base pref; do MGVS_LIB_PATH=gpurun_variants/lib_$v.so timeout 300 python scripts/time_kernels.py c2 c4 2>&1 | tail -1; done; done | tee gpurun_out/${TAG}_bwd_prefetch.txt
MGVS_LIB_PATH=gpurun_variants/lib_pref.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "golden_backward or against_oracle" 2>&1 | tail -2 | tee -a gpurun_out/${TAG}_bwd_prefetch.txt
