#!/bin/bash
TAG=r02l
mkdir -p gpurun_out
for rep in 1 2; do for v in base pair pairlite; do MGVS_LIB_PATH=gpurun_variants/lib_$v.so timeout 300 python scripts/time_kernels.py c2 c4 2>&1 | tail -1; done; done | tee gpurun_out/${TAG}_bwd_pair.txt
MGVS_LIB_PATH=gpurun_variants/lib_pairlite.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "golden_backward or against_oracle or deterministic" 2>&1 | tail -3 | tee -a gpurun_out/${TAG}_bwd_pair.txt
