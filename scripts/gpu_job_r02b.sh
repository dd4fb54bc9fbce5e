#!/bin/bash
# Round-2 job B: new GPU tests (ABI v6: pose matrices, geometry API, un-snapped poses), forward variants (timing only)
TAG=r02b
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_gpu.txt
for rep in 1 2; do for v in base roll unr2 rollunr2 pipe abl2 abl6; do MGVS_LIB_PATH=gpurun_variants/lib_$v.so timeout 300 python scripts/time_kernels.py c2 c4 2>&1 | tail -1; done; done | tee gpurun_out/${TAG}_variants.txt
