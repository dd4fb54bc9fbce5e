#!/bin/bash
# run_variants.sh <workloads...> -- times every gpurun_variants/lib_*.so named in $VARIANTS (same box, back to back, twice)
WL="${@:-c2}"
for rep in 1 2; do for v in $VARIANTS; do MGVS_LIB_PATH=gpurun_variants/lib_$v.so python scripts/time_kernels.py $WL 2>&1 | tail -1; done; done | tee gpurun_out/variants_last.txt
