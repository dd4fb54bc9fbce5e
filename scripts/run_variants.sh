for v in base s1 s2 s12 a b c d abcd pipe; do MGVS_LIB_PATH=gpurun_variants/lib_$v.so python scripts/time_kernels.py c2 c4 2>&1 | tail -1; done | tee gpurun_out/variants_phase.txt
