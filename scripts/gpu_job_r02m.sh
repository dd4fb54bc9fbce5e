#!/bin/bash
TAG=r02m
mkdir -p gpurun_out
(cd scripts/microbench && timeout 120 bash build.sh tma_align) > gpurun_out/${TAG}_tma_align.txt 2>&1; tail -8 gpurun_out/${TAG}_tma_align.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/tools/sanitize_all.py > gpurun_out/${TAG}_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/${TAG}_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tests/tools/sanitize_all.py > gpurun_out/${TAG}_racecheck.txt 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/${TAG}_racecheck.txt
# DRAM traffic of the two big kernels at the driver's workload (B=64 1024x2048): one full-set capture each
timeout 1500 ncu --set full --clock-control none -k regex:'fwd_kernel|bwd_stash' -s 6 -c 2 -f -o gpurun_out/${TAG}_prof_c4s \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_c4s.log 2>&1
ls -la gpurun_out/${TAG}_prof_c4s.ncu-rep
