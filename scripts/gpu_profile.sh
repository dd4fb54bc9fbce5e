#!/bin/bash
# ncu launch list + one full capture of the forward and backward kernels of the default (stash) path.
# Usage (under gpurun): bash scripts/gpu_profile.sh <tag> [extra bench flags]
TAG=${1:-r01e}; shift
mkdir -p gpurun_out
for m in stash recompute; do MGVS_BACKWARD=$m timeout 300 python scripts/time_kernels.py c2 c3 c4 2>&1 | tail -1; done | tee gpurun_out/${TAG}_time_kernels.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fwd_kernel|bwd_' -s 6 -c 2 -f -o gpurun_out/${TAG}_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/${TAG}_prof.ncu-rep
