#!/bin/bash
# GPU job for the stash backward: parity (verbose + asserting), kernel timings of both backward modes, bench line.
TAG=${1:-r01e}
mkdir -p gpurun_out
timeout 600 python tests/tools/gpu_check.py 2>&1 | tail -15 | tee gpurun_out/${TAG}_gpu_check.txt
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_gpu.txt
for m in stash recompute; do MGVS_BACKWARD=$m timeout 300 python scripts/time_kernels.py c2 c3 c4 2>&1 | tail -1; done | tee gpurun_out/${TAG}_time_kernels.txt
timeout 600 python bench.py --steps 30 --warmup 5 2>gpurun_out/${TAG}_bench_c2.err | tail -1 | tee gpurun_out/${TAG}_bench_c2.json
