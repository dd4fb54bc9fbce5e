#!/bin/bash
TAG=r02n
mkdir -p gpurun_out
timeout 1400 python -m pytest tests -q -m gpu -x 2>&1 | grep -v "^$" | grep -v Warning | tail -8 | tee gpurun_out/${TAG}_pytest_gpu.txt
timeout 300 python scripts/time_fused_upsample.py 2>&1 | tail -4 | tee gpurun_out/${TAG}_fused_upsample.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/${TAG}_bench_c4s.err | tail -1 | tee gpurun_out/${TAG}_bench_c4s.json
timeout 600 python bench.py --workload c2 --steps 200 --warmup 10 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/${TAG}_bench_c2.json
