#!/bin/bash
TAG=r02i
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -s 2>&1 | grep -v "^$" | tail -25 | tee gpurun_out/${TAG}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.txt
timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/${TAG}_bench_c4s.err | tail -1 | tee gpurun_out/${TAG}_bench_c4s.json
timeout 600 python bench.py --workload c2 --steps 200 --warmup 10 --no-cpu-baseline 2>gpurun_out/${TAG}_bench_c2.err | tail -1 | tee gpurun_out/${TAG}_bench_c2.json
timeout 600 python bench.py --workload c1 --steps 200 --warmup 10 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/${TAG}_bench_c1.json
