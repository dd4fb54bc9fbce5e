#!/bin/bash
# Round-2 job A: full GPU test-suite (incl. the new oracle comparisons at the benchmarked shapes), forward ablations and
# tile-shape variants (timing only), the new default bench line (config[3] on one GPU).
TAG=r02a
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/${TAG}_smi.txt
python -c "import os; print('cpus', os.cpu_count())"; free -g | head -2
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_gpu.txt
for v in base abl1 abl2 abl4 abl8 abl1; do MGVS_LIB_PATH=gpurun_variants/lib_$v.so timeout 300 python scripts/time_kernels.py c2 c4 2>&1 | tail -1; done | tee gpurun_out/${TAG}_ablations.txt
for v in base t32x32 t32x16c4 t64x32 t128x8 t32x24c3 base; do MGVS_BACKWARD=recompute MGVS_LIB_PATH=gpurun_variants/lib_$v.so timeout 300 python scripts/time_kernels.py c2 c4 2>&1 | tail -1; done | tee gpurun_out/${TAG}_tiles.txt
timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/${TAG}_bench_c4s.err | tail -1 | tee gpurun_out/${TAG}_bench_c4s.json
timeout 600 python bench.py --workload c2 --steps 30 --warmup 5 2>gpurun_out/${TAG}_bench_c2.err | tail -1 | tee gpurun_out/${TAG}_bench_c2.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | tee gpurun_out/${TAG}_bench_ref.json
