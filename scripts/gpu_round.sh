#!/bin/bash
# One GPU-box job: parity tests, smoke, bench lines, ncu launch list and one full capture per hot kernel.
# Usage (from the repo root, under gpurun): bash scripts/gpu_round.sh <tag> [quick]
TAG=${1:-r01}
MODE=${2:-full}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.txt
python bench.py --steps 30 --warmup 5 2>gpurun_out/${TAG}_bench_c2.err | tail -1 | tee gpurun_out/${TAG}_bench_c2.json
python bench.py --workload c3 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/${TAG}_bench_c3.json
python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/${TAG}_bench_c4.json
if [ "$MODE" = "full" ]; then
  python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | tee gpurun_out/${TAG}_bench_ref.json
  # launch list (cold-cache, serialised: compare shares)
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_launch.log 2>&1
  # one full capture of each hot kernel (skip warm-up launches)
  ncu --set full --clock-control none --import-source on -k regex:'fwd_kernel|bwd_' -s 6 -c 2 -f -o gpurun_out/${TAG}_prof \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_full.log 2>&1
fi
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/${TAG}_smi.txt
