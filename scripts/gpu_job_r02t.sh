#!/bin/bash
TAG=r02t
timeout 600 ncu --metrics gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:'fwd_kernel|bwd_stash' -s 6 -c 2 --csv --log-file gpurun_out/${TAG}_conflicts.csv python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
cat gpurun_out/${TAG}_conflicts.csv | tail -12
timeout 1400 python -m pytest tests -q -m gpu 2>&1 | grep -v "^$" | grep -v Warning | tail -3 | tee gpurun_out/${TAG}_pytest_gpu.txt
