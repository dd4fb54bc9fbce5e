#!/bin/bash
# Round-2 job C: gated forward vs exact forward (new tests), the existing parity suites under the gated default, timings of both modes
TAG=r02c
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gated_forward.py -x -q -s 2>&1 | tail -40 | tee gpurun_out/${TAG}_pytest_gated.txt
timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_gated_forward.py 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_gpu.txt
for rep in 1 2; do for m in exact gated; do MGVS_FORWARD_MODE=$m timeout 300 python scripts/time_kernels.py c2 c4 2>&1 | tail -1; done; done | tee gpurun_out/${TAG}_modes.txt
for m in exact gated; do MGVS_BACKWARD=recompute MGVS_FORWARD_MODE=$m timeout 300 python scripts/time_kernels.py c2 c4 2>&1 | tail -1; done | tee -a gpurun_out/${TAG}_modes.txt
