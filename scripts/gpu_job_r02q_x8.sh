#!/bin/bash
# final multi-GPU evidence at HEAD: 2-GPU sharded parity test, config[3] strong-scaling lines at N=2/4/8 (device-timed + e2e)
TAG=r02q
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_sharded_gpu.py -q -m gpu -k sharded 2>&1 | tail -2 | tee gpurun_out/${TAG}_pytest_sharded.txt
for N in 8 4 2; do
  timeout 900 $TR --nproc-per-node $N --master-port $((29700 + N)) bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/${TAG}_bench_c4s_x$N.err | tail -1 | tee gpurun_out/${TAG}_bench_c4s_x$N.json | cut -c1-200
done
