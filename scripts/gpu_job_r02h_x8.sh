#!/bin/bash
# 8-GPU box: sharded-vs-full-batch parity (NCCL + peer exchange), config[3] strong-scaling line at N=8 and N=4, config[4] DDP training step at N=2/4/8
TAG=r02h
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m 2>/dev/null | head -14 > gpurun_out/${TAG}_topo.txt
timeout 600 python -m pytest tests/test_sharded_gpu.py -q -m gpu -k sharded 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest_sharded.txt
timeout 900 $TR --nproc-per-node 8 --master-port 29601 bench.py --gpus 8 --steps 20 --warmup 5 2>gpurun_out/${TAG}_bench_c4s_x8.err | tail -1 | tee gpurun_out/${TAG}_bench_c4s_x8.json
timeout 900 $TR --nproc-per-node 4 --master-port 29602 bench.py --gpus 4 --steps 12 --warmup 4 2>gpurun_out/${TAG}_bench_c4s_x4.err | tail -1 | tee gpurun_out/${TAG}_bench_c4s_x4.json
rm -f gpurun_out/${TAG}_train_step.jsonl
timeout 600 $TR --nproc-per-node 8 --master-port 29603 tests/tools/train_step_mgnet.py --crop 1024x1024 --batch 2 --iters 8 --out gpurun_out/${TAG}_train_step.jsonl 2>&1 | grep TRAIN_STEP
timeout 600 $TR --nproc-per-node 8 --master-port 29604 tests/tools/train_step_mgnet.py --crop 512x1024 --batch 4 --iters 8 --out gpurun_out/${TAG}_train_step.jsonl 2>&1 | grep TRAIN_STEP
timeout 600 $TR --nproc-per-node 4 --master-port 29605 tests/tools/train_step_mgnet.py --crop 1024x1024 --batch 2 --iters 8 --variants fused,fused_upsample --out gpurun_out/${TAG}_train_step.jsonl 2>&1 | grep TRAIN_STEP
timeout 600 $TR --nproc-per-node 2 --master-port 29606 tests/tools/train_step_mgnet.py --crop 1024x1024 --batch 2 --iters 8 --variants fused,fused_upsample --out gpurun_out/${TAG}_train_step.jsonl 2>&1 | grep TRAIN_STEP
