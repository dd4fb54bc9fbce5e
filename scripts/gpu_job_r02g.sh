#!/bin/bash
TAG=r02g
mkdir -p gpurun_out
rm -f gpurun_out/${TAG}_train_step.jsonl
timeout 600 python tests/tools/train_step_mgnet.py --crop 512x1024 --batch 4 --out gpurun_out/${TAG}_train_step.jsonl 2>&1 | tail -3
timeout 600 python tests/tools/train_step_mgnet.py --crop 1024x1024 --batch 2 --out gpurun_out/${TAG}_train_step.jsonl 2>&1 | tail -3
