"""Multi-GPU boxes only (skipped on one GPU): the batch-sharded loss over NCCL and over the fused peer-memory exchange
(mgvs_exchange_finalize) against the single-GPU full-batch loss, by launching scripts/check_sharded_nccl.py under torchrun."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_sharded_batch_nccl_and_peer_exchange_match_full_batch():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(ROOT, "scripts", "check_sharded_nccl.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "-> OK" in r.stdout and "bit-identical" in r.stdout


@pytest.mark.gpu
def test_full_cityscapes_batch_on_one_gpu():
    """BASELINE config[3] at its full size on ONE GPU (B=64, 1024x2048: 31 GB with the coefficient stash): size-independent
    property -- a batch of 8 distinct images repeated 8 times has the loss of the 8, 1/8 of their per-image gradients and
    repeating selection maps (scripts/check_big_batch.py).  Exercises every >2^31-byte offset of the kernels."""
    free, _ = torch.cuda.mem_get_info(0)
    if free < 48 * 2**30:
        pytest.skip("needs 48 GB of free device memory")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "check_big_batch.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "-> OK" in r.stdout
