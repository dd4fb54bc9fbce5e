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


@pytest.mark.gpu
def test_exchange_timeout_reports_instead_of_trapping():
    """A rank that never arrives (unequal call sequences): mgvs_exchange_finalize must end its wait, write NaN losses and the step
    number into the status word -- no device trap, the context stays usable.  One GPU plays rank 0 of a
    2-rank world whose peer buffer nobody ever writes; the wait bound is shortened through MgvsPeerExchange.max_spins."""
    import ctypes
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from mgnet_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda:0")
    nbytes = int(L.mgvs_exchange_bytes())
    mine = torch.zeros(nbytes // 8, dtype=torch.int64, device=dev)
    ghost = torch.zeros(nbytes // 8, dtype=torch.int64, device=dev)        # the peer that never calls
    x = _lib.MgvsPeerExchange()
    x.rank, x.world, x.max_spins = 0, 2, 1 << 12
    x.peer_base[0], x.peer_base[1] = mine.data_ptr(), ghost.data_ptr()
    prob = _lib.MgvsProblem()
    prob.n = 3
    prob.photometric_weight, prob.smoothing_weight = 1.0, 1e-3
    sums = torch.ones(12, dtype=torch.float64, device=dev)
    losses = torch.zeros(2, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(L.mgvs_exchange_finalize(ctypes.byref(prob), ctypes.byref(x), sums.data_ptr(), losses.data_ptr(), st))
    torch.cuda.synchronize()                                               # would raise after a device trap
    assert torch.isnan(losses).all()
    assert int(mine[1].item()) == 1                                        # MGVS_EXCHANGE_STATUS_OFFSET: first step that timed out
    assert float((torch.ones(4, device=dev) * 2).sum().item()) == 8.0      # the context is alive
    # a complete world of one still works afterwards
    x1 = _lib.MgvsPeerExchange()
    x1.rank, x1.world = 0, 1
    solo = torch.zeros(nbytes // 8, dtype=torch.int64, device=dev)
    x1.peer_base[0] = solo.data_ptr()
    _lib.check(L.mgvs_exchange_finalize(ctypes.byref(prob), ctypes.byref(x1), sums.data_ptr(), losses.data_ptr(), st))
    torch.cuda.synchronize()
    assert torch.isfinite(losses).all() and int(solo[1].item()) == 0
