"""Batch sharding of the loss across ranks.

The path shards by image: every per-pixel quantity is local to one image, and the only batch-wide
quantities are the masked-mean numerators and denominators (loss.py:245,285-286).  Each rank therefore
runs the fused kernels on its contiguous slice of the batch and the ranks exchange ONE vector of 3n+3
doubles (layout: include/mgvs.h, mgvs_num_sums) with a sum all-reduce -- NCCL over NVLink on the GPU
box, gloo in the CPU tests.  Everything else (depth and pose gradients) stays local.

Two modes, selected on ``MultiViewPhotometricLoss``:
  process_group=None            every rank normalises by its local mask counts -- bit-for-bit what the
                                reference does under DDP (no exchange at all)
  process_group=g               global counts: the returned loss is the full-batch loss on every rank;
                                with ddp_grad_scale=True local gradients are multiplied by the world size
                                so that DDP's 1/G averaging reproduces the full-batch gradient exactly
"""
from __future__ import annotations

import torch

__all__ = ["batch_slice", "allreduce_sums", "losses_from_sums"]


def batch_slice(B: int, world_size: int, rank: int) -> slice:
    """Contiguous slice of the batch owned by ``rank`` (sizes differ by at most one image)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, extra = divmod(B, world_size)
    start = rank * base + min(rank, extra)
    return slice(start, start + base + (1 if rank < extra else 0))


def allreduce_sums(sums: torch.Tensor, group) -> int:
    """In-place sum all-reduce of the partial-sum vector; returns the world size of ``group``."""
    import torch.distributed as dist
    if sums.dtype != torch.float64:
        raise TypeError("partial sums travel as float64 (mask counts exceed 2^24)")
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return dist.get_world_size(group)


def losses_from_sums(sums, n: int, photometric_loss_weight: float, smoothing_loss_weight: float):
    """Host mirror of finalize_kernel (loss.py:151-154,252-254,274-294) -- used by the CPU tests of the
    sharding logic; the product path runs finalize_kernel on the device."""
    s = [float(x) for x in sums]
    N, Nx, Ny = s[n], s[3 * n + 1], s[3 * n + 2]
    lp = sum(s[i] / N for i in range(n)) / n
    ls = sum((s[n + 1 + i] / Nx + s[2 * n + 1 + i] / Ny) / (1 << i) for i in range(n)) / n
    return lp * photometric_loss_weight, ls * smoothing_loss_weight
