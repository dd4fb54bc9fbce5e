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

__all__ = ["batch_slice", "allreduce_sums", "losses_from_sums", "PeerExchange"]


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


class PeerExchange:
    """Symmetric exchange buffers for ``mgvs_exchange_finalize`` (include/mgvs.h): the partial sums cross GPUs by NVLink
    P2P stores inside one single-CTA kernel that also finalizes the losses -- no NCCL launch on the step's critical path.

    Pass an instance as ``MultiViewPhotometricLoss(..., process_group=g, exchange=PeerExchange(g))``.  Construction is
    collective (every rank of the group): it allocates the buffer with torch's symmetric memory, maps the peers'
    buffers and zero-fills.  Single node only (NVLink / NVSwitch peers); raises if the mapping is unavailable.
    """

    def __init__(self, group=None, device=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > _lib.MAX_RANKS:
            raise ValueError("PeerExchange supports up to %d ranks" % _lib.MAX_RANKS)
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        nbytes = int(_lib.lib().mgvs_exchange_bytes())
        self.buffer = symm.empty(nbytes // 8, dtype=torch.int64, device=dev)
        self.handle = symm.rendezvous(self.buffer, self.group)
        self.buffer.zero_()
        torch.cuda.synchronize(dev)
        dist.barrier(group=self.group)          # nobody pushes before every buffer is zeroed
        x = _lib.MgvsPeerExchange()
        x.rank, x.world = self.rank, self.world
        for r, ptr in enumerate(self.handle.buffer_ptrs):
            x.peer_base[r] = int(ptr)
        self.struct = x

    def check(self) -> None:
        """Raises if an exchange on this rank ever timed out (MGVS_EXCHANGE_STATUS_OFFSET, include/mgvs.h): the ranks made
        unequal call sequences.  Synchronises the device -- call it outside the step's critical path (e.g. once per epoch)."""
        step = int(self.buffer[1].item())
        if step != 0:
            raise RuntimeError("mgvs_exchange_finalize timed out at step %d on rank %d: not every rank called it (losses were NaN)" % (step, self.rank))
