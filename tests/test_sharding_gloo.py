"""CPU, world_size 2, gloo: the N>1 path.  Each rank takes its slice of the batch, produces the
partial-sum vector (here with the CPU oracle standing in for the kernels -- same 3n+3 layout), the ranks
exchange it through mgnet_b200.sharding.allreduce_sums, and the results must equal the single-process
full-batch run: loss to 1e-6 relative, gradients identical up to rounding (global counts)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import l2rel, relerr
from mgnet_b200.sharding import allreduce_sums, batch_slice, losses_from_sums
from mgnet_b200.synthetic import make_inputs


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _slice_inputs(pred, tgt, sl):
    p = {"depth": [d[sl].contiguous() for d in pred["depth"]], "poses": pred["poses"][sl].contiguous()}
    t = {k: v[sl].contiguous() for k, v in tgt.items()}
    return p, t


def _worker(rank, world, port, B, H, W, n, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.oracle import Oracle
        torch.set_num_threads(1)
        pred, tgt = make_inputs(B, H, W, n, seed=31)
        sl = batch_slice(B, world, rank)
        p, t = _slice_inputs(pred, tgt, sl)
        o = Oracle(p, t)
        f = o.forward()
        sums = torch.from_numpy(f["sums"].copy())
        ws = allreduce_sums(sums, dist.group.WORLD)
        assert ws == world
        lp, ls = losses_from_sums(sums.tolist(), n, 1.0, 1e-3)
        g = o.backward(1.0, 1.0, sums=sums.numpy())
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), lp=lp, ls=ls, gp=g["grad_poses"],
                 **{"gd%d" % i: g["grad_depth"][i] for i in range(n)})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [4, 3])
def test_two_rank_sharding_matches_full_batch(tmp_path, B):
    H, W, n, world = 32, 64, 2, 2
    mp.spawn(_worker, args=(world, _free_port(), B, H, W, n, str(tmp_path)), nprocs=world, join=True)
    from oracle.oracle import Oracle
    pred, tgt = make_inputs(B, H, W, n, seed=31)
    o = Oracle(pred, tgt)
    f = o.forward()
    g = o.backward(1.0, 1.0)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    for r in range(world):
        assert relerr(parts[r]["lp"], f["loss_photometric"]) <= 1e-6
        assert relerr(parts[r]["ls"], f["loss_smoothness"]) <= 1e-6
    for i in range(n):
        cat = np.concatenate([parts[r]["gd%d" % i] for r in range(world)], 0)
        assert l2rel(cat, g["grad_depth"][i]) <= 1e-6
    assert l2rel(np.concatenate([parts[r]["gp"] for r in range(world)], 0), g["grad_poses"]) <= 1e-6


def test_batch_slice_partitions_exactly():
    for B in (1, 2, 7, 64):
        for world in (1, 2, 3, 8):
            idx = []
            for r in range(world):
                sl = batch_slice(B, world, r)
                idx += list(range(B))[sl]
            assert idx == list(range(B))
    with pytest.raises(ValueError):
        batch_slice(4, 2, 2)


def test_allreduce_requires_float64():
    with pytest.raises(TypeError):
        allreduce_sums(torch.zeros(4, dtype=torch.float32), None)
