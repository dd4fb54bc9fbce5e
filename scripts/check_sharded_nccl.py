"""Run under torchrun on N GPUs: the batch-sharded loss (NCCL all-reduce of 3n+3 doubles) must equal the
single-GPU full-batch loss; concatenated local gradients must equal the full-batch gradients.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_sharded_nccl.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from mgnet_b200 import MultiViewPhotometricLoss
from mgnet_b200.sharding import batch_slice
from mgnet_b200.synthetic import make_inputs

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
HP = dict(ssim_loss_weight=0.85, photometric_loss_weight=1.0, smoothing_loss_weight=1e-3, automask_loss=True,
          photometric_reduce_op="min", padding_mode="zeros")
B, H, W, n = 2 * world + 1, 96, 320, 3      # odd batch: uneven slices
pred, tgt = make_inputs(B, H, W, n, seed=77)

from mgnet_b200.sharding import PeerExchange
peer = PeerExchange(dist.group.WORLD)

def run(p, t, group, exchange=None):
    mod = MultiViewPhotometricLoss(process_group=group, exchange=exchange, **HP)
    pd = {"depth": [d.to(dev).requires_grad_(True) for d in p["depth"]], "poses": p["poses"].to(dev).requires_grad_(True)}
    td = {k: v.to(dev) for k, v in t.items()}
    out = mod(pd, td)
    (out["loss_photometric"] + out["loss_smoothness"]).backward()
    torch.cuda.synchronize()
    return out["loss_photometric"].item(), out["loss_smoothness"].item(), [d.grad for d in pd["depth"]], pd["poses"].grad, mod.last_selection

sl = batch_slice(B, world, rank)
ps = {"depth": [d[sl].contiguous() for d in pred["depth"]], "poses": pred["poses"][sl].contiguous()}
ts = {k: v[sl].contiguous() for k, v in tgt.items()}
lp, ls, gd, gp, sel = run(ps, ts, dist.group.WORLD)
flp, fls, fgd, fgp, fsel = run(pred, tgt, None)     # full batch on every rank (single-GPU path)
ok = abs(lp - flp) <= 1e-6 * abs(flp) and abs(ls - fls) <= 1e-6 * abs(fls)
for i in range(n):
    a, b = gd[i].double(), fgd[i][sl].double()
    ok = ok and float((a - b).norm() / b.norm()) <= 1e-6
ok = ok and float((gp.double() - fgp[sl].double()).norm() / fgp[sl].double().norm()) <= 1e-6
ok = ok and bool((sel == fsel[:, sl]).all())
# the same through the fused peer-memory exchange (mgvs_exchange_finalize): bit-identical to the NCCL path on every rank,
# repeatedly (generations alternate), and all ranks must agree on the loss bits
ok_peer = True
for rep in range(5):
    plp, pls, pgd, pgp, psel = run(ps, ts, dist.group.WORLD, peer)
    ok_peer = ok_peer and (plp == lp) and (pls == ls) and all(torch.equal(a, b) for a, b in zip(pgd, gd)) and torch.equal(pgp, gp)
both = torch.tensor([plp, pls], device=dev, dtype=torch.float64)
gathered = [torch.zeros_like(both) for _ in range(world)]
dist.all_gather(gathered, both)
ok_peer = ok_peer and all(torch.equal(g, gathered[0]) for g in gathered)
if rank == 0:
    print("peer-memory exchange x%d vs NCCL all-reduce: loss %.10f/%.10f -> %s" % (world, plp, lp, "bit-identical" if ok_peer else "MISMATCH"))
ok = ok and ok_peer
# time the forward (the part the exchange sits in) both ways
def time_fwd(exchange, reps=200):
    mod = MultiViewPhotometricLoss(process_group=dist.group.WORLD, exchange=exchange, **HP)
    pd = {"depth": [d.to(dev) for d in ps["depth"]], "poses": ps["poses"].to(dev)}
    td = {k: v.to(dev) for k, v in ts.items()}
    for _ in range(20): mod(pd, td)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): mod(pd, td)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
t_nccl, t_peer = time_fwd(None), time_fwd(peer)
if rank == 0:
    print("forward-only step on a small slice (B=%d/rank %dx%d): NCCL %.1f us, peer exchange %.1f us" % (B // world, H, W, t_nccl * 1e3, t_peer * 1e3))
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("sharded x%d vs full batch: loss %.8f/%.8f smooth %.4e/%.4e -> %s" % (world, lp, flp, ls, fls, "OK" if flag.item() else "MISMATCH"))
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
