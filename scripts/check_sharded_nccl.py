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

def run(p, t, group):
    mod = MultiViewPhotometricLoss(process_group=group, **HP)
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
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("sharded x%d vs full batch: loss %.8f/%.8f smooth %.4e/%.4e -> %s" % (world, lp, flp, ls, fls, "OK" if flag.item() else "MISMATCH"))
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
