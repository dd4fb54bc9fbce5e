"""BASELINE config[4] in miniature, under torchrun: a small depth + pose network wrapped in DDP, stepped with the drop-in
loss (batch-sharded, global mask counts through one NCCL all-reduce of 3n+3 doubles, ddp_grad_scale=True so that DDP's
1/G averaging yields the full-batch gradient).  Checks on every rank that the DDP-averaged parameter gradients equal the
single-process full-batch gradients, then times the step.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 scripts/train_step_ddp.py [H W B_per_gpu]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP
from mgnet_b200 import MultiViewPhotometricLoss
from mgnet_b200.sharding import batch_slice
from mgnet_b200.synthetic import make_inputs
from test_train_step_gpu import HP, TinyDepthNet, TinyPoseNet

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
H, W, Bg = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (512, 1024, 4)))
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
B = Bg * world
_, tgt = make_inputs(B, H, W, 3, seed=63, snap_trig=False)
torch.manual_seed(0)
nets = torch.nn.ModuleList([TinyDepthNet(), TinyPoseNet()]).to(dev)

def run(mods, loss, t, fuse):
    for p in mods.parameters(): p.grad = None
    m = mods.module if isinstance(mods, DDP) else mods
    # (forward through the DDP wrapper so that its gradient hooks fire)
    pred = mods(t, fuse)
    out = loss(pred, t)
    (out["loss_photometric"] + out["loss_smoothness"]).backward()
    return out

class Both(torch.nn.Module):
    def __init__(self, nets): super().__init__(); self.d, self.p = nets[0], nets[1]
    def forward(self, t, fuse): return {"depth": self.d(t["image_orig"], not fuse), "poses": self.p(t["image_orig"], t["image_prev_orig"], t["image_next_orig"])}

model = Both(nets)
full_t = {k: v.to(dev) for k, v in tgt.items()}
ref_loss = MultiViewPhotometricLoss(photometric_reduce_op="min", padding_mode="zeros", fuse_upsample=True, **HP)
o_full = run(model, ref_loss, full_t, True)                       # single process, full batch
g_full = [p.grad.detach().clone() for p in model.parameters()]
ddp = DDP(model, device_ids=[lr])
sl = batch_slice(B, world, rank)
my_t = {k: v[sl].contiguous() for k, v in full_t.items()}
from mgnet_b200.sharding import PeerExchange
exchange = PeerExchange(dist.group.WORLD) if os.environ.get("MGVS_EXCHANGE", "peer") == "peer" else None   # fused P2P exchange + finalize
loss = MultiViewPhotometricLoss(photometric_reduce_op="min", padding_mode="zeros", fuse_upsample=True, process_group=dist.group.WORLD,
                                exchange=exchange, ddp_grad_scale=True, **HP)
o = run(ddp, loss, my_t, True)
torch.cuda.synchronize()
num = sum(float((a.double() - b.grad.double()).pow(2).sum()) for a, b in zip(g_full, model.parameters()))
den = sum(float(a.double().pow(2).sum()) for a in g_full)
rel = (num / den) ** 0.5
lrel = abs(o["loss_photometric"].item() - o_full["loss_photometric"].item()) / abs(o_full["loss_photometric"].item())
ok = torch.tensor([1 if (rel <= 1e-4 and lrel <= 1e-6) else 0], device=dev); dist.all_reduce(ok, op=dist.ReduceOp.MIN)
for _ in range(3): run(ddp, loss, my_t, True)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 10; e0.record()
for _ in range(K): run(ddp, loss, my_t, True)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / K], device=dev); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
# where the step time goes (events on the compute stream): network forward | loss forward (incl. the exchange) | backward
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
for it in range(K):
    for p_ in ddp.parameters(): p_.grad = None
    ev[it][0].record(); pred = ddp(my_t, True)
    ev[it][1].record(); out = loss(pred, my_t)
    ev[it][2].record(); (out["loss_photometric"] + out["loss_smoothness"]).backward()
    ev[it][3].record()
torch.cuda.synchronize()
parts = [sum(e[j].elapsed_time(e[j + 1]) for e in ev) / K for j in range(3)]
if rank == 0:
    print("DDP_PARTS net_fwd %.3f loss_fwd %.3f backward %.3f ms (%s)" % (parts[0], parts[1], parts[2], "peer" if exchange is not None else "nccl"))
if rank == 0:
    print("DDP_STEP " + json.dumps({"gpus": world, "B_per_gpu": Bg, "H": H, "W": W, "step_ms": ms.item(), "images_per_s": B / (ms.item() * 1e-3),
                                   "exchange": "peer" if exchange is not None else "nccl", "ddp_grad_vs_full_batch_l2rel": rel, "loss_rel": lrel, "parity": "OK" if ok.item() else "MISMATCH"}))
dist.destroy_process_group()
sys.exit(0 if ok.item() else 1)
