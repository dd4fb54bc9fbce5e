"""SURVEY 8f-1 measurement: loss fwd+bwd from the depth head's LOW-resolution maps, (a) the reference's way -- F.interpolate
(bilinear, align_corners=True) to full resolution in PyTorch, then the (unfused) drop-in loss, autograd back through the
interpolate -- vs (b) fuse_upsample=True.  Usage: python scripts/time_fused_upsample.py [c2 c4]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
from mgnet_b200 import MultiViewPhotometricLoss
from mgnet_b200.synthetic import make_inputs
W = {"c2": (16, 192, 640), "c3": (8, 512, 1024), "c4": (8, 1024, 2048)}
HP = dict(ssim_loss_weight=0.85, photometric_loss_weight=1.0, smoothing_loss_weight=1e-3, automask_loss=True, photometric_reduce_op="min", padding_mode="zeros")
STR = (8, 16, 32)
dev = torch.device("cuda:0")
out = {}
for name in (sys.argv[1:] or ["c2", "c4"]):
    B, H, Wd = W[name]
    _, tgt = make_inputs(B, H, Wd, 3, seed=5)
    t = {k: v.to(dev) for k, v in tgt.items()}
    g = torch.Generator().manual_seed(7)
    lows = [(0.05 + 1.9 * torch.rand(B, 1, H // s, Wd // s, generator=g)).to(dev).requires_grad_(True) for s in STR]
    poses = (0.01 * torch.randn(B, 2, 6, generator=g)).to(dev).requires_grad_(True)
    plain, fused = MultiViewPhotometricLoss(**HP), MultiViewPhotometricLoss(fuse_upsample=True, **HP)
    def step_a():
        fulls = [F.interpolate(x, scale_factor=s, mode="bilinear", align_corners=True) for x, s in zip(lows, STR)]
        o = plain({"depth": fulls, "poses": poses}, t); (o["loss_photometric"] + o["loss_smoothness"]).backward(); return o
    def step_b():
        o = fused({"depth": lows, "poses": poses}, t); (o["loss_photometric"] + o["loss_smoothness"]).backward(); return o
    res = {}
    for nm, fn in (("interpolate+loss", step_a), ("fused", step_b)):
        for _ in range(5): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 20
        e0.record()
        for _ in range(K):
            for x in lows + [poses]: x.grad = None
            o = fn()
        e1.record(); torch.cuda.synchronize()
        res[nm] = {"ms": e0.elapsed_time(e1) / K, "loss": o["loss_photometric"].item()}
    out[name] = res
    print(name, json.dumps(res))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "fused_upsample_timing.json"), "w"), indent=1)
