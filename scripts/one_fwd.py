import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mgnet_b200 import MultiViewPhotometricLoss
from mgnet_b200.synthetic import make_inputs
dev = torch.device("cuda:0")
B, H, W, n = [int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (1, 32, 64, 2))]
pred, tgt = make_inputs(B, H, W, n, seed=5)
hp = dict(ssim_loss_weight=0.85, photometric_loss_weight=1.0, smoothing_loss_weight=1e-3, automask_loss=True, photometric_reduce_op="min", padding_mode="zeros")
mod = MultiViewPhotometricLoss(**hp)
p = {"depth": [d.to(dev).requires_grad_(True) for d in pred["depth"]], "poses": pred["poses"].to(dev).requires_grad_(True)}
t = {k: v.to(dev) for k, v in tgt.items()}
out = mod(p, t)
torch.cuda.synchronize()
print("fwd ok", out["loss_photometric"].item(), out["loss_smoothness"].item())
(out["loss_photometric"] + out["loss_smoothness"]).backward()
torch.cuda.synchronize()
print("bwd ok", p["poses"].grad.abs().sum().item())
