"""BASELINE config[3] on ONE GPU: the full 64-image 1024x2048 batch (6.6 GB of inputs, 19 GB of coefficient stash) in the
B200's 180 GB.  Size-independent property: a batch made of 8 distinct images repeated 8 times has the loss of the 8 and, per
image, 1/8 of their gradients; the selection maps repeat.  Exercises every >2^31-byte offset of the kernels."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mgnet_b200 import MultiViewPhotometricLoss
from mgnet_b200.synthetic import make_inputs
HP = dict(ssim_loss_weight=0.85, photometric_loss_weight=1.0, smoothing_loss_weight=1e-3, automask_loss=True, photometric_reduce_op="min", padding_mode="zeros")
dev = torch.device("cuda:0")
H, W, n, base, rep = 1024, 2048, 3, 8, 8
pred, tgt = make_inputs(base, H, W, n, seed=5)

def run(p, t):
    mod = MultiViewPhotometricLoss(**HP)
    pd = {"depth": [d.to(dev).requires_grad_(True) for d in p["depth"]], "poses": p["poses"].to(dev).requires_grad_(True)}
    td = {k: v.to(dev) for k, v in t.items()}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = mod(pd, td)
    (out["loss_photometric"] + out["loss_smoothness"]).backward()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return out["loss_photometric"].item(), out["loss_smoothness"].item(), [d.grad for d in pd["depth"]], pd["poses"].grad, mod.last_selection, dt

lp, ls, gd, gp, sel, _ = run(pred, tgt)
big_p = {"depth": [d.repeat(rep, 1, 1, 1) for d in pred["depth"]], "poses": pred["poses"].repeat(rep, 1, 1)}
big_t = {k: v.repeat(rep, *([1] * (v.dim() - 1))) for k, v in tgt.items()}
blp, bls, bgd, bgp, bsel, dt = run(big_p, big_t)
ok = abs(blp - lp) <= 1e-6 * abs(lp) and abs(bls - ls) <= 1e-6 * abs(ls)
for i in range(n):
    for r in (0, rep - 1):
        a, b = bgd[i][r * base:(r + 1) * base].double() * rep, gd[i].double()
        ok = ok and float((a - b).norm() / b.norm()) <= 1e-6
ok = ok and float((bgp[-base:].double() * rep - gp.double()).norm() / gp.double().norm()) <= 1e-6
ok = ok and bool((bsel[:, -base:] == sel).all()) and bool((bsel[:, :base] == sel).all())
print("B=%d %dx%d on one GPU: loss %.8f vs %.8f (B=%d), peak memory %.1f GB, first call %.1f ms -> %s"
      % (base * rep, H, W, blp, lp, base, torch.cuda.max_memory_allocated() / 2**30, dt * 1e3, "OK" if ok else "MISMATCH"))
sys.exit(0 if ok else 1)
