"""Verbose GPU parity report (diagnostics; the asserting version is tests/test_gpu_parity.py)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import golden_names, load_golden, l2rel, maxrel, relerr
from mgnet_b200 import MultiViewPhotometricLoss
from mgnet_b200.synthetic import make_inputs
from oracle.oracle import Oracle

dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0))

def run(pred, tgt, hp):
    mod = MultiViewPhotometricLoss(**hp)
    p = {"depth": [d.to(dev).requires_grad_(True) for d in pred["depth"]], "poses": pred["poses"].to(dev).requires_grad_(True)}
    t = {k: v.to(dev) for k, v in tgt.items()}
    out = mod(p, t)
    (out["loss_photometric"] + out["loss_smoothness"]).backward()
    torch.cuda.synchronize()
    return out["loss_photometric"].item(), out["loss_smoothness"].item(), mod.last_selection.cpu().numpy(), [d.grad.cpu().numpy() for d in p["depth"]], p["poses"].grad.cpu().numpy()

for name in golden_names():
    pred, tgt, hp, ref = load_golden(name)
    lp, ls, sel, gd, gp = run(pred, tgt, hp)
    n = len(pred["depth"])
    mism = [int((sel[i] != ref["sel_%d" % i][:, 0]).sum()) for i in range(n)]
    print("%-20s Lp rel %.1e Ls rel %.1e sel mism %s | gdepth L2 %s max %s | gpose L2 %.1e max %.1e" % (
        name, relerr(lp, ref["loss_photometric"]), relerr(ls, ref["loss_smoothness"]), mism,
        ["%.1e" % l2rel(gd[i], ref["grad_depth_%d" % i]) for i in range(n)],
        ["%.1e" % maxrel(gd[i], ref["grad_depth_%d" % i]) for i in range(n)],
        l2rel(gp, ref["grad_poses"]), maxrel(gp, ref["grad_poses"])))

hp = dict(ssim_loss_weight=0.85, photometric_loss_weight=1.0, smoothing_loss_weight=1e-3, automask_loss=True, photometric_reduce_op="min", padding_mode="zeros")
for (B, H, W, n, noise, shift) in [(2, 192, 640, 3, 0.2, False), (1, 96, 320, 4, 0.0, True), (3, 50, 70, 2, 0.2, False)]:
    pred, tgt = make_inputs(B, H, W, n, seed=11, noise=noise, shift_sources=shift)
    o = Oracle(pred, tgt); t0 = time.time(); f = o.forward(); g = o.backward(); to = time.time() - t0
    lp, ls, sel, gd, gp = run(pred, tgt, hp)
    print("oracle %s (%.2fs): Lp rel %.1e Ls rel %.1e sel mism %d | gdepth L2 %s max %s | gpose L2 %.1e" % (
        (B, H, W, n), to, relerr(lp, f["loss_photometric"]), relerr(ls, f["loss_smoothness"]), int((sel != f["sel"]).sum()),
        ["%.1e" % l2rel(gd[i], g["grad_depth"][i]) for i in range(n)], ["%.1e" % maxrel(gd[i], g["grad_depth"][i]) for i in range(n)],
        l2rel(gp, g["grad_poses"])))

# quick timing
for (B, H, W, n) in [(16, 192, 640, 3), (8, 512, 1024, 4)]:
    pred, tgt = make_inputs(B, H, W, n, seed=3)
    mod = MultiViewPhotometricLoss(**hp)
    p = {"depth": [d.to(dev).requires_grad_(True) for d in pred["depth"]], "poses": pred["poses"].to(dev).requires_grad_(True)}
    t = {k: v.to(dev) for k, v in tgt.items()}
    for it in range(3):
        out = mod(p, t); (out["loss_photometric"] + out["loss_smoothness"]).backward()
    torch.cuda.synchronize()
    e0, e1, e2 = torch.cuda.Event(True), torch.cuda.Event(True), torch.cuda.Event(True)
    K = 10
    tf = tb = 0.0
    for it in range(K):
        e0.record(); out = mod(p, t); e1.record(); (out["loss_photometric"] + out["loss_smoothness"]).backward(); e2.record()
        torch.cuda.synchronize(); tf += e0.elapsed_time(e1); tb += e1.elapsed_time(e2)
    px = B * H * W
    print("timing %s: fwd %.3f ms bwd %.3f ms -> %.2f Gpx/s fwd+bwd" % ((B, H, W, n), tf / K, tb / K, px / ((tf + tb) / K * 1e-3) / 1e9))
