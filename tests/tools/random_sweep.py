"""Extended randomised parity sweep (not part of the pytest suite): CUDA path vs the C oracle on N random configurations --
shapes from 8x8 to 200x320 (TMA-eligible and ragged widths), 1..4 scales, batch 1..3, every padding mode, automask on/off,
ssim weights {0.85, 0.5, 0}, pose scales {0.01, 0.05, 0.2}, mask / no mask, shifted-copy sources, both backward kernels.
    python tests/tools/random_sweep.py [N] [seed]        -> summary on stdout (kept under profiles/)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import l2rel, maxrel, relerr
from mgnet_b200.synthetic import make_inputs
from oracle.oracle import Oracle
from test_gpu_parity import OR_KEYS, _run_cuda

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
dev = torch.device("cuda:0")
worst = dict(loss_p=0.0, loss_s=0.0, grad_l2=0.0, grad_max=0.0, pose_l2=0.0)
worst_pose_case = None
decisions = mism = 0
for k in range(N):
    big = k % 4 == 0
    H = int(rng.randint(40, 201)) if big else int(rng.randint(8, 65))
    W = int(rng.randint(16, 81)) * 4 if (big or k % 3 == 0) else int(rng.randint(8, 97))
    c = dict(B=int(rng.randint(1, 4)), H=H, W=W, n=int(rng.randint(1, 5)), seed=1000 + k, pad=["zeros", "border", "reflection"][k % 3],
             automask=bool(rng.randint(0, 2)), ssim=[0.85, 0.85, 0.5, 0.0][int(rng.randint(0, 4))], pose_scale=[0.01, 0.05, 0.2][int(rng.randint(0, 3))],
             with_mask=bool(rng.randint(0, 2)), shift=bool(rng.randint(0, 2)))
    if c["ssim"] == 0.0:
        c["with_mask"] = True          # the reference itself raises without a mask in the raw-L1 branch (loss.py:237-238)
    pred, tgt = make_inputs(c["B"], c["H"], c["W"], c["n"], seed=c["seed"], noise=0.0 if c["shift"] else 0.15, pose_scale=c["pose_scale"],
                            with_mask=c["with_mask"], shift_sources=c["shift"])
    hp = dict(ssim_loss_weight=c["ssim"], photometric_loss_weight=1.0, smoothing_loss_weight=1e-3, automask_loss=c["automask"],
              photometric_reduce_op="min", padding_mode=c["pad"])
    o = Oracle(pred, tgt, **{kk: hp[kk] for kk in OR_KEYS})
    f = o.forward(); g = o.backward(1.0, 1.0)
    for bw in ("stash", "recompute"):
        r = _run_cuda(pred, tgt, hp, dev, backward=bw)
        worst["loss_p"] = max(worst["loss_p"], relerr(r["loss_photometric"], f["loss_photometric"]))
        worst["loss_s"] = max(worst["loss_s"], relerr(r["loss_smoothness"], f["loss_smoothness"]))
        m = int((r["sel"] != f["sel"]).sum()); mism += m; decisions += r["sel"].size
        for i in range(c["n"]):
            worst["grad_l2"] = max(worst["grad_l2"], l2rel(r["grad_depth"][i], g["grad_depth"][i]))
            worst["grad_max"] = max(worst["grad_max"], maxrel(r["grad_depth"][i], g["grad_depth"][i]))
        if float(np.abs(g["grad_poses"]).max()) > 0:
            e = l2rel(r["grad_poses"], g["grad_poses"])
            if e > worst["pose_l2"]:
                worst["pose_l2"], worst_pose_case = e, (dict(c), bw, hp, r["grad_poses"].copy(), g["grad_poses"].copy())
        if m:
            print("MISMATCH", c, bw, m)
print("random sweep: %d configurations x 2 backward kernels, %d selection decisions, %d mismatches" % (N, decisions, mism))
print("worst relative errors vs the oracle: loss_photometric %.2e, loss_smoothness %.2e (bar 1e-5); depth gradients L2 %.2e, max-norm %.2e, pose gradients L2 %.2e (bar 1e-4)"
      % (worst["loss_p"], worst["loss_s"], worst["grad_l2"], worst["grad_max"], worst["pose_l2"]))
if worst_pose_case is not None:
    # how far is the REFERENCE's own fp32 autograd (the ATen port on the CPU, bit-identical to the reference) from the oracle's adjoint
    # (evaluated in double) on that configuration, and how far is the CUDA path from the reference?
    from oracle.torch_port import reference_loss
    c, bw, hp, got, want = worst_pose_case
    pred, tgt = make_inputs(c["B"], c["H"], c["W"], c["n"], seed=c["seed"], noise=0.0 if c["shift"] else 0.15, pose_scale=c["pose_scale"],
                            with_mask=c["with_mask"], shift_sources=c["shift"])
    inv = [d.clone().requires_grad_(True) for d in pred["depth"]]
    poses = pred["poses"].clone().requires_grad_(True)
    out = reference_loss({"depth": inv, "poses": poses}, tgt, ssim_loss_weight=c["ssim"], automask_loss=c["automask"], padding_mode=c["pad"])
    (out["loss_photometric"] + out["loss_smoothness"]).backward()
    ref = poses.grad.numpy()
    print("worst pose-gradient case %s (%s backward): CUDA vs oracle %.2e | reference fp32 autograd vs oracle %.2e | CUDA vs reference %.2e | max |grad| %.2e"
          % (c, bw, l2rel(got, want), l2rel(ref, want), l2rel(got, ref), float(np.abs(want).max())))
sys.exit(1 if mism or worst["loss_p"] > 1e-5 or worst["loss_s"] > 1e-5 or worst["grad_l2"] > 1e-4 or worst["pose_l2"] > 1e-4 else 0)
