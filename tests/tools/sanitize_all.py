"""Tiny invocations of every kernel family for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tests/tools/sanitize_all.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from mgnet_b200 import MultiViewPhotometricLoss
from mgnet_b200.postprocessing import dgc_rescale
from mgnet_b200.synthetic import make_dgc_inputs, make_inputs, quantize_images
from mgnet_b200.uncertainty import apply_uncertainty
dev = torch.device("cuda:0")

def run(hp, pred, tgt, **kw):
    mod = MultiViewPhotometricLoss(**hp, **kw)
    p = {"depth": [d.to(dev).requires_grad_(True) for d in pred["depth"]], "poses": pred["poses"].to(dev).requires_grad_(True)}
    t = {k: v.to(dev) for k, v in tgt.items()}
    out = mod(p, t)
    w = apply_uncertainty(out, torch.zeros(2, device=dev, requires_grad=True))
    (w["loss_photometric"] + w["loss_smoothness"]).backward()
    torch.cuda.synchronize()
    return float(out["loss_photometric"])

base = dict(ssim_loss_weight=0.85, photometric_loss_weight=1.0, smoothing_loss_weight=1e-3, automask_loss=True, photometric_reduce_op="min", padding_mode="zeros")
for (B, H, W, n) in ((2, 48, 128, 2), (1, 37, 75, 1), (1, 16, 64, 3)):        # TMA path, ragged (manual loader), exactly one tile
    pred, tgt = make_inputs(B, H, W, n, seed=3, pose_scale=0.05)
    for pad in ("zeros", "border", "reflection"):
        for bw in ("stash", "recompute"):
            print(B, H, W, n, pad, bw, run(dict(base, padding_mode=pad), pred, tgt, backward=bw))
    print("l1only", run(dict(base, ssim_loss_weight=0.0), pred, tgt))
    print("mean", run(dict(base, automask_loss=False, photometric_reduce_op="mean"), pred, tgt))
    tu, _ = quantize_images(tgt)
    print("uint8", run(base, pred, tu))
pred, tgt = make_inputs(2, 64, 128, 3, seed=4)
g = torch.Generator().manual_seed(1)
lows = {"depth": [(0.05 + 1.9 * torch.rand(2, 1, 64 // s, 128 // s, generator=g)) for s in (8, 16, 32)], "poses": pred["poses"]}
print("fused upsample", run(base, lows, tgt, fuse_upsample=True))
for (H, W, pan) in ((48, 160, True), (37, 75, True), (40, 64, False)):
    d = make_dgc_inputs(H, W, seed=2)
    dep = d["depth"].to(dev)
    pts, scale, cnt = dgc_rescale(dep, d["camera_matrix"].to(dev), d["real_camera_height"].to(dev), d["panoptic_seg"].to(dev) if pan else None,
                                  0 if pan else -1, [10000] if pan else [])
    torch.cuda.synchronize()
    print("dgc", H, W, pan, float(scale[0]), int(cnt[0]))
# round 2: pose-matrix input, bit-packed mask, PoseCNN tail, stand-alone geometry kernels with two cameras
from mgnet_b200.geometry import Camera, Pose, view_synthesis
from mgnet_b200.pose_tail import pose_tail
from mgnet_b200.synthetic import pack_mask
pred, tgt = make_inputs(2, 37, 75, 2, seed=6, pose_scale=0.05)
mats = torch.stack([Pose.from_vec(pred["poses"][:, s], "euler").mat for s in range(2)], 1)
print("pose matrices", run(base, {"depth": pred["depth"], "poses": mats}, tgt))
print("packed mask", run(base, pred, dict(tgt, reprojection_mask=pack_mask(tgt["reprojection_mask"]))))
feat = torch.randn(2, 12, 5, 9, device=dev, requires_grad=True)
pose_tail(feat).sum().backward()
K = tgt["camera_matrix"][:, :3, :3].contiguous().to(dev)
cam, ref_cam = Camera(K.clone()).to(dev), Camera(K.clone(), Tcw=Pose(mats[:, 0].to(dev))).scaled(0.9, 1.05)
depth = (1.0 / pred["depth"][0].clamp(min=1e-6)).to(dev)
for pad in ("zeros", "border", "reflection"):
    view_synthesis(tgt["image_prev_orig"].to(dev), depth, ref_cam, cam, padding_mode=pad, return_coords=True)
X = cam.reconstruct(depth, frame="c")
ref_cam.project(X, frame="w")
torch.cuda.synchronize()
print("geometry ops ok")
print("SANITIZE_DONE")
