"""Generates the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference, imported through tests/golden/ref_loader.py) on seeded synthetic inputs.

Run in the build container only:   python tests/golden/make_golden.py
The GPU box has no /root/reference; the committed .npz files are what travels.

Every fixture stores the inputs (so nothing depends on re-generating them bit-identically on another
host), the hyper-parameters, and the reference outputs: the two losses, autograd gradients of
(loss_photometric + loss_smoothness) w.r.t. every inverse-depth map and the pose vectors, the
per-scale argmin selection, and -- for the first case -- scale-0 intermediates (coordinates, warped
image, per-pixel photometric maps) obtained by calling the reference's own methods.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_loader  # noqa: E402
from mgnet_b200.synthetic import make_inputs, snap_pose_trig  # noqa: E402

CASES = {
    # name: (make_inputs kwargs, hyper-parameter overrides, keep scale-0 intermediates)
    "kitti_small_mask": (dict(B=2, H=48, W=64, n=3, seed=1, noise=0.2), {}, True),
    "grad_smooth_shift": (dict(B=2, H=48, W=64, n=3, seed=2, noise=0.0, shift_sources=True), {}, False),
    "nomask_n1": (dict(B=1, H=32, W=96, n=1, seed=3, noise=0.2, with_mask=False), {}, False),
    "ragged_n4_bigpose": (dict(B=2, H=40, W=72, n=4, seed=4, noise=0.1, pose_scale=0.05), {}, False),
    "automask_off": (dict(B=1, H=32, W=64, n=2, seed=5, noise=0.2), {"automask_loss": False}, False),
    "weights_alpha": (dict(B=1, H=36, W=68, n=2, seed=6, noise=0.0, shift_sources=True),
                      {"ssim_loss_weight": 0.5, "photometric_loss_weight": 2.0, "smoothing_loss_weight": 0.05}, False),
    # photometric_reduce_op="mean" (loss.py:242-243; automask must be off, :106-109).  File name starts with "mean_" so the
    # selection-mask tests skip it (there is no argmin under "mean")
    "mean_reduce": (dict(B=2, H=40, W=72, n=2, seed=12, noise=0.0, shift_sources=True),
                    {"automask_loss": False, "photometric_reduce_op": "mean"}, False),
    # ssim_loss_weight == 0: raw 3-channel L1, min over 3 channels per list entry (loss.py:195-196)
    "l1_only": (dict(B=2, H=40, W=72, n=2, seed=13, noise=0.1, shift_sources=True), {"ssim_loss_weight": 0.0}, False),
    "l1_only_noautomask": (dict(B=1, H=36, W=68, n=2, seed=14, noise=0.0, shift_sources=True),
                           {"ssim_loss_weight": 0.0, "automask_loss": False}, False),
    # grid_sample padding modes (camera_utils.py:52-54; config.py:116-117): large poses push many samples outside the image
    "pad_border": (dict(B=2, H=40, W=72, n=2, seed=9, noise=0.1, pose_scale=0.06), {"padding_mode": "border"}, True),
    "pad_reflection": (dict(B=2, H=40, W=72, n=2, seed=10, noise=0.0, pose_scale=0.06, shift_sources=True),
                       {"padding_mode": "reflection"}, True),
}


def adversarial(pred, tgt):
    """ragged_n4_bigpose: points behind the camera, fully out-of-bounds warps, clamped inverse depth,
    an all-false mask row block."""
    poses = pred["poses"].clone()
    poses[0, 0] = torch.tensor([0.3, -0.1, -4.0, 0.02, 0.4, -0.03])   # tz=-4: most points behind the camera
    poses[1, 1] = torch.tensor([30.0, 2.0, 0.5, 0.0, 0.0, 0.7])       # far out of bounds
    pred["poses"] = snap_pose_trig(poses)
    pred["depth"][1][0, 0, :5, :7] = 0.0          # clamp(min=1e-6) active
    pred["depth"][2][1, 0, 10:12, :] = 1e-7
    tgt["reprojection_mask"][0, 0, 20:30, :] = False
    tgt["reprojection_mask"][1, 0, :, 60:] = False
    return pred, tgt


def far_out(pred):
    """pad_*: one source far outside the image (several reflections), one behind the camera."""
    poses = pred["poses"].clone()
    poses[1, 1] = torch.tensor([3.0, 0.5, 0.2, 0.0, 0.02, 0.3])
    poses[0, 0] = torch.tensor([-0.8, 0.1, -0.3, 0.05, -0.2, 0.01])
    pred["poses"] = snap_pose_trig(poses)
    return pred


def main(only=None):
    for name, (kw, hp_over, keep) in CASES.items():
        if only and name not in only:
            continue
        pred, tgt = make_inputs(**kw)
        if name == "ragged_n4_bigpose":
            pred, tgt = adversarial(pred, tgt)
        if name.startswith("pad_"):
            pred = far_out(pred)
        hp = dict(ref_loader.DEFAULT_HP)
        hp.update(hp_over)
        res = ref_loader.run_reference(pred, tgt, hp=hp, want_grads=True, want_intermediates=True)
        n = len(pred["depth"])
        out = {
            "in_image_orig": tgt["image_orig"].numpy(),
            "in_image_prev_orig": tgt["image_prev_orig"].numpy(),
            "in_image_next_orig": tgt["image_next_orig"].numpy(),
            "in_camera_matrix": tgt["camera_matrix"].numpy(),
            "in_poses": pred["poses"].numpy(),
            "hp_ssim_loss_weight": np.float64(hp["ssim_loss_weight"]),
            "hp_photometric_loss_weight": np.float64(hp["photometric_loss_weight"]),
            "hp_smoothing_loss_weight": np.float64(hp["smoothing_loss_weight"]),
            "hp_automask_loss": np.bool_(hp["automask_loss"]),
            "hp_padding_mode": np.array(hp["padding_mode"]),
            "hp_photometric_reduce_op": np.array(hp["photometric_reduce_op"]),
            "loss_photometric": res["loss_photometric"],
            "loss_smoothness": res["loss_smoothness"],
            "grad_poses": res["grad_poses"],
            "pose_mat": np.stack([res["pose_mat_0"], res["pose_mat_1"]], 1),
        }
        if "reprojection_mask" in tgt:
            out["in_reprojection_mask"] = tgt["reprojection_mask"].numpy()
        for i in range(n):
            out["in_depth_%d" % i] = pred["depth"][i].numpy()
            out["grad_depth_%d" % i] = res["grad_depth_%d" % i]
            out["sel_%d" % i] = res["sel_%d" % i]
        if keep:
            for s in range(2):
                for k in ("coords", "warped", "photo"):
                    out["%s_0_%d" % (k, s)] = res["%s_0_%d" % (k, s)]
                out["identity_%d" % s] = res["identity_%d" % s]
            out["minmap_0"] = res["minmap_0"]
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print("%-20s %7.1f KB  Lp=%.8f Ls=%.8e" % (name, os.path.getsize(path) / 1024.0,
                                                  float(res["loss_photometric"]), float(res["loss_smoothness"])))


def make_fused_upsample():
    """SURVEY 8f-1 fixture: the depth head's low-resolution maps (strides 8/16/32), upsampled with the head's own
    F.interpolate(scale_factor=stride, mode="bilinear", align_corners=True) (mg_net.py:803-806) and fed to the
    unmodified reference loss; gradients are chained back through the interpolate to the low-resolution maps.
    File name starts with "fused_" so the full-resolution tests skip it."""
    import torch.nn.functional as F
    B, H, W, strides = 2, 64, 128, (8, 16, 32)
    pred, tgt = make_inputs(B=B, H=H, W=W, n=len(strides), seed=8, noise=0.2)
    g = torch.Generator(device="cpu").manual_seed(8080)
    lows = [(0.05 + 1.9 * torch.rand(B, 1, H // s, W // s, generator=g)).contiguous() for s in strides]
    leaves = [lo.clone().requires_grad_(True) for lo in lows]
    fulls = [F.interpolate(x, scale_factor=s, mode="bilinear", align_corners=True) for x, s in zip(leaves, strides)]
    hp = dict(ref_loader.DEFAULT_HP)
    res = ref_loader.run_reference({"depth": [f.detach() for f in fulls], "poses": pred["poses"]}, tgt, hp=hp,
                                   want_grads=True, want_intermediates=True)
    out = {
        "in_image_orig": tgt["image_orig"].numpy(), "in_image_prev_orig": tgt["image_prev_orig"].numpy(),
        "in_image_next_orig": tgt["image_next_orig"].numpy(), "in_camera_matrix": tgt["camera_matrix"].numpy(),
        "in_poses": pred["poses"].numpy(), "in_reprojection_mask": tgt["reprojection_mask"].numpy(),
        "hp_ssim_loss_weight": np.float64(hp["ssim_loss_weight"]), "hp_photometric_loss_weight": np.float64(hp["photometric_loss_weight"]),
        "hp_smoothing_loss_weight": np.float64(hp["smoothing_loss_weight"]), "hp_automask_loss": np.bool_(hp["automask_loss"]),
        "loss_photometric": res["loss_photometric"], "loss_smoothness": res["loss_smoothness"], "grad_poses": res["grad_poses"],
    }
    for i, (leaf, full) in enumerate(zip(leaves, fulls)):
        full.backward(torch.from_numpy(res["grad_depth_%d" % i]))
        out["in_depth_%d" % i] = lows[i].numpy()                 # LOW resolution
        out["grad_depth_%d" % i] = leaf.grad.numpy()             # LOW resolution
        out["sel_%d" % i] = res["sel_%d" % i]
    path = os.path.join(HERE, "fused_upsample_n3.npz")
    np.savez_compressed(path, **out)
    print("%-20s %7.1f KB  Lp=%.8f Ls=%.8e" % ("fused_upsample_n3", os.path.getsize(path) / 1024.0,
                                              float(res["loss_photometric"]), float(res["loss_smoothness"])))


if __name__ == "__main__":
    torch.set_num_threads(1)   # the reference on CPU is thread-count independent; be safe
    if len(sys.argv) > 1 and sys.argv[1] == "fused":
        make_fused_upsample()      # leaves the other (committed) fixtures untouched
    elif len(sys.argv) > 1:
        main(only=sys.argv[1:])    # e.g. `make_golden.py pad_border pad_reflection`
    else:
        main()
        make_fused_upsample()
