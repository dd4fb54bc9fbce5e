"""Shared helpers for the parity tests: golden-fixture loading and error metrics."""
from __future__ import annotations

import glob
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# Parity bars from BASELINE.json north_star
LOSS_RTOL = 1e-5
GRAD_RTOL = 1e-4


def golden_names():
    """Full-resolution fixtures (the reference's contract).  "fused_*" fixtures hold low-resolution maps."""
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
    return [n for n in names if not n.startswith(("fused_", "dgc_", "mean_"))]


def load_golden(name):
    """Returns (predictions, targets, hyper-parameters, reference outputs) with CPU tensors."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))

    def image(key):
        # compact fixtures (make_golden_c1.py) hold the data loader's uint8; the loss sees u8.float() / 255 (mg_net.py:320-335)
        if "in_u8_" + key in z.files:
            return (torch.from_numpy(z["in_u8_" + key]).float() / 255.0).contiguous()
        return torch.from_numpy(z["in_" + key])

    def depth(i):
        if "in_f16_depth_%d" % i in z.files:       # fp16-representable values, stored as fp16
            return torch.from_numpy(z["in_f16_depth_%d" % i]).float().contiguous()
        return torch.from_numpy(z["in_depth_%d" % i])

    n = len([k for k in z.files if k.startswith(("in_depth_", "in_f16_depth_"))])
    pred = {
        "depth": [depth(i) for i in range(n)],
        "poses": torch.from_numpy(z["in_poses"]),
    }
    tgt = {
        "image_orig": image("image_orig"),
        "image_prev_orig": image("image_prev_orig"),
        "image_next_orig": image("image_next_orig"),
        "camera_matrix": torch.from_numpy(z["in_camera_matrix"]),
    }
    if "in_reprojection_mask" in z.files:
        tgt["reprojection_mask"] = torch.from_numpy(z["in_reprojection_mask"])
    hp = dict(
        ssim_loss_weight=float(z["hp_ssim_loss_weight"]),
        photometric_loss_weight=float(z["hp_photometric_loss_weight"]),
        smoothing_loss_weight=float(z["hp_smoothing_loss_weight"]),
        automask_loss=bool(z["hp_automask_loss"]),
        photometric_reduce_op=str(z["hp_photometric_reduce_op"]) if "hp_photometric_reduce_op" in z.files else "min",
        padding_mode=str(z["hp_padding_mode"]) if "hp_padding_mode" in z.files else "zeros",
    )
    ref = {k: z[k] for k in z.files if not k.startswith("in_") and not k.startswith("hp_")}
    return pred, tgt, hp, ref


def l2rel(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def maxrel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def relerr(a, b):
    return abs(float(a) - float(b)) / max(abs(float(b)), 1e-300)
