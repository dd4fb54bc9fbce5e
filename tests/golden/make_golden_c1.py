"""Generates tests/golden/c1_kitti_192x640.npz: BASELINE.json config[0] ("the parity gate") run through the
UNMODIFIED reference (/root/reference via tests/golden/ref_loader.py) -- B1, 192x640, 2 source frames, n=3 and n=1, so that at
least one reference-made fixture has interior tiles (12 x 10 tiles of 64x16).

Run in the build container only:   python tests/golden/make_golden_c1.py

Compact encoding (tests/helpers.load_golden decodes it): images are stored as the data loader's uint8 (`in_u8_*`; the loss
sees exactly `u8.float() / 255`, mg_net.py:320-335) and the inverse-depth maps as fp16 (`in_f16_*`; every value is exactly
representable, the loss sees `.float()`).  Three result sets share those inputs:
  (default keys)  n=3, poses snapped to angles whose MKL sin/cos is correctly rounded (synthetic.snap_pose_trig)
  n1_*            n=1 (first map only)
  us_*            n=3 with UN-snapped poses `us_in_poses`: the reference's own rotation matrices `us_pose_mat` differ from the
                  correctly rounded ones by 1 ulp in some entries -- tests/test_unsnapped_poses.py reports the selection
                  mismatches of the Euler path and checks that the pose-matrix input reproduces the reference bit for bit
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
from mgnet_b200.synthetic import make_inputs, quantize_images  # noqa: E402


def main():
    torch.set_num_threads(1)
    B, H, W, n = 1, 192, 640, 3
    pred, tgt = make_inputs(B, H, W, n, seed=41, noise=0.2, snap_trig=True)
    tu, tf = quantize_images(tgt)
    depth16 = [d.half() for d in pred["depth"]]
    pred = {"depth": [d.float() for d in depth16], "poses": pred["poses"]}
    hp = dict(ref_loader.DEFAULT_HP)
    out = {
        "in_camera_matrix": tgt["camera_matrix"].numpy(), "in_poses": pred["poses"].numpy(),
        "in_reprojection_mask": tgt["reprojection_mask"].numpy(),
        "hp_ssim_loss_weight": np.float64(hp["ssim_loss_weight"]), "hp_photometric_loss_weight": np.float64(hp["photometric_loss_weight"]),
        "hp_smoothing_loss_weight": np.float64(hp["smoothing_loss_weight"]), "hp_automask_loss": np.bool_(hp["automask_loss"]),
        "hp_padding_mode": np.array(hp["padding_mode"]), "hp_photometric_reduce_op": np.array(hp["photometric_reduce_op"]),
    }
    for k in ("image_orig", "image_prev_orig", "image_next_orig"):
        out["in_u8_" + k] = tu[k].numpy()
    for i, d in enumerate(depth16):
        out["in_f16_depth_%d" % i] = d.numpy()

    def run(p, prefix, n_):
        res = ref_loader.run_reference({"depth": p["depth"][:n_], "poses": p["poses"]}, tf, hp=hp, want_grads=True, want_intermediates=True)
        out[prefix + "loss_photometric"] = res["loss_photometric"]
        out[prefix + "loss_smoothness"] = res["loss_smoothness"]
        out[prefix + "grad_poses"] = res["grad_poses"]
        out[prefix + "pose_mat"] = np.stack([res["pose_mat_0"], res["pose_mat_1"]], 1)
        for i in range(n_):
            out[prefix + "sel_%d" % i] = res["sel_%d" % i]
            out[prefix + "grad_depth_%d" % i] = res["grad_depth_%d" % i]
        print("%-4s n=%d Lp=%.8f Ls=%.8e" % (prefix or "n3", n_, float(res["loss_photometric"]), float(res["loss_smoothness"])))
        return res

    run(pred, "", 3)
    run(pred, "n1_", 1)
    raw, _ = make_inputs(B, H, W, n, seed=41, noise=0.2, snap_trig=False)
    # several un-snapped pose draws so that some angle certainly hits an MKL-vs-correctly-rounded difference
    g = torch.Generator().manual_seed(4141)
    cand = [raw["poses"]] + [(0.01 * torch.randn(B, 2, 6, generator=g)).contiguous() for _ in range(255)]
    chosen, most = None, 0
    for c in cand:
        nbad = 0
        for s in range(2):      # same slicing as the reference (Pose.from_vec(poses[:, s].float()) -> euler2mat(vec[:, 3:]))
            rot = c[:, s].float()[:, 3:]
            for k in range(3):
                a = rot[:, k]
                nbad += int(((torch.cos(a) != torch.cos(a.double()).float()) | (torch.sin(a) != torch.sin(a.double()).float())).sum())
        if nbad > most:
            chosen, most = c, nbad
    print("un-snapped draw with %d of 6 angles whose MKL sin/cos differs from the correctly rounded value" % most)
    assert chosen is not None, "no pose draw with an MKL / correctly-rounded trig difference"
    out["us_in_poses"] = chosen.numpy()
    res = run({"depth": pred["depth"], "poses": chosen}, "us_", 3)
    for i in range(3):
        del out["us_grad_depth_%d" % i]          # the gradient bars are covered by the snapped set; keep the file small
    path = os.path.join(HERE, "c1_kitti_192x640.npz")
    np.savez_compressed(path, **out)
    print("c1_kitti_192x640 %.1f KB" % (os.path.getsize(path) / 1024.0))


if __name__ == "__main__":
    main()
