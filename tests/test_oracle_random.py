"""Randomised pinning of the C oracle (oracle/mgvs_oracle.c) against the ATen-level port (oracle/torch_port.py, itself bit-identical
to the unmodified reference -- tests/test_torch_port.py): random ragged shapes, 1..3 scales, every grid_sample padding mode, automask
on/off, ssim_loss_weight in {0.85, 0.5, 0 (raw 3-channel L1)}, small and large poses.  CPU only; runs on the GPU box as well."""
import numpy as np
import pytest
import torch

from helpers import GRAD_RTOL, LOSS_RTOL, l2rel, relerr
from mgnet_b200.synthetic import make_inputs
from oracle.oracle import Oracle
from oracle.torch_port import reference_loss

CASES = []
_rng = np.random.RandomState(2024)
for _k in range(12):
    CASES.append(dict(
        B=int(_rng.randint(1, 3)), H=int(_rng.randint(8, 41)), W=int(_rng.randint(8, 73)), n=int(_rng.randint(1, 4)),
        seed=100 + _k, pad=["zeros", "border", "reflection"][_k % 3], automask=bool(_rng.randint(0, 2)),
        ssim=[0.85, 0.5, 0.0][int(_rng.randint(0, 3))], pose_scale=[0.01, 0.05, 0.2][int(_rng.randint(0, 3))],
        with_mask=bool(_rng.randint(0, 2)), shift=bool(_rng.randint(0, 2))))
    # the reference itself raises for ssim_loss_weight == 0 without a reprojection mask (its default mask takes the 3-channel shape
    # of the raw L1 map, loss.py:237-238, and cannot index the 1-channel min): only the masked form is a valid reference configuration
    if CASES[-1]["ssim"] == 0.0:
        CASES[-1]["with_mask"] = True


@pytest.mark.parametrize("c", CASES, ids=lambda c: "%dx%dx%d_n%d_%s_am%d_a%.2f_p%.2f" % (c["B"], c["H"], c["W"], c["n"], c["pad"], c["automask"], c["ssim"], c["pose_scale"]))
def test_oracle_matches_aten_port(c):
    torch.set_num_threads(1)
    pred, tgt = make_inputs(c["B"], c["H"], c["W"], c["n"], seed=c["seed"], noise=0.0 if c["shift"] else 0.15,
                            pose_scale=c["pose_scale"], with_mask=c["with_mask"], shift_sources=c["shift"])
    inv = [d.clone().requires_grad_(True) for d in pred["depth"]]
    poses = pred["poses"].clone().requires_grad_(True)
    out = reference_loss({"depth": inv, "poses": poses}, tgt, ssim_loss_weight=c["ssim"], automask_loss=c["automask"],
                         padding_mode=c["pad"], return_selection=True)
    (out["loss_photometric"] + out["loss_smoothness"]).backward()
    o = Oracle(pred, tgt, ssim_loss_weight=c["ssim"], automask_loss=c["automask"], padding_mode=c["pad"])
    f = o.forward()
    g = o.backward(1.0, 1.0)
    assert relerr(f["loss_photometric"], out["loss_photometric"].item()) <= LOSS_RTOL
    assert relerr(f["loss_smoothness"], out["loss_smoothness"].item()) <= LOSS_RTOL
    sel = out["selection"].numpy()          # [n,B,H,W] uint8
    assert int((f["sel"] != sel).sum()) == 0
    for i in range(c["n"]):
        assert l2rel(g["grad_depth"][i], inv[i].grad.numpy()) <= GRAD_RTOL
    if float(poses.grad.abs().max()) > 0:
        assert l2rel(g["grad_poses"], poses.grad.numpy()) <= GRAD_RTOL
