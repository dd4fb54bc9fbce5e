"""CPU: the ATen-level port (oracle/torch_port.py) against (a) the golden fixtures and (b) the live
reference when /root/reference is mounted (bit-identical losses, selection and gradients expected --
same operator sequence on the same host)."""
import os
import sys

import numpy as np
import pytest
import torch

from helpers import GRAD_RTOL, LOSS_RTOL, golden_names, l2rel, load_golden, relerr
from oracle.torch_port import reference_loss

PORT_KEYS = ("ssim_loss_weight", "photometric_loss_weight", "smoothing_loss_weight", "automask_loss", "padding_mode")


def _run_port(pred, tgt, hp):
    inv = [d.clone().requires_grad_(True) for d in pred["depth"]]
    poses = pred["poses"].clone().requires_grad_(True)
    out = reference_loss({"depth": inv, "poses": poses}, tgt, return_selection=True, **{k: hp[k] for k in PORT_KEYS})
    (out["loss_photometric"] + out["loss_smoothness"]).backward()
    return out, [d.grad.numpy() for d in inv], poses.grad.numpy()


@pytest.mark.parametrize("name", golden_names())
def test_port_matches_golden(name):
    pred, tgt, hp, ref = load_golden(name)
    out, gd, gp = _run_port(pred, tgt, hp)
    assert relerr(out["loss_photometric"].item(), ref["loss_photometric"]) <= LOSS_RTOL
    assert relerr(out["loss_smoothness"].item(), ref["loss_smoothness"]) <= LOSS_RTOL
    for i in range(len(gd)):
        assert np.array_equal(out["selection"][i].numpy(), ref["sel_%d" % i][:, 0])
        assert l2rel(gd[i], ref["grad_depth_%d" % i]) <= GRAD_RTOL
    assert l2rel(gp, ref["grad_poses"]) <= GRAD_RTOL


def test_port_is_bit_identical_to_live_reference():
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import ref_loader
    if not ref_loader.available():
        pytest.skip("/root/reference not mounted")
    from mgnet_b200.synthetic import make_inputs
    pred, tgt = make_inputs(2, 64, 96, 3, seed=21)
    hp = dict(ref_loader.DEFAULT_HP)
    ref = ref_loader.run_reference(pred, tgt, hp=hp, want_intermediates=False)
    out, gd, gp = _run_port(pred, tgt, hp)
    assert out["loss_photometric"].item() == float(ref["loss_photometric"])
    assert out["loss_smoothness"].item() == float(ref["loss_smoothness"])
    for i in range(3):
        assert np.array_equal(gd[i], ref["grad_depth_%d" % i])
    assert np.array_equal(gp, ref["grad_poses"])
