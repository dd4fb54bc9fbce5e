"""CPU: pins the oracle (oracle/mgvs_oracle.c) against the golden fixtures made by the reference.

Bars (BASELINE.json): losses 1e-5 relative, gradients 1e-4 relative (L2 and max-norm), selection
mask and every forward intermediate bit-exact.
"""
import numpy as np
import pytest

from helpers import GRAD_RTOL, LOSS_RTOL, golden_names, l2rel, load_golden, maxrel, relerr
from oracle.oracle import Oracle

OR_KEYS = ("ssim_loss_weight", "photometric_loss_weight", "smoothing_loss_weight", "automask_loss", "padding_mode")


def _oracle(name):
    pred, tgt, hp, ref = load_golden(name)
    return Oracle(pred, tgt, **{k: hp[k] for k in OR_KEYS}), ref, len(pred["depth"])


@pytest.mark.parametrize("name", golden_names())
def test_forward_losses_and_selection(name):
    o, ref, n = _oracle(name)
    f = o.forward()
    assert relerr(f["loss_photometric"], ref["loss_photometric"]) <= LOSS_RTOL
    assert relerr(f["loss_smoothness"], ref["loss_smoothness"]) <= LOSS_RTOL
    for i in range(n):
        assert np.array_equal(f["sel"][i], ref["sel_%d" % i][:, 0]), "selection mask differs at scale %d" % i
    B = f["posemat"].shape[0]
    assert np.array_equal(f["posemat"].reshape(B, 2, 3, 4), ref["pose_mat"][:, :, :3, :4])


def test_forward_intermediates_bit_exact():
    o, ref, n = _oracle("kitti_small_mask")
    f = o.forward(dumps=True)
    for s in range(2):
        assert np.array_equal(f["coords"][0, s], ref["coords_0_%d" % s])
        assert np.array_equal(f["warped"][0, s], ref["warped_0_%d" % s])
        assert np.array_equal(f["photo"][0, s], ref["photo_0_%d" % s][:, 0])
        assert np.array_equal(f["identity"][s], ref["identity_%d" % s][:, 0])
    assert np.array_equal(f["minmap"][0], ref["minmap_0"][:, 0])


@pytest.mark.parametrize("name", golden_names())
def test_backward_matches_reference_autograd(name):
    o, ref, n = _oracle(name)
    o.forward()
    g = o.backward(1.0, 1.0)
    for i in range(n):
        assert l2rel(g["grad_depth"][i], ref["grad_depth_%d" % i]) <= GRAD_RTOL
        assert maxrel(g["grad_depth"][i], ref["grad_depth_%d" % i]) <= GRAD_RTOL
    assert l2rel(g["grad_poses"], ref["grad_poses"]) <= GRAD_RTOL
    assert maxrel(g["grad_poses"], ref["grad_poses"]) <= GRAD_RTOL


def test_backward_is_linear_in_upstream():
    o, ref, n = _oracle("grad_smooth_shift")
    o.forward()
    g10 = o.backward(1.0, 0.0)
    g01 = o.backward(0.0, 1.0)
    g23 = o.backward(2.0, 3.0)
    for i in range(n):
        comb = 2.0 * g10["grad_depth"][i].astype(np.float64) + 3.0 * g01["grad_depth"][i]
        assert l2rel(g23["grad_depth"][i], comb) <= 1e-6
    assert np.abs(g01["grad_poses"]).max() == 0.0   # smoothness does not depend on the poses
