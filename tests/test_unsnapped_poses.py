"""Un-snapped Euler angles and the pose-matrix input (VERDICT r01 weak #1b; SURVEY 8b "pose [B,S,6] or [B,S,3,4]").

torch-CPU evaluates sin/cos with MKL-VML, which is 1 ulp off the correctly rounded value for a few percent of small
arguments; the kernels (and the oracle) use the correctly rounded value.  Every other parity input is "snapped" to angles
where the two agree (synthetic.snap_pose_trig).  Here the poses are NOT snapped (fixture set `us_*` of
tests/golden/c1_kitti_192x640.npz, made by the unmodified reference at BASELINE config[0]'s shape):

 * Euler input: the rotation matrices differ from the reference's in the last bit of some entries, so a handful of
   near-tie pixels may select another candidate.  The test REPORTS the mismatch count and the margin (second-best minus
   best loss) of every flipped pixel -- profiles/r02_unsnapped_selection.json -- and bounds both.
 * pose-matrix input (MgvsProblem.pose_mats / predictions["poses"] of shape [B,S,4,4]): the caller hands over the matrices the
   reference itself built (Pose.from_vec -> pose_vec2mat) and the selection is bit-exact unconditionally.
"""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GRAD_RTOL, LOSS_RTOL, ROOT, l2rel, load_golden, maxrel, relerr
from oracle.oracle import Oracle

NAME = "c1_kitti_192x640"
MAX_FLIPS = 64            # of 3 x 122 880 decisions
MAX_MARGIN = 2e-5         # a flipped pixel must be a near-tie: |second best - best| of the reference's own loss maps


def _us():
    pred, tgt, hp, ref = load_golden(NAME)
    poses = torch.from_numpy(ref["us_in_poses"])
    mats = torch.from_numpy(ref["us_pose_mat"])          # [B,S,4,4] as built by the reference
    return pred, tgt, hp, ref, poses, mats


def _margins(f_mat, ref, n):
    """Per scale: (second best - best) over [warp_prev, id_prev, warp_next, id_next] from the bit-exact maps of the matrix-input run."""
    out = []
    for i in range(n):
        stack = np.stack([f_mat["photo"][i, 0], f_mat["identity"][0], f_mat["photo"][i, 1], f_mat["identity"][1]], 0).astype(np.float64)
        srt = np.sort(stack, 0)
        out.append(srt[1] - srt[0])
    return out


def test_oracle_unsnapped_euler_report_and_matrix_input_exact():
    pred, tgt, hp, ref, poses, mats = _us()
    n = len(pred["depth"])
    # (a) matrix input: the reference's own R|t -> bit-exact selection, exact rotation pass-through
    om = Oracle({"depth": pred["depth"], "poses": mats}, tgt)
    fm = om.forward(dumps=True)
    for i in range(n):
        assert np.array_equal(fm["sel"][i], ref["us_sel_%d" % i][:, 0]), "matrix input: selection differs at scale %d" % i
    assert np.array_equal(fm["posemat"].reshape(-1, 2, 3, 4), ref["us_pose_mat"][:, :, :3, :4])
    assert relerr(fm["loss_photometric"], ref["us_loss_photometric"]) <= LOSS_RTOL
    assert relerr(fm["loss_smoothness"], ref["us_loss_smoothness"]) <= LOSS_RTOL
    # (b) Euler input with correctly rounded trig: report
    oe = Oracle({"depth": pred["depth"], "poses": poses}, tgt)
    fe = oe.forward()
    ulp_diff = int((fe["posemat"].reshape(-1, 2, 3, 4) != ref["us_pose_mat"][:, :, :3, :4]).sum())
    assert ulp_diff > 0, "fixture lost its purpose: MKL and correctly rounded trig agree on these angles"
    marg = _margins(fm, ref, n)
    report = {"fixture": NAME, "shape": "B1 192x640 n=3", "rotation_entries_differing_from_reference": ulp_diff,
              "decisions": int(fe["sel"].size), "flips": [], "loss_photometric_relerr": relerr(fe["loss_photometric"], ref["us_loss_photometric"])}
    total = 0
    for i in range(n):
        bad = fe["sel"][i] != ref["us_sel_%d" % i][:, 0]
        m = marg[i][bad]
        total += int(bad.sum())
        report["flips"].append({"scale": i, "count": int(bad.sum()), "max_margin": float(m.max()) if m.size else 0.0,
                                "margins": [float(x) for x in np.sort(m)[:32]]})
        assert m.size == 0 or float(m.max()) <= MAX_MARGIN, "a flipped pixel is not a near-tie (margin %.3e)" % float(m.max())
    allm = np.concatenate([x.ravel() for x in marg])
    edges = [0, 1e-8, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1.0]
    report["margin_histogram_all_pixels"] = {"edges": edges, "counts": [int(c) for c in np.histogram(allm, bins=edges)[0]],
                                             "exact_ties": int((allm == 0).sum())}
    report["total_flips"] = total
    assert total <= MAX_FLIPS, "%d selection flips with un-snapped Euler angles" % total
    assert report["loss_photometric_relerr"] <= LOSS_RTOL
    try:
        os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
        with open(os.path.join(ROOT, "profiles", "r02_unsnapped_selection.json"), "w") as f:
            json.dump(report, f, indent=1)
    except OSError:
        pass


def test_oracle_c1_n1_matches_reference():
    pred, tgt, hp, ref = load_golden(NAME)
    o = Oracle({"depth": pred["depth"][:1], "poses": pred["poses"]}, tgt)
    f = o.forward()
    g = o.backward(1.0, 1.0)
    assert relerr(f["loss_photometric"], ref["n1_loss_photometric"]) <= LOSS_RTOL
    assert relerr(f["loss_smoothness"], ref["n1_loss_smoothness"]) <= LOSS_RTOL
    assert np.array_equal(f["sel"][0], ref["n1_sel_0"][:, 0])
    assert l2rel(g["grad_depth"][0], ref["n1_grad_depth_0"]) <= GRAD_RTOL
    assert maxrel(g["grad_depth"][0], ref["n1_grad_depth_0"]) <= GRAD_RTOL
    assert l2rel(g["grad_poses"], ref["n1_grad_poses"]) <= GRAD_RTOL


# ---- the CUDA path -------------------------------------------------------------------------------------------------------
def _cuda(pred, tgt, hp, backward="stash"):
    from test_gpu_parity import _run_cuda
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return _run_cuda(pred, tgt, hp, torch.device("cuda:0"), backward=backward)


@pytest.mark.gpu
def test_cuda_c1_n1_matches_reference():
    pred, tgt, hp, ref = load_golden(NAME)
    r = _cuda({"depth": pred["depth"][:1], "poses": pred["poses"]}, tgt, hp)
    assert relerr(r["loss_photometric"], ref["n1_loss_photometric"]) <= LOSS_RTOL
    assert relerr(r["loss_smoothness"], ref["n1_loss_smoothness"]) <= LOSS_RTOL
    assert np.array_equal(r["sel"][0], ref["n1_sel_0"][:, 0])
    assert l2rel(r["grad_depth"][0], ref["n1_grad_depth_0"]) <= GRAD_RTOL
    assert maxrel(r["grad_depth"][0], ref["n1_grad_depth_0"]) <= GRAD_RTOL
    assert l2rel(r["grad_poses"], ref["n1_grad_poses"]) <= GRAD_RTOL


@pytest.mark.gpu
@pytest.mark.parametrize("backward", ["stash", "recompute"])
def test_cuda_pose_matrix_input_is_bit_exact_for_unsnapped_angles(backward):
    pred, tgt, hp, ref, poses, mats = _us()
    n = len(pred["depth"])
    r = _cuda({"depth": pred["depth"], "poses": mats}, tgt, hp, backward=backward)
    for i in range(n):
        assert int((r["sel"][i] != ref["us_sel_%d" % i][:, 0]).sum()) == 0
    assert relerr(r["loss_photometric"], ref["us_loss_photometric"]) <= LOSS_RTOL
    assert relerr(r["loss_smoothness"], ref["us_loss_smoothness"]) <= LOSS_RTOL
    # gradient w.r.t. the matrices = dL/d(R|t), against the oracle's fp64 sums; the bottom row carries none
    o = Oracle({"depth": pred["depth"], "poses": mats}, tgt)
    o.forward()
    g = o.backward(1.0, 1.0)
    assert r["grad_poses"].shape == (1, 2, 4, 4)
    assert l2rel(r["grad_poses"][:, :, :3, :], g["grad_Rt"].reshape(1, 2, 3, 4)) <= GRAD_RTOL
    assert np.abs(r["grad_poses"][:, :, 3, :]).max() == 0.0
    for i in range(n):
        assert l2rel(r["grad_depth"][i], g["grad_depth"][i]) <= GRAD_RTOL


@pytest.mark.gpu
def test_cuda_unsnapped_euler_equals_oracle_and_is_close_to_reference():
    """The Euler path computes what the oracle computes (correctly rounded trig), bit for bit; against the reference's MKL trig
    the flips stay within the bound the CPU test above reports."""
    pred, tgt, hp, ref, poses, mats = _us()
    n = len(pred["depth"])
    r = _cuda({"depth": pred["depth"], "poses": poses}, tgt, hp)
    fe = Oracle({"depth": pred["depth"], "poses": poses}, tgt).forward()
    assert int((r["sel"] != fe["sel"]).sum()) == 0
    flips = sum(int((r["sel"][i] != ref["us_sel_%d" % i][:, 0]).sum()) for i in range(n))
    assert flips <= MAX_FLIPS
    assert relerr(r["loss_photometric"], ref["us_loss_photometric"]) <= LOSS_RTOL


@pytest.mark.gpu
def test_cuda_pose_matrix_through_torch_autograd_matches_euler_gradients():
    """A caller that builds the matrices with torch (pose_vec2mat) and lets autograd chain through them gets the Euler-vector
    gradient of the 6-vector path."""
    from mgnet_b200 import MultiViewPhotometricLoss
    from mgnet_b200.geometry import Pose
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    dev = torch.device("cuda:0")
    pred, tgt, hp, ref = load_golden(NAME)
    t = {k: v.to(dev) for k, v in tgt.items()}
    res = []
    for mode in ("euler", "matrix"):
        vec = pred["poses"].to(dev).requires_grad_(True)
        inv = [d.to(dev).requires_grad_(True) for d in pred["depth"]]
        if mode == "matrix":
            poses = torch.stack([Pose.from_vec(vec[:, s], "euler").mat for s in range(2)], 1)
        else:
            poses = vec
        out = MultiViewPhotometricLoss(**hp)({"depth": inv, "poses": poses}, t)
        (out["loss_photometric"] + out["loss_smoothness"]).backward()
        res.append((out["loss_photometric"].item(), vec.grad.cpu().numpy(), inv[0].grad.cpu().numpy()))
    assert relerr(res[1][0], res[0][0]) <= LOSS_RTOL
    assert l2rel(res[1][1], res[0][1]) <= GRAD_RTOL
    assert l2rel(res[1][2], res[0][2]) <= GRAD_RTOL
    assert l2rel(res[0][1], ref["grad_poses"]) <= GRAD_RTOL
