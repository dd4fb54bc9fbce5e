"""GPU parity tests (run on the B200 box with -m gpu).  Everything goes through the C ABI
(mgnet_b200/_lib.py -> libmgvs.so); the checker is the golden fixtures made by the reference and the
CPU oracle (oracle/), never the thing under test.

Bars (BASELINE.json north_star): losses <= 1e-5 relative, depth/pose gradients <= 1e-4 relative
(L2 norm and max norm), selection mask bit-exact.
"""
import ctypes

import numpy as np
import pytest
import torch

from helpers import GRAD_RTOL, LOSS_RTOL, golden_names, l2rel, load_golden, maxrel, relerr

pytestmark = pytest.mark.gpu

OR_KEYS = ("ssim_loss_weight", "photometric_loss_weight", "smoothing_loss_weight", "automask_loss", "padding_mode")


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _to_dev(pred, tgt, dev, grad=True):
    p = {"depth": [d.to(dev).requires_grad_(grad) for d in pred["depth"]], "poses": pred["poses"].to(dev).requires_grad_(grad)}
    t = {k: v.to(dev) for k, v in tgt.items()}
    return p, t


BACKWARDS = ("stash", "recompute")   # both backward kernels must meet the same bars


def _run_cuda(pred, tgt, hp, dev, g=(1.0, 1.0), backward="stash"):
    from mgnet_b200 import MultiViewPhotometricLoss
    mod = MultiViewPhotometricLoss(backward=backward, **hp)
    p, t = _to_dev(pred, tgt, dev)
    out = mod(p, t)
    (g[0] * out["loss_photometric"] + g[1] * out["loss_smoothness"]).backward()
    torch.cuda.synchronize()
    return {
        "loss_photometric": out["loss_photometric"].item(),
        "loss_smoothness": out["loss_smoothness"].item(),
        "sel": mod.last_selection.cpu().numpy() if mod.last_selection is not None else None,
        "grad_depth": [d.grad.cpu().numpy() for d in p["depth"]],
        "grad_poses": p["poses"].grad.cpu().numpy(),
    }


def test_library_loaded_is_in_tree():
    _dev()
    from mgnet_b200 import _lib
    L = _lib.lib()
    assert L.mgvs_abi_version() == _lib.ABI_VERSION == 7
    assert _lib.LIB_PATH.endswith("mgnet_b200/libmgvs.so")


@pytest.mark.parametrize("name", golden_names())
def test_golden_forward(name):
    dev = _dev()
    pred, tgt, hp, ref = load_golden(name)
    r = _run_cuda(pred, tgt, hp, dev)
    assert relerr(r["loss_photometric"], ref["loss_photometric"]) <= LOSS_RTOL
    assert relerr(r["loss_smoothness"], ref["loss_smoothness"]) <= LOSS_RTOL
    for i in range(len(pred["depth"])):
        mism = int((r["sel"][i] != ref["sel_%d" % i][:, 0]).sum())
        assert mism == 0, "scale %d: %d selection mismatches vs the reference" % (i, mism)


@pytest.mark.parametrize("backward", BACKWARDS)
@pytest.mark.parametrize("name", golden_names())
def test_golden_backward(name, backward):
    dev = _dev()
    pred, tgt, hp, ref = load_golden(name)
    r = _run_cuda(pred, tgt, hp, dev, backward=backward)
    for i in range(len(pred["depth"])):
        assert l2rel(r["grad_depth"][i], ref["grad_depth_%d" % i]) <= GRAD_RTOL
        assert maxrel(r["grad_depth"][i], ref["grad_depth_%d" % i]) <= GRAD_RTOL
    assert l2rel(r["grad_poses"], ref["grad_poses"]) <= GRAD_RTOL
    assert maxrel(r["grad_poses"], ref["grad_poses"]) <= GRAD_RTOL


@pytest.mark.parametrize("backward", BACKWARDS)
@pytest.mark.parametrize("shape", [(2, 192, 640, 3, 0.2, False, "zeros", 0.01), (1, 96, 320, 4, 0.0, True, "zeros", 0.01),
                                   (3, 50, 70, 2, 0.2, False, "zeros", 0.01), (2, 192, 640, 3, 0.0, True, "border", 0.05),
                                   (2, 96, 320, 2, 0.2, False, "reflection", 0.05), (2, 50, 70, 2, 0.0, True, "reflection", 0.08)])
def test_against_oracle(shape, backward):
    """Sizes the oracle finishes in seconds, incl. H/W that are not multiples of the 64x16 tile and grid_sample's
    border / reflection padding with poses large enough to push many samples outside the image."""
    dev = _dev()
    from mgnet_b200.synthetic import make_inputs
    from oracle.oracle import Oracle
    B, H, W, n, noise, shift, pad, pose_scale = shape
    pred, tgt = make_inputs(B, H, W, n, seed=11, noise=noise, shift_sources=shift, pose_scale=pose_scale)
    hp = dict(ssim_loss_weight=0.85, photometric_loss_weight=1.0, smoothing_loss_weight=1e-3, automask_loss=True,
              photometric_reduce_op="min", padding_mode=pad)
    o = Oracle(pred, tgt, **{k: hp[k] for k in OR_KEYS})
    f = o.forward()
    g = o.backward(1.0, 1.0)
    r = _run_cuda(pred, tgt, hp, dev, backward=backward)
    assert relerr(r["loss_photometric"], f["loss_photometric"]) <= LOSS_RTOL
    assert relerr(r["loss_smoothness"], f["loss_smoothness"]) <= LOSS_RTOL
    assert int((r["sel"] != f["sel"]).sum()) == 0
    for i in range(n):
        assert l2rel(r["grad_depth"][i], g["grad_depth"][i]) <= GRAD_RTOL
        assert maxrel(r["grad_depth"][i], g["grad_depth"][i]) <= GRAD_RTOL
    assert l2rel(r["grad_poses"], g["grad_poses"]) <= GRAD_RTOL


@pytest.mark.parametrize("backward", BACKWARDS)
def test_backward_deterministic_and_linear(backward):
    dev = _dev()
    pred, tgt, hp, ref = load_golden("grad_smooth_shift")
    import functools
    _run = functools.partial(_run_cuda, backward=backward)
    a = _run(pred, tgt, hp, dev)
    b = _run(pred, tgt, hp, dev)
    for x, y in zip(a["grad_depth"], b["grad_depth"]):
        assert np.array_equal(x, y)
    assert np.array_equal(a["grad_poses"], b["grad_poses"])
    c = _run(pred, tgt, hp, dev, g=(2.0, 3.0))
    p10 = _run(pred, tgt, hp, dev, g=(1.0, 0.0))
    p01 = _run(pred, tgt, hp, dev, g=(0.0, 1.0))
    for i in range(len(pred["depth"])):
        assert l2rel(c["grad_depth"][i], 2.0 * p10["grad_depth"][i].astype(np.float64) + 3.0 * p01["grad_depth"][i]) <= 1e-5
    assert np.abs(p01["grad_poses"]).max() == 0.0


def test_exact_division_matches_ieee():
    dev = _dev()
    from mgnet_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device="cpu").manual_seed(5)
    n = 1 << 22
    cases = []
    a = (torch.rand(n, generator=g) * 4 - 2) * torch.exp(torch.randn(n, generator=g) * 6)
    b = (torch.rand(n, generator=g) + 1e-3) * torch.exp(torch.randn(n, generator=g) * 6)
    cases.append((a, b))
    for d in (63.0, 95.0, 191.0, 639.0, 1023.0, 2047.0, 511.0, 71.0, 39.0):
        cases.append(((torch.rand(n, generator=g) * 2 - 1) * 4000.0, torch.full((n,), d)))
    for a, b in cases:
        a, b = a.float().to(dev), b.float().to(dev)
        out = torch.empty_like(a)
        _lib.check(L.mgvs_test_div(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(),
                                   ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        ref = (a.double() / b.double()).float()      # correctly rounded (double rounding is safe: 53 >= 2*24+2)
        assert int((out != ref).sum()) == 0


@pytest.mark.parametrize("fixture", ["kitti_small_mask", "pad_border", "pad_reflection"])
def test_view_synthesis_matches_reference_intermediates(fixture):
    dev = _dev()
    from mgnet_b200.geometry import Camera, Pose, inv2depth, view_synthesis
    pred, tgt, hp, ref = load_golden(fixture)
    K = tgt["camera_matrix"][:, :3, :3].to(dev)
    for s, key in enumerate(("image_prev_orig", "image_next_orig")):
        pose = Pose.from_vec(pred["poses"][:, s].to(dev), "euler")
        # the pose matrix itself must match the reference's bit for bit (trig boundary: snapped angles)
        depth = inv2depth(pred["depth"][0].to(dev))
        pose_ref = Pose(torch.from_numpy(ref["pose_mat"][:, s]).to(dev))
        warped, coords = view_synthesis(tgt[key].to(dev), depth, Camera(K, Tcw=pose_ref), Camera(K).to(dev),
                                        padding_mode=hp["padding_mode"], return_coords=True)
        assert np.array_equal(coords.cpu().numpy(), ref["coords_0_%d" % s])
        assert np.array_equal(warped.cpu().numpy(), ref["warped_0_%d" % s])


@pytest.mark.parametrize("shape", [(2, 96, 320, 3), (1, 50, 70, 2)])
def test_uint8_images_match_float_path(shape):
    """uint8 images (what the data loader produces) must give bit-identical results to the
    reference caller's `x.float() / 255.0` followed by the float path (mg_net.py:320-335), also when
    W % 4 != 0 (scalar loader)."""
    dev = _dev()
    from mgnet_b200.synthetic import make_inputs, quantize_images
    B, H, W, n = shape
    pred, tgt = make_inputs(B, H, W, n, seed=23)
    tu, tf = quantize_images(tgt)
    hp = dict(ssim_loss_weight=0.85, photometric_loss_weight=1.0, smoothing_loss_weight=1e-3, automask_loss=True,
              photometric_reduce_op="min", padding_mode="zeros")
    a = _run_cuda(pred, tu, hp, dev)
    b = _run_cuda(pred, tf, hp, dev)
    assert a["loss_photometric"] == b["loss_photometric"] and a["loss_smoothness"] == b["loss_smoothness"]
    assert np.array_equal(a["sel"], b["sel"])
    for x, y in zip(a["grad_depth"], b["grad_depth"]):
        assert np.array_equal(x, y)
    assert np.array_equal(a["grad_poses"], b["grad_poses"])
    # and the float path on the quantised images is itself checked against the oracle
    from oracle.oracle import Oracle
    f = Oracle(pred, tf, **{k: hp[k] for k in OR_KEYS}).forward()
    assert relerr(a["loss_photometric"], f["loss_photometric"]) <= LOSS_RTOL
    assert int((a["sel"] != f["sel"]).sum()) == 0


def test_cpu_tensors_raise():
    _dev()
    from mgnet_b200 import MultiViewPhotometricLoss
    pred, tgt, hp, ref = load_golden("nomask_n1")
    with pytest.raises(RuntimeError):
        MultiViewPhotometricLoss(**hp)(pred, tgt)


def test_stash_and_recompute_backward_agree_at_full_size():
    """BASELINE config[1] size (B16 192x640 n=3): the two backward kernels share no code above the per-output chain
    (stash: forward-emitted SSIM-adjoint coefficients + box adjoint; recompute: tile+2 warps + statistics), so their
    agreement is an independent cross-check on top of the oracle comparison at this size (tests/test_gpu_bench_shapes.py).
    Forward outputs are bit-identical (the stash only adds stores)."""
    dev = _dev()
    from mgnet_b200.synthetic import make_inputs
    pred, tgt = make_inputs(16, 192, 640, 3, seed=3)
    hp = dict(ssim_loss_weight=0.85, photometric_loss_weight=1.0, smoothing_loss_weight=1e-3, automask_loss=True,
              photometric_reduce_op="min", padding_mode="zeros")
    a = _run_cuda(pred, tgt, hp, dev, backward="stash")
    b = _run_cuda(pred, tgt, hp, dev, backward="recompute")
    assert a["loss_photometric"] == b["loss_photometric"] and a["loss_smoothness"] == b["loss_smoothness"]
    assert np.array_equal(a["sel"], b["sel"])
    for x, y in zip(a["grad_depth"], b["grad_depth"]):
        assert l2rel(x, y) <= 5e-5 and maxrel(x, y) <= 5e-5      # each is within 1e-4 of the reference; fp32 summation order differs
    assert l2rel(a["grad_poses"], b["grad_poses"]) <= 1e-5


@pytest.mark.parametrize("backward", BACKWARDS)
def test_mean_reduce_matches_reference_fixture(backward):
    """photometric_reduce_op="mean" (loss.py:242-243, automask off): fixture from the unmodified reference; also against the
    oracle composed the same way (one "min" evaluation per source frame, that frame in both slots)."""
    dev = _dev()
    from oracle.oracle import Oracle
    pred, tgt, hp, ref = load_golden("mean_reduce")
    assert hp["photometric_reduce_op"] == "mean" and not hp["automask_loss"]
    r = _run_cuda(pred, tgt, hp, dev, backward=backward)
    assert relerr(r["loss_photometric"], ref["loss_photometric"]) <= LOSS_RTOL
    assert relerr(r["loss_smoothness"], ref["loss_smoothness"]) <= LOSS_RTOL
    for i in range(len(pred["depth"])):
        assert l2rel(r["grad_depth"][i], ref["grad_depth_%d" % i]) <= GRAD_RTOL
    assert l2rel(r["grad_poses"], ref["grad_poses"]) <= GRAD_RTOL
    lp = []
    for s, key in enumerate(("image_prev_orig", "image_next_orig")):
        t2 = dict(tgt, image_prev_orig=tgt[key], image_next_orig=tgt[key])
        p2 = {"depth": pred["depth"], "poses": pred["poses"][:, [s, s]].contiguous()}
        lp.append(float(Oracle(p2, t2, automask_loss=False).forward()["loss_photometric"]))
    assert relerr(r["loss_photometric"], 0.5 * (lp[0] + lp[1])) <= LOSS_RTOL


def _random_cases():
    import test_oracle_random
    return test_oracle_random.CASES


@pytest.mark.parametrize("backward", BACKWARDS)
@pytest.mark.parametrize("c", _random_cases(), ids=lambda c: "%dx%dx%d_n%d_%s_am%d_a%.2f_p%.2f" % (c["B"], c["H"], c["W"], c["n"], c["pad"], c["automask"], c["ssim"], c["pose_scale"]))
def test_random_configurations_against_oracle(c, backward):
    """The randomised configurations of tests/test_oracle_random.py (ragged shapes down to 8 pixels, every padding mode, automask on/off,
    ssim weights incl. 0, small and large poses) through the CUDA path."""
    dev = _dev()
    from mgnet_b200.synthetic import make_inputs
    from oracle.oracle import Oracle
    pred, tgt = make_inputs(c["B"], c["H"], c["W"], c["n"], seed=c["seed"], noise=0.0 if c["shift"] else 0.15,
                            pose_scale=c["pose_scale"], with_mask=c["with_mask"], shift_sources=c["shift"])
    hp = dict(ssim_loss_weight=c["ssim"], photometric_loss_weight=1.0, smoothing_loss_weight=1e-3, automask_loss=c["automask"],
              photometric_reduce_op="min", padding_mode=c["pad"])
    o = Oracle(pred, tgt, **{k: hp[k] for k in OR_KEYS})
    f = o.forward()
    g = o.backward(1.0, 1.0)
    r = _run_cuda(pred, tgt, hp, dev, backward=backward)
    assert relerr(r["loss_photometric"], f["loss_photometric"]) <= LOSS_RTOL
    assert relerr(r["loss_smoothness"], f["loss_smoothness"]) <= LOSS_RTOL
    assert int((r["sel"] != f["sel"]).sum()) == 0
    for i in range(c["n"]):
        assert l2rel(r["grad_depth"][i], g["grad_depth"][i]) <= GRAD_RTOL
    if float(np.abs(g["grad_poses"]).max()) > 0:
        assert l2rel(r["grad_poses"], g["grad_poses"]) <= GRAD_RTOL


@pytest.mark.parametrize("shape", [(2, 64, 128), (3, 37, 75), (1, 192, 640)])
def test_bitpacked_mask_matches_bool_mask(shape):
    """targets["reprojection_mask"] as numpy.packbits bytes (mgvs_unpack_mask): identical results, 1/8 of the mask's H2D bytes."""
    dev = _dev()
    from mgnet_b200.synthetic import make_inputs, pack_mask
    B, H, W = shape
    pred, tgt = make_inputs(B, H, W, 2, seed=23)
    hp = dict(ssim_loss_weight=0.85, photometric_loss_weight=1.0, smoothing_loss_weight=1e-3, automask_loss=True,
              photometric_reduce_op="min", padding_mode="zeros")
    a = _run_cuda(pred, tgt, hp, dev)
    packed = dict(tgt, reprojection_mask=pack_mask(tgt["reprojection_mask"]))
    assert packed["reprojection_mask"].shape == (B, 1, H, (W + 7) // 8) and packed["reprojection_mask"].dtype == torch.uint8
    b = _run_cuda(pred, packed, hp, dev)
    assert a["loss_photometric"] == b["loss_photometric"] and a["loss_smoothness"] == b["loss_smoothness"]
    assert np.array_equal(a["sel"], b["sel"]) and np.array_equal(a["grad_poses"], b["grad_poses"])
    for x, y in zip(a["grad_depth"], b["grad_depth"]):
        assert np.array_equal(x, y)
