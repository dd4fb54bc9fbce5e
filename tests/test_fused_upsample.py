"""fused head-side upsample.  The depth head's low-resolution inverse-depth maps go straight into the
loss, which applies the head's own F.interpolate(scale_factor=stride, mode="bilinear", align_corners=True)
(mg_net.py:803-806) inside the kernels and returns low-resolution gradients.

CPU part (no GPU): the fixture tests/golden/fused_upsample_n3.npz (made by the unmodified reference behind the
head's interpolate, tests/golden/make_golden.py) is re-derived from the ATen port, and the arithmetic the kernels
use for the upsample (mgvs_device.cuh: upsample_at) is pinned bit for bit against ATen's CPU kernel through a numpy
restatement of the same formula.
GPU part: the fused module against that fixture and, at a larger size, against port + C oracle.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import GOLDEN, GRAD_RTOL, LOSS_RTOL, l2rel, maxrel, relerr

STRIDES = (8, 16, 32)
HP = dict(ssim_loss_weight=0.85, photometric_loss_weight=1.0, smoothing_loss_weight=1e-3, automask_loss=True)


def _load():
    z = np.load(os.path.join(GOLDEN, "fused_upsample_n3.npz"))
    lows = [torch.from_numpy(z["in_depth_%d" % i]) for i in range(3)]
    tgt = {k[3:]: torch.from_numpy(z[k]) for k in ("in_image_orig", "in_image_prev_orig", "in_image_next_orig", "in_camera_matrix", "in_reprojection_mask")}
    return z, lows, torch.from_numpy(z["in_poses"]), tgt


def upsample_restated(x, s):
    """numpy restatement of upsample_at() (mgnet_b200/csrc/mgvs_device.cuh) == ATen CPU upsample_bilinear2d, align_corners=True."""
    x = np.asarray(x, np.float32)
    B, _, h, w = x.shape
    H, W = h * s, w * s
    f32 = np.float32

    def axis(n_in, n_out):
        r = f32(np.float64(n_in - 1) / np.float64(n_out - 1)) if n_out > 1 else f32(0)
        real = (r * np.arange(n_out, dtype=np.float32)).astype(np.float32)
        i0 = np.minimum(real.astype(np.int64), n_in - 1)
        i1 = np.minimum(i0 + 1, n_in - 1)
        l1 = np.clip((real - i0.astype(np.float32)).astype(np.float32), 0, 1).astype(np.float32)
        return i0, i1, (f32(1) - l1).astype(np.float32), l1

    def fma(a, b, c):
        return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)   # exact product, one rounding

    y0, y1, ly0, ly1 = axis(h, H)
    x0, x1, lx0, lx1 = axis(w, W)
    a, b = x[:, :, y0][:, :, :, x0], x[:, :, y0][:, :, :, x1]
    c, d = x[:, :, y1][:, :, :, x0], x[:, :, y1][:, :, :, x1]
    LX0, LX1 = np.broadcast_to(lx0[None, None, None, :], a.shape), np.broadcast_to(lx1[None, None, None, :], a.shape)
    LY0, LY1 = np.broadcast_to(ly0[None, None, :, None], a.shape), np.broadcast_to(ly1[None, None, :, None], a.shape)
    t = fma(LX0, a, (LX1 * b).astype(np.float32))
    u = fma(LX0, c, (LX1 * d).astype(np.float32))
    return fma(LY0, t, (LY1 * u).astype(np.float32))


# ATen's CPU kernel switches arithmetic with the output size: for small outputs (roughly H*W < 4096 and W < 128, probed
# on torch 2.11 / AVX-512) it evaluates the flat sum fma(w11,d, fma(w10,c, fma(w00,a, w01*b))) with product weights, for
# larger ones the nested form restated above.  Every resolution MGNet trains at (192x640 ... 1024x2048) is far on the
# "large" side; bit-exactness of the fused path is claimed (and tested) there.
@pytest.mark.parametrize("shape", [(24, 80, 8), (12, 40, 16), (6, 20, 32), (8, 16, 8), (64, 128, 8), (16, 64, 2), (128, 4, 16)])
def test_upsample_formula_matches_aten_bit_for_bit(shape):
    h, w, s = shape
    g = torch.Generator().manual_seed(h * 100 + s)
    x = torch.sigmoid(torch.randn(2, 1, h, w, generator=g)) / 0.5
    ref = F.interpolate(x, scale_factor=s, mode="bilinear", align_corners=True).numpy()
    assert np.array_equal(upsample_restated(x.numpy(), s), ref)


def test_fixture_is_reproduced_by_the_port():
    from oracle.torch_port import reference_loss
    z, lows, poses, tgt = _load()
    leaves = [x.clone().requires_grad_(True) for x in lows]
    pl = poses.clone().requires_grad_(True)
    fulls = [F.interpolate(x, scale_factor=s, mode="bilinear", align_corners=True) for x, s in zip(leaves, STRIDES)]
    out = reference_loss({"depth": fulls, "poses": pl}, tgt, return_selection=True, **HP)
    (out["loss_photometric"] + out["loss_smoothness"]).backward()
    assert relerr(out["loss_photometric"].item(), z["loss_photometric"]) <= LOSS_RTOL
    assert relerr(out["loss_smoothness"].item(), z["loss_smoothness"]) <= LOSS_RTOL
    for i in range(3):
        assert np.array_equal(out["selection"][i].numpy(), z["sel_%d" % i][:, 0])
        assert l2rel(leaves[i].grad.numpy(), z["grad_depth_%d" % i]) <= GRAD_RTOL
    assert l2rel(pl.grad.numpy(), z["grad_poses"]) <= GRAD_RTOL


def _run_fused(lows, poses, tgt, dev):
    from mgnet_b200 import MultiViewPhotometricLoss
    mod = MultiViewPhotometricLoss(photometric_reduce_op="min", padding_mode="zeros", fuse_upsample=True, **HP)
    p = {"depth": [x.to(dev).requires_grad_(True) for x in lows], "poses": poses.to(dev).requires_grad_(True)}
    t = {k: v.to(dev) for k, v in tgt.items()}
    out = mod(p, t)
    (out["loss_photometric"] + out["loss_smoothness"]).backward()
    torch.cuda.synchronize()
    return (out["loss_photometric"].item(), out["loss_smoothness"].item(), mod.last_selection.cpu().numpy(),
            [d.grad.cpu().numpy() for d in p["depth"]], p["poses"].grad.cpu().numpy())


@pytest.mark.gpu
def test_fused_upsample_matches_reference_fixture():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    z, lows, poses, tgt = _load()
    lp, ls, sel, gd, gp = _run_fused(lows, poses, tgt, torch.device("cuda:0"))
    assert relerr(lp, z["loss_photometric"]) <= LOSS_RTOL
    assert relerr(ls, z["loss_smoothness"]) <= LOSS_RTOL
    for i in range(3):
        assert int((sel[i] != z["sel_%d" % i][:, 0]).sum()) == 0        # bit-exact: needs the bit-exact upsample
        assert gd[i].shape == z["grad_depth_%d" % i].shape               # low resolution
        assert l2rel(gd[i], z["grad_depth_%d" % i]) <= GRAD_RTOL
        assert maxrel(gd[i], z["grad_depth_%d" % i]) <= GRAD_RTOL
    assert l2rel(gp, z["grad_poses"]) <= GRAD_RTOL


@pytest.mark.gpu
def test_fused_upsample_kitti_size_against_port_and_oracle():
    """B2 192x640 (the KITTI resolution of BASELINE config[1]), strides 8/16/32: loss / selection against the C oracle
    on the ATen-upsampled maps, low-resolution gradients against autograd through interpolate + port; and the fused
    result must equal the unfused module fed with the upsampled maps (same kernels, same tiles)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from mgnet_b200 import MultiViewPhotometricLoss
    from mgnet_b200.synthetic import make_inputs
    from oracle.oracle import Oracle
    from oracle.torch_port import reference_loss
    dev = torch.device("cuda:0")
    B, H, W = 2, 192, 640
    pred, tgt = make_inputs(B, H, W, 3, seed=19)
    g = torch.Generator().manual_seed(1919)
    lows = [(0.05 + 1.9 * torch.rand(B, 1, H // s, W // s, generator=g)).contiguous() for s in STRIDES]
    leaves = [x.clone().requires_grad_(True) for x in lows]
    pl = pred["poses"].clone().requires_grad_(True)
    fulls = [F.interpolate(x, scale_factor=s, mode="bilinear", align_corners=True) for x, s in zip(leaves, STRIDES)]
    out = reference_loss({"depth": fulls, "poses": pl}, tgt, **HP)
    (out["loss_photometric"] + out["loss_smoothness"]).backward()
    f = Oracle({"depth": [x.detach() for x in fulls], "poses": pred["poses"]}, tgt, **HP).forward()
    lp, ls, sel, gd, gp = _run_fused(lows, pred["poses"], tgt, dev)
    assert relerr(lp, f["loss_photometric"]) <= LOSS_RTOL and relerr(ls, f["loss_smoothness"]) <= LOSS_RTOL
    assert int((sel != f["sel"]).sum()) == 0
    for i in range(3):
        assert l2rel(gd[i], leaves[i].grad.numpy()) <= GRAD_RTOL
        assert maxrel(gd[i], leaves[i].grad.numpy()) <= GRAD_RTOL
    assert l2rel(gp, pl.grad.numpy()) <= GRAD_RTOL
    # unfused module on the upsampled maps: identical forward, and its full-resolution gradients pushed through ATen's
    # (CPU, deterministic) interpolate backward must give the fused low-resolution gradients
    mod = MultiViewPhotometricLoss(photometric_reduce_op="min", padding_mode="zeros", **HP)
    fd = [x.detach().to(dev).requires_grad_(True) for x in fulls]
    o2 = mod({"depth": fd, "poses": pred["poses"].to(dev).requires_grad_(True)}, {k: v.to(dev) for k, v in tgt.items()})
    (o2["loss_photometric"] + o2["loss_smoothness"]).backward()
    assert o2["loss_photometric"].item() == lp and o2["loss_smoothness"].item() == ls
    assert np.array_equal(mod.last_selection.cpu().numpy(), sel)
    for i, s in enumerate(STRIDES):
        leaf = lows[i].clone().requires_grad_(True)
        F.interpolate(leaf, scale_factor=s, mode="bilinear", align_corners=True).backward(fd[i].grad.cpu())
        assert l2rel(gd[i], leaf.grad.numpy()) <= 1e-5
    # deterministic
    again = _run_fused(lows, pred["poses"], tgt, dev)
    for x, y in zip(gd, again[3]):
        assert np.array_equal(x, y)


@pytest.mark.gpu
def test_fused_upsample_rejects_bad_shapes():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from mgnet_b200 import MultiViewPhotometricLoss
    from mgnet_b200.synthetic import make_inputs
    dev = torch.device("cuda:0")
    pred, tgt = make_inputs(1, 64, 128, 1, seed=2)
    t = {k: v.to(dev) for k, v in tgt.items()}
    fused = MultiViewPhotometricLoss(photometric_reduce_op="min", padding_mode="zeros", fuse_upsample=True, **HP)
    with pytest.raises(ValueError):      # 64x128 -> 10x16 is not an integer stride
        fused({"depth": [torch.rand(1, 1, 10, 16, device=dev)], "poses": pred["poses"].to(dev)}, t)
    plain = MultiViewPhotometricLoss(photometric_reduce_op="min", padding_mode="zeros", **HP)
    with pytest.raises(ValueError):      # low-resolution maps without fuse_upsample are a shape error, never silently resized
        plain({"depth": [torch.rand(1, 1, 8, 16, device=dev)], "poses": pred["poses"].to(dev)}, t)
