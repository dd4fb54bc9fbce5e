"""PoseCNN tail (reference layers.py:164-166) as one kernel each way, against the reference's own three ATen calls."""
import pytest
import torch


def _ref_tail(x, num_context=2):
    out = x.mean(3).mean(2)                               # layers.py:164
    return 0.01 * out.view(out.size(0), num_context, 6)   # layers.py:165-166


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(4, 12, 6, 20), (2, 12, 32, 64), (1, 12, 1, 1), (3, 18, 16, 33), (2, 12, 300, 7)])
def test_pose_tail_matches_reference_tail(shape):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from mgnet_b200.pose_tail import pose_tail
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    B, C, h, w = shape
    nctx = C // 6
    x = torch.randn(B, C, h, w, generator=g)
    up = torch.randn(B, nctx, 6, generator=g)
    xr = x.clone().double().requires_grad_(True)          # fp64 reference of the same three calls
    ref = _ref_tail(xr, nctx)
    (ref * up.double()).sum().backward()
    xd = x.to(dev).requires_grad_(True)
    out = pose_tail(xd, nctx)
    assert out.shape == (B, nctx, 6) and out.dtype == torch.float32
    (out * up.to(dev)).sum().backward()
    # within one fp32 ulp of the exactly rounded value (the fp32 ATen sequence itself is a few ulps off)
    assert torch.allclose(out.cpu().double(), ref.detach(), rtol=2e-7, atol=1e-12)
    assert torch.allclose(xd.grad.cpu().double(), xr.grad, rtol=2e-7, atol=1e-12)
    # and the fp32 reference tail on the same device agrees to its own rounding
    assert torch.allclose(out, _ref_tail(x.to(dev), nctx), rtol=2e-5, atol=1e-9)
    # deterministic
    assert torch.equal(out, pose_tail(x.to(dev), nctx))


@pytest.mark.gpu
def test_pose_tail_feeds_the_loss_and_accepts_half():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from mgnet_b200 import MultiViewPhotometricLoss
    from mgnet_b200.pose_tail import pose_tail
    from mgnet_b200.synthetic import make_inputs
    dev = torch.device("cuda:0")
    pred, tgt = make_inputs(2, 64, 128, 2, seed=9)
    feat = (100.0 * pred["poses"].view(2, 12, 1, 1)).expand(2, 12, 2, 4).contiguous().to(dev).half().requires_grad_(True)
    poses = pose_tail(feat)
    assert torch.allclose(poses.cpu(), pred["poses"], rtol=2e-3, atol=1e-6)
    mod = MultiViewPhotometricLoss(0.85, 1.0, 1e-3, True, "min", "zeros")
    out = mod({"depth": [d.to(dev) for d in pred["depth"]], "poses": poses}, {k: v.to(dev) for k, v in tgt.items()})
    out["loss_photometric"].backward()
    assert feat.grad is not None and feat.grad.dtype == torch.float16 and torch.isfinite(feat.grad).all() and float(feat.grad.abs().sum()) > 0


def test_pose_tail_refuses_cpu_and_bad_shapes():
    from mgnet_b200.pose_tail import pose_tail
    with pytest.raises(RuntimeError):
        pose_tail(torch.zeros(1, 12, 2, 2))
    with pytest.raises(ValueError):
        pose_tail(torch.zeros(1, 11, 2, 2))
