"""The `mgnet.geometry`-compatible surface of the drop-in (reference mgnet/geometry/*):
Camera.reconstruct / project / scaled, scale_intrinsics, construct_K, view_synthesis with two different cameras, calc_smoothness,
match_scales, Pose, the fp16 / autocast caller contract (mg_net.py:827), and the forward-only honesty of the stand-alone kernels.

Checker: the reference's formulas written with plain torch ops on the CPU (the stand-alone ops are a handful of ATen calls each:
camera.py:107-182, camera_utils.py:10-54) -- bit-exact where the CPU reference is deterministic (bmm with K=3 is an ascending FMA
chain), which is the same bar the fused kernels are held to.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.torch_port import _apply44, _invert44, _kinv, _pixel_grid, _pose44, _synthesize


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _inputs(B=2, H=40, W=72, seed=3, pose_scale=0.02):
    from mgnet_b200.synthetic import make_inputs
    pred, tgt = make_inputs(B, H, W, 1, seed=seed, noise=0.1, pose_scale=pose_scale)
    K = tgt["camera_matrix"][:, :3, :3].contiguous()
    depth = 1.0 / pred["depth"][0].clamp(min=1e-6)
    return pred, tgt, K, depth


# ---- CPU: pure host logic -------------------------------------------------------------------------------------------------
def test_intrinsics_helpers_and_scaled_camera():
    from mgnet_b200.geometry import Camera, Pose, construct_K, scale_intrinsics
    K = construct_K(100.0, 120.0, 31.5, 23.5)
    assert K.shape == (3, 3) and K.dtype == torch.float32
    assert K.tolist() == [[100.0, 0.0, 31.5], [0.0, 120.0, 23.5], [0.0, 0.0, 1.0]]
    Ks = scale_intrinsics(K.clone().unsqueeze(0), 0.5, 0.25)          # camera_utils.py:14-21: (c + 0.5) * s - 0.5
    assert Ks[0].tolist() == [[50.0, 0.0, (31.5 + 0.5) * 0.5 - 0.5], [0.0, 30.0, (23.5 + 0.5) * 0.25 - 0.5], [0.0, 0.0, 1.0]]
    pose = Pose.from_vec(torch.tensor([[0.1, 0.2, 0.3, 0.01, -0.02, 0.03]]), "euler")
    cam = Camera(K.unsqueeze(0), Tcw=pose)
    assert cam.scaled(1.0) is cam                                      # camera.py:96-97
    c2 = cam.scaled(0.5)
    assert c2 is not cam and c2.Tcw is cam.Tcw and torch.equal(cam.K[0], K)     # the original K is not modified
    assert torch.equal(c2.K, scale_intrinsics(K.clone().unsqueeze(0), 0.5, 0.5))
    c3 = cam.scaled(0.5, 0.25)
    assert torch.equal(c3.K, Ks)
    assert len(cam) == 1 and float(cam.fx) == 100.0 and float(cam.fy) == 120.0 and float(cam.cx) == 31.5 and float(cam.cy) == 23.5
    assert torch.equal(cam.Kinv, _kinv(K.unsqueeze(0)))                # closed form, camera.py:72-81


def test_pose_class_matches_the_port():
    from mgnet_b200.geometry import Pose, euler2mat, invert_pose, pose_vec2mat
    g = torch.Generator().manual_seed(0)
    vec = 0.1 * torch.randn(3, 6, generator=g)
    P = Pose.from_vec(vec, "euler")
    assert torch.equal(P.mat, _pose44(vec)) and P.shape == (3, 4, 4) and P.item() is P.mat and len(P) == 3
    assert torch.equal(pose_vec2mat(vec, "euler"), _pose44(vec)[:, :3])
    assert torch.equal(P.inverse().mat, _invert44(P.mat)) and torch.equal(invert_pose(P.mat), _invert44(P.mat))
    assert torch.equal(euler2mat(vec[:, 3:]), _pose44(vec)[:, :3, :3])
    pts = torch.randn(3, 3, 4, 5, generator=g)
    assert torch.equal(P @ pts, _apply44(P.mat, pts))
    assert torch.equal((P @ P.inverse()).mat, P.mat.bmm(_invert44(P.mat)))
    I = Pose.identity(2)
    assert I.mat.dtype == torch.float32 and torch.equal(I.mat, torch.eye(4).repeat(2, 1, 1))
    assert Pose.identity(1).repeat(4, 1, 1).shape == (4, 4, 4)
    with pytest.raises(ValueError):
        P @ torch.zeros(3, 5)
    with pytest.raises(ValueError):
        pose_vec2mat(vec, "quaternion")


def test_smoothness_and_scale_helpers():
    from mgnet_b200.geometry import calc_smoothness, gradient_x, gradient_y, image_grid, inv2depth, match_scales, same_shape
    g = torch.Generator().manual_seed(1)
    img = torch.rand(2, 3, 12, 20, generator=g)
    inv = [0.05 + 1.9 * torch.rand(2, 1, 12, 20, generator=g) for _ in range(2)]
    sx, sy = calc_smoothness(inv, img, 2)
    for i in range(2):
        nrm = inv[i] / inv[i].mean(2, True).mean(3, True).clamp(min=1e-6)          # depth.py:18-51
        wx = torch.exp(-(img[..., :-1] - img[..., 1:]).abs().mean(1, keepdim=True))
        wy = torch.exp(-(img[..., :-1, :] - img[..., 1:, :]).abs().mean(1, keepdim=True))
        assert torch.equal(sx[i], (nrm[..., :-1] - nrm[..., 1:]) * wx)
        assert torch.equal(sy[i], (nrm[..., :-1, :] - nrm[..., 1:, :]) * wy)
    assert torch.equal(gradient_x(img), img[..., :-1] - img[..., 1:]) and torch.equal(gradient_y(img), img[..., :-1, :] - img[..., 1:, :])
    same = match_scales(img, inv, 2)
    assert same[0] is img and same[1] is img                                       # equal H x W: the same tensor (image.py:91-95)
    small = [torch.zeros(2, 1, 6, 10)]
    assert torch.equal(match_scales(img, small, 1)[0], F.interpolate(img, size=(6, 10), mode="bilinear", align_corners=True))
    assert same_shape((1, 2), (1, 2)) and not same_shape((1, 2), (1, 2, 3))
    d = inv2depth(inv)
    assert isinstance(d, list) and torch.equal(d[0], 1.0 / inv[0].clamp(min=1e-6))
    assert torch.equal(inv2depth(torch.zeros(1, 1, 2, 2)), (1.0 / torch.tensor(1e-6)).expand(1, 1, 2, 2))     # clamp(min=1e-6), depth.py:15
    grid = image_grid(2, 3, 4, torch.float32, torch.device("cpu"))
    assert torch.equal(grid.view(2, 3, -1), _pixel_grid(2, 3, 4, torch.float32, torch.device("cpu")).view(2, 3, -1))


# ---- GPU: the stand-alone kernels -----------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 40, 72), (1, 33, 50)])
def test_reconstruct_matches_reference_formula(shape):
    dev = _dev()
    from mgnet_b200.geometry import Camera, Pose
    B, H, W = shape
    pred, tgt, K, depth = _inputs(B, H, W)
    ref_c = (_kinv(K).bmm(_pixel_grid(B, H, W, torch.float32, torch.device("cpu")).view(B, 3, -1)) * depth.view(B, 1, -1)).view(B, 3, H, W)
    cam = Camera(K.to(dev)).to(dev)                                                # as the reference's callers do: the identity Tcw is made on the CPU
    out_c = cam.reconstruct(depth.to(dev), frame="c")
    assert np.array_equal(out_c.cpu().numpy(), ref_c.numpy())                      # camera.py:129-136, bit for bit
    assert torch.equal(cam.reconstruct(depth.to(dev), frame="w"), out_c)           # identity Tcw: Twc @ X is exact
    pose = Pose.from_vec(pred["poses"][:, 0], "euler")
    cam_w = Camera(K.to(dev), Tcw=Pose(pose.mat.to(dev)))
    ref_w = _apply44(_invert44(pose.mat), ref_c)
    assert torch.allclose(cam_w.reconstruct(depth.to(dev), frame="w").cpu(), ref_w, rtol=1e-6, atol=1e-6)   # Twc @ X is a torch (cuBLAS) bmm
    with pytest.raises(ValueError):
        cam.reconstruct(depth.to(dev), frame="x")


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 40, 72), (1, 33, 50)])
def test_project_matches_reference_formula(shape):
    dev = _dev()
    from mgnet_b200.geometry import Camera, Pose
    B, H, W = shape
    pred, tgt, K, depth = _inputs(B, H, W)
    X = (_kinv(K).bmm(_pixel_grid(B, H, W, torch.float32, torch.device("cpu")).view(B, 3, -1)) * depth.view(B, 1, -1)).view(B, 3, H, W)
    pose = Pose.from_vec(pred["poses"][:, 1], "euler")

    def ref_project(Xw, T):            # camera.py:157-182
        Xc = K.bmm((_apply44(T, Xw) if T is not None else Xw).view(B, 3, -1))
        Z = Xc[:, 2].clamp(min=1e-5)
        xn = 2 * (Xc[:, 0] / Z) / (W - 1) - 1.0
        yn = 2 * (Xc[:, 1] / Z) / (H - 1) - 1.0
        return torch.stack([xn, yn], dim=-1).view(B, H, W, 2)

    cam = Camera(K.to(dev), Tcw=Pose(pose.mat.to(dev)))
    assert np.array_equal(cam.project(X.to(dev), frame="w").cpu().numpy(), ref_project(X, pose.mat).numpy())
    assert np.array_equal(cam.project(X.to(dev), frame="c").cpu().numpy(), ref_project(X, None).numpy())
    with pytest.raises(ValueError):
        cam.project(X.to(dev), frame="x")


@pytest.mark.gpu
@pytest.mark.parametrize("padding_mode", ["zeros", "border"])
def test_view_synthesis_uses_both_cameras(padding_mode):
    """camera_utils.py:48-50 lifts with `cam` and projects with `ref_cam`: a Camera.scaled() pair must not be confused."""
    dev = _dev()
    from mgnet_b200.geometry import Camera, Pose, view_synthesis
    B, H, W = 2, 40, 72
    pred, tgt, K, depth = _inputs(B, H, W, pose_scale=0.03)
    pose = Pose.from_vec(pred["poses"][:, 0], "euler")
    cam = Camera(K.clone())                                    # target camera (lifts)
    ref_cam = Camera(K.clone(), Tcw=pose).scaled(0.9, 1.05)    # reference camera with other intrinsics (projects)
    assert not torch.equal(cam.K, ref_cam.K)
    src = tgt["image_prev_orig"]
    want = _synthesize(src, depth, ref_cam.K, _kinv(cam.K), pose.mat, torch.eye(4).repeat(B, 1, 1), padding_mode=padding_mode)
    got = view_synthesis(src.to(dev), depth.to(dev), Camera(ref_cam.K.to(dev), Tcw=Pose(pose.mat.to(dev))), Camera(cam.K.to(dev)),
                         padding_mode=padding_mode)
    assert np.array_equal(got.cpu().numpy(), want.numpy())
    same_k = view_synthesis(src.to(dev), depth.to(dev), Camera(cam.K.to(dev), Tcw=Pose(pose.mat.to(dev))), Camera(cam.K.to(dev)),
                            padding_mode=padding_mode)
    assert not torch.equal(same_k, got)


@pytest.mark.gpu
def test_standalone_ops_refuse_to_drop_gradients():
    dev = _dev()
    from mgnet_b200.geometry import Camera, Pose, view_synthesis
    B, H, W = 1, 33, 50
    pred, tgt, K, depth = _inputs(B, H, W)
    cam = Camera(K.to(dev))
    d = depth.to(dev).requires_grad_(True)
    with pytest.raises(NotImplementedError, match="forward-only"):
        cam.reconstruct(d, frame="c")
    with torch.no_grad():
        X = cam.reconstruct(d, frame="c")              # fine without autograd
    with pytest.raises(NotImplementedError, match="forward-only"):
        cam.project(X.clone().requires_grad_(True), frame="c")
    vec = pred["poses"][:, 0].to(dev).requires_grad_(True)
    ref_cam = Camera(K.to(dev), Tcw=Pose.from_vec(vec, "euler"))
    with pytest.raises(NotImplementedError, match="forward-only"):
        view_synthesis(tgt["image_prev_orig"].to(dev), depth.to(dev), ref_cam, cam)
    with pytest.raises(NotImplementedError, match="forward-only"):
        view_synthesis(tgt["image_prev_orig"].to(dev), d, Camera(K.to(dev)), cam)
    assert view_synthesis(tgt["image_prev_orig"].to(dev), d.detach(), Camera(K.to(dev)), cam).shape == (B, 3, H, W)


@pytest.mark.gpu
def test_warp_ref_image_helper_and_fused_away_helpers():
    dev = _dev()
    from mgnet_b200 import MultiViewPhotometricLoss
    from mgnet_b200.geometry import Camera, Pose
    B, H, W = 2, 40, 72
    pred, tgt, K, depth = _inputs(B, H, W)
    mod = MultiViewPhotometricLoss(0.85, 1.0, 1e-3, True, "min", "zeros")
    pose = Pose.from_vec(pred["poses"][:, 0], "euler")
    out = mod.warp_ref_image([depth.to(dev)], tgt["image_prev_orig"].to(dev), [Camera(K.to(dev))], K.to(dev), Pose(pose.mat.clone()))
    want = _synthesize(tgt["image_prev_orig"], depth, K, _kinv(K), pose.mat, torch.eye(4).repeat(B, 1, 1))
    assert len(out) == 1 and np.array_equal(out[0].cpu().numpy(), want.numpy())
    for name, args in (("ssim", (depth, depth)), ("calc_photometric_loss", ([depth], [depth])), ("reduce_photometric_loss", ([[depth]],)),
                       ("calc_smoothness_loss", ([depth], [depth]))):
        with pytest.raises(NotImplementedError, match="fused"):
            getattr(mod, name)(*args)


@pytest.mark.gpu
def test_fp16_inputs_under_autocast_follow_custom_fwd_contract():
    """mg_net.py:827 decorates the head's losses() with custom_fwd(cast_inputs=torch.float32): fp16 head outputs are cast up, the loss
    computes in fp32 and the gradients come back in the input dtype."""
    dev = _dev()
    from mgnet_b200 import MultiViewPhotometricLoss
    from mgnet_b200.synthetic import make_inputs
    pred, tgt = make_inputs(2, 64, 128, 2, seed=17)
    hp = dict(ssim_loss_weight=0.85, photometric_loss_weight=1.0, smoothing_loss_weight=1e-3, automask_loss=True,
              photometric_reduce_op="min", padding_mode="zeros")
    mod = MultiViewPhotometricLoss(**hp)
    t = {k: v.to(dev) for k, v in tgt.items()}
    d16 = [d.to(dev).half().requires_grad_(True) for d in pred["depth"]]
    p16 = pred["poses"].to(dev).half().requires_grad_(True)
    with torch.autocast(device_type="cuda", dtype=torch.float16):
        out16 = mod({"depth": d16, "poses": p16}, t)
        assert out16["loss_photometric"].dtype == torch.float32
        (out16["loss_photometric"] + out16["loss_smoothness"]).backward()
    sel16 = mod.last_selection.clone()
    d32 = [d.detach().float().requires_grad_(True) for d in d16]
    p32 = p16.detach().float().requires_grad_(True)
    out32 = mod({"depth": d32, "poses": p32}, t)
    (out32["loss_photometric"] + out32["loss_smoothness"]).backward()
    assert out16["loss_photometric"].item() == out32["loss_photometric"].item()
    assert out16["loss_smoothness"].item() == out32["loss_smoothness"].item()
    assert torch.equal(sel16, mod.last_selection)
    for a, b in zip(d16, d32):
        assert a.grad.dtype == torch.float16 and torch.equal(a.grad, b.grad.half())
    assert p16.grad.dtype == torch.float16 and torch.equal(p16.grad, p32.grad.half())


@pytest.mark.gpu
def test_backward_twice_with_retain_graph_is_identical():
    """the coefficient stash lives with the saved tensors, so a second backward over the same graph runs the same kernel."""
    dev = _dev()
    from mgnet_b200 import MultiViewPhotometricLoss
    from mgnet_b200.synthetic import make_inputs
    pred, tgt = make_inputs(1, 48, 64, 2, seed=5)
    mod = MultiViewPhotometricLoss(0.85, 1.0, 1e-3, True, "min", "zeros")
    d = [x.to(dev).requires_grad_(True) for x in pred["depth"]]
    p = pred["poses"].to(dev).requires_grad_(True)
    out = mod({"depth": d, "poses": p}, {k: v.to(dev) for k, v in tgt.items()})
    loss = out["loss_photometric"] + out["loss_smoothness"]
    loss.backward(retain_graph=True)
    g1 = [x.grad.clone() for x in d] + [p.grad.clone()]
    for x in d + [p]:
        x.grad = None
    loss.backward()
    g2 = [x.grad for x in d] + [p.grad]
    for a, b in zip(g1, g2):
        assert torch.equal(a, b)


def test_mean_reduce_with_l1_only_is_refused():
    from mgnet_b200 import MultiViewPhotometricLoss
    with pytest.raises(NotImplementedError):
        MultiViewPhotometricLoss(0.0, 1.0, 1e-3, False, "mean", "zeros")
