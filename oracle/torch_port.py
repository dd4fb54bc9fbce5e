"""ATen-level CPU port of the reference loss.  TEST / BASELINE INFRASTRUCTURE ONLY.

A second restatement of the reference path -- this one issues the same sequence of PyTorch (ATen)
operators as ``mgnet/modeling/loss.py:111-294`` + ``mgnet/geometry/*`` (bmm, grid_sample, reflect pad +
avg_pool2d, boolean-mask means) so that (a) its results are bit-identical to the reference on the same
host (checked in tests/test_torch_port.py whenever /root/reference is mounted) and (b) its run time is the
run time of the reference's own CPU implementation.  bench.py uses it for ``cpu_baseline`` and
``--impl reference`` (kind "port": the Python reference itself cannot travel to the GPU box); the GPU
tests use its autograd gradients as a second checker next to oracle/mgvs_oracle.c.

Written as flat functions (the reference is a class hierarchy); never imported by mgnet_b200/.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

__all__ = ["reference_loss"]

_grid_cache = {}


def _pixel_grid(B, H, W, dtype, device):
    """[B,3,H*W] homogeneous pixel coordinates (u, v, 1) (image.py:138-199)."""
    key = (B, H, W, dtype, str(device))
    if key not in _grid_cache:
        us = torch.linspace(0, W - 1, W, device=device, dtype=dtype)
        vs = torch.linspace(0, H - 1, H, device=device, dtype=dtype)
        vv, uu = torch.meshgrid(vs, us, indexing="ij")
        g = torch.stack([uu, vv, torch.ones_like(uu)], 0).reshape(1, 3, H * W).repeat(B, 1, 1)
        _grid_cache[key] = g
    return _grid_cache[key]


def _rotation(angles):
    """euler2mat (pose_utils.py:9-38): Rx @ Ry @ Rz from [B,3]."""
    x, y, z = angles[:, 0], angles[:, 1], angles[:, 2]
    zero = z.detach() * 0
    one = zero.detach() + 1
    cz, sz, cy, sy, cx, sx = torch.cos(z), torch.sin(z), torch.cos(y), torch.sin(y), torch.cos(x), torch.sin(x)
    B = angles.size(0)
    zm = torch.stack([cz, -sz, zero, sz, cz, zero, zero, zero, one], 1).view(B, 3, 3)
    ym = torch.stack([cy, zero, sy, zero, one, zero, -sy, zero, cy], 1).view(B, 3, 3)
    xm = torch.stack([one, zero, zero, zero, cx, -sx, zero, sx, cx], 1).view(B, 3, 3)
    return xm.bmm(ym).bmm(zm)


def _pose44(vec):
    """Pose.from_vec(vec, "euler") (pose.py:41-47)."""
    m34 = torch.cat([_rotation(vec[:, 3:]), vec[:, :3].unsqueeze(-1)], 2)
    T = torch.eye(4, device=vec.device, dtype=vec.dtype).repeat([len(vec), 1, 1])
    T[:, :3, :3] = m34[:, :3, :3]
    T[:, :3, -1] = m34[:, :3, -1]
    return T


def _invert44(T):
    Ti = torch.eye(4, device=T.device, dtype=T.dtype).repeat([len(T), 1, 1])
    Ti[:, :3, :3] = T[:, :3, :3].transpose(-2, -1)
    Ti[:, :3, -1] = torch.bmm(-1.0 * Ti[:, :3, :3], T[:, :3, -1].unsqueeze(-1)).squeeze(-1)
    return Ti


def _apply44(T, pts):
    B, _, H, W = pts.shape
    return (T[:, :3, :3].bmm(pts.view(B, 3, -1)) + T[:, :3, -1].unsqueeze(-1)).view(B, 3, H, W)


def _kinv(K):
    Ki = K.clone()
    Ki[:, 0, 0] = 1.0 / K[:, 0, 0]
    Ki[:, 1, 1] = 1.0 / K[:, 1, 1]
    Ki[:, 0, 2] = -1.0 * K[:, 0, 2] / K[:, 0, 0]
    Ki[:, 1, 2] = -1.0 * K[:, 1, 2] / K[:, 1, 1]
    return Ki


def _synthesize(src, depth, K, Ki, T_src, T_tgt_inv, padding_mode="zeros"):
    """view_synthesis (camera_utils.py:24-54) = reconstruct("w") -> project("w") -> grid_sample."""
    B, _, H, W = depth.shape
    rays = Ki.bmm(_pixel_grid(B, H, W, depth.dtype, depth.device)).view(B, 3, H, W)
    world = _apply44(T_tgt_inv, rays * depth)
    cam_pts = K.bmm(_apply44(T_src, world).view(B, 3, -1))
    X, Y, Z = cam_pts[:, 0], cam_pts[:, 1], cam_pts[:, 2].clamp(min=1e-5)
    xn = 2 * (X / Z) / (W - 1) - 1.0
    yn = 2 * (Y / Z) / (H - 1) - 1.0
    coords = torch.stack([xn, yn], dim=-1).view(B, H, W, 2)
    return F.grid_sample(src, coords, mode="bilinear", padding_mode=padding_mode, align_corners=True)


def _ssim(x, y, c1=1e-4, c2=9e-4):
    x, y = F.pad(x, [1, 1, 1, 1], "reflect"), F.pad(y, [1, 1, 1, 1], "reflect")
    mu_x, mu_y = F.avg_pool2d(x, 3, stride=1), F.avg_pool2d(y, 3, stride=1)
    mxy, mxx, myy = mu_x * mu_y, mu_x.pow(2), mu_y.pow(2)
    sx = F.avg_pool2d(x.pow(2), 3, stride=1) - mxx
    sy = F.avg_pool2d(y.pow(2), 3, stride=1) - myy
    sxy = F.avg_pool2d(x * y, 3, stride=1) - mxy
    val = (2 * mxy + c1) * (2 * sxy + c2) / ((mxx + myy + c1) * (sx + sy + c2))
    return torch.clamp((1.0 - val) / 2.0, 0.0, 1.0)


def _photometric(est, tgt, w_ssim):
    l1 = torch.abs(est - tgt)
    if not w_ssim > 0.0:
        return l1       # the raw 3-channel map (loss.py:195-196): the min then runs over 3 channels per list entry
    return w_ssim * _ssim(est, tgt).mean(1, True) + (1 - w_ssim) * l1.mean(1, True)


def reference_loss(predictions, targets, ssim_loss_weight=0.85, photometric_loss_weight=1.0,
                   smoothing_loss_weight=1e-3, automask_loss=True, return_selection=False, padding_mode="zeros"):
    """Same dictionaries in, same dictionary out as the reference's forward (loss.py:111-154)."""
    inv = predictions["depth"]
    n = len(inv)
    poses = predictions["poses"]
    sources = [targets["image_prev_orig"], targets["image_next_orig"]]
    tgt = targets["image_orig"]
    K = targets["camera_matrix"][:, :3, :3].float()
    Ki = _kinv(K)
    depths = [1.0 / d.clamp(min=1e-6) for d in inv]
    ident_inv = _invert44(torch.eye(4, dtype=torch.float, device=tgt.device).repeat([len(K), 1, 1]))
    maps = [[] for _ in range(n)]
    for s, src in enumerate(sources):
        T = _pose44(poses[:, s].float())
        for i in range(n):
            maps[i].append(_photometric(_synthesize(src, depths[i], K, Ki, T, ident_inv, padding_mode), tgt, ssim_loss_weight))
            if i == 0 and automask_loss:
                ident = _photometric(src, tgt, ssim_loss_weight)
            if automask_loss:
                maps[i].append(ident)
    mask = targets["reprojection_mask"] if "reprojection_mask" in targets else torch.ones_like(maps[0][0], dtype=torch.bool)
    sel = []
    photo = 0
    for i in range(n):
        mn, idx = torch.cat(maps[i], 1).min(1, True)
        photo = photo + mn[mask].mean()
        sel.append(idx)
    photo = photo / n
    # smoothness (depth.py:18-51, loss.py:257-294)
    wx = torch.exp(-torch.mean(torch.abs(tgt[:, :, :, :-1] - tgt[:, :, :, 1:]), 1, keepdim=True))
    wy = torch.exp(-torch.mean(torch.abs(tgt[:, :, :-1, :] - tgt[:, :, 1:, :]), 1, keepdim=True))
    smooth = 0
    for i in range(n):
        dn = inv[i] / inv[i].mean(2, True).mean(3, True).clamp(min=1e-6)
        gx = (dn[:, :, :, :-1] - dn[:, :, :, 1:]) * wx
        gy = (dn[:, :, :-1, :] - dn[:, :, 1:, :]) * wy
        smooth = smooth + (gx[mask[:, :, :, :-1]].abs().mean() + gy[mask[:, :, :-1, :]].abs().mean()) / 2 ** i
    smooth = smooth / n
    out = {"loss_photometric": photo * photometric_loss_weight, "loss_smoothness": smooth * smoothing_loss_weight}
    if return_selection:
        out["selection"] = torch.stack([s[:, 0].to(torch.uint8) for s in sel], 0)
    return out


# ---- DGC depth rescaling (reference mgnet/postprocessing/depth_post_proc.py:11-185), same ATen operator sequence ------
def _dgc_surface_normal(P):
    """_get_surface_normal (depth_post_proc.py:107-151), nei = 1."""
    ctr = P[:, :, 1:-1, 1:-1]
    def nb(dy, dx):
        H, W = P.shape[-2:]
        return P[:, :, 1 + dy:H - 1 + dy, 1 + dx:W - 1 + dx] - ctr
    pairs = ((nb(0, -1), nb(-1, 0)), (nb(0, 1), nb(1, 0)), (nb(-1, -1), nb(1, -1)), (nb(-1, 1), nb(1, 1)))
    ns = [F.normalize(torch.cross(a, b, dim=1), dim=1).unsqueeze(0) for a, b in pairs]
    n = F.normalize(torch.cat(ns, dim=0).mean(0), dim=1)
    return F.pad(n, (1, 1, 1, 1), "replicate")


def reference_dgc(depth_logits, camera_matrix, real_camera_height, panoptic_seg=None, road_class_id=-1,
                  depth_filter_class_ids=None):
    """get_depth_prediction(..., use_dgc_scaling=True) as flat ATen calls on whatever device the inputs live on.
    depth_logits [1,1,H,W] is modified in place like the reference does.  Returns (depth [H,W], points [3,H,W], scale)."""
    import math
    B, _, H, W = depth_logits.shape
    K = camera_matrix.reshape(-1, 3, 3).to(depth_logits.device).float()
    Kinv = K.clone()
    fx, fy, cx, cy = K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2]
    Kinv[:, 0, 0] = 1.0 / fx
    Kinv[:, 1, 1] = 1.0 / fy
    Kinv[:, 0, 2] = -1.0 * cx / fx
    Kinv[:, 1, 2] = -1.0 * cy / fy
    grid = _pixel_grid(B, H, W, depth_logits.dtype, depth_logits.device)
    P = Kinv.bmm(grid).view(B, 3, H, W) * depth_logits
    N = _dgc_surface_normal(P)
    if panoptic_seg is not None:
        ground = panoptic_seg == road_class_id
    else:
        thr = math.cos(math.radians(5))
        vertical = torch.cat((torch.zeros_like(depth_logits), torch.ones_like(depth_logits), torch.zeros_like(depth_logits)), 1)
        cs = torch.nn.CosineSimilarity(dim=1, eps=1e-6)(N, vertical).unsqueeze(1)
        ground = ((cs > thr) | (cs < -thr)).masked_fill(P[:, 1].unsqueeze(1) <= 0, False)
    heights = (P * N).sum(1).abs().unsqueeze(1)
    med = torch.median(torch.masked_select(heights, ground)).unsqueeze(0)
    scale = torch.reciprocal(med).mul_(real_camera_height.to(depth_logits.device))
    depth_logits *= scale
    P *= scale
    P = P.squeeze(0)
    out = depth_logits.squeeze(0).squeeze(0)
    if panoptic_seg is not None:
        for cid in (depth_filter_class_ids or []):
            out[panoptic_seg == cid] = 0
            P[:, panoptic_seg == cid] = float("nan")
    return out, P, scale


def reference_uncertainty(losses, log_vars):
    """The uncertainty epilogue of MGNet.forward (mg_net.py:360-372) as the same eager expression, minus detectron2's event
    storage: returns (weighted dict, log dict with the floats the reference passes to put_scalar)."""
    import math
    out, log = {}, {}
    idx = 0
    for key, value in losses.items():
        log[key + "_raw"] = value.detach().item()
        tau = 1.0 if key == "loss_sem_seg" else 0.5
        out[key] = tau * torch.exp(-log_vars[idx]) * value + 0.5 * log_vars[idx]
        log[key + "_uncertainty"] = math.exp(log_vars[idx].detach().item())
        idx = idx + 1
    return out, log
