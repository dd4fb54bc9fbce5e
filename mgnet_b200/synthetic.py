"""Seeded synthetic inputs for the view-synthesis loss.

Every tensor is drawn on the CPU from its own ``torch.Generator`` so that a CPU run and a
GPU run (after ``.to(device)``) see identical bits.  Shapes and value ranges follow what
``MGNet.forward`` hands to the depth-head loss (reference ``mgnet/modeling/mg_net.py:318-349``,
``:799-823``; ``mgnet/modeling/layers.py:165-166`` for the ``0.01 *`` pose scale).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

__all__ = ["make_inputs", "kitti_like_K", "bytes_per_pixel", "snap_pose_trig", "quantize_images", "pack_mask"]


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def _smooth_field(shape, g, cell=16):
    """Band-limited field in [0,1]: bilinear upsample of a coarse uniform grid."""
    b, c, h, w = shape
    coarse = torch.rand(b, c, h // cell + 2, w // cell + 2, generator=g, dtype=torch.float32)
    return F.interpolate(coarse, size=(h, w), mode="bilinear", align_corners=True)


def kitti_like_K(B: int, H: int, W: int) -> torch.Tensor:
    """4x4 camera matrix with KITTI-like normalised intrinsics (dataset_mapper.py:247-251)."""
    K = torch.eye(4, dtype=torch.float32).repeat(B, 1, 1)
    K[:, 0, 0] = 0.58 * W
    K[:, 1, 1] = 1.92 * H
    K[:, 0, 2] = 0.5 * W - 0.5
    K[:, 1, 2] = 0.5 * H - 0.5
    return K


def snap_pose_trig(poses: torch.Tensor) -> torch.Tensor:
    """Nudges every Euler angle (by ~2^-12 relative steps) to a value whose fp32 sin/cos on THIS host's torch-CPU
    equals the correctly rounded value.

    torch-CPU evaluates sin/cos with MKL VML (1-ulp accurate, CPU-dispatch dependent), torch-CUDA with
    libdevice, and the kernels here with fp64 rounded to fp32; the three agree on ~95% of arguments.
    Bit-exact selection-mask parity is only defined where they agree, so parity fixtures use snapped
    angles (see DESIGN.md "trig boundary").
    """
    p = poses.detach().clone().cpu().float().contiguous()
    assert p.dim() == 3 and p.shape[-1] == 6
    for _ in range(64):
        bad_any = False
        for s in range(p.shape[1]):
            # same slicing as the reference: Pose.from_vec(poses[:, s].float()) -> euler2mat(vec[:, 3:])
            # (MKL VML results can depend on the position inside the strided view, so mirror it exactly)
            rot = p[:, s].float()[:, 3:]
            for k in range(3):
                a = rot[:, k]
                ok = (torch.cos(a) == torch.cos(a.double()).float()) & (torch.sin(a) == torch.sin(a.double()).float())
                if not bool(ok.all()):
                    bad_any = True
                    col = p[:, s, 3 + k]
                    # cos barely moves per ulp of a small angle: step by ~2^-12 relative instead
                    col[~ok] = col[~ok] + (col[~ok].abs() * (2.0 ** -12) + 1e-9)
        if not bad_any:
            break
    return p


def make_inputs(
    B: int,
    H: int,
    W: int,
    n: int = 3,
    seed: int = 0,
    noise: float = 0.2,
    pose_scale: float = 0.01,
    with_mask: bool = True,
    shift_sources: bool = False,
    mask_keep: float = 0.9,
    snap_trig: bool = True,
):
    """Returns (predictions, targets) dicts of CPU fp32 tensors in the reference's layout.

    noise          fraction of white noise blended on top of the band-limited images
                   (0.2 for forward / throughput tests, 0.0 for gradient-parity tests).
    shift_sources  make the source frames horizontally shifted copies of the target
                   (plus a little independent structure) so the warp is informative.
    """
    def image(seed_k):
        g = _gen(seed_k)
        smooth = _smooth_field((B, 3, H, W), g)
        if noise > 0:
            smooth = (1.0 - noise) * smooth + noise * torch.rand(B, 3, H, W, generator=g)
        return smooth.clamp_(0.0, 1.0).contiguous()

    tgt = image(seed * 1000 + 1)
    if shift_sources:
        prev = (0.9 * torch.roll(tgt, 2, dims=3) + 0.1 * image(seed * 1000 + 2)).contiguous()
        nxt = (0.9 * torch.roll(tgt, -3, dims=3) + 0.1 * image(seed * 1000 + 3)).contiguous()
    else:
        prev = image(seed * 1000 + 2)
        nxt = image(seed * 1000 + 3)

    inv_depths = []
    for i in range(n):
        g = _gen(seed * 1000 + 10 + i)
        f = _smooth_field((B, 1, H, W), g, cell=8)
        inv_depths.append((0.05 + 1.9 * f).contiguous())

    g = _gen(seed * 1000 + 50)
    poses = (pose_scale * torch.randn(B, 2, 6, generator=g, dtype=torch.float32)).contiguous()
    if snap_trig:
        poses = snap_pose_trig(poses)

    targets = {
        "image_orig": tgt,
        "image_prev_orig": prev,
        "image_next_orig": nxt,
        "camera_matrix": kitti_like_K(B, H, W),
    }
    if with_mask:
        g = _gen(seed * 1000 + 60)
        targets["reprojection_mask"] = torch.rand(B, 1, H, W, generator=g) < mask_keep
    predictions = {"depth": inv_depths, "poses": poses}
    return predictions, targets


IMAGE_KEYS = ("image_orig", "image_prev_orig", "image_next_orig")


def quantize_images(targets):
    """Returns (targets_u8, targets_f32): the three images as the data loader's uint8 tensors and as what the
    reference's caller makes of them, ``x.float() / 255.0`` (mg_net.py:320-335).  Other entries are shared."""
    tu, tf = dict(targets), dict(targets)
    for k in IMAGE_KEYS:
        u = (targets[k] * 255.0).round().clamp_(0, 255).to(torch.uint8).contiguous()
        tu[k] = u
        tf[k] = (u.float() / 255.0).contiguous()
    return tu, tf


def pack_mask(mask: torch.Tensor) -> torch.Tensor:
    """bool [B,1,H,W] -> bit-packed uint8 [B,1,H,ceil(W/8)] (numpy.packbits along the row, most significant bit first): the
    form ``MultiViewPhotometricLoss`` accepts as ``targets["reprojection_mask"]`` to cut its host-to-device bytes by 8."""
    import numpy as np
    return torch.from_numpy(np.packbits(mask.cpu().numpy().astype(bool), axis=-1)).contiguous()


def bytes_per_pixel(n: int, S: int = 2, mask: bool = True, fwd_only: bool = False) -> int:
    """Algorithmic HBM bytes per target pixel."""
    m = 1 if mask else 0
    fwd = 12 + 12 * S + 4 * n + m
    return fwd if fwd_only else 2 * fwd + 4 * n


def make_dgc_inputs(H: int, W: int, seed: int = 0, scale_true: float = 7.5, cam_height: float = 1.65,
                    with_panoptic: bool = True, road_class_id: int = 0, label_divisor: int = 1000):
    """Synthetic inference-time scene for the DGC depth rescaling (reference depth_post_proc.py:11-104).

    A pinhole camera ``cam_height`` metres above a slightly bumpy ground plane (y points down), a far "sky" above the
    horizon and two box-like obstacles; the network's depth is the metric depth divided by ``scale_true`` (what an
    unscaled self-supervised depth head produces), so the recovered scale factor should be close to ``scale_true``.
    Returns a dict of CPU tensors: depth [1,1,H,W], camera_matrix [1,3,3], real_camera_height [1], and (optionally)
    panoptic_seg [H,W] int64 in the reference's ``trainId * label_divisor`` convention (road / obstacle / sky).
    """
    g = _gen(seed * 1000 + 70)
    K = kitti_like_K(1, H, W)[:, :3, :3].contiguous()
    fx, fy, cx, cy = K[0, 0, 0], K[0, 1, 1], K[0, 0, 2], K[0, 1, 2]
    v = torch.arange(H, dtype=torch.float32).view(H, 1).expand(H, W)
    u = torch.arange(W, dtype=torch.float32).view(1, W).expand(H, W)
    ry = (v - cy) / fy
    bump = 1.0 + 0.02 * (_smooth_field((1, 1, H, W), g, cell=8)[0, 0] - 0.5)
    z_ground = cam_height * bump / ry.clamp(min=1e-3)
    far = 60.0 + 20.0 * _smooth_field((1, 1, H, W), g, cell=16)[0, 0]
    z = torch.where(ry > 0.02, torch.minimum(z_ground, far), far)
    pan = torch.full((H, W), 10 * label_divisor, dtype=torch.int64)          # sky / far
    pan[(ry > 0.02) & (z_ground < far)] = road_class_id
    # two fronto-parallel obstacles standing on the ground
    for k, (u0, u1, zobj) in enumerate(((0.15, 0.3, 9.0), (0.6, 0.8, 15.0))):
        cols = (u >= u0 * W) & (u < u1 * W)
        top = cy - 1.2 * fy / zobj
        rows = (v >= top) & (z > zobj)
        sel = cols & rows
        z = torch.where(sel, torch.full_like(z, zobj), z)
        pan[sel] = (13 + k) * label_divisor + 1 + k
    z = z * (1.0 + 0.004 * (torch.rand(H, W, generator=g) - 0.5))
    depth = (z / scale_true).view(1, 1, H, W).contiguous()
    out = {
        "depth": depth,
        "camera_matrix": K,
        "real_camera_height": torch.tensor([cam_height], dtype=torch.float32),
    }
    if with_panoptic:
        out["panoptic_seg"] = pan.contiguous()
    return out
