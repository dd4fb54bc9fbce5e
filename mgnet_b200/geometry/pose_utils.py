"""Euler / SE(3) helpers with the reference's signatures (mgnet/geometry/pose_utils.py:9-59).

Small differentiable torch expressions on [B,...] tensors: host-side plumbing, not the hot path
(inside the fused loss the same arithmetic runs in prep_kernel / pose_reduce_kernel).
"""
import torch

__all__ = ["euler2mat", "pose_vec2mat", "invert_pose"]


def _axis_rotation(axis, c, s):
    o, z = torch.ones_like(c), torch.zeros_like(c)
    rows = {
        "x": (o, z, z, z, c, -s, z, s, c),
        "y": (c, z, s, z, o, z, -s, z, c),
        "z": (c, -s, z, s, c, z, z, z, o),
    }[axis]
    return torch.stack(rows, dim=1).view(-1, 3, 3)


def euler2mat(angle):
    """[B,3] (rx, ry, rz) -> [B,3,3] rotation Rx @ Ry @ Rz (right-handed)."""
    mats = [_axis_rotation(a, torch.cos(angle[:, k]), torch.sin(angle[:, k])) for k, a in enumerate("xyz")]
    return mats[0].bmm(mats[1]).bmm(mats[2])


def pose_vec2mat(vec, mode="euler"):
    """[B,6] (tx,ty,tz,rx,ry,rz) -> [B,3,4] (R|t)."""
    if mode is None:
        return vec
    if mode != "euler":
        raise ValueError("Rotation mode not supported {}".format(mode))
    return torch.cat([euler2mat(vec[:, 3:]), vec[:, :3].unsqueeze(-1)], dim=2)


def invert_pose(T):
    """[B,4,4] rigid transform -> its inverse [R^T, -R^T t]."""
    Rt = T[:, :3, :3].transpose(-2, -1)
    out = torch.eye(4, device=T.device, dtype=T.dtype).repeat(len(T), 1, 1)
    out[:, :3, :3] = Rt
    out[:, :3, 3] = torch.bmm(-1.0 * Rt, T[:, :3, 3:4]).squeeze(-1)
    return out
