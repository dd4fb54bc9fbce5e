"""Inverse-depth helpers with the reference's signatures (mgnet/geometry/depth.py:11-51)."""
import torch

from .image import gradient_x, gradient_y

__all__ = ["inv2depth", "calc_smoothness"]


def inv2depth(inv_depth):
    if isinstance(inv_depth, (tuple, list)):
        return [inv2depth(d) for d in inv_depth]
    return 1.0 / inv_depth.clamp(min=1e-6)


def calc_smoothness(inv_depths, image, num_scales):
    """Edge-aware first-order smoothness terms of the mean-normalised inverse depths.
    (Inside the fused loss this is evaluated in factorised form by fwd_kernel / bwd_kernel.)"""
    normed = [d / d.mean(2, True).mean(3, True).clamp(min=1e-6) for d in inv_depths]
    wx = torch.exp(-gradient_x(image).abs().mean(1, keepdim=True))
    wy = torch.exp(-gradient_y(image).abs().mean(1, keepdim=True))
    sx = [gradient_x(normed[i]) * wx for i in range(num_scales)]
    sy = [gradient_y(normed[i]) * wy for i in range(num_scales)]
    return sx, sy
