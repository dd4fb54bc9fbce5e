"""Image-grid and finite-difference helpers with the reference's signatures (mgnet/geometry/image.py)."""
from functools import lru_cache

import torch
import torch.nn.functional as F

__all__ = ["same_shape", "gradient_x", "gradient_y", "interpolate_image", "match_scales", "meshgrid", "image_grid"]


def same_shape(shape1, shape2):
    return len(shape1) == len(shape2) and all(a == b for a, b in zip(shape1, shape2))


def gradient_x(image):
    """[B,C,H,W] -> [B,C,H,W-1]: left minus right neighbour."""
    return image[..., :-1] - image[..., 1:]


def gradient_y(image):
    """[B,C,H,W] -> [B,C,H-1,W]: upper minus lower neighbour."""
    return image[..., :-1, :] - image[..., 1:, :]


def interpolate_image(image, shape, mode="bilinear", align_corners=True):
    shape = tuple(shape)[-2:]
    if same_shape(tuple(image.shape[-2:]), shape):
        return image
    return F.interpolate(image, size=shape, mode=mode, align_corners=align_corners)


def match_scales(image, targets, num_scales, mode="bilinear", align_corners=True):
    """One (possibly resized) copy of ``image`` per target resolution.  With MGNet's head every scale is
    already full resolution (mg_net.py:799-807), so this returns the same tensor ``num_scales`` times."""
    return [interpolate_image(image, targets[i].shape, mode=mode, align_corners=align_corners)
            for i in range(num_scales)]


@lru_cache(maxsize=None)
def meshgrid(B, H, W, dtype, device, normalized=False):
    lo_x, hi_x, lo_y, hi_y = (-1, 1, -1, 1) if normalized else (0, W - 1, 0, H - 1)
    xs = torch.linspace(lo_x, hi_x, W, device=device, dtype=dtype)
    ys = torch.linspace(lo_y, hi_y, H, device=device, dtype=dtype)
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    return gx.repeat(B, 1, 1), gy.repeat(B, 1, 1)


@lru_cache(maxsize=None)
def image_grid(B, H, W, dtype, device, normalized=False):
    gx, gy = meshgrid(B, H, W, dtype, device, normalized=normalized)
    return torch.stack([gx, gy, torch.ones_like(gx)], dim=1)
