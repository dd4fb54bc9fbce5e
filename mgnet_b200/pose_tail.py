"""The tail of the reference's ``PoseCNN.forward`` (mgnet/modeling/layers.py:164-166):

    out = out.mean(3).mean(2)
    out = 0.01 * out.view(out.size(0), self.num_context_images, 6)

as one sm_100a kernel forward and one backward (``mgvs_pose_tail_forward / _backward``) instead of three ATen launches each way.
Deterministic (fixed-order fp64 accumulation).  The result is ``predictions["poses"]`` of the view-synthesis loss.  CUDA only.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .ops import launch_counter

__all__ = ["pose_tail"]


class _PoseTail(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale):
        B, C, h, w = x.shape
        with torch.cuda.device(x.device):
            out = torch.empty(B, C, dtype=torch.float32, device=x.device)
            _lib.check(_lib.lib().mgvs_pose_tail_forward(B * C, h, w, x.data_ptr(), scale, out.data_ptr(),
                                                        ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)), "mgvs_pose_tail_forward")
        launch_counter.n += 1
        ctx.shape, ctx.scale = (B, C, h, w), scale
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        B, C, h, w = ctx.shape
        g = g.float().contiguous()
        with torch.cuda.device(g.device):
            gx = torch.empty(B, C, h, w, dtype=torch.float32, device=g.device)
            _lib.check(_lib.lib().mgvs_pose_tail_backward(B * C, h, w, g.data_ptr(), ctx.scale, gx.data_ptr(),
                                                         ctypes.c_void_p(torch.cuda.current_stream(g.device).cuda_stream)), "mgvs_pose_tail_backward")
        launch_counter.n += 1
        return gx, None


def pose_tail(features: torch.Tensor, num_context_images: int = 2, scale: float = 0.01) -> torch.Tensor:
    """``features``: the [B, 6 * num_context_images, h, w] output of PoseCNN.conv4 (any float dtype; computed in fp32 like the loss
    that consumes it, mg_net.py:827).  Returns [B, num_context_images, 6] fp32 -- ``predictions["poses"]``."""
    if features.dim() != 4 or features.shape[1] != 6 * num_context_images:
        raise ValueError("expected [B, %d, h, w], got %s" % (6 * num_context_images, tuple(features.shape)))
    if not features.is_cuda:
        raise RuntimeError("pose_tail runs only on CUDA (sm_100a); there is no CPU fallback")
    out = _PoseTail.apply(features.float().contiguous(), float(scale))
    return out.view(features.shape[0], num_context_images, 6)
