"""Pinhole ``Camera`` with the reference's interface (mgnet/geometry/camera.py:16-182).

``reconstruct`` and ``project`` run sm_100a kernels (reconstruct_kernel / project_kernel).  They are forward
only and raise when an input requires grad (the reference ops are differentiable; gradients of the training path
flow through the fused loss, where the same arithmetic is part of fwd_kernel / bwd_kernel).  CPU tensors raise."""
import ctypes
from functools import lru_cache

import torch
import torch.nn as nn

from .. import _lib
from .camera_utils import require_no_grad, scale_intrinsics
from .pose import Pose

__all__ = ["Camera"]


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


class Camera(nn.Module):
    def __init__(self, K, Tcw=None):
        super().__init__()
        self.K = K
        self.Tcw = Pose.identity(len(K)) if Tcw is None else Tcw

    def __len__(self):
        return len(self.K)

    def to(self, *args, **kwargs):
        self.K = self.K.to(*args, **kwargs)
        self.Tcw = self.Tcw.to(*args, **kwargs)
        return self

    fx = property(lambda self: self.K[:, 0, 0])
    fy = property(lambda self: self.K[:, 1, 1])
    cx = property(lambda self: self.K[:, 0, 2])
    cy = property(lambda self: self.K[:, 1, 2])

    @property
    @lru_cache()
    def Twc(self):
        return self.Tcw.inverse()

    @property
    @lru_cache()
    def Kinv(self):
        """Closed-form inverse of a skew-free K (the other entries are passed through)."""
        Ki = self.K.clone()
        Ki[:, 0, 0] = 1.0 / self.fx
        Ki[:, 1, 1] = 1.0 / self.fy
        Ki[:, 0, 2] = -1.0 * self.cx / self.fx
        Ki[:, 1, 2] = -1.0 * self.cy / self.fy
        return Ki

    def scaled(self, x_scale, y_scale=None):
        y_scale = x_scale if y_scale is None else y_scale
        if x_scale == 1.0 and y_scale == 1.0:
            return self
        return Camera(scale_intrinsics(self.K.clone(), x_scale, y_scale), Tcw=self.Tcw)

    def reconstruct(self, depth, frame="w"):
        """depth [B,1,H,W] -> 3-D points [B,3,H,W] in the camera ("c") or world ("w") frame."""
        B, C, H, W = depth.shape
        assert C == 1
        if frame not in ("c", "w"):
            raise ValueError("Unknown reference frame {}".format(frame))
        if not depth.is_cuda:
            raise RuntimeError("Camera.reconstruct runs only on CUDA (sm_100a); there is no CPU fallback")
        require_no_grad("Camera.reconstruct", depth, self.K, self.Tcw.mat)
        depth = depth.float().contiguous()
        K = self.K.to(depth.device).float().contiguous()
        pts = torch.empty(B, 3, H, W, device=depth.device, dtype=torch.float32)
        with torch.cuda.device(depth.device):
            _lib.check(_lib.lib().mgvs_reconstruct(B, H, W, depth.data_ptr(), K.data_ptr(), K.stride(0), K.stride(1),
                                                   pts.data_ptr(), _stream(depth)), "mgvs_reconstruct")
        return pts if frame == "c" else self.Twc @ pts

    def project(self, X, frame="w"):
        """3-D points [B,3,H,W] -> sample coordinates [B,H,W,2] normalised to [-1,1] (align_corners=True)."""
        B, C, H, W = X.shape
        assert C == 3
        if frame not in ("c", "w"):
            raise ValueError("Unknown reference frame {}".format(frame))
        if not X.is_cuda:
            raise RuntimeError("Camera.project runs only on CUDA (sm_100a); there is no CPU fallback")
        require_no_grad("Camera.project", X, self.K, self.Tcw.mat)
        X = X.float().contiguous()
        K = self.K.to(X.device).float().contiguous()
        pose34 = self.Tcw.mat[:, :3, :4].to(X.device).float().contiguous() if frame == "w" else None
        coords = torch.empty(B, H, W, 2, device=X.device, dtype=torch.float32)
        with torch.cuda.device(X.device):
            _lib.check(_lib.lib().mgvs_project(B, H, W, X.data_ptr(), K.data_ptr(), K.stride(0), K.stride(1),
                                               pose34.data_ptr() if pose34 is not None else None, coords.data_ptr(),
                                               _stream(X)), "mgvs_project")
        return coords
