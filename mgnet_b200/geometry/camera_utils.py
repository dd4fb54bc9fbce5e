"""Intrinsics helpers and ``view_synthesis`` with the reference's signatures
(mgnet/geometry/camera_utils.py:10-54).  ``view_synthesis`` runs the sm_100a kernel
``view_synthesis_kernel``.  It is FORWARD ONLY -- gradients of the training path flow through the fused
loss -- and says so: an input that requires grad raises instead of silently returning a tensor without
``grad_fn`` (the reference op is an ordinary autograd graph).  CPU tensors raise, there is no fallback."""
import ctypes

import torch

from .. import _lib

__all__ = ["construct_K", "scale_intrinsics", "view_synthesis"]


def construct_K(fx, fy, cx, cy, dtype=torch.float, device=None):
    return torch.tensor([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=dtype, device=device)


def scale_intrinsics(K, x_scale, y_scale):
    """In-place rescale for a resized image; principal point follows the pixel-centre convention."""
    K[..., 0, 0] *= x_scale
    K[..., 1, 1] *= y_scale
    K[..., 0, 2] = (K[..., 0, 2] + 0.5) * x_scale - 0.5
    K[..., 1, 2] = (K[..., 1, 2] + 0.5) * y_scale - 0.5
    return K


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def require_no_grad(what, *tensors):
    """The stand-alone geometry kernels have no backward: refuse to drop a gradient silently."""
    if torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors):
        raise NotImplementedError(
            "%s is forward-only in mgnet_b200 (the reference op is differentiable): an input requires grad.  Call it under "
            "torch.no_grad() / on detached tensors, or take gradients through MultiViewPhotometricLoss, whose fused backward "
            "covers this arithmetic." % what)


def view_synthesis(ref_image, depth, ref_cam, cam, mode="bilinear", padding_mode="zeros", return_coords=False):
    assert depth.size(1) == 1
    if mode != "bilinear":
        raise NotImplementedError("view_synthesis kernel implements mode='bilinear' only")
    if padding_mode not in _lib.PADDING_MODES:
        raise ValueError("padding_mode must be 'zeros', 'border' or 'reflection', got %r" % (padding_mode,))
    if not (ref_image.is_cuda and depth.is_cuda):
        raise RuntimeError("view_synthesis runs only on CUDA (sm_100a); there is no CPU fallback")
    B, _, H, W = depth.shape
    # target camera pose must be the identity as in the loss (cams[0], loss.py:127): the kernel lifts
    # with K^-1 only and applies ref_cam.Tcw
    tcw = cam.Tcw.mat.float()                    # may still live on the CPU (Camera(K) builds its identity pose there)
    if not torch.equal(tcw, torch.eye(4, device=tcw.device, dtype=torch.float32).expand_as(tcw)):
        raise NotImplementedError("view_synthesis kernel expects the target camera at the identity pose")
    require_no_grad("view_synthesis", ref_image, depth, ref_cam.K, ref_cam.Tcw.mat, cam.K)
    ref_image = ref_image.float().contiguous()
    depth = depth.float().contiguous()
    dev = depth.device                           # (a Camera built without .to(device) keeps its identity pose on the CPU)
    K = ref_cam.K.to(dev).float().contiguous()   # projects (camera_utils.py:50)
    Kl = cam.K.to(dev).float().contiguous()      # back-projects (camera_utils.py:48); differs from K for Camera.scaled users
    if Kl.shape[0] != B or K.shape[0] != B:
        raise ValueError("camera batch sizes (%d, %d) do not match depth (%d)" % (Kl.shape[0], K.shape[0], B))
    pose34 = ref_cam.Tcw.mat[:, :3, :4].to(dev).float().contiguous()
    warped = torch.empty_like(ref_image)
    coords = torch.empty(B, H, W, 2, device=depth.device, dtype=torch.float32) if return_coords else None
    with torch.cuda.device(depth.device):
        _lib.check(_lib.lib().mgvs_view_synthesis_ex(
            B, H, W, ref_image.data_ptr(), depth.data_ptr(), K.data_ptr(), K.stride(0), K.stride(1),
            Kl.data_ptr(), Kl.stride(0), Kl.stride(1), pose34.data_ptr(),
            _lib.PADDING_MODES[padding_mode], warped.data_ptr(), coords.data_ptr() if coords is not None else None, _stream(depth)), "mgvs_view_synthesis")
    return (warped, coords) if return_coords else warped
