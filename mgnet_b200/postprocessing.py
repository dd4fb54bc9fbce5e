"""Drop-in for ``mgnet.postprocessing.get_depth_prediction`` (reference depth_post_proc.py:11-71): DGC depth
rescaling at inference time.

Same call signature, same in-place behaviour (``depth_logits`` is rescaled in place and returned squeezed), same
assertions -- but Camera.reconstruct(frame="c"), the surface normals, the ground mask, ``masked_select`` + ``median``
and the rescaling run in four small sm_100a kernels behind the C ABI (include/mgvs.h: ``mgvs_dgc_rescale``), with
no host synchronisation (the reference's ``masked_select`` syncs) and results bit-identical to the reference on CPU.
CUDA only: CPU tensors raise (no fallback).
"""
from __future__ import annotations

import ctypes
from typing import List, Optional

import torch

from . import _lib
from .ops import launch_counter

__all__ = ["get_depth_prediction", "dgc_rescale", "dgc_camera_heights"]

DGC_LAUNCHES = 4   # dgc_heights_kernel, dgc_refine_kernel<2>, dgc_refine_kernel<3>, dgc_apply_kernel (+ one memset node)


def _cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("%s is on %s: the DGC post-processing runs only on CUDA (sm_100a); there is no CPU fallback" % (name, t.device))
    return t


def _problem(depth, camera, height, panoptic, road_class_id, filter_ids, use_dgc, camera_is_inverse, points, scale, count, ws):
    B, _, H, W = depth.shape
    p = _lib.MgvsDgcProblem()
    p.B, p.H, p.W = B, H, W
    p.depth = depth.data_ptr()
    if camera is not None:
        p.camera = camera.data_ptr()
        p.cam_batch_stride = camera.stride(0) if camera.shape[0] > 1 else 0
        p.cam_row_stride = camera.stride(1)
    p.camera_is_inverse = int(bool(camera_is_inverse))
    if height is not None:
        p.real_camera_height = height.data_ptr()
        p.height_stride = 1 if height.numel() > 1 else 0
    if panoptic is not None:
        p.panoptic = panoptic.data_ptr()
        p.panoptic_dtype = _lib.PANOPTIC_I64 if panoptic.dtype == torch.int64 else _lib.PANOPTIC_I32
    else:
        p.panoptic_dtype = _lib.PANOPTIC_NONE
    p.use_dgc = int(bool(use_dgc))
    p.road_class_id = int(road_class_id)
    ids = list(filter_ids or [])
    if len(ids) > _lib.DGC_MAX_FILTER:
        raise ValueError("at most %d depth_filter_class_ids are supported" % _lib.DGC_MAX_FILTER)
    for k, v in enumerate(ids):
        p.filter_ids[k] = int(v)
    p.n_filter = len(ids)
    p.points = points.data_ptr() if points is not None else None
    p.scale = scale.data_ptr() if scale is not None else None
    p.count = count.data_ptr() if count is not None else None
    p.workspace = ws.data_ptr() if ws is not None else None
    p.workspace_bytes = ws.numel() if ws is not None else 0
    return p


def _prepare(depth_logits, camera_matrix, real_camera_height, panoptic_seg, use_dgc):
    depth_logits = _cuda_f32(depth_logits, "depth_logits")
    if depth_logits.dim() != 4 or depth_logits.shape[1] != 1:
        raise ValueError("depth_logits must be [B,1,H,W], got %s" % (tuple(depth_logits.shape),))
    if depth_logits.dtype != torch.float32:
        raise TypeError("depth_logits must be float32 (it is rescaled in place)")
    B, _, H, W = depth_logits.shape
    dev = depth_logits.device
    cam = hgt = pan = None
    if use_dgc:
        cam = camera_matrix.to(dev, torch.float32)          # `Camera(K=camera_matrix).to(device)`, depth_post_proc.py:52
        if cam.dim() == 2:
            cam = cam.unsqueeze(0)
        if cam.dim() != 3 or cam.shape[0] not in (1, B) or cam.shape[1] < 3 or cam.shape[2] < 3:
            raise ValueError("camera_matrix must be [1 or B, >=3, >=3]")
        if cam.stride(2) != 1:
            cam = cam.contiguous()
        hgt = real_camera_height.to(dev, torch.float32).reshape(-1).contiguous()   # depth_post_proc.py:51
        if hgt.numel() not in (1, B):
            raise ValueError("real_camera_height must hold 1 or B values")
    if panoptic_seg is not None:
        pan = panoptic_seg
        if not pan.is_cuda:
            raise RuntimeError("panoptic_seg must be on CUDA")
        if pan.dtype not in (torch.int64, torch.int32):
            pan = pan.long()
        if pan.numel() != B * H * W:
            raise ValueError("panoptic_seg must be [H,W] (or [B,H,W])")
        pan = pan.contiguous()
    return depth_logits, cam, hgt, pan


def dgc_rescale(depth_logits, camera_matrix, real_camera_height, panoptic_seg=None, road_class_id=-1,
                depth_filter_class_ids=None, use_dgc_scaling=True, camera_is_inverse=False, want_points=True):
    """Batched functional form.  depth_logits [B,1,H,W] float32 CUDA is rescaled IN PLACE.

    Returns (points [B,3,H,W] or None, scale [B] or None, count [B] int64 or None); ``count[b] == 0`` iff the ground
    mask of image b is empty (then ``scale[b]`` is NaN, exactly what ``torch.median`` of an empty selection gives
    the reference).  Nothing is synchronised with the host.
    """
    L = _lib.lib()
    depth_logits, cam, hgt, pan = _prepare(depth_logits, camera_matrix, real_camera_height, panoptic_seg, use_dgc_scaling)
    B, _, H, W = depth_logits.shape
    dev = depth_logits.device
    work = depth_logits if depth_logits.is_contiguous() else depth_logits.contiguous()
    with torch.cuda.device(dev):
        points = scale = count = ws = None
        if use_dgc_scaling:
            points = torch.empty((B, 3, H, W), dtype=torch.float32, device=dev) if want_points else None
            scale = torch.empty(B, dtype=torch.float32, device=dev)
            count = torch.empty(B, dtype=torch.int64, device=dev)
            ws = torch.empty(int(L.mgvs_dgc_workspace_bytes(B, H, W)), dtype=torch.uint8, device=dev)
        prob = _problem(work, cam, hgt, pan, road_class_id, depth_filter_class_ids, use_dgc_scaling, camera_is_inverse,
                        points, scale, count, ws)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(L.mgvs_dgc_rescale(ctypes.byref(prob), stream), "mgvs_dgc_rescale")
        if use_dgc_scaling:
            launch_counter.n += DGC_LAUNCHES
        elif pan is not None and prob.n_filter:
            launch_counter.n += 1
    if work is not depth_logits:
        depth_logits.copy_(work)
    return points, scale, count


def dgc_camera_heights(depth_logits, camera_matrix, panoptic_seg=None, road_class_id=-1, camera_is_inverse=False):
    """Diagnostics for the parity tests: per-pixel camera heights |P.N| [B,H,W] (depth_post_proc.py:96) and the ground
    mask [B,H,W] bool the median runs over.  ``depth_logits`` is not modified."""
    L = _lib.lib()
    dummy_h = torch.ones(1, device=depth_logits.device)
    depth_logits, cam, hgt, pan = _prepare(depth_logits, camera_matrix, dummy_h, panoptic_seg, True)
    B, _, H, W = depth_logits.shape
    dev = depth_logits.device
    work = depth_logits.contiguous()
    with torch.cuda.device(dev):
        heights = torch.empty((B, H, W), dtype=torch.float32, device=dev)
        ground = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
        scale = torch.empty(B, dtype=torch.float32, device=dev)
        ws = torch.empty(int(L.mgvs_dgc_workspace_bytes(B, H, W)), dtype=torch.uint8, device=dev)
        prob = _problem(work, cam, hgt, pan, road_class_id, None, True, camera_is_inverse, None, scale, None, ws)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(L.mgvs_dgc_heights(ctypes.byref(prob), heights.data_ptr(), ground.data_ptr(), stream), "mgvs_dgc_heights")
        launch_counter.n += 1
    return heights, ground.bool()


def get_depth_prediction(
    depth_logits: torch.Tensor,
    use_dgc_scaling: bool,
    camera_matrix: torch.Tensor = None,
    real_camera_height: torch.Tensor = None,
    panoptic_seg: torch.Tensor = None,
    road_class_id: int = -1,
    depth_filter_class_ids: Optional[List[int]] = None,
):
    """Same contract as the reference (depth_post_proc.py:11-71).

    Args:
        depth_logits: [1, 1, H, W] predicted depth (CUDA float32); rescaled in place.
        use_dgc_scaling: whether to rescale to metric depth with the DGC module.
        camera_matrix: [1, 3, 3] intrinsics.  real_camera_height: [1] mounting height over ground.
        panoptic_seg: [H, W] panoptic label or None (then the ground mask comes from the surface normals).
        road_class_id: id of the road class in panoptic_seg.  depth_filter_class_ids: classes whose depth is zeroed.
    Returns:
        depth_logits [H, W] (the same storage, squeezed) and cam_xyz_points [3, H, W] (None without DGC).
    """
    cam_xyz_points = None
    if use_dgc_scaling:
        assert camera_matrix is not None, "camera_matrix is necessary for dgc rescaling!"
        assert real_camera_height is not None, "real_camera_height is necessary for dgc rescaling!"
        if panoptic_seg is not None:
            assert (
                road_class_id != -1
            ), "road_class_id is necessary for dgc rescaling using panoptic prediction!"
    if depth_logits.dim() != 4 or depth_logits.shape[0] != 1:
        raise ValueError("get_depth_prediction handles one image per call ([1,1,H,W]) like the reference; use dgc_rescale for batches")
    points, _, _ = dgc_rescale(depth_logits, camera_matrix, real_camera_height, panoptic_seg, road_class_id,
                               depth_filter_class_ids if panoptic_seg is not None else None, use_dgc_scaling)
    if points is not None:
        cam_xyz_points = points.squeeze_()
    depth_logits.squeeze_()
    return depth_logits, cam_xyz_points
