"""torch.autograd glue around the C ABI (include/mgvs.h).  PyTorch is used here for device memory,
streams and torch.distributed only; all arithmetic of the loss runs in libmgvs.so.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from .sharding import allreduce_sums

__all__ = ["LossConfig", "view_synthesis_loss", "launch_counter"]


class _Counter:
    """Counts kernel launches issued through the library (for bench.py's gpu_launches)."""

    def __init__(self):
        self.n = 0


launch_counter = _Counter()
FWD_LAUNCHES = 3   # pack_sources_kernel (+camera table), fwd_kernel, reduce_kernel (+ one upsample_kernel per scale with fuse_upsample)
FIN_LAUNCHES = 1   # finalize_kernel
BWD_LAUNCHES = 2   # bwd_kernel, pose_reduce_kernel


@dataclass(frozen=True)
class LossConfig:
    """The six constructor arguments of the reference loss (loss.py:87-109) plus sharding options."""
    ssim_loss_weight: float = 0.85
    photometric_loss_weight: float = 1.0
    smoothing_loss_weight: float = 1e-3
    automask_loss: bool = True
    photometric_reduce_op: str = "min"
    padding_mode: str = "zeros"
    process_group: object = None      # torch.distributed group for the partial-sum all-reduce (None = local)
    exchange: object = None           # sharding.PeerExchange: the sums cross GPUs inside ONE kernel (NVLink P2P stores +
                                      # flags + finalize) instead of an NCCL all-reduce + finalize_kernel; needs process_group
    ddp_grad_scale: bool = False      # multiply local grads by world size so DDP's 1/G averaging yields the
                                      # full-batch gradient; only with process_group
    fuse_upsample: bool = False       # predictions["depth"][i] are the depth head's LOW-resolution maps
                                      # [B,1,H/s,W/s] (after sigmoid()/0.5, before its F.interpolate(scale_factor=s,
                                      # mode="bilinear", align_corners=True), mg_net.py:803-806,823); the library upsamples them
                                      # itself (an HBM-bound pre-pass, bit-identical to ATen's CPU kernel) and returns
                                      # low-resolution gradients through a deterministic adjoint.  Needs backward="stash".
    backward: str = "stash"           # "stash": the forward also writes the SSIM-adjoint coefficient texels of the
                                      # selected source (48 B/px/scale) and the backward consumes them (fastest);
                                      # "recompute": nothing but the uint8 selection is carried over and the backward
                                      # recomputes the forward from the same tiles (lean-memory mode).  Same results
                                      # up to fp32 rounding of the gradients; forward outputs are bit-identical.


def _require_cuda_f32(t: torch.Tensor, name: str, shape=None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("%s is on %s: the view-synthesis loss runs only on CUDA (sm_100a); there is no CPU fallback" % (name, t.device))
    if t.dtype != torch.float32:
        t = t.float()
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError("%s has shape %s, expected %s" % (name, tuple(t.shape), tuple(shape)))
    return t.contiguous()


def _require_cuda_image(t: torch.Tensor, name: str, shape=None) -> torch.Tensor:
    """Images may arrive as float (what the reference loss receives) or as the data loader's uint8
    (mg_net.py:320-335 converts with ``.float() / 255.0``; the library does that conversion itself)."""
    if isinstance(t, torch.Tensor) and t.dtype == torch.uint8:
        if not t.is_cuda:
            raise RuntimeError("%s is on %s: the view-synthesis loss runs only on CUDA (sm_100a); there is no CPU fallback" % (name, t.device))
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError("%s has shape %s, expected %s" % (name, tuple(t.shape), tuple(shape)))
        return t.contiguous()
    return _require_cuda_f32(t, name, shape)


def _fill_problem(prob, cfg, tgt, prev, nxt, inv, camera, poses, mask, ws, stash=None):
    B, _, H, W = tgt.shape
    if not (tgt.dtype == prev.dtype == nxt.dtype):
        raise TypeError("image_orig, image_prev_orig and image_next_orig must share one dtype (float32 or uint8)")
    prob.image_dtype = _lib.IMAGE_U8 if tgt.dtype == torch.uint8 else _lib.IMAGE_F32
    prob.B, prob.H, prob.W, prob.n = B, H, W, len(inv)
    prob.target = tgt.data_ptr()
    prob.source[0] = prev.data_ptr()
    prob.source[1] = nxt.data_ptr()
    for i, d in enumerate(inv):
        prob.inv_depth[i] = d.data_ptr()
    prob.camera = camera.data_ptr()
    prob.cam_batch_stride = camera.stride(0)
    prob.cam_row_stride = camera.stride(1)
    if poses.dim() == 4:                      # [B,S,3,4] pose matrices the caller built (MgvsProblem.pose_mats)
        prob.poses, prob.pose_mats = None, poses.data_ptr()
    else:
        prob.poses, prob.pose_mats = poses.data_ptr(), None
    prob.mask = mask.data_ptr() if mask is not None else None
    prob.ssim_weight = float(cfg.ssim_loss_weight)
    prob.one_minus_ssim_weight = 1.0 - float(cfg.ssim_loss_weight)   # Python double, rounded to fp32 by ctypes
    prob.photometric_weight = float(cfg.photometric_loss_weight)
    prob.smoothing_weight = float(cfg.smoothing_loss_weight)
    prob.automask = int(bool(cfg.automask_loss))
    prob.reduce_op = 0
    prob.padding_mode = _lib.PADDING_MODES[cfg.padding_mode]
    prob.workspace = ws.data_ptr()
    prob.workspace_bytes = ws.numel()
    prob.stash = stash.data_ptr() if stash is not None else None
    prob.stash_bytes = stash.numel() if stash is not None else 0
    for i, d in enumerate(inv):
        lowres = tuple(d.shape[-2:]) != (H, W)
        prob.inv_height[i] = d.shape[-2] if lowres else 0
        prob.inv_width[i] = d.shape[-1] if lowres else 0


class _ViewSynthesisLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg: LossConfig, poses, camera, tgt, prev, nxt, mask, *inv):
        L = _lib.lib()
        if cfg.photometric_reduce_op != "min":
            if cfg.photometric_reduce_op == "mean":
                raise NotImplementedError("photometric_reduce_op='mean' is not implemented by the fused kernels")
            raise NotImplementedError("Unknown photometric_reduce_op: {}".format(cfg.photometric_reduce_op))
        if cfg.padding_mode not in _lib.PADDING_MODES:
            raise ValueError("padding_mode must be 'zeros', 'border' or 'reflection', got %r" % (cfg.padding_mode,))
        n = len(inv)
        if n < 1 or n > _lib.MAX_SCALES:
            raise ValueError("need 1..%d inverse-depth maps, got %d" % (_lib.MAX_SCALES, n))
        tgt = _require_cuda_image(tgt, "image_orig")
        B, C, H, W = tgt.shape
        if C != 3:
            raise ValueError("image_orig must be [B,3,H,W]")
        prev = _require_cuda_image(prev, "image_prev_orig", (B, 3, H, W))
        nxt = _require_cuda_image(nxt, "image_next_orig", (B, 3, H, W))
        if not (tgt.dtype == prev.dtype == nxt.dtype):
            raise TypeError("image_orig, image_prev_orig and image_next_orig must share one dtype (float32 or uint8)")
        img_dtype = _lib.IMAGE_U8 if tgt.dtype == torch.uint8 else _lib.IMAGE_F32
        if cfg.fuse_upsample:
            if cfg.backward != "stash":
                raise NotImplementedError("fuse_upsample needs backward='stash'")
            checked = []
            for i, d in enumerate(inv):
                d = _require_cuda_f32(d, "depth[%d]" % i)
                if d.dim() != 4 or d.shape[0] != B or d.shape[1] != 1:
                    raise ValueError("depth[%d] must be [B,1,h,w]" % i)
                h, w = d.shape[-2:]
                if h < 1 or w < 1 or H % h or W % w or H // h != W // w or (h, w) == (H, W):
                    raise ValueError("fuse_upsample: depth[%d] is %dx%d, expected the image size %dx%d divided by one integer stride > 1" % (i, h, w, H, W))
                checked.append(d)
            inv = checked
        else:
            inv = [_require_cuda_f32(d, "depth[%d]" % i, (B, 1, H, W)) for i, d in enumerate(inv)]
        # [B,S,6] Euler vectors (the reference's contract, loss.py:117-119) or [B,S,3,4] matrices = rows 0..2 of what
        # Pose.from_vec(vec, "euler") builds (pose.py:41-47): then the kernels use the caller's rotation bits as they are
        poses = _require_cuda_f32(poses, "poses", (B, 2, 3, 4) if poses.dim() == 4 else (B, 2, 6))
        if camera.dim() != 3 or camera.shape[0] != B or camera.shape[1] < 3 or camera.shape[2] < 3:
            raise ValueError("camera_matrix must be [B,>=3,>=3]")
        camera = _require_cuda_f32(camera, "camera_matrix")
        if mask is not None:
            if not mask.is_cuda:
                raise RuntimeError("reprojection_mask must be on CUDA")
            if mask.dtype == torch.uint8 and tuple(mask.shape) == (B, 1, H, (W + 7) // 8):
                # bit-packed mask (numpy.packbits(mask, axis=-1)): 1/8 of the bytes on the host-to-device link; unpacked here
                bits = mask.contiguous()
                mask = torch.empty((B, 1, H, W), dtype=torch.bool, device=bits.device)
                with torch.cuda.device(bits.device):
                    _lib.check(L.mgvs_unpack_mask(B * H, W, bits.data_ptr(), mask.data_ptr(),
                                                  ctypes.c_void_p(torch.cuda.current_stream(bits.device).cuda_stream)), "mgvs_unpack_mask")
                launch_counter.n += 1
            if tuple(mask.shape) != (B, 1, H, W):
                raise ValueError("reprojection_mask must be [B,1,H,W] (bool / integer) or the bit-packed uint8 [B,1,H,ceil(W/8)]")
            if mask.dtype != torch.bool:
                mask = mask != 0
            mask = mask.contiguous()
        dev = tgt.device
        if cfg.backward not in ("stash", "recompute"):
            raise ValueError("backward must be 'stash' or 'recompute', got %r" % (cfg.backward,))
        # the stash is only worth writing when a backward pass will follow
        # (ssim_loss_weight == 0, the raw-L1 branch, has no SSIM adjoint to stash: the library runs its recompute backward)
        want_stash = cfg.backward == "stash" and cfg.ssim_loss_weight > 0 and any(ctx.needs_input_grad[i] for i in (1,) + tuple(range(7, 7 + n)))
        with torch.cuda.device(dev):
            ws = torch.empty(int(L.mgvs_workspace_bytes_ex2(B, H, W, n, img_dtype, int(cfg.fuse_upsample))), dtype=torch.uint8, device=dev)
            stash = torch.empty(int(L.mgvs_stash_bytes_ex(B, H, W, n, int(cfg.fuse_upsample))), dtype=torch.uint8, device=dev) if want_stash else None
            sel = torch.empty((n, B, H, W), dtype=torch.uint8, device=dev)
            sums = torch.empty(3 * n + 3, dtype=torch.float64, device=dev)
            losses = torch.empty(2, dtype=torch.float32, device=dev)
            prob = _lib.MgvsProblem()
            _fill_problem(prob, cfg, tgt, prev, nxt, inv, camera, poses, mask, ws, stash)
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            world = 1
            if cfg.process_group is None:
                # single rank: the reduction's last block also writes the two losses (no finalize launch)
                _lib.check(L.mgvs_forward_losses(ctypes.byref(prob), sel.data_ptr(), sums.data_ptr(), losses.data_ptr(), stream),
                           "mgvs_forward_losses")
                launch_counter.n += FWD_LAUNCHES + (n if cfg.fuse_upsample else 0)
            else:
                _lib.check(L.mgvs_forward(ctypes.byref(prob), sel.data_ptr(), sums.data_ptr(), stream), "mgvs_forward")
                launch_counter.n += FWD_LAUNCHES + (n if cfg.fuse_upsample else 0)
                if cfg.exchange is not None:
                    # the only inter-GPU exchange of the path, fused with the finalize: P2P pushes over NVLink in one kernel
                    world = cfg.exchange.world
                    _lib.check(L.mgvs_exchange_finalize(ctypes.byref(prob), ctypes.byref(cfg.exchange.struct), sums.data_ptr(),
                                                        losses.data_ptr(), stream), "mgvs_exchange_finalize")
                else:
                    world = allreduce_sums(sums, cfg.process_group)    # same exchange through NCCL
                    _lib.check(L.mgvs_finalize(ctypes.byref(prob), sums.data_ptr(), losses.data_ptr(), stream), "mgvs_finalize")
                launch_counter.n += FIN_LAUNCHES
        ctx.cfg = cfg
        ctx.n = n
        ctx.has_mask = mask is not None
        ctx.has_stash = stash is not None
        ctx.grad_scale = float(world) if (cfg.ddp_grad_scale and cfg.process_group is not None) else 1.0
        # the stash rides with the saved tensors: freed with them after backward (retain_graph=False), alive for a second
        # backward over the same graph (retain_graph=True) -- the same kernel runs both times
        saved = [poses, camera, tgt, prev, nxt, sel, sums, ws] + ([mask] if mask is not None else []) + ([stash] if stash is not None else []) + list(inv)
        ctx.save_for_backward(*saved)
        ctx.mark_non_differentiable(sel)
        ctx.set_materialize_grads(False)
        lp, ls = losses.unbind(0)
        return lp, ls, sel

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_photo, g_smooth, _g_sel):
        L = _lib.lib()
        saved = ctx.saved_tensors
        poses, camera, tgt, prev, nxt, sel, sums, ws = saved[:8]
        k = 8
        mask = None
        if ctx.has_mask:
            mask = saved[k]
            k += 1
        stash = None
        if ctx.has_stash:
            stash = saved[k]
            k += 1
        inv = list(saved[k:])
        dev = tgt.device
        with torch.cuda.device(dev):
            # the two upstream gradients as one [2] device vector: one launch when both are there (the usual case)
            if g_photo is not None and g_smooth is not None:
                g = torch.stack((g_photo.reshape(()).float(), g_smooth.reshape(()).float()))
            else:
                g = torch.zeros(2, dtype=torch.float32, device=dev)
                if g_photo is not None:
                    g[0] = g_photo
                if g_smooth is not None:
                    g[1] = g_smooth
            if ctx.grad_scale != 1.0:
                g = g * ctx.grad_scale
            grads = [torch.empty_like(d) for d in inv]
            gp = torch.empty_like(poses)
            prob = _lib.MgvsProblem()
            _fill_problem(prob, ctx.cfg, tgt, prev, nxt, inv, camera, poses, mask, ws, stash)
            arr = (ctypes.c_void_p * len(grads))(*[x.data_ptr() for x in grads])
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(L.mgvs_backward(ctypes.byref(prob), sel.data_ptr(), sums.data_ptr(), g.data_ptr(), arr,
                                       gp.data_ptr(), stream), "mgvs_backward")
            launch_counter.n += BWD_LAUNCHES + (2 * len(inv) if ctx.cfg.fuse_upsample else 0)   # + two upsample-adjoint kernels per scale
        return (None, gp, None, None, None, None, None) + tuple(grads)


def view_synthesis_loss(inv_depths, poses, image, image_prev, image_next, camera_matrix,
                        reprojection_mask: Optional[torch.Tensor] = None, cfg: LossConfig = LossConfig()):
    """Functional form.  Returns (loss_photometric, loss_smoothness, selection uint8 [n,B,H,W])."""
    return _ViewSynthesisLoss.apply(cfg, poses, camera_matrix, image, image_prev, image_next, reprojection_mask,
                                    *inv_depths)
