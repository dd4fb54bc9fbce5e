"""Drop-in replacement of ``mgnet.modeling.loss.MultiViewPhotometricLoss`` (reference loss.py:84-294).

Same constructor arguments, same ``forward(predictions, targets)`` dictionary contract, same scales,
automask / min-reprojection semantics, grid_sample padding ("zeros") and align_corners=True behaviour --
but the arithmetic runs in two fused sm_100a kernels (mgnet_b200/csrc) reached through the C ABI in
include/mgvs.h.  CUDA only: CPU tensors raise (no fallback).

Plug-in seam (reference mg_net.py:744,757,772-779,792): pass an instance as ``loss=`` to
``MGNetSelfSupervisedDepthHead`` or build it in ``from_config`` from ``cfg.MODEL.DEPTH_HEAD.*``.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .ops import LossConfig, view_synthesis_loss

__all__ = ["MultiViewPhotometricLoss"]


class MultiViewPhotometricLoss(nn.Module):
    def __init__(
        self,
        ssim_loss_weight,
        photometric_loss_weight,
        smoothing_loss_weight,
        automask_loss,
        photometric_reduce_op,
        padding_mode,
        process_group=None,
        exchange=None,
        ddp_grad_scale=False,
        backward="stash",
        fuse_upsample=False,
    ):
        super().__init__()
        self.n = None
        self.ssim_loss_weight = ssim_loss_weight
        self.photometric_loss_weight = photometric_loss_weight
        self.smoothing_loss_weight = smoothing_loss_weight
        self.automask_loss = automask_loss
        self.photometric_reduce_op = photometric_reduce_op
        self.padding_mode = padding_mode
        self.process_group = process_group
        self.exchange = exchange     # sharding.PeerExchange: fused peer-memory exchange + finalize instead of NCCL (see ops.LossConfig)
        if exchange is not None and process_group is None:
            raise ValueError("exchange= needs process_group=")
        self.ddp_grad_scale = ddp_grad_scale
        self.fuse_upsample = fuse_upsample   # predictions["depth"] are the head's low-resolution maps; see ops.LossConfig
        self.backward = backward     # "stash" (fast, +48 B/px/scale of scratch) or "recompute" (lean memory), see ops.LossConfig
        self.last_selection = None   # uint8 [n,B,H,W] argmin of the most recent forward (new side output)
        # same assertion as the reference (loss.py:106-109)
        if self.automask_loss:
            assert self.photometric_reduce_op == "min", \
                "For automasking only the min photometric_reduce_op is supported."
        # legal-but-unimplemented combinations fail loudly at construction time
        if padding_mode not in ("zeros", "border", "reflection"):   # what F.grid_sample accepts (camera_utils.py:52-54)
            raise ValueError("padding_mode must be 'zeros', 'border' or 'reflection', got %r" % (padding_mode,))
        if photometric_reduce_op not in ("min", "mean"):
            raise NotImplementedError("Unknown photometric_reduce_op: {}".format(photometric_reduce_op))
        if photometric_reduce_op == "mean" and not (float(ssim_loss_weight) > 0):
            # the reference averages all 3 channels of the raw L1 maps here (loss[mask].mean() on [B,3,H,W], and raises an
            # IndexError when a [B,1,H,W] mask is passed); the two-evaluation composition below would return a per-pixel
            # channel minimum instead -- a different number
            raise NotImplementedError("photometric_reduce_op='mean' with ssim_loss_weight == 0 is not implemented")
        # "mean" (loss.py:242-243; only legal with automask off, :106-109) runs the fused path once per source frame with that
        # frame in both source slots -- min(L, L) = L, index 0 -- and averages the two photometric losses: exact semantics,
        # twice the cost of "min" (no shipped config selects it)
        # ssim_loss_weight == 0 is the reference's raw 3-channel L1 branch (loss.py:195-196): the min then runs over 3 channels
        # per list entry; the kernels implement it (selection index = entry * 3 + channel), always with the recompute backward

    def _config(self) -> LossConfig:
        return LossConfig(
            ssim_loss_weight=float(self.ssim_loss_weight),
            photometric_loss_weight=float(self.photometric_loss_weight),
            smoothing_loss_weight=float(self.smoothing_loss_weight),
            automask_loss=bool(self.automask_loss),
            photometric_reduce_op=self.photometric_reduce_op,
            padding_mode=self.padding_mode,
            process_group=self.process_group,
            exchange=self.exchange,
            ddp_grad_scale=bool(self.ddp_grad_scale),
            backward=self.backward,
            fuse_upsample=bool(self.fuse_upsample),
        )

    def forward(self, predictions, targets):
        inv_depths = predictions["depth"]
        pose_results = predictions["poses"]
        self.n = len(inv_depths)
        assert pose_results.shape[1] == 2, "Context and poses lists must be of same length"
        # extension: [B,S,4,4] / [B,S,3,4] pose matrices built by the caller (e.g. torch's pose_vec2mat) instead of [B,S,6] Euler
        # vectors -- the kernels then use the caller's rotation bits (unconditionally bit-exact selection)
        if pose_results.dim() == 4:
            pose_results = pose_results[:, :, :3, :4]
        # custom_fwd(cast_inputs=torch.float32) equivalent (mg_net.py:827): the op computes in fp32
        # images: float in [0,1] as in the reference, or the data loader's uint8 tensors -- then the kernels apply
        # the caller's `x.float() / 255.0` (mg_net.py:320-335) themselves, bit-identically
        def img(t):
            return t if t.dtype == torch.uint8 else t.float()

        if self.photometric_reduce_op == "mean":
            return self._forward_mean(predictions, targets, img)
        with torch.autocast(device_type="cuda", enabled=False):
            lp, ls, sel = view_synthesis_loss(
                [d.float() for d in inv_depths],
                pose_results.float(),
                img(targets["image_orig"]),
                img(targets["image_prev_orig"]),
                img(targets["image_next_orig"]),
                targets["camera_matrix"].float(),
                targets["reprojection_mask"] if "reprojection_mask" in targets else None,
                self._config(),
            )
        self.last_selection = sel
        return {"loss_photometric": lp, "loss_smoothness": ls}

    def graphed(self, predictions, targets, num_warmup_iters=3):
        """CUDA-graph mode for launch-bound shapes (C1, B1 192x640: the eager call is ~3x host overhead).

        Captures this module's forward and backward for the SHAPES of the sample ``predictions`` / ``targets`` with
        ``torch.cuda.make_graphed_callables`` (the C ABI is capturable: no allocation, no synchronisation, launches on the
        caller's stream only) and returns a callable with the same ``(predictions, targets) -> dict`` contract that copies its
        arguments into the static buffers and replays the graphs.  It takes part in autograd like the module itself; results are
        bit-identical to the eager call (tests/test_cuda_graph.py).  Single-rank only (no process_group)."""
        if self.process_group is not None:
            raise NotImplementedError("graphed(): not with process_group (the exchange must see the same call sequence on every rank)")
        if self.photometric_reduce_op != "min":
            raise NotImplementedError("graphed(): photometric_reduce_op='min' only")
        n = len(predictions["depth"])
        has_mask = "reprojection_mask" in targets
        keys = ("image_orig", "image_prev_orig", "image_next_orig", "camera_matrix") + (("reprojection_mask",) if has_mask else ())
        cfg = self._config()

        def fn(*args):
            inv, poses = list(args[:n]), args[n]
            tgt, prev, nxt, K = args[n + 1:n + 5]
            mask = args[n + 5] if has_mask else None
            with torch.autocast(device_type="cuda", enabled=False):
                lp, ls, sel = view_synthesis_loss(inv, poses, tgt, prev, nxt, K, mask, cfg)
            return lp, ls, sel

        def flatten(pred, tgt):
            poses = pred["poses"]
            if poses.dim() == 4:
                poses = poses[:, :, :3, :4]
            out = [d.float() for d in pred["depth"]] + [poses.float()]
            for k in keys:
                v = tgt[k]
                if k == "camera_matrix" or (k != "reprojection_mask" and v.dtype != torch.uint8):
                    v = v.float()
                out.append(v)
            return tuple(out)

        sample = tuple(a.detach().clone().requires_grad_(a.requires_grad) for a in flatten(predictions, targets))
        replay = torch.cuda.make_graphed_callables(fn, sample, num_warmup_iters=num_warmup_iters)
        module = self

        def call(pred, tgt):
            lp, ls, sel = replay(*flatten(pred, tgt))
            module.last_selection = sel
            return {"loss_photometric": lp, "loss_smoothness": ls}

        return call

    # The reference's helper methods (loss.py:156-294) have no stand-alone counterpart: their arithmetic is fused into the two
    # kernels and their intermediates (warped images, per-pixel SSIM / photometric maps) never exist in memory.
    def _fused_away(self, name, hint):
        raise NotImplementedError(
            "MultiViewPhotometricLoss.%s is fused into the view-synthesis kernels of mgnet_b200 and is not available as a separate "
            "step; %s" % (name, hint))

    def warp_ref_image(self, depths, ref_image, cams, ref_camera_matrix, pose):
        """loss.py:156-167, same arguments (depths are METRIC depths, cams = [target Camera]) -- through the stand-alone
        kernel (forward only; bit-identical to the reference on CPU)."""
        from .geometry import Camera, view_synthesis
        ref_cam = Camera(K=ref_camera_matrix.float(), Tcw=pose.to(ref_image.device))
        assert len(cams) == 1
        return [view_synthesis(ref_image, d, ref_cam, cams[0], padding_mode=self.padding_mode) for d in depths]

    def calc_photometric_loss(self, t_est, images):
        self._fused_away("calc_photometric_loss", "call forward(); per-pixel maps are available from the CPU oracle (oracle/) for debugging")

    def ssim(self, x, y, kernel_size=3, c1=1e-4, c2=9e-4):
        self._fused_away("ssim", "call forward()")

    def reduce_photometric_loss(self, photometric_losses, mask=None):
        self._fused_away("reduce_photometric_loss", "call forward(); the per-scale argmin is in self.last_selection")

    def calc_smoothness_loss(self, inv_depths, images, mask=None):
        self._fused_away("calc_smoothness_loss", "call forward() and read 'loss_smoothness'")

    def _forward_mean(self, predictions, targets, img):
        """photometric_reduce_op="mean" (loss.py:242-243): sum_s masked_mean(L_s) / S per scale, averaged over scales -- linear in
        the per-source losses, so it is the average of two fused "min" evaluations that each see ONE source frame in both slots
        (identical maps tie, the strict `<` scan keeps index 0, so slot 0 carries the whole gradient).  The smoothness term does
        not depend on the sources and is taken from the first evaluation."""
        import dataclasses
        cfg = dataclasses.replace(self._config(), photometric_reduce_op="min", automask_loss=False)
        inv = [d.float() for d in predictions["depth"]]
        poses = predictions["poses"].float()
        if poses.dim() == 4:
            poses = poses[:, :, :3, :4]
        mask = targets["reprojection_mask"] if "reprojection_mask" in targets else None
        tgt = img(targets["image_orig"])
        K = targets["camera_matrix"].float()
        photo, smooth = [], None
        with torch.autocast(device_type="cuda", enabled=False):
            for s, key in enumerate(("image_prev_orig", "image_next_orig")):
                src = img(targets[key])
                lp, ls, _ = view_synthesis_loss(inv, poses[:, [s, s]], tgt, src, src, K, mask, cfg)
                photo.append(lp)
                smooth = ls if smooth is None else smooth
        self.last_selection = None      # there is no argmin under "mean"
        return {"loss_photometric": (photo[0] + photo[1]) / 2, "loss_smoothness": smooth}
