"""Homoscedastic uncertainty weighting of the task losses: the epilogue of ``MGNet.forward``
(reference mgnet/modeling/mg_net.py:360-372) without its per-loss ``.item()`` host synchronisations.

    losses[key] = tau * exp(-log_vars[idx]) * value + 0.5 * log_vars[idx]          tau = 1.0 for "loss_sem_seg" else 0.5

The reference walks the loss dict in insertion order with a running index into ``self.log_vars`` and pushes
``key + "_raw"`` and ``key + "_uncertainty"`` into detectron2's event storage through two ``.item()`` calls per loss
(10 device synchronisations per training step with all five losses).  Here one tiny kernel weights every loss and
writes the logging copies to device memory; the caller fetches them with a single asynchronous copy when it logs.
CUDA only (no CPU fallback).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .ops import launch_counter

__all__ = ["apply_uncertainty"]


class _Uncertainty(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw, log_vars, tau):
        L = _lib.lib()
        k = raw.numel()
        dev = raw.device
        with torch.cuda.device(dev):
            weighted = torch.empty(k, dtype=torch.float32, device=dev)
            log_out = torch.empty(2 * k, dtype=torch.float32, device=dev)
            arr = (ctypes.c_float * k)(*tau)
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(L.mgvs_uncertainty_forward(k, raw.data_ptr(), log_vars.data_ptr(), arr, weighted.data_ptr(),
                                                  log_out.data_ptr(), stream), "mgvs_uncertainty_forward")
            launch_counter.n += 1
        ctx.tau = tau
        ctx.save_for_backward(raw, log_vars)
        ctx.mark_non_differentiable(log_out)
        return weighted, log_out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_weighted, _g_log):
        L = _lib.lib()
        raw, log_vars = ctx.saved_tensors
        k = raw.numel()
        dev = raw.device
        with torch.cuda.device(dev):
            g = g_weighted.float().contiguous()
            g_raw = torch.empty(k, dtype=torch.float32, device=dev)
            g_s = torch.empty(k, dtype=torch.float32, device=dev)
            arr = (ctypes.c_float * k)(*ctx.tau)
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(L.mgvs_uncertainty_backward(k, raw.data_ptr(), log_vars.data_ptr(), arr, g.data_ptr(), g_raw.data_ptr(),
                                                   g_s.data_ptr(), stream), "mgvs_uncertainty_backward")
            launch_counter.n += 1
        return g_raw, g_s, None


def apply_uncertainty(losses: dict, log_vars: torch.Tensor, log: dict = None) -> dict:
    """Mirror of mg_net.py:360-372.  ``losses``: ordered dict of 0-d CUDA tensors; ``log_vars``: the model's parameter
    (at least ``len(losses)`` entries, indexed in dict order like the reference's running ``idx``).

    Returns a new dict with the weighted losses (gradients flow to every loss and to ``log_vars``).  When ``log`` is a
    dict it receives ``key + "_raw"`` and ``key + "_uncertainty"`` as 0-d DEVICE tensors (views of one buffer) -- the
    values the reference hands to ``storage.put_scalar`` after a ``.item()`` each.
    """
    keys = list(losses.keys())
    k = len(keys)
    if k < 1 or k > _lib.MAX_LOSSES:
        raise ValueError("1..%d losses supported, got %d" % (_lib.MAX_LOSSES, k))
    if log_vars.numel() < k:
        raise IndexError("index %d is out of bounds for log_vars with size %d" % (k - 1, log_vars.numel()))
    vals = [losses[key] for key in keys]
    for key, v in zip(keys, vals):
        if not v.is_cuda:
            raise RuntimeError("%s is on %s: the uncertainty epilogue runs only on CUDA; there is no CPU fallback" % (key, v.device))
    raw = torch.stack([v.float().reshape(()) for v in vals])
    tau = tuple(1.0 if key == "loss_sem_seg" else 0.5 for key in keys)
    weighted, log_out = _Uncertainty.apply(raw, log_vars[:k].float().contiguous(), tau)
    out = {key: w for key, w in zip(keys, weighted.unbind(0))}
    if log is not None:
        for i, key in enumerate(keys):
            log[key + "_raw"] = log_out[i]
            log[key + "_uncertainty"] = log_out[k + i]
    return out
