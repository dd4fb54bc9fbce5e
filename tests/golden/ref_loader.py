"""Loads the UNMODIFIED reference modules from /root/reference (build container only).

Used by ``make_golden.py`` and by the ``-m "not gpu"`` tests that validate the oracle
live against the reference when ``/root/reference`` is mounted.  Nothing here is copied
from the reference: the files are imported where they lie.  Never imported by the product,
by ``bench.py`` or by any ``-m gpu`` test (the GPU box has no /root/reference).

Two shims:
  1. ``mgnet/__init__.py`` imports detectron2, so ``mgnet`` and ``mgnet.modeling`` are
     registered as empty namespace stubs and only ``mgnet.geometry`` (torch-only) and
     ``mgnet/modeling/loss.py`` are executed.
  2. ``MultiViewPhotometricLoss.warp_ref_image`` (loss.py:160-161) calls
     ``ref_image.get_device()`` which is -1 on CPU; the subclass below replaces only the
     device argument with ``ref_image.device``.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("MGNET_REF", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "mgnet", "modeling", "loss.py"))


_cache = {}


def load():
    """Returns (geometry_module, loss_module, CpuLossClass)."""
    if "mods" in _cache:
        return _cache["mods"]
    if not available():
        raise RuntimeError("reference not mounted at %s" % REF_ROOT)
    pkg = types.ModuleType("mgnet")
    pkg.__path__ = [os.path.join(REF_ROOT, "mgnet")]
    sys.modules.setdefault("mgnet", pkg)
    mp = types.ModuleType("mgnet.modeling")
    mp.__path__ = [os.path.join(REF_ROOT, "mgnet", "modeling")]
    sys.modules.setdefault("mgnet.modeling", mp)
    import mgnet.geometry as geo  # noqa: E402  (torch-only)

    spec = importlib.util.spec_from_file_location(
        "mgnet.modeling.loss", os.path.join(REF_ROOT, "mgnet", "modeling", "loss.py")
    )
    lossmod = importlib.util.module_from_spec(spec)
    sys.modules["mgnet.modeling.loss"] = lossmod
    spec.loader.exec_module(lossmod)

    class CpuLoss(lossmod.MultiViewPhotometricLoss):
        """Device shim only: identical body except ``pose.to(ref_image.device)``."""

        def warp_ref_image(self, depths, ref_image, cams, ref_camera_matrix, pose):
            ref_cam = geo.Camera(K=ref_camera_matrix.float(), Tcw=pose.to(ref_image.device))
            return [
                geo.view_synthesis(
                    ref_image, depths[i], ref_cam, cams[0], padding_mode=self.padding_mode
                )
                for i in range(self.n)
            ]

    _cache["mods"] = (geo, lossmod, CpuLoss)
    return _cache["mods"]


DEFAULT_HP = dict(
    ssim_loss_weight=0.85,
    photometric_loss_weight=1.0,
    smoothing_loss_weight=1e-3,
    automask_loss=True,
    photometric_reduce_op="min",
    padding_mode="zeros",
)


def run_reference(predictions, targets, hp=None, want_grads=True, want_intermediates=True):
    """Runs the reference loss on CPU and returns a dict of numpy arrays.

    Intermediates are obtained by calling the reference's own methods (no re-derivation):
    per (scale, source) sample coordinates, warped image and photometric map; identity maps;
    per-scale argmin over [warp_prev, id_prev, warp_next, id_next] (loss.py:136-144,245).
    """
    geo, lossmod, CpuLoss = load()
    hp = dict(DEFAULT_HP if hp is None else hp)
    loss = CpuLoss(**hp)
    inv = [d.detach().clone().requires_grad_(want_grads) for d in predictions["depth"]]
    poses = predictions["poses"].detach().clone().requires_grad_(want_grads)
    out = loss({"depth": inv, "poses": poses}, targets)
    res = {
        "loss_photometric": out["loss_photometric"].detach().numpy().copy(),
        "loss_smoothness": out["loss_smoothness"].detach().numpy().copy(),
    }
    if want_grads:
        (out["loss_photometric"] + out["loss_smoothness"]).backward()
        for i, d in enumerate(inv):
            res["grad_depth_%d" % i] = d.grad.numpy().copy()
        res["grad_poses"] = poses.grad.numpy().copy()
        # separate upstream weights (1, 0) and (0, 1) are covered by linearity tests
    if want_intermediates:
        with torch.no_grad():
            n = len(inv)
            loss.n = n
            K = targets["camera_matrix"][:, :3, :3].float()
            context = [targets["image_prev_orig"], targets["image_next_orig"]]
            tgt = targets["image_orig"]
            depths = [geo.inv2depth(d.detach()) for d in inv]
            cam = geo.Camera(K=K)
            maps = [[] for _ in range(n)]
            for s, ref_image in enumerate(context):
                pose = geo.Pose.from_vec(poses.detach()[:, s].float(), "euler")
                res["pose_mat_%d" % s] = pose.mat.numpy().copy()
                ref_cam = geo.Camera(K=K, Tcw=pose)
                for i in range(n):
                    world = cam.reconstruct(depths[i], frame="w")
                    coords = ref_cam.project(world, frame="w")
                    warped = torch.nn.functional.grid_sample(
                        ref_image, coords, mode="bilinear", padding_mode=hp["padding_mode"],
                        align_corners=True,
                    )
                    pm = loss.calc_photometric_loss([warped], [tgt])[0]
                    res["coords_%d_%d" % (i, s)] = coords.numpy().copy()
                    res["warped_%d_%d" % (i, s)] = warped.numpy().copy()
                    res["photo_%d_%d" % (i, s)] = pm.numpy().copy()
                    maps[i].append(pm)
                if hp["automask_loss"]:
                    idm = loss.calc_photometric_loss([ref_image], [tgt])[0]
                    res["identity_%d" % s] = idm.numpy().copy()
                    for i in range(n):
                        maps[i].append(idm)
            if True:   # also for ssim_loss_weight == 0: the maps are 3-channel then and the index is entry * 3 + channel
                for i in range(n):
                    stack = torch.cat(maps[i], 1)
                    mn, idx = stack.min(1, True)
                    res["sel_%d" % i] = idx.to(torch.uint8).numpy().copy()
                    res["minmap_%d" % i] = mn.numpy().copy()
            res["Kinv"] = cam.Kinv.numpy().copy()
    return res
