"""ctypes front-end of the CPU oracle (oracle/mgvs_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module; the product package ``mgnet_b200`` never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmgvs_oracle.so")
_DGC_LIB_PATH = os.path.join(_HERE, "libdgc_oracle.so")
MAX_SCALES = 8
S = 2


def _compile(src_name: str, lib_path: str, force: bool) -> str:
    src = os.path.join(_HERE, src_name)
    if (not force) and os.path.isfile(lib_path) and os.path.getmtime(lib_path) >= os.path.getmtime(src):
        return lib_path
    base = ["-O2", "-mfma", "-ffp-contract=off", "-fPIC", "-std=gnu11", "-shared", "-o", lib_path, src, "-lm"]
    last = None
    for cc in ("/usr/bin/gcc", "gcc", "cc"):
        for omp in (["-fopenmp"], []):
            try:
                subprocess.run([cc] + omp + base, check=True, capture_output=True, text=True)
                return lib_path
            except (subprocess.CalledProcessError, FileNotFoundError) as e:  # try next
                last = e
    raise RuntimeError("could not build the oracle: %s" % (getattr(last, "stderr", last),))


def build(force: bool = False) -> str:
    """Compiles the C restatements in place (gcc, OpenMP if available): the loss oracle and the DGC oracle."""
    _compile("dgc_oracle.c", _DGC_LIB_PATH, force)
    return _compile("mgvs_oracle.c", _LIB_PATH, force)


class _OrcIn(ctypes.Structure):
    _fields_ = [
        ("B", ctypes.c_int), ("H", ctypes.c_int), ("W", ctypes.c_int), ("n", ctypes.c_int),
        ("target", ctypes.c_void_p),
        ("source", ctypes.c_void_p * S),
        ("inv", ctypes.c_void_p * MAX_SCALES),
        ("cam", ctypes.c_void_p),
        ("cam_bs", ctypes.c_long), ("cam_rs", ctypes.c_long),
        ("poses", ctypes.c_void_p),
        ("mask", ctypes.c_void_p),
        ("ssim_w", ctypes.c_float), ("photo_w", ctypes.c_float), ("smooth_w", ctypes.c_float),
        ("automask", ctypes.c_int),
        ("padding_mode", ctypes.c_int),
        ("pose_mats", ctypes.c_void_p),
    ]


class _OrcOut(ctypes.Structure):
    _fields_ = [
        ("coords", ctypes.c_void_p), ("warped", ctypes.c_void_p), ("photo", ctypes.c_void_p),
        ("identity", ctypes.c_void_p), ("minmap", ctypes.c_void_p), ("sel", ctypes.c_void_p),
        ("posemat", ctypes.c_void_p), ("kinv", ctypes.c_void_p), ("sums", ctypes.c_void_p),
        ("loss_photo", ctypes.c_float), ("loss_smooth", ctypes.c_float),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_forward.restype = ctypes.c_int
        _lib.orc_forward.argtypes = [ctypes.POINTER(_OrcIn), ctypes.c_double, ctypes.c_int, ctypes.POINTER(_OrcOut)]
        _lib.orc_backward.restype = ctypes.c_int
        _lib.orc_backward.argtypes = [
            ctypes.POINTER(_OrcIn), ctypes.c_double, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_double, ctypes.c_double, ctypes.c_void_p * MAX_SCALES, ctypes.c_void_p, ctypes.c_void_p,
        ]
    return _lib


def _np(x, dtype=np.float32):
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x, dtype=dtype)


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


class Oracle:
    """Holds numpy copies of one problem instance; ``forward`` / ``backward`` mirror the C entry points.

    The constructor takes the reference's own dictionaries (loss.py:111-147).
    EULER_FMA selects how the two tiny 3x3 bmm's of euler2mat round (decided by the golden tests).
    """

    EULER_FMA = 0

    def __init__(self, predictions, targets, ssim_loss_weight=0.85, photometric_loss_weight=1.0,
                 smoothing_loss_weight=1e-3, automask_loss=True, padding_mode="zeros"):
        self.inv = [_np(d) for d in predictions["depth"]]
        self.n = len(self.inv)
        self.B, _, self.H, self.W = self.inv[0].shape
        self.poses = _np(predictions["poses"])
        # [B,S,6] Euler vectors (the reference's contract) or [B,S,3,4] / [B,S,4,4] pose matrices the caller built
        self.pose_mats = None
        if self.poses.ndim == 4:
            self.pose_mats = np.ascontiguousarray(self.poses[:, :, :3, :4])
            self.poses = np.zeros((self.poses.shape[0], S, 6), np.float32)
        self.tgt = _np(targets["image_orig"])
        self.src = [_np(targets["image_prev_orig"]), _np(targets["image_next_orig"])]
        self.cam = _np(targets["camera_matrix"])
        m = targets.get("reprojection_mask", None)
        self.mask = None if m is None else _np(m, np.uint8)
        self.ssim_w = float(ssim_loss_weight)
        self.photo_w = float(photometric_loss_weight)
        self.smooth_w = float(smoothing_loss_weight)
        self.automask = bool(automask_loss)
        self._in = _OrcIn()
        i = self._in
        i.B, i.H, i.W, i.n = self.B, self.H, self.W, self.n
        i.target = _ptr(self.tgt)
        for s in range(S):
            i.source[s] = _ptr(self.src[s])
        for k in range(self.n):
            i.inv[k] = _ptr(self.inv[k])
        i.cam = _ptr(self.cam)
        i.cam_bs = self.cam.shape[1] * self.cam.shape[2]
        i.cam_rs = self.cam.shape[2]
        i.poses = _ptr(self.poses)
        i.mask = _ptr(self.mask) if self.mask is not None else None
        i.ssim_w, i.photo_w, i.smooth_w = self.ssim_w, self.photo_w, self.smooth_w
        i.automask = int(self.automask)
        i.padding_mode = {"zeros": 0, "border": 1, "reflection": 2}[padding_mode]
        i.pose_mats = _ptr(self.pose_mats) if self.pose_mats is not None else None
        self.sums = None
        self.sel = None

    def forward(self, dumps=False):
        n, B, H, W = self.n, self.B, self.H, self.W
        out = _OrcOut()
        res = {}
        self.sel = np.zeros((n, B, H, W), np.uint8)
        self.sums = np.zeros(3 * n + 3, np.float64)
        res["posemat"] = np.zeros((B, S, 12), np.float32)
        res["kinv"] = np.zeros((B, 9), np.float32)
        out.sel, out.sums = _ptr(self.sel), _ptr(self.sums)
        out.posemat, out.kinv = _ptr(res["posemat"]), _ptr(res["kinv"])
        if dumps:
            res["coords"] = np.zeros((n, S, B, H, W, 2), np.float32)
            res["warped"] = np.zeros((n, S, B, 3, H, W), np.float32)
            res["photo"] = np.zeros((n, S, B, H, W), np.float32)
            res["identity"] = np.zeros((S, B, H, W), np.float32)
            res["minmap"] = np.zeros((n, B, H, W), np.float32)
            out.coords, out.warped, out.photo = _ptr(res["coords"]), _ptr(res["warped"]), _ptr(res["photo"])
            out.identity, out.minmap = _ptr(res["identity"]), _ptr(res["minmap"])
        rc = lib().orc_forward(ctypes.byref(self._in), self.ssim_w, self.EULER_FMA, ctypes.byref(out))
        if rc != 0:
            raise RuntimeError("orc_forward failed: %d" % rc)
        res["sel"] = self.sel
        res["sums"] = self.sums
        res["loss_photometric"] = np.float32(out.loss_photo)
        res["loss_smoothness"] = np.float32(out.loss_smooth)
        return res

    def backward(self, g_photo=1.0, g_smooth=1.0, sums=None, sel=None):
        if self.sel is None:
            self.forward()
        sums = self.sums if sums is None else np.ascontiguousarray(sums, np.float64)
        sel = self.sel if sel is None else np.ascontiguousarray(sel, np.uint8)
        grads = [np.zeros_like(d) for d in self.inv]
        gp = np.zeros((self.B, S, 6), np.float32)
        gRt = np.zeros((self.B, S, 12), np.float64)
        arr = (ctypes.c_void_p * MAX_SCALES)()
        for k in range(self.n):
            arr[k] = grads[k].ctypes.data
        rc = lib().orc_backward(ctypes.byref(self._in), self.ssim_w, self.EULER_FMA, _ptr(sel), _ptr(sums),
                                float(g_photo), float(g_smooth), arr, _ptr(gp), _ptr(gRt))
        if rc != 0:
            raise RuntimeError("orc_backward failed: %d" % rc)
        return {"grad_depth": grads, "grad_poses": gp, "grad_Rt": gRt}


# ---- DGC depth rescaling (oracle/dgc_oracle.c; reference depth_post_proc.py:11-185) ---------------------------
_dgc = None


def dgc_lib():
    global _dgc
    if _dgc is None:
        build()
        _dgc = ctypes.CDLL(_DGC_LIB_PATH)
        vp = ctypes.c_void_p
        _dgc.orc_dgc.restype = ctypes.c_int
        _dgc.orc_dgc.argtypes = [ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_long, ctypes.c_int, ctypes.c_float, vp,
                                 ctypes.c_longlong, vp, ctypes.c_int, vp, vp, vp, vp, vp, vp, vp]
    return _dgc


def dgc_depth_prediction(depth, camera_matrix, real_camera_height, panoptic_seg=None, road_class_id=-1,
                         depth_filter_class_ids=None, camera_is_inverse=False):
    """CPU oracle of ``get_depth_prediction(..., use_dgc_scaling=True)`` for one image.

    depth [H,W] (or [1,1,H,W]), camera_matrix [3,3] (or [1,3,3]).  Returns a dict with depth [H,W], points [3,H,W],
    normals [3,H,W], heights [H,W], ground [H,W] uint8, scale (np.float32), count, empty (bool: torch.median would raise).
    """
    d = _np(depth)
    H, W = d.shape[-2:]
    d = d.reshape(H, W)
    cam = _np(camera_matrix).reshape(-1)[-9:].reshape(3, 3) if _np(camera_matrix).size == 9 else _np(camera_matrix)[..., :3, :3].reshape(3, 3).copy()
    cam = np.ascontiguousarray(cam, np.float32)
    pan = None if panoptic_seg is None else _np(panoptic_seg, np.int64).reshape(H, W)
    ids = np.asarray(list(depth_filter_class_ids or []), np.int64)
    out = {
        "depth": np.zeros((H, W), np.float32), "points": np.zeros((3, H, W), np.float32),
        "normals": np.zeros((3, H, W), np.float32), "heights": np.zeros((H, W), np.float32),
        "ground": np.zeros((H, W), np.uint8),
    }
    scale = ctypes.c_float(0.0)
    count = ctypes.c_longlong(0)
    rh = float(_np(real_camera_height).reshape(-1)[0])
    rc = dgc_lib().orc_dgc(H, W, _ptr(d), _ptr(cam), 3, int(bool(camera_is_inverse)), rh,
                           _ptr(pan) if pan is not None else None, int(road_class_id),
                           _ptr(ids) if ids.size else None, int(ids.size), _ptr(out["depth"]), _ptr(out["points"]),
                           _ptr(out["normals"]), _ptr(out["heights"]), _ptr(out["ground"]), ctypes.byref(scale),
                           ctypes.byref(count))
    out["scale"] = np.float32(scale.value)
    out["count"] = int(count.value)
    out["empty"] = rc == 1
    return out
