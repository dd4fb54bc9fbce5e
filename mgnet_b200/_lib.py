"""ctypes binding of libmgvs.so (include/mgvs.h).  No torch types cross this boundary.

The library is built in-tree by ``build()`` (nvcc, sm_100a only) so that it travels with the source
snapshot.  There is NO CPU fallback: if the library is missing or does not load, importing callers get
a RuntimeError.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("MGVS_LIB_PATH") or os.path.join(_HERE, "libmgvs.so")   # override: kernel-variant experiments only
MAX_SCALES = 8
ABI_VERSION = 7
IMAGE_F32, IMAGE_U8 = 0, 1
NUM_SOURCES = 2

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


class MgvsProblem(ctypes.Structure):
    _fields_ = [
        ("B", ctypes.c_int), ("H", ctypes.c_int), ("W", ctypes.c_int), ("n", ctypes.c_int),
        ("target", ctypes.c_void_p),
        ("source", ctypes.c_void_p * NUM_SOURCES),
        ("inv_depth", ctypes.c_void_p * MAX_SCALES),
        ("camera", ctypes.c_void_p),
        ("cam_batch_stride", ctypes.c_longlong), ("cam_row_stride", ctypes.c_longlong),
        ("poses", ctypes.c_void_p),
        ("mask", ctypes.c_void_p),
        ("ssim_weight", ctypes.c_float), ("one_minus_ssim_weight", ctypes.c_float),
        ("photometric_weight", ctypes.c_float), ("smoothing_weight", ctypes.c_float),
        ("automask", ctypes.c_int), ("reduce_op", ctypes.c_int), ("padding_mode", ctypes.c_int),
        ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t),
        ("image_dtype", ctypes.c_int),
        ("stash", ctypes.c_void_p), ("stash_bytes", ctypes.c_size_t),
        ("inv_height", ctypes.c_int * MAX_SCALES), ("inv_width", ctypes.c_int * MAX_SCALES),
        ("pose_mats", ctypes.c_void_p),
    ]


DGC_MAX_FILTER = 16
MAX_LOSSES = 16
PADDING_MODES = {"zeros": 0, "border": 1, "reflection": 2}
PANOPTIC_NONE, PANOPTIC_I64, PANOPTIC_I32 = 0, 1, 2


class MgvsDgcProblem(ctypes.Structure):
    """include/mgvs.h: MgvsDgcProblem (DGC depth rescaling, reference depth_post_proc.py:11-185)."""
    _fields_ = [
        ("B", ctypes.c_int), ("H", ctypes.c_int), ("W", ctypes.c_int),
        ("depth", ctypes.c_void_p),
        ("camera", ctypes.c_void_p),
        ("cam_batch_stride", ctypes.c_longlong), ("cam_row_stride", ctypes.c_longlong),
        ("camera_is_inverse", ctypes.c_int),
        ("real_camera_height", ctypes.c_void_p),
        ("height_stride", ctypes.c_longlong),
        ("panoptic", ctypes.c_void_p),
        ("panoptic_dtype", ctypes.c_int),
        ("use_dgc", ctypes.c_int),
        ("road_class_id", ctypes.c_longlong),
        ("filter_ids", ctypes.c_longlong * DGC_MAX_FILTER),
        ("n_filter", ctypes.c_int),
        ("points", ctypes.c_void_p),
        ("scale", ctypes.c_void_p),
        ("count", ctypes.c_void_p),
        ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t),
    ]


MAX_RANKS = 16


class MgvsPeerExchange(ctypes.Structure):
    """include/mgvs.h: MgvsPeerExchange (peer-memory exchange of the partial sums)."""
    _fields_ = [("rank", ctypes.c_int), ("world", ctypes.c_int), ("peer_base", ctypes.c_void_p * MAX_RANKS),
                ("max_spins", ctypes.c_ulonglong)]


def _sources():
    return sorted(os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cu", ".cuh"))) + [
        os.path.join(os.path.dirname(_HERE), "include", "mgvs.h")
    ]


def needs_build() -> bool:
    if os.environ.get("MGVS_LIB_PATH"):
        return False
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> mgnet_b200/libmgvs.so (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, os.path.join(_CSRC, "mgvs_api.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), r.stderr[-4000:]))
    if verbose:
        print(r.stderr)
    return LIB_PATH


_lib = None


def lib():
    """Loads libmgvs.so (building it first if the sources are newer) and declares the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if needs_build():
        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            build()
        elif not os.path.isfile(LIB_PATH):
            raise RuntimeError("libmgvs.so is missing and nvcc is not available; there is no CPU fallback")
    try:
        L = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise RuntimeError("cannot load %s: %s (there is no CPU fallback)" % (LIB_PATH, e))
    vp, ll, ci = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int
    PP = ctypes.POINTER(MgvsProblem)
    L.mgvs_abi_version.restype = ci
    L.mgvs_abi_version.argtypes = []
    L.mgvs_last_error.restype = ctypes.c_char_p
    L.mgvs_last_error.argtypes = []
    L.mgvs_num_sums.restype = ci
    L.mgvs_num_sums.argtypes = [ci]
    L.mgvs_workspace_bytes.restype = ctypes.c_size_t
    L.mgvs_workspace_bytes.argtypes = [ci, ci, ci, ci]
    L.mgvs_workspace_bytes_ex.restype = ctypes.c_size_t
    L.mgvs_workspace_bytes_ex.argtypes = [ci, ci, ci, ci, ci]
    L.mgvs_workspace_bytes_ex2.restype = ctypes.c_size_t
    L.mgvs_workspace_bytes_ex2.argtypes = [ci, ci, ci, ci, ci, ci]
    L.mgvs_stash_bytes.restype = ctypes.c_size_t
    L.mgvs_stash_bytes.argtypes = [ci, ci, ci, ci]
    L.mgvs_stash_bytes_ex.restype = ctypes.c_size_t
    L.mgvs_stash_bytes_ex.argtypes = [ci, ci, ci, ci, ci]
    L.mgvs_forward.restype = ci
    L.mgvs_forward.argtypes = [PP, vp, vp, vp]
    L.mgvs_forward_losses.restype = ci
    L.mgvs_forward_losses.argtypes = [PP, vp, vp, vp, vp]
    L.mgvs_finalize.restype = ci
    L.mgvs_finalize.argtypes = [PP, vp, vp, vp]
    L.mgvs_backward.restype = ci
    L.mgvs_backward.argtypes = [PP, vp, vp, vp, ctypes.POINTER(vp), vp, vp]
    L.mgvs_view_synthesis.restype = ci
    L.mgvs_view_synthesis.argtypes = [ci, ci, ci, vp, vp, vp, ll, ll, vp, vp, vp, vp]
    L.mgvs_view_synthesis_ex.restype = ci
    L.mgvs_view_synthesis_ex.argtypes = [ci, ci, ci, vp, vp, vp, ll, ll, vp, ll, ll, vp, ci, vp, vp, vp]
    L.mgvs_reconstruct.restype = ci
    L.mgvs_reconstruct.argtypes = [ci, ci, ci, vp, vp, ll, ll, vp, vp]
    L.mgvs_project.restype = ci
    L.mgvs_project.argtypes = [ci, ci, ci, vp, vp, ll, ll, vp, vp, vp]
    DP = ctypes.POINTER(MgvsDgcProblem)
    L.mgvs_dgc_workspace_bytes.restype = ctypes.c_size_t
    L.mgvs_dgc_workspace_bytes.argtypes = [ci, ci, ci]
    L.mgvs_dgc_rescale.restype = ci
    L.mgvs_dgc_rescale.argtypes = [DP, vp]
    L.mgvs_dgc_heights.restype = ci
    L.mgvs_dgc_heights.argtypes = [DP, vp, vp, vp]
    L.mgvs_uncertainty_forward.restype = ci
    L.mgvs_uncertainty_forward.argtypes = [ci, vp, vp, ctypes.POINTER(ctypes.c_float), vp, vp, vp]
    L.mgvs_uncertainty_backward.restype = ci
    L.mgvs_uncertainty_backward.argtypes = [ci, vp, vp, ctypes.POINTER(ctypes.c_float), vp, vp, vp, vp]
    L.mgvs_exchange_bytes.restype = ctypes.c_size_t
    L.mgvs_exchange_bytes.argtypes = []
    L.mgvs_exchange_finalize.restype = ci
    L.mgvs_exchange_finalize.argtypes = [PP, ctypes.POINTER(MgvsPeerExchange), vp, vp, vp]
    L.mgvs_pose_tail_forward.restype = ci
    L.mgvs_pose_tail_forward.argtypes = [ci, ci, ci, vp, ctypes.c_float, vp, vp]
    L.mgvs_pose_tail_backward.restype = ci
    L.mgvs_pose_tail_backward.argtypes = [ci, ci, ci, vp, ctypes.c_float, vp, vp]
    L.mgvs_unpack_mask.restype = ci
    L.mgvs_unpack_mask.argtypes = [ll, ci, vp, vp, vp]
    L.mgvs_test_div.restype = ci
    L.mgvs_test_div.argtypes = [vp, vp, vp, ll, vp]
    if L.mgvs_abi_version() != ABI_VERSION:
        raise RuntimeError("libmgvs.so ABI version mismatch")
    _lib = L
    return L


EXPORTED_SYMBOLS = (
    "mgvs_abi_version", "mgvs_last_error", "mgvs_num_sums", "mgvs_workspace_bytes", "mgvs_workspace_bytes_ex", "mgvs_workspace_bytes_ex2", "mgvs_stash_bytes", "mgvs_stash_bytes_ex", "mgvs_forward", "mgvs_forward_losses",
    "mgvs_finalize", "mgvs_backward", "mgvs_view_synthesis", "mgvs_view_synthesis_ex", "mgvs_reconstruct", "mgvs_project", "mgvs_test_div", "mgvs_unpack_mask", "mgvs_pose_tail_forward", "mgvs_pose_tail_backward",
    "mgvs_dgc_workspace_bytes", "mgvs_dgc_rescale", "mgvs_dgc_heights",
    "mgvs_uncertainty_forward", "mgvs_uncertainty_backward",
    "mgvs_exchange_bytes", "mgvs_exchange_finalize",
)


def check(rc: int, what: str = "mgvs"):
    """Maps C return codes onto the exception types the reference raises for the same condition."""
    if rc == 0:
        return
    msg = lib().mgvs_last_error().decode("utf-8", "replace")
    if rc == -2:
        raise NotImplementedError("%s: %s" % (what, msg))
    if rc == -1:
        raise ValueError("%s: %s" % (what, msg))
    raise RuntimeError("%s failed (%d): %s" % (what, rc, msg))
