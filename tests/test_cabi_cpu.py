"""CPU: the C-ABI library builds for sm_100a without a GPU, loads, and exports every symbol that
include/mgvs.h declares (no compute calls here); host-side argument validation."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "mgvs.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mgvs_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported():
    from mgnet_b200 import _lib
    path = _lib.build()
    L = ctypes.CDLL(path)
    decl = _declared_symbols()
    assert len(decl) >= 10
    for sym in decl:
        assert hasattr(L, sym), "include/mgvs.h declares %s but libmgvs.so does not export it" % sym
    assert sorted(_lib.EXPORTED_SYMBOLS) == decl


def test_host_side_queries():
    from mgnet_b200 import _lib
    L = _lib.lib()
    assert L.mgvs_abi_version() == _lib.ABI_VERSION == 7
    assert L.mgvs_num_sums(3) == 12
    ws = L.mgvs_workspace_bytes(16, 192, 640, 3)
    assert ws > 0 and ws % 256 == 0
    assert L.mgvs_stash_bytes(16, 192, 640, 3) == 3 * 16 * 3 * 192 * 640 * 16 + 2 * 16 * 192 * 640 * 4   # 48 B per (pixel, scale) + 8 B per pixel
    assert L.mgvs_stash_bytes(1, 50, 70, 2) >= 2 * 1 * 3 * 50 * 4 * 18 * 16 + 2 * 50 * 72 * 4   # ragged W: ceil(70/4) column groups
    # fused upsample keeps the n full-resolution inverse-depth maps of its pre-pass in the workspace (256-byte aligned maps)
    assert L.mgvs_workspace_bytes_ex2(16, 192, 640, 3, 0, 0) == ws
    assert L.mgvs_workspace_bytes_ex2(16, 192, 640, 3, 0, 1) == ws + 3 * 16 * 192 * 640 * 4
    assert L.mgvs_workspace_bytes(0, 192, 640, 3) == 0
    assert L.mgvs_workspace_bytes(1, 192, 640, 9) == 0


def test_struct_layout_matches_header():
    """ctypes mirror must have the size the C compiler gives the struct (checked by compiling a probe)."""
    import subprocess
    import tempfile
    from mgnet_b200 import _lib
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "p.c")
        open(src, "w").write('#include <stdio.h>\n#include "mgvs.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(MgvsProblem), sizeof(MgvsDgcProblem), sizeof(MgvsPeerExchange));return 0;}\n')
        exe = os.path.join(d, "p")
        subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, src], check=True)
        size, dgc_size, xch_size = map(int, subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split())
    assert ctypes.sizeof(_lib.MgvsProblem) == size
    assert ctypes.sizeof(_lib.MgvsDgcProblem) == dgc_size
    assert ctypes.sizeof(_lib.MgvsPeerExchange) == xch_size
    assert _lib.lib().mgvs_exchange_bytes() == 1024 + 2 * 16 * 32 * 8


def test_module_rejects_unsupported_configs_loudly():
    from mgnet_b200 import MultiViewPhotometricLoss
    with pytest.raises(ValueError):
        MultiViewPhotometricLoss(0.85, 1.0, 1e-3, True, "min", "mirror")       # not a grid_sample padding mode
    for mode in ("zeros", "border", "reflection"):                             # all three of the reference's are implemented
        MultiViewPhotometricLoss(0.85, 1.0, 1e-3, True, "min", mode)
    with pytest.raises(AssertionError):
        MultiViewPhotometricLoss(0.85, 1.0, 1e-3, True, "mean", "zeros")
    MultiViewPhotometricLoss(0.85, 1.0, 1e-3, False, "mean", "zeros")           # legal with automask off (loss.py:106-109); implemented
    with pytest.raises(NotImplementedError):
        MultiViewPhotometricLoss(0.85, 1.0, 1e-3, False, "median", "zeros")
    MultiViewPhotometricLoss(0.0, 1.0, 1e-3, True, "min", "zeros")              # raw 3-channel L1 branch (loss.py:195-196): implemented


def test_cpu_tensors_fail_loudly_no_fallback():
    import torch
    from mgnet_b200 import MultiViewPhotometricLoss
    from mgnet_b200.synthetic import make_inputs
    pred, tgt = make_inputs(1, 32, 64, 1, seed=0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        MultiViewPhotometricLoss(0.85, 1.0, 1e-3, True, "min", "zeros")(pred, tgt)


def test_pack_mask_layout_is_numpy_packbits():
    """mgnet_b200.synthetic.pack_mask: row-wise numpy.packbits (MSB first, rows padded to bytes) -- the layout mgvs_unpack_mask expects."""
    import numpy as np
    import torch
    from mgnet_b200.synthetic import pack_mask
    g = torch.Generator().manual_seed(0)
    for W in (8, 13, 64, 75):
        m = torch.rand(2, 1, 5, W, generator=g) < 0.5
        p = pack_mask(m)
        assert p.dtype == torch.uint8 and tuple(p.shape) == (2, 1, 5, (W + 7) // 8)
        back = np.unpackbits(p.numpy(), axis=-1)[..., :W].astype(bool)
        assert np.array_equal(back, m.numpy())
        assert int(p[0, 0, 0, 0]) >> 7 == int(m[0, 0, 0, 0])          # most significant bit = first pixel of the row
