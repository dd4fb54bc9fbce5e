"""Times mgvs_forward(+finalize) and mgvs_backward through the C ABI for a few workloads (experiments).
Usage: [MGVS_LIB_PATH=variant.so] python scripts/time_kernels.py [c2 c3 c4]"""
import ctypes, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mgnet_b200 import _lib
from mgnet_b200.ops import LossConfig, _fill_problem
from mgnet_b200.synthetic import make_inputs
W = {"c1": (1, 192, 640, 3), "c2": (16, 192, 640, 3), "c3": (8, 512, 1024, 4), "c4": (8, 1024, 2048, 3), "c2n1": (16, 192, 640, 1)}
names = sys.argv[1:] or ["c2", "c3"]
dev = torch.device("cuda:0"); L = _lib.lib(); cfg = LossConfig()
res = []
for name in names:
    B, H, Wd, n = W[name]
    sets = []
    for k in range(3):
        pred, tgt = make_inputs(B, H, Wd, n, seed=100 + k)
        sets.append(({"depth": [d.to(dev) for d in pred["depth"]], "poses": pred["poses"].to(dev)}, {kk: v.to(dev) for kk, v in tgt.items()}))
    ws = torch.empty(int(L.mgvs_workspace_bytes(B, H, Wd, n)), dtype=torch.uint8, device=dev)
    stash = torch.empty(int(L.mgvs_stash_bytes(B, H, Wd, n)), dtype=torch.uint8, device=dev) if os.environ.get("MGVS_BACKWARD", "stash") == "stash" else None
    sel = torch.empty((n, B, H, Wd), dtype=torch.uint8, device=dev); sums = torch.empty(3 * n + 3, dtype=torch.float64, device=dev)
    losses = torch.empty(2, device=dev); g = torch.ones(2, device=dev)
    grads = [torch.empty_like(d) for d in sets[0][0]["depth"]]; gp = torch.empty_like(sets[0][0]["poses"])
    arr = (ctypes.c_void_p * n)(*[x.data_ptr() for x in grads]); st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    tf, tb = [], []
    for it in range(13):
        p, t = sets[it % 3]
        prob = _lib.MgvsProblem(); _fill_problem(prob, cfg, t["image_orig"], t["image_prev_orig"], t["image_next_orig"], p["depth"], t["camera_matrix"], p["poses"], t.get("reprojection_mask"), ws, stash)
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record(); _lib.check(L.mgvs_forward_losses(ctypes.byref(prob), sel.data_ptr(), sums.data_ptr(), losses.data_ptr(), st)); b.record()
        _lib.check(L.mgvs_backward(ctypes.byref(prob), sel.data_ptr(), sums.data_ptr(), g.data_ptr(), arr, gp.data_ptr(), st)); c.record(); torch.cuda.synchronize()
        if it >= 3: tf.append(a.elapsed_time(b)); tb.append(b.elapsed_time(c))
    f, bb = statistics.median(tf), statistics.median(tb)
    res.append("%s: fwd %.3f bwd %.3f ms -> %.3f Gpx/s (loss %.8f)" % (name, f, bb, B * H * Wd / ((f + bb) * 1e-3) / 1e9, losses[0].item()))
print(os.environ.get("MGVS_LIB_PATH", "default"), os.environ.get("MGVS_BACKWARD", "stash"), " | ".join(res))
