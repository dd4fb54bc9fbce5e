"""CUDA path vs the CPU oracle at the BENCHMARKED shapes (BASELINE.json configs[0..3]) -- interior tiles, many tiles per
image, many images per launch.  The oracle (oracle/mgvs_oracle.c, OpenMP) needs ~2 s per 1024x2048 image fwd+bwd, so each
case is a few seconds of CPU work; its result is computed once and shared by the two backward kernels.

  C1  B1  192x640   n=3 and n=1   ("the parity gate")
  C2  B16 192x640   n=3
  C3  B8  512x1024  n=4
  C4  B8  1024x2048 n=3           (the per-GPU slice of config[3] at 8 GPUs; the B=64 batch on one GPU is checked through a
                                   repeat-the-batch property in test_big_batch_property)

Bars (north_star): losses 1e-5 relative, gradients 1e-4 relative (L2 over the map and max-norm = max|a-b| / max|b|, see
helpers.maxrel), selection bit-exact.
"""
import functools

import numpy as np
import pytest
import torch

from helpers import GRAD_RTOL, LOSS_RTOL, l2rel, maxrel, relerr

pytestmark = pytest.mark.gpu

HP = dict(ssim_loss_weight=0.85, photometric_loss_weight=1.0, smoothing_loss_weight=1e-3, automask_loss=True,
          photometric_reduce_op="min", padding_mode="zeros")
SHAPES = {
    "c1_n3": (1, 192, 640, 3),
    "c1_n1": (1, 192, 640, 1),
    "c2": (16, 192, 640, 3),
    "c3": (8, 512, 1024, 4),
    "c4": (8, 1024, 2048, 3),
}


@functools.lru_cache(maxsize=1)          # one shape at a time: C4's inputs + outputs are ~1 GB of host memory
def _oracle(name):
    from mgnet_b200.synthetic import make_inputs
    from oracle.oracle import Oracle
    B, H, W, n = SHAPES[name]
    pred, tgt = make_inputs(B, H, W, n, seed=31)
    o = Oracle(pred, tgt)
    f = o.forward()
    g = o.backward(1.0, 1.0)
    return pred, tgt, {"loss_photometric": f["loss_photometric"], "loss_smoothness": f["loss_smoothness"], "sel": f["sel"].copy(),
                       "grad_depth": g["grad_depth"], "grad_poses": g["grad_poses"]}


@pytest.mark.parametrize("name", list(SHAPES))
@pytest.mark.parametrize("backward", ["stash", "recompute"])
def test_benchmarked_shape_against_oracle(name, backward):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from test_gpu_parity import _run_cuda
    dev = torch.device("cuda:0")
    pred, tgt, ref = _oracle(name)
    r = _run_cuda(pred, tgt, HP, dev, backward=backward)
    n = len(pred["depth"])
    assert relerr(r["loss_photometric"], ref["loss_photometric"]) <= LOSS_RTOL
    assert relerr(r["loss_smoothness"], ref["loss_smoothness"]) <= LOSS_RTOL
    mism = int((r["sel"] != ref["sel"]).sum())
    assert mism == 0, "%s: %d selection mismatches of %d" % (name, mism, ref["sel"].size)
    for i in range(n):
        assert l2rel(r["grad_depth"][i], ref["grad_depth"][i]) <= GRAD_RTOL
        assert maxrel(r["grad_depth"][i], ref["grad_depth"][i]) <= GRAD_RTOL
    assert l2rel(r["grad_poses"], ref["grad_poses"]) <= GRAD_RTOL
    assert maxrel(r["grad_poses"], ref["grad_poses"]) <= GRAD_RTOL


def test_big_batch_property():
    """BASELINE config[3] on ONE GPU (B=64, 1024x2048, n=3; ~31 GB): the batch is the C4 batch (checked against the oracle
    above) repeated 8 times, so the loss must equal the B=8 loss, every selection map must repeat and every per-image
    gradient must be 1/8 of the B=8 one (the masked means divide by the batch-wide counts)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs ~31 GB of device memory")
    from test_gpu_parity import _run_cuda
    dev = torch.device("cuda:0")
    pred, tgt, ref = _oracle("c4")
    rep = 8
    pred64 = {"depth": [d.repeat(rep, 1, 1, 1) for d in pred["depth"]], "poses": pred["poses"].repeat(rep, 1, 1)}
    tgt64 = {k: v.repeat(rep, *([1] * (v.dim() - 1))) for k, v in tgt.items()}
    r = _run_cuda(pred64, tgt64, HP, dev, backward="stash")
    assert relerr(r["loss_photometric"], ref["loss_photometric"]) <= LOSS_RTOL
    assert relerr(r["loss_smoothness"], ref["loss_smoothness"]) <= LOSS_RTOL
    B = pred["poses"].shape[0]
    sel = r["sel"].reshape(r["sel"].shape[0], rep, B, *r["sel"].shape[2:])
    assert int((sel != ref["sel"][:, None]).sum()) == 0
    for i in range(len(pred["depth"])):
        g = r["grad_depth"][i].reshape(rep, B, *r["grad_depth"][i].shape[1:])
        assert np.array_equal(g[0], g[rep - 1])
        assert l2rel(g[0] * rep, ref["grad_depth"][i]) <= GRAD_RTOL
    gp = r["grad_poses"].reshape(rep, B, 2, 6)
    assert l2rel(gp[3] * rep, ref["grad_poses"]) <= GRAD_RTOL
