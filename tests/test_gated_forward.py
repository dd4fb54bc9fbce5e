"""The margin-gated forward (MgvsProblem.forward_mode, include/mgvs.h) against the exact one on the same GPU.

The exact kernel (forward_mode="exact", the reference's fp32 rounding sequence at every pixel) is pinned to the oracle and the
reference's fixtures by test_gpu_parity / test_gpu_bench_shapes; here it is the checker for the gated kernel:

  * the selection mask must be IDENTICAL (north_star: bit-exact integer selection), on benign inputs and on inputs built to
    produce near-ties and exact ties (identical source frames, constant / very smooth images, everything out of bounds, the
    warped source equal to the un-warped one);
  * "recheck_all" (every pixel goes through the warp-level exact re-evaluation) must reproduce the exact kernel's sums bit for
    bit -- this pins the re-evaluation machinery itself;
  * the photometric loss must agree to 1e-6 relative (bar: 1e-5), the gradients to 5e-5 (bar: 1e-4; the gated forward writes its stash coefficients from the fast evaluation);
  * the per-pixel error bound must hold with room: max |fast - exact| / bound over EVERY pixel (recheck_all) stays below 1/2.
"""
import numpy as np
import pytest
import torch

from helpers import l2rel, maxrel, relerr

pytestmark = pytest.mark.gpu

HP = dict(ssim_loss_weight=0.85, photometric_loss_weight=1.0, smoothing_loss_weight=1e-3, automask_loss=True,
          photometric_reduce_op="min", padding_mode="zeros")


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _three_modes(pred, tgt, hp, dev, backward="stash"):
    from test_gpu_parity import _run_cuda
    return {m: _run_cuda(pred, tgt, hp, dev, backward=backward, forward_mode=m, diag=(m != "exact")) for m in ("exact", "gated", "recheck_all")}


def _check(r, n, npix, max_recheck_frac=None):
    ex, ga, ra = r["exact"], r["gated"], r["recheck_all"]
    assert np.array_equal(ga["sel"], ex["sel"]), "%d selection mismatches gated vs exact" % int((ga["sel"] != ex["sel"]).sum())
    assert np.array_equal(ra["sel"], ex["sel"])
    # every valid pixel re-evaluated: same per-pixel values as the exact kernel, same fixed-order sums
    assert ra["loss_photometric"] == ex["loss_photometric"] and ra["loss_smoothness"] == ex["loss_smoothness"]
    assert ra["diag"][0] == n * npix
    assert relerr(ga["loss_photometric"], ex["loss_photometric"]) <= 1e-6
    assert ga["loss_smoothness"] == ex["loss_smoothness"]
    for a, b in zip(ga["grad_depth"], ex["grad_depth"]):
        assert l2rel(a, b) <= 5e-5 and maxrel(a, b) <= 5e-5
    assert l2rel(ga["grad_poses"], ex["grad_poses"]) <= 5e-5
    # the bound: measured on every pixel of every scale
    assert ra["diag"][2] < 0.5, "max |fast - exact| / bound = %.3f" % ra["diag"][2]
    if max_recheck_frac is not None:
        assert ga["diag"][0] <= max_recheck_frac * n * npix, "re-evaluated %.3f%% of the pixels" % (100.0 * ga["diag"][0] / (n * npix))
    return ga["diag"], ra["diag"]


@pytest.mark.parametrize("backward", ["stash", "recompute"])
@pytest.mark.parametrize("case", [
    dict(B=2, H=192, W=640, n=3, noise=0.2),
    dict(B=2, H=192, W=640, n=3, noise=0.0, shift=True),                      # smooth images: d1/d2 large, many near-ties
    dict(B=3, H=50, W=70, n=2, noise=0.2),                                    # ragged: manual loader, partial tiles
    dict(B=2, H=96, W=320, n=4, noise=0.05, pad="border", pose_scale=0.05),
    dict(B=2, H=96, W=320, n=2, noise=0.2, pad="reflection", pose_scale=0.05),
    dict(B=2, H=96, W=320, n=2, noise=0.2, automask=False),
    dict(B=2, H=96, W=320, n=2, noise=0.2, mask=False, alpha=0.5),
])
def test_gated_matches_exact(case, backward):
    dev = _dev()
    from mgnet_b200.synthetic import make_inputs
    pred, tgt = make_inputs(case["B"], case["H"], case["W"], case["n"], seed=7, noise=case["noise"], shift_sources=case.get("shift", False),
                            pose_scale=case.get("pose_scale", 0.01), with_mask=case.get("mask", True))
    hp = dict(HP, padding_mode=case.get("pad", "zeros"), automask_loss=case.get("automask", True), ssim_loss_weight=case.get("alpha", 0.85))
    _check(_three_modes(pred, tgt, hp, dev, backward), case["n"], case["B"] * case["H"] * case["W"])


@pytest.mark.parametrize("kind", ["same_sources", "constant", "tiny_texture", "all_out_of_bounds", "identity_pose", "dark", "saturated"])
def test_gated_on_ties_and_near_ties(kind):
    """Inputs built so that candidates of the minimum coincide or nearly coincide."""
    dev = _dev()
    from mgnet_b200.synthetic import make_inputs
    B, H, W, n = 2, 64, 128, 2
    pred, tgt = make_inputs(B, H, W, n, seed=5, noise=0.1)
    g = torch.Generator().manual_seed(1)
    if kind == "same_sources":          # every pixel: warp_prev == warp_next and id_prev == id_next exactly (lowest index must win)
        tgt["image_next_orig"] = tgt["image_prev_orig"].clone()
        pred["poses"][:, 1] = pred["poses"][:, 0]
    elif kind == "constant":            # zero variance: d2 = c2, the bound is at its loosest
        for k in ("image_orig", "image_prev_orig", "image_next_orig"):
            tgt[k] = torch.full_like(tgt[k], 0.5)
    elif kind == "tiny_texture":        # 1-2 grey levels of texture on a flat background, like sky / road in real frames
        base = torch.full((B, 3, H, W), 0.4)
        for k in ("image_orig", "image_prev_orig", "image_next_orig"):
            tgt[k] = (base + torch.randint(0, 3, (B, 3, H, W), generator=g).float() / 255.0).contiguous()
    elif kind == "all_out_of_bounds":   # both warps sample only the zero padding: exact ties between the two warped candidates
        pred["poses"][:, :, 0] = 50.0
    elif kind == "identity_pose":       # warp == un-warped source up to rounding: warp vs identity near-ties everywhere
        pred["poses"].zero_()
    elif kind == "dark":
        for k in ("image_orig", "image_prev_orig", "image_next_orig"):
            tgt[k] = (tgt[k] * 0.02).contiguous()
    elif kind == "saturated":
        for k in ("image_orig", "image_prev_orig", "image_next_orig"):
            tgt[k] = (0.98 + tgt[k] * 0.02).contiguous()
    for automask in (True, False):
        d_g, d_r = _check(_three_modes(pred, tgt, dict(HP, automask_loss=automask), dev), n, B * H * W)
        print("%s automask=%d: re-evaluated %.2f%% of the pixels, %d selections corrected, max error/bound %.3f"
              % (kind, automask, 100.0 * d_g[0] / (n * B * H * W), d_g[1], d_r[2]))


def test_gated_recheck_rate_at_the_benchmarked_shape():
    """C2's synthetic frames (B16 192x640 n=3): the exact re-evaluation must stay a small fraction of the pixels."""
    dev = _dev()
    from mgnet_b200.synthetic import make_inputs
    B, H, W, n = 16, 192, 640, 3
    pred, tgt = make_inputs(B, H, W, n, seed=100)
    d_g, d_r = _check(_three_modes(pred, tgt, HP, dev), n, B * H * W, max_recheck_frac=0.02)
    print("C2: re-evaluated %.3f%% of the pixels, %d selections corrected, max error/bound %.3f" % (100.0 * d_g[0] / (n * B * H * W), d_g[1], d_r[2]))
