"""DGC depth rescaling (reference mgnet/postprocessing/depth_post_proc.py:11-185).

CPU (``-m "not gpu"``): the C oracle (oracle/dgc_oracle.c) against fixtures produced by the unmodified reference
(tests/golden/make_golden_dgc.py) and, when /root/reference is mounted, against the live reference on a fresh scene.
GPU (``-m gpu``): the sm_100a kernels through the C ABI (mgvs_dgc_rescale / mgvs_dgc_heights) against the oracle and
the fixtures.  The bar is bit-exactness everywhere: the ground mask and the median are selections, and every
floating-point step follows the reference's fp32 rounding sequence (NaN payloads excepted: filtered points are NaN).
"""
import glob
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLDEN)

from mgnet_b200.synthetic import make_dgc_inputs  # noqa: E402
from oracle.oracle import dgc_depth_prediction  # noqa: E402

DGC_FIXTURES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "dgc_*.npz")))


def bits_differ(a, b):
    """Number of elements whose bit patterns differ, NaNs compared as equal to NaNs."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.dtype == np.float32:
        both_nan = np.isnan(a) & np.isnan(b)
        return int(((a.view(np.uint32) != b.view(np.uint32)) & ~both_nan).sum())
    return int((a != b).sum())


def load_fixture(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    use_pan = bool(g["use_panoptic"])
    return g, use_pan, [int(x) for x in g["filter_ids"]]


def test_fixtures_present():
    assert len(DGC_FIXTURES) >= 3


@pytest.mark.parametrize("name", DGC_FIXTURES)
def test_oracle_matches_reference_fixture(name):
    g, use_pan, ids = load_fixture(name)
    o = dgc_depth_prediction(g["in_depth"], g["in_camera"], g["in_height"], g["in_panoptic"] if use_pan else None,
                             0 if use_pan else -1, ids)
    assert not o["empty"]
    assert bits_differ(o["normals"], g["ref_normals"]) == 0
    assert bits_differ(o["heights"], g["ref_heights"]) == 0
    assert bits_differ(o["ground"], g["ref_ground"]) == 0
    assert bits_differ(np.float32(o["scale"]).reshape(1), g["ref_scale"]) == 0
    assert bits_differ(o["depth"], g["ref_depth"]) == 0
    assert bits_differ(o["points"], g["ref_points"]) == 0
    if ids:
        assert np.isnan(g["ref_points"]).any() and (g["ref_depth"] == 0).any()


@pytest.mark.parametrize("name", DGC_FIXTURES)
def test_torch_port_matches_reference_fixture(name):
    """oracle/torch_port.reference_dgc (the comparator tests/tools/time_dgc.py times) issues the reference's ATen sequence."""
    from oracle.torch_port import reference_dgc
    g, use_pan, ids = load_fixture(name)
    out, P, s = reference_dgc(torch.from_numpy(g["in_depth"]).clone(), torch.from_numpy(g["in_camera"]),
                              torch.from_numpy(g["in_height"]), torch.from_numpy(g["in_panoptic"]) if use_pan else None,
                              0 if use_pan else -1, ids)
    assert bits_differ(out.numpy(), g["ref_depth"]) == 0 and bits_differ(P.numpy(), g["ref_points"]) == 0
    assert bits_differ(s.numpy(), g["ref_scale"]) == 0


def test_oracle_matches_live_reference_large():
    """A KITTI-sized scene through the unmodified reference, when it is mounted (build container only)."""
    import ref_loader
    if not ref_loader.available():
        pytest.skip("reference not mounted")
    import make_golden_dgc as mk
    geo, _, _ = ref_loader.load()
    m = mk.load_post_proc()
    d = make_dgc_inputs(192, 640, seed=11, scale_true=5.0)
    ref = mk.run_reference(m, geo, d, True, [10000])
    o = dgc_depth_prediction(d["depth"], d["camera_matrix"], d["real_camera_height"], d["panoptic_seg"], 0, [10000])
    assert bits_differ(o["heights"], ref["ref_heights"]) == 0
    assert bits_differ(np.float32(o["scale"]).reshape(1), ref["ref_scale"]) == 0
    assert bits_differ(o["depth"], ref["ref_depth"]) == 0
    assert bits_differ(o["points"], ref["ref_points"]) == 0
    assert abs(float(o["scale"]) - 5.0) < 0.05 * 5.0     # the synthetic scene's true scale is recovered


def test_oracle_edge_cases():
    d = make_dgc_inputs(24, 40, seed=4)
    # empty ground mask: torch.median of an empty selection is NaN, so is the scale and everything it multiplies
    o = dgc_depth_prediction(d["depth"], d["camera_matrix"], d["real_camera_height"], d["panoptic_seg"], 424242, [])
    assert o["empty"] and o["count"] == 0 and np.isnan(o["scale"]) and np.isnan(o["depth"]).all()
    # a NaN height inside the ground mask poisons the median (torch.median propagates NaN)
    dep = d["depth"].clone()
    ys, xs = np.nonzero(d["panoptic_seg"].numpy() == 0)
    dep[0, 0, ys[0], xs[0]] = float("nan")
    o = dgc_depth_prediction(dep, d["camera_matrix"], d["real_camera_height"], d["panoptic_seg"], 0, [])
    assert not o["empty"] and np.isnan(o["scale"])
    # inverse camera matrix given (exportable_post_proc.py:66-68): same result when it is Camera.Kinv
    K = d["camera_matrix"][0].numpy()
    Kinv = K.copy()
    Kinv[0, 0], Kinv[1, 1] = np.float32(1) / K[0, 0], np.float32(1) / K[1, 1]
    Kinv[0, 2], Kinv[1, 2] = (np.float32(-1) * K[0, 2]) / K[0, 0], (np.float32(-1) * K[1, 2]) / K[1, 1]
    a = dgc_depth_prediction(d["depth"], d["camera_matrix"], d["real_camera_height"], d["panoptic_seg"], 0, [])
    b = dgc_depth_prediction(d["depth"], Kinv, d["real_camera_height"], d["panoptic_seg"], 0, [], camera_is_inverse=True)
    assert bits_differ(a["points"], b["points"]) == 0 and a["scale"] == b["scale"]


def test_module_rejects_cpu_tensors_and_bad_arguments():
    from mgnet_b200.postprocessing import get_depth_prediction
    d = make_dgc_inputs(24, 40, seed=4)
    with pytest.raises(AssertionError, match="camera_matrix is necessary"):
        get_depth_prediction(d["depth"].clone(), True, None, d["real_camera_height"])
    with pytest.raises(AssertionError, match="real_camera_height is necessary"):
        get_depth_prediction(d["depth"].clone(), True, d["camera_matrix"], None)
    with pytest.raises(AssertionError, match="road_class_id is necessary"):
        get_depth_prediction(d["depth"].clone(), True, d["camera_matrix"], d["real_camera_height"], d["panoptic_seg"])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        get_depth_prediction(d["depth"].clone(), True, d["camera_matrix"], d["real_camera_height"], d["panoptic_seg"], 0)


# ---------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------
def _dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _run_gpu(depth, cam, height, pan, road, ids):
    from mgnet_b200.postprocessing import dgc_camera_heights, get_depth_prediction
    dev = _dev()
    dep = torch.as_tensor(depth).clone().to(dev)
    camt = torch.as_tensor(cam).to(dev)
    hgt = torch.as_tensor(height).to(dev)
    pant = torch.as_tensor(pan).to(dev) if pan is not None else None
    heights, ground = dgc_camera_heights(dep, camt, pant, road)
    before = dep.clone()
    out_d, out_p = get_depth_prediction(dep, True, camt, hgt, pant, road, ids)
    assert out_d.data_ptr() == dep.data_ptr() and out_d.shape == before.shape[-2:]      # in place, squeezed
    return {"depth": out_d.cpu().numpy(), "points": out_p.cpu().numpy(), "heights": heights[0].cpu().numpy(),
            "ground": ground[0].cpu().numpy().astype(np.uint8)}


@pytest.mark.gpu
@pytest.mark.parametrize("name", DGC_FIXTURES)
def test_gpu_matches_reference_fixture(name):
    g, use_pan, ids = load_fixture(name)
    r = _run_gpu(g["in_depth"], g["in_camera"], g["in_height"], g["in_panoptic"] if use_pan else None,
                 0 if use_pan else -1, ids)
    assert bits_differ(r["heights"], g["ref_heights"]) == 0
    assert bits_differ(r["ground"], g["ref_ground"]) == 0
    assert bits_differ(r["depth"], g["ref_depth"]) == 0
    assert bits_differ(r["points"], g["ref_points"]) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,seed,use_pan,pan_dtype", [
    (192, 640, 21, True, torch.int64), (192, 640, 22, False, None), (512, 1024, 23, True, torch.int32),
    (45, 131, 24, True, torch.int64), (3, 3, 25, False, None), (1024, 2048, 26, True, torch.int64),
])
def test_gpu_matches_oracle(H, W, seed, use_pan, pan_dtype):
    d = make_dgc_inputs(H, W, seed=seed, scale_true=2.0 + seed % 5)
    pan = d["panoptic_seg"] if use_pan else None
    ids = [10000, 14002] if use_pan else []
    o = dgc_depth_prediction(d["depth"], d["camera_matrix"], d["real_camera_height"], pan, 0 if use_pan else -1, ids)
    r = _run_gpu(d["depth"], d["camera_matrix"], d["real_camera_height"], pan.to(pan_dtype) if use_pan else None,
                 0 if use_pan else -1, ids)
    assert bits_differ(r["heights"], o["heights"]) == 0
    assert bits_differ(r["ground"], o["ground"]) == 0
    assert bits_differ(r["depth"], o["depth"]) == 0
    assert bits_differ(r["points"], o["points"]) == 0


@pytest.mark.gpu
def test_gpu_batched_scale_count_and_edge_cases():
    from mgnet_b200.postprocessing import dgc_rescale, get_depth_prediction
    dev = _dev()
    scenes = [make_dgc_inputs(96, 160, seed=31 + k, scale_true=3.0 + k) for k in range(3)]
    dep = torch.cat([s["depth"] for s in scenes]).to(dev)
    pan = torch.stack([s["panoptic_seg"] for s in scenes]).to(dev)
    pan[2] = 77                                            # image 2: empty ground mask
    dep[1, 0, 95, 80] = float("nan")                       # image 1: a NaN inside the road region
    assert int(pan[1, 95, 80]) == 0
    cam = torch.cat([s["camera_matrix"] for s in scenes]).to(dev)
    hgt = torch.cat([s["real_camera_height"] for s in scenes]).to(dev)
    ref = [dgc_depth_prediction(dep[k].cpu(), cam[k].cpu(), hgt[k:k + 1].cpu(), pan[k].cpu(), 0, []) for k in range(3)]
    points, scale, count = dgc_rescale(dep, cam, hgt, pan, 0, [])
    scale, count = scale.cpu().numpy(), count.cpu().numpy()
    assert bits_differ(scale[:1], np.float32(ref[0]["scale"]).reshape(1)) == 0
    assert np.isnan(scale[1]) and np.isnan(ref[1]["scale"]) and count[1] > 0
    assert np.isnan(scale[2]) and count[2] == 0 and ref[2]["empty"]
    assert bits_differ(dep[0, 0].cpu().numpy(), ref[0]["depth"]) == 0
    assert bits_differ(points[0].cpu().numpy(), ref[0]["points"]) == 0
    assert torch.isnan(dep[1:]).all()
    # use_dgc_scaling=False: only the class filter runs, no points (depth_post_proc.py:42, 59-69)
    s = scenes[0]
    d0 = s["depth"].clone().to(dev)
    out, pts = get_depth_prediction(d0, False, None, None, s["panoptic_seg"].to(dev), 0, [10000])
    expect = s["depth"][0, 0].clone()
    expect[s["panoptic_seg"] == 10000] = 0
    assert pts is None and torch.equal(out.cpu(), expect)
    # inverse camera matrix given (exportable_post_proc.py:66-68)
    from mgnet_b200.geometry import Camera
    kinv = Camera(K=s["camera_matrix"].to(dev)).Kinv
    d1, d2 = s["depth"].clone().to(dev), s["depth"].clone().to(dev)
    p1, s1, _ = dgc_rescale(d1, s["camera_matrix"].to(dev), s["real_camera_height"].to(dev), s["panoptic_seg"].to(dev), 0, [])
    p2, s2, _ = dgc_rescale(d2, kinv, s["real_camera_height"].to(dev), s["panoptic_seg"].to(dev), 0, [], camera_is_inverse=True)
    assert torch.equal(s1, s2) and torch.equal(p1, p2) and torch.equal(d1, d2)


@pytest.mark.gpu
def test_gpu_median_is_order_statistic_at_full_size():
    """Size-independent property at the Cityscapes resolution: the recovered scale equals real_height / (lower median of
    the ground heights), with the heights taken from the kernel's own diagnostics and the median from torch.sort."""
    from mgnet_b200.postprocessing import dgc_camera_heights, dgc_rescale
    dev = _dev()
    d = make_dgc_inputs(1024, 2048, seed=41, scale_true=6.0)
    dep = d["depth"].to(dev)
    cam, hgt = d["camera_matrix"].to(dev), d["real_camera_height"].to(dev)
    for pan in (d["panoptic_seg"].to(dev), None):
        heights, ground = dgc_camera_heights(dep, cam, pan, 0 if pan is not None else -1)
        sel = heights[ground]
        med = torch.sort(sel)[0][(sel.numel() - 1) // 2]
        work = dep.clone()
        _, scale, count = dgc_rescale(work, cam, hgt, pan, 0 if pan is not None else -1, [])
        assert int(count[0]) == sel.numel()
        assert torch.equal(scale, torch.reciprocal(med).unsqueeze(0) * hgt)
        assert torch.equal(work, dep * scale)
        assert abs(float(scale[0]) - 6.0) < 0.3
    # graph capture: no allocation-free guarantee is claimed for the Python wrapper, but the C entry point is capturable
    # (one memset node + four kernel nodes, no host synchronisation)
