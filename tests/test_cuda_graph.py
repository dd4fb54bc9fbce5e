"""The C ABI claims every entry point is CUDA-graph capturable (no allocation, no host synchronisation, launches only on
the caller's stream -- include/mgvs.h).  Capture a whole fwd+bwd of the loss module and the DGC post-processing in
torch.cuda.CUDAGraph, replay them on fresh input values and compare with the eager calls bit for bit."""
import pytest
import torch

HP = dict(ssim_loss_weight=0.85, photometric_loss_weight=1.0, smoothing_loss_weight=1e-3, automask_loss=True,
          photometric_reduce_op="min", padding_mode="zeros")


def _eager(mod, pred, tgt, dev):
    p = {"depth": [d.to(dev).requires_grad_(True) for d in pred["depth"]], "poses": pred["poses"].to(dev).requires_grad_(True)}
    t = {k: v.to(dev) for k, v in tgt.items()}
    out = mod(p, t)
    (out["loss_photometric"] + out["loss_smoothness"]).backward()
    return ([out["loss_photometric"].detach().clone(), out["loss_smoothness"].detach().clone()],
            [d.grad.clone() for d in p["depth"]] + [p["poses"].grad.clone()], mod.last_selection.clone())


@pytest.mark.gpu
@pytest.mark.parametrize("backward,fuse", [("stash", False), ("recompute", False)])
def test_loss_fwd_bwd_replays_in_a_cuda_graph(backward, fuse):
    from mgnet_b200 import MultiViewPhotometricLoss
    from mgnet_b200.synthetic import make_inputs
    dev = torch.device("cuda:0")
    mod = MultiViewPhotometricLoss(backward=backward, **HP)
    predA, tgtA = make_inputs(2, 64, 128, 3, seed=51)
    predB, tgtB = make_inputs(2, 64, 128, 3, seed=52)
    # static buffers
    sp = {"depth": [d.to(dev).requires_grad_(True) for d in predA["depth"]], "poses": predA["poses"].to(dev).requires_grad_(True)}
    st = {k: v.to(dev) for k, v in tgtA.items()}
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            for x in sp["depth"] + [sp["poses"]]:
                x.grad = None
            o = mod(sp, st)
            (o["loss_photometric"] + o["loss_smoothness"]).backward()
    torch.cuda.current_stream().wait_stream(side)
    for x in sp["depth"] + [sp["poses"]]:
        x.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        o = mod(sp, st)
        (o["loss_photometric"] + o["loss_smoothness"]).backward()
    sel_static = mod.last_selection
    for pred, tgt in ((predB, tgtB), (predA, tgtA)):
        with torch.no_grad():
            for a, b in zip(sp["depth"], pred["depth"]):
                a.copy_(b)
            sp["poses"].copy_(pred["poses"])
            for k in st:
                st[k].copy_(tgt[k])
        graph.replay()
        torch.cuda.synchronize()
        got_losses = [o["loss_photometric"].clone(), o["loss_smoothness"].clone()]
        got_grads = [d.grad.clone() for d in sp["depth"]] + [sp["poses"].grad.clone()]
        got_sel = sel_static.clone()
        ref_losses, ref_grads, ref_sel = _eager(MultiViewPhotometricLoss(backward=backward, **HP), pred, tgt, dev)
        assert all(torch.equal(a, b) for a, b in zip(got_losses, ref_losses))
        assert all(torch.equal(a, b) for a, b in zip(got_grads, ref_grads))
        assert torch.equal(got_sel, ref_sel)


@pytest.mark.gpu
def test_module_graphed_mode_matches_eager_and_is_faster_when_launch_bound():
    """MultiViewPhotometricLoss.graphed(): the built-in CUDA-graph mode (make_graphed_callables) -- same results as the eager call on
    new input values, inside an autograd graph with non-trivial upstream gradients; and at C1 (B1 192x640) a shorter step."""
    import time
    from mgnet_b200 import MultiViewPhotometricLoss
    from mgnet_b200.synthetic import make_inputs
    dev = torch.device("cuda:0")
    mod = MultiViewPhotometricLoss(**HP)
    predA, tgtA = make_inputs(1, 192, 640, 3, seed=71)
    predB, tgtB = make_inputs(1, 192, 640, 3, seed=72)

    def dev_inputs(pred, tgt):
        return ({"depth": [d.to(dev).requires_grad_(True) for d in pred["depth"]], "poses": pred["poses"].to(dev).requires_grad_(True)},
                {k: v.to(dev) for k, v in tgt.items()})

    pA, tA = dev_inputs(predA, tgtA)
    graphed = mod.graphed(pA, tA)
    for pred, tgt in ((predB, tgtB), (predA, tgtA)):
        res = []
        for f in (graphed, MultiViewPhotometricLoss(**HP)):
            p, t = dev_inputs(pred, tgt)
            out = f(p, t)
            (2.0 * out["loss_photometric"] + 3.0 * out["loss_smoothness"]).backward()
            res.append(([out["loss_photometric"].detach().clone(), out["loss_smoothness"].detach().clone()],
                        [d.grad.clone() for d in p["depth"]] + [p["poses"].grad.clone()]))
        assert all(torch.equal(a, b) for a, b in zip(res[0][0], res[1][0]))
        assert all(torch.equal(a, b) for a, b in zip(res[0][1], res[1][1]))
    # launch-bound shape: the replayed step is shorter than the eager one
    times = {}
    for name, f in (("eager", mod), ("graphed", graphed)):
        p, t = dev_inputs(predA, tgtA)
        for _ in range(5):
            out = f(p, t); (out["loss_photometric"] + out["loss_smoothness"]).backward()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            for x in p["depth"] + [p["poses"]]:
                x.grad = None
            out = f(p, t); (out["loss_photometric"] + out["loss_smoothness"]).backward()
        torch.cuda.synchronize()
        times[name] = (time.perf_counter() - t0) / 50 * 1e3
    print("C1 step: eager %.3f ms, graphed %.3f ms" % (times["eager"], times["graphed"]))
    assert times["graphed"] < times["eager"]


@pytest.mark.gpu
def test_dgc_rescale_replays_in_a_cuda_graph():
    from mgnet_b200.postprocessing import dgc_rescale
    from mgnet_b200.synthetic import make_dgc_inputs
    dev = torch.device("cuda:0")
    a = make_dgc_inputs(96, 160, seed=61, scale_true=4.0)
    b = make_dgc_inputs(96, 160, seed=62, scale_true=9.0)
    depth = a["depth"].clone().to(dev)
    cam, hgt, pan = a["camera_matrix"].to(dev), a["real_camera_height"].to(dev), a["panoptic_seg"].to(dev)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        dgc_rescale(depth.clone(), cam, hgt, pan, 0, [10000])
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        points, scale, count = dgc_rescale(depth, cam, hgt, pan, 0, [10000])
    for scene in (b, a):
        depth.copy_(scene["depth"])
        pan.copy_(scene["panoptic_seg"])
        graph.replay()
        torch.cuda.synchronize()
        d2 = scene["depth"].clone().to(dev)
        p2, s2, c2 = dgc_rescale(d2, cam, hgt, scene["panoptic_seg"].to(dev), 0, [10000])
        assert torch.equal(scale, s2) and torch.equal(count, c2)
        assert torch.equal(depth, d2)
        assert torch.equal(torch.nan_to_num(points, nan=-1.0), torch.nan_to_num(p2, nan=-1.0))
