"""Launch-bound regime ("also CUDA-graph replay for C1/C2"): loss fwd+bwd through the public module, eager
launches vs one torch.cuda.CUDAGraph replay per step.  Usage: python scripts/time_graph.py [c1 c2]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mgnet_b200 import MultiViewPhotometricLoss
from mgnet_b200.synthetic import make_inputs
W = {"c1": (1, 192, 640), "c2": (16, 192, 640)}
HP = dict(ssim_loss_weight=0.85, photometric_loss_weight=1.0, smoothing_loss_weight=1e-3, automask_loss=True, photometric_reduce_op="min", padding_mode="zeros")
dev = torch.device("cuda:0")
out = {}
for name in (sys.argv[1:] or ["c1", "c2"]):
    B, H, Wd = W[name]
    pred, tgt = make_inputs(B, H, Wd, 3, seed=5)
    p = {"depth": [d.to(dev).requires_grad_(True) for d in pred["depth"]], "poses": pred["poses"].to(dev).requires_grad_(True)}
    t = {k: v.to(dev) for k, v in tgt.items()}
    mod = MultiViewPhotometricLoss(**HP)
    def step():
        for x in p["depth"] + [p["poses"]]:
            x.grad = None
        o = mod(p, t)
        (o["loss_photometric"] + o["loss_smoothness"]).backward()
        return o
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(5): step()
    torch.cuda.current_stream().wait_stream(side)
    for x in p["depth"] + [p["poses"]]: x.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        o = mod(p, t); (o["loss_photometric"] + o["loss_smoothness"]).backward()
    def timed(fn, reps=200):
        for _ in range(10): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    te, tg = timed(step), timed(graph.replay)
    px = B * H * Wd
    out[name] = {"eager_ms": te, "graph_ms": tg, "eager_gpx_s": px / te / 1e6, "graph_gpx_s": px / tg / 1e6}
    print(name, json.dumps(out[name]))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "graph_timing.json"), "w"), indent=1)
