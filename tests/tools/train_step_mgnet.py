"""BASELINE config[4]: an end-to-end training step around the fused depth loss -- ResNet18 encoder + three decoder/head groups +
PoseCNN stand-in (tests/tools/standin_mgnet.py, plain PyTorch / cuDNN, random-init weights), panoptic losses + the depth loss,
backward, SGD step -- with the depth loss computed three ways from identical weights and inputs:

  eager           the reference's ATen operator sequence on the GPU (oracle/torch_port.py; per-rank mask normalisation as in the
                  reference under DDP)
  fused           mgnet_b200.MultiViewPhotometricLoss (batch-sharded, global normalisation through the fused peer exchange,
                  ddp_grad_scale=True)
  fused_upsample  the same with fuse_upsample=True: the head hands over its low-resolution maps

Single GPU:   python tests/tools/train_step_mgnet.py --crop 512x1024 --batch 4
DDP:          python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 \
                  tests/tools/train_step_mgnet.py --crop 1024x1024 --batch 2
Prints one JSON line per (crop, amp) on rank 0 (TRAIN_STEP ...): step ms of the three variants (device-timed, max over ranks),
the depth loss's own fwd+bwd time on the head's outputs, and -- single GPU, fp32 -- the agreement of the parameter gradients
fused vs eager.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP

from standin_mgnet import StandInMGNet, panoptic_losses, synthetic_targets

HP = dict(ssim_loss_weight=0.85, photometric_loss_weight=1.0, smoothing_loss_weight=1e-3, automask_loss=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--crop", default="512x1024")
    ap.add_argument("--batch", type=int, default=4, help="images per GPU")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--amp", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--variants", default="eager,fused,fused_upsample")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    H, W = (int(x) for x in args.crop.split("x"))
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", lr)
    torch.cuda.set_device(dev)
    group = exchange = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    from mgnet_b200 import MultiViewPhotometricLoss
    from oracle.torch_port import reference_loss
    if world > 1:
        from mgnet_b200.sharding import PeerExchange
        try:
            exchange = PeerExchange(group)
        except Exception:
            exchange = None
        ok = torch.tensor([1 if exchange is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        exchange = exchange if int(ok.item()) else None
    torch.backends.cudnn.benchmark = True
    t = synthetic_targets(args.batch, H, W, 200 + rank, dev)

    def depth_loss_fn(name):
        if name == "eager":
            return lambda pred, tt: reference_loss(pred, tt, **HP)
        return MultiViewPhotometricLoss(photometric_reduce_op="min", padding_mode="zeros", fuse_upsample=(name == "fused_upsample"),
                                        process_group=group, exchange=exchange, ddp_grad_scale=world > 1, **HP)

    def build():
        torch.manual_seed(0)
        return StandInMGNet().to(dev).to(memory_format=torch.channels_last)

    def step(model, opt, loss_fn, upsample, amp):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            out = model(t, upsample)
            pan = panoptic_losses(out, t)
        dl = loss_fn({"depth": out["depth"], "poses": out["poses"]}, t)          # fp32 (custom_fwd(cast_inputs=float32), mg_net.py:827)
        total = pan + dl["loss_photometric"] + dl["loss_smoothness"]
        total.backward()
        opt.step()
        return total

    res = {}
    amp = args.amp == "bf16"
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name in args.variants.split(","):
        net = build()
        model = DDP(net, device_ids=[lr]) if world > 1 else net
        opt = torch.optim.SGD(model.parameters(), lr=1e-4, momentum=0.9)
        fn = depth_loss_fn(name)
        up = name != "fused_upsample"
        for _ in range(3):
            step(model, opt, fn, up, amp)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0.record()
        for _ in range(args.iters):
            total = step(model, opt, fn, up, amp)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / args.iters], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        res[name] = {"step_ms": float(ms.item()), "images_per_s": args.batch * world / (float(ms.item()) * 1e-3), "loss": float(total.item())}
        # the depth loss's own fwd+bwd on the (detached) head outputs of this model
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            out = model(t, up)
        d = [x.detach().float().requires_grad_(True) for x in out["depth"]]
        p = out["poses"].detach().float().requires_grad_(True)
        for _ in range(2):
            o = fn({"depth": d, "poses": p}, t)
            (o["loss_photometric"] + o["loss_smoothness"]).backward()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0.record()
        for _ in range(args.iters):
            o = fn({"depth": d, "poses": p}, t)
            (o["loss_photometric"] + o["loss_smoothness"]).backward()
        e1.record()
        torch.cuda.synchronize()
        lm = torch.tensor([e0.elapsed_time(e1) / args.iters], device=dev)
        if world > 1:
            dist.all_reduce(lm, op=dist.ReduceOp.MAX)
        res[name]["depth_loss_fwd_bwd_ms"] = float(lm.item())
        del model, opt, net
        torch.cuda.empty_cache()

    agree = None
    if world == 1 and "eager" in res and "fused" in res:
        # parameter-gradient agreement, fp32 network, TF32 off, same weights: only the depth loss implementation differs
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.benchmark = False
        grads = {}
        for name in ("eager", "fused", "fused_upsample"):
            net = build()
            fn = depth_loss_fn(name)
            out = net(t, name != "fused_upsample")
            dl = fn({"depth": out["depth"], "poses": out["poses"]}, t)
            (dl["loss_photometric"] + dl["loss_smoothness"]).backward()          # the depth loss alone: its gradient is what is compared
            grads[name] = ([q.grad.detach().double().clone() if q.grad is not None else None for q in net.parameters()],
                           float(dl["loss_photometric"].item()), float(dl["loss_smoothness"].item()))
            del net

        def rel(a, b):
            num = sum(float((x - y).pow(2).sum()) for x, y in zip(a, b) if x is not None and y is not None)
            den = sum(float(y.pow(2).sum()) for y in b if y is not None)
            return (num / max(den, 1e-300)) ** 0.5

        agree = {"param_grad_l2rel_fused_vs_eager": rel(grads["fused"][0], grads["eager"][0]),
                 "param_grad_l2rel_fused_upsample_vs_fused": rel(grads["fused_upsample"][0], grads["fused"][0]),
                 "loss_photometric_rel": abs(grads["fused"][1] - grads["eager"][1]) / abs(grads["eager"][1]),
                 "loss_smoothness_rel": abs(grads["fused"][2] - grads["eager"][2]) / abs(grads["eager"][2])}
    if rank == 0:
        line = {"workload": "training step: ResNet18 encoder + 3 decoder/head groups + PoseCNN stand-in, panoptic + depth losses, SGD",
                "gpus": world, "B_per_gpu": args.batch, "H": H, "W": W, "amp": args.amp,
                "exchange": "peer" if exchange is not None else ("nccl" if world > 1 else None), "variants": res, "agreement_fp32": agree}
        print("TRAIN_STEP " + json.dumps(line))
        if args.out:
            with open(args.out, "a") as f:
                f.write(json.dumps(line) + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
