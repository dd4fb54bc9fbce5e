"""Plain-PyTorch stand-in for the parts of MGNet that surround the depth loss in a training step (BASELINE config[4]).

NOT part of the product and not a re-implementation of the reference's meta-architecture (out of scope): it
only has to put the same KIND and AMOUNT of cuDNN work and the same tensor interfaces around the fused loss that the reference
has -- a ResNet18 encoder (reference res_net.py:11-165: strides 4/8/16/32, 64..512 channels), a BiSeNet-like context/fusion
decoder at the common stride 8 with three output groups (semantic logits, instance centre + offsets, inverse depth at strides
8/16/32 with `sigmoid()/0.5`, mg_net.py:796-824), and PoseCNN (a second ResNet18 over the 9-channel frame triplet, four
convolutions, `mean(3).mean(2) * 0.01`, layers.py:130-167).  Random-init weights, synthetic targets.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _cbr(cin, cout, k=3, s=1):
    return nn.Sequential(nn.Conv2d(cin, cout, k, s, k // 2, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class BasicBlock(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.c1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.b1 = nn.BatchNorm2d(cout)
        self.c2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.b2 = nn.BatchNorm2d(cout)
        self.short = None
        if stride != 1 or cin != cout:
            self.short = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        y = F.relu(self.b1(self.c1(x)), inplace=True)
        y = self.b2(self.c2(y))
        return F.relu(y + (x if self.short is None else self.short(x)), inplace=True)


class ResNet18(nn.Module):
    """res2..res5 at strides 4/8/16/32 with 64/128/256/512 channels."""

    def __init__(self, in_channels=3):
        super().__init__()
        self.stem = nn.Sequential(nn.Conv2d(in_channels, 64, 7, 2, 3, bias=False), nn.BatchNorm2d(64), nn.ReLU(inplace=True),
                                  nn.MaxPool2d(3, 2, 1))
        chans, layers, cin = (64, 128, 256, 512), [], 64
        for i, c in enumerate(chans):
            layers.append(nn.Sequential(BasicBlock(cin, c, 1 if i == 0 else 2), BasicBlock(c, c, 1)))
            cin = c
        self.layers = nn.ModuleList(layers)

    def forward(self, x):
        x = self.stem(x)
        out = {}
        for name, layer in zip(("res2", "res3", "res4", "res5"), self.layers):
            x = layer(x)
            out[name] = x
        return out


class Decoder(nn.Module):
    """Context path over res4/res5 (attention-refined, like the reference's ARM), fused with res3 at stride 8."""

    def __init__(self, arm=(128, 128), ffm=128):
        super().__init__()
        self.arm5, self.att5 = _cbr(512, arm[0]), nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Conv2d(arm[0], arm[0], 1), nn.Sigmoid())
        self.arm4, self.att4 = _cbr(256, arm[1]), nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Conv2d(arm[1], arm[1], 1), nn.Sigmoid())
        self.ref5, self.ref4 = _cbr(arm[0], arm[0]), _cbr(arm[1], arm[1])
        self.ffm = _cbr(128 + arm[1], ffm, 1)

    def forward(self, f):
        a5 = self.arm5(f["res5"])
        a5 = self.ref5(a5 * self.att5(a5))                                             # stride 32
        a4 = self.arm4(f["res4"])
        a4 = a4 * self.att4(a4) + F.interpolate(a5, size=a4.shape[-2:], mode="nearest")
        a4 = self.ref4(a4)                                                             # stride 16
        y = self.ffm(torch.cat([f["res3"], F.interpolate(a4, size=f["res3"].shape[-2:], mode="nearest")], 1))   # stride 8
        return y, (a5, a4)


class Head(nn.Module):
    def __init__(self, cin, mid, cout):
        super().__init__()
        self.body, self.out = _cbr(cin, mid), nn.Conv2d(mid, cout, 1)

    def forward(self, x):
        return self.out(self.body(x))


class StandInMGNet(nn.Module):
    """forward(targets, upsample_depth) -> {"sem", "center", "offset", "depth": [3 maps], "poses": [B,2,6]}"""

    def __init__(self, num_classes=19):
        super().__init__()
        self.backbone = ResNet18(3)
        self.sem_dec, self.ins_dec, self.depth_dec = Decoder(), Decoder(), Decoder()
        self.sem_head = Head(128, 128, num_classes)
        self.center_head, self.offset_head = Head(128, 32, 1), Head(128, 32, 2)
        self.depth_heads = nn.ModuleList([Head(128, 32, 1), Head(128, 32, 1), Head(128, 32, 1)])     # ffm (1/8), arm 1/16, arm 1/32
        self.pose_encoder = ResNet18(9)
        self.pose_convs = nn.ModuleList([nn.Conv2d(512, 256, 1), nn.Conv2d(256, 256, 3, padding=1), nn.Conv2d(256, 256, 3, padding=1),
                                         nn.Conv2d(256, 12, 1)])

    def forward(self, t, upsample_depth=True):
        img = t["image"]
        f = self.backbone(img)
        ys, _ = self.sem_dec(f)
        yi, _ = self.ins_dec(f)
        yd, (a5, a4) = self.depth_dec(f)
        depth = []
        for head, feat, stride in zip(self.depth_heads, (yd, a4, a5), (8, 16, 32)):
            y = head(feat).float().sigmoid() / 0.5                                      # mg_net.py:823
            depth.append(F.interpolate(y, scale_factor=stride, mode="bilinear", align_corners=True) if upsample_depth else y)
        p = self.pose_encoder(torch.cat([img, t["image_prev"], t["image_next"]], 1))["res5"]
        for i, conv in enumerate(self.pose_convs):
            p = conv(p)
            if i < 3:
                p = F.relu(p, inplace=True)
        poses = 0.01 * p.float().mean(3).mean(2).view(-1, 2, 6)                         # layers.py:164-166
        return {"sem": self.sem_head(ys), "center": self.center_head(yi), "offset": self.offset_head(yi), "depth": depth, "poses": poses}


def panoptic_losses(out, t):
    """Semantic CE + centre MSE + offset L1 at stride 8 against synthetic targets (the panoptic heads' share of the backward)."""
    sem = F.cross_entropy(out["sem"].float(), t["sem_gt"], ignore_index=255)
    center = F.mse_loss(out["center"].float(), t["center_gt"])
    offset = F.l1_loss(out["offset"].float(), t["offset_gt"])
    return sem + 200.0 * center + 0.01 * offset


def synthetic_targets(B, H, W, seed, device):
    """Frame triplet + camera for the depth loss (mgnet_b200.synthetic) and stride-8 panoptic targets."""
    from mgnet_b200.synthetic import make_inputs
    _, tgt = make_inputs(B, H, W, 3, seed=seed, snap_trig=False)
    g = torch.Generator().manual_seed(seed)
    t = {k: v.to(device) for k, v in tgt.items()}
    t["image"], t["image_prev"], t["image_next"] = t["image_orig"], t["image_prev_orig"], t["image_next_orig"]
    t["sem_gt"] = torch.randint(0, 19, (B, H // 8, W // 8), generator=g).to(device)
    t["center_gt"] = torch.rand(B, 1, H // 8, W // 8, generator=g).to(device)
    t["offset_gt"] = (torch.randn(B, 2, H // 8, W // 8, generator=g) * 8).to(device)
    return t
