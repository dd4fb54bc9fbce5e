"""A few DGC rescale calls at one size for ncu captures.  Usage: python scripts/one_dgc.py [H W reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mgnet_b200.postprocessing import dgc_rescale
from mgnet_b200.synthetic import make_dgc_inputs
H, W, reps = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1024, 2048, 3)
dev = torch.device("cuda:0")
d = make_dgc_inputs(H, W, seed=3, scale_true=6.0)
cam, hgt, pan, src = d["camera_matrix"].to(dev), d["real_camera_height"].to(dev), d["panoptic_seg"].to(dev), d["depth"].to(dev)
for _ in range(reps):
    work = src.clone()
    _, scale, _ = dgc_rescale(work, cam, hgt, pan, 0, [10000])
torch.cuda.synchronize()
print("scale", float(scale[0]))
