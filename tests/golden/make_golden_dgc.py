"""Generates tests/golden/dgc_*.npz by running the UNMODIFIED reference post-processing
(/root/reference/mgnet/postprocessing/depth_post_proc.py, imported where it lies) on seeded synthetic scenes.

Run in the build container only:   python tests/golden/make_golden_dgc.py

Shims (the reference file is executed unmodified): ``mgnet`` is a namespace stub (tests/golden/ref_loader.py), and --
only for the "auto ground mask" case -- ``torch.Tensor.cuda`` is the identity, because ``_get_ground_mask``
(depth_post_proc.py:175) hard-codes ``.cuda()`` for two constant tensors.  Each fixture stores the inputs and the
reference outputs: rescaled depth, scaled camera points, scale factor, and the intermediates obtained by calling the
reference's own helpers (surface normals, per-pixel camera heights, ground mask).
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_loader  # noqa: E402
from mgnet_b200.synthetic import make_dgc_inputs  # noqa: E402

CASES = {
    # name: (make_dgc_inputs kwargs, use panoptic ground mask, filter ids)
    "dgc_panoptic": (dict(H=48, W=160, seed=1), True, [10000]),
    "dgc_auto_mask": (dict(H=64, W=96, seed=2, scale_true=3.25), False, []),
    "dgc_ragged": (dict(H=37, W=75, seed=3, scale_true=11.0, cam_height=1.2), True, [10000, 13001]),
}


def load_post_proc():
    ref_loader.load()
    spec = importlib.util.spec_from_file_location(
        "mgnet.postprocessing.depth_post_proc",
        os.path.join(ref_loader.REF_ROOT, "mgnet", "postprocessing", "depth_post_proc.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def run_reference(m, geo, d, use_panoptic, filter_ids):
    pan = d["panoptic_seg"] if use_panoptic else None
    if not use_panoptic:
        torch.Tensor.cuda = lambda self, *a, **k: self      # depth_post_proc.py:175 hard-codes .cuda()
    depth = d["depth"].clone()
    out_depth, out_points = m.get_depth_prediction(depth, True, d["camera_matrix"], d["real_camera_height"], pan,
                                                   0 if use_panoptic else -1, filter_ids)
    P = geo.Camera(K=d["camera_matrix"]).reconstruct(d["depth"], frame="c")
    N = m._get_surface_normal(P)
    heights = (P * N).sum(1).abs()[0]
    ground = (pan == 0) if use_panoptic else m._get_ground_mask(P, N)[0, 0]
    scale = m._get_scale_recovery(P, d["real_camera_height"], ground_mask=(pan == 0) if use_panoptic else None)
    return {
        "ref_depth": out_depth.numpy().copy(), "ref_points": out_points.numpy().copy(),
        "ref_normals": N[0].numpy().copy(), "ref_heights": heights.numpy().copy(),
        "ref_ground": ground.numpy().astype(np.uint8), "ref_scale": scale.numpy().copy(),
    }


def main():
    geo, _, _ = ref_loader.load()
    m = load_post_proc()
    for name, (kw, use_pan, ids) in CASES.items():
        d = make_dgc_inputs(**kw)
        res = run_reference(m, geo, d, use_pan, ids)
        res.update({"in_depth": d["depth"].numpy(), "in_camera": d["camera_matrix"].numpy(),
                    "in_height": d["real_camera_height"].numpy(), "in_panoptic": d["panoptic_seg"].numpy(),
                    "use_panoptic": np.array(use_pan), "filter_ids": np.asarray(ids, np.int64)})
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **res)
        print(name, "scale", res["ref_scale"], "ground px", int(res["ref_ground"].sum()), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
