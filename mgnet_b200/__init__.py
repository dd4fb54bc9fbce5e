"""mgnet_b200 -- B200-native (sm_100a) view-synthesis loss for MGNet-style self-supervised depth.

Only what the hot path needs: the CUDA kernels + C ABI (csrc/, include/mgvs.h), the autograd glue
(ops.py), the drop-in loss module (loss.py) and the mirror of the reference's mgnet.geometry API
(geometry/).  See DESIGN.md.
"""
from .loss import MultiViewPhotometricLoss  # noqa: F401
from .ops import LossConfig, view_synthesis_loss  # noqa: F401

__version__ = "0.1.0"
