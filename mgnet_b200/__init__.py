"""mgnet_b200 -- B200-native (sm_100a) view-synthesis loss for MGNet-style self-supervised depth.

Only what the hot path and its two neighbouring rows need: the CUDA kernels + C ABI (csrc/, include/mgvs.h),
the autograd glue (ops.py), the drop-in loss module (loss.py), the mirror of the reference's mgnet.geometry API (geometry/), batch
sharding with the NCCL or the fused peer-memory exchange (sharding.py), the DGC depth rescaling of the inference path
(postprocessing.py: get_depth_prediction) and the uncertainty-weighting epilogue of the training step (uncertainty.py).
CUDA only: there is no CPU fallback anywhere.  See DESIGN.md and INTEGRATION.md.
"""
from .loss import MultiViewPhotometricLoss  # noqa: F401
from .ops import LossConfig, view_synthesis_loss  # noqa: F401

__version__ = "0.2.0"
