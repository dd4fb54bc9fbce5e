"""Mirror of the reference's ``mgnet.geometry`` namespace (mgnet/geometry/__init__.py:1-16)."""
from .camera import Camera
from .camera_utils import construct_K, scale_intrinsics, view_synthesis
from .depth import calc_smoothness, inv2depth
from .image import (
    gradient_x,
    gradient_y,
    image_grid,
    interpolate_image,
    match_scales,
    meshgrid,
    same_shape,
)
from .pose import Pose
from .pose_utils import euler2mat, invert_pose, pose_vec2mat

__all__ = [k for k in globals().keys() if not k.startswith("_")]
