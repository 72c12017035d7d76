"""B200-native (sm_100a) implementation of the geometric hot path of
hongrui16/3DHandPoseEstimation: the batched MANO layer, the RHD 21-joint
forward-kinematics layer, the pinhole projection and the visible-joint
MPJPE / L2 reductions — same ``nn.Module`` signatures as the reference's
``network/sub_modules``, arithmetic in hand-written CUDA behind a C ABI
(include/mano_b200.h).  No CPU fallback.

The directory name is not a Python identifier; import it with
``importlib.import_module("3dhandposeestimation_b200")`` or through the
``handpose_b200`` alias module at the repository root.
"""
from . import assets, dropin, fitting  # noqa: F401
from .dropin import install_into_reference  # noqa: F401
from ._cabi import ManoB200Error, lib as load_library  # noqa: F401
from .criterions import L2Loss, MPJPE, compute_hand_mask_loss, compute_regularization_loss  # noqa: F401
from .fk_layer import ForwardKinematics, ForwardKinematicsLoss, batch_project_xyz_to_uv, mano_joints_to_rhd_uv, match_mano_to_RHD  # noqa: F401
from .keypoint_trafo import bone_rel_trafo, bone_rel_trafo_inv, canonical_trafo, flip_right_hand, mirror_left_hand  # noqa: F401
from .mano_layer import ManoLayer  # noqa: F401
from .head_loss import ManoHeadLoss  # noqa: F401
from .viewpoint import _get_rot_mat, viewpoint_transform  # noqa: F401

__all__ = ["ManoLayer", "ManoHeadLoss", "ForwardKinematics", "ForwardKinematicsLoss", "batch_project_xyz_to_uv", "match_mano_to_RHD", "mano_joints_to_rhd_uv",
           "bone_rel_trafo", "bone_rel_trafo_inv", "canonical_trafo", "flip_right_hand", "mirror_left_hand", "_get_rot_mat", "viewpoint_transform", "MPJPE", "L2Loss",
           "compute_regularization_loss", "compute_hand_mask_loss", "install_into_reference", "ManoB200Error", "assets", "load_library"]
