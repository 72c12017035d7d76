"""ForwardKinematics and batch_project_xyz_to_uv — drop-ins for the reference's
``network/sub_modules/forwardKinematicsLayer.py:ForwardKinematics`` (:142-358) and
``utils/coordinate_trans.py:batch_project_xyz_to_uv`` (:29-73), backed by the sm_100a
kernels of libmano_b200 (fk.cu).  No CPU path.
"""
from __future__ import annotations

import sys

import torch
import torch.nn as nn

from . import _cabi
from .mano_layer import _as_f32_cuda


def _reference_joint_order_switched(default=True) -> bool:
    """The reference reads the mutable global ``config.joint_order_switched`` at call
    time (forwardKinematicsLayer.py:324; the RHD loader sets it, dataloaderRHD.py:528).
    When this layer is used as a drop-in inside the reference tree, honour the same global."""
    cfg = sys.modules.get("config.config")
    if cfg is not None and hasattr(cfg, "joint_order_switched"):
        return bool(cfg.joint_order_switched)
    return default


class _FKFunction(torch.autograd.Function):
    @staticmethod
    @_cabi.on_tensor_device
    def forward(ctx, root_angles, other_angles, bone_lengths, K, scale, root, swap):
        lib = _cabi.lib()
        B = root_angles.shape[0]
        dev = root_angles.device
        xyz = torch.empty((B, 21, 3), dtype=torch.float32, device=dev)
        uv = torch.empty((B, 21, 2), dtype=torch.float32, device=dev)
        _cabi.check(lib.mb_fk_forward(root_angles.data_ptr(), other_angles.data_ptr(), bone_lengths.data_ptr(),
                                      K.data_ptr(), scale.data_ptr(), root.data_ptr(), B, int(swap),
                                      xyz.data_ptr(), uv.data_ptr(), _cabi.stream_handle(dev)), "mb_fk_forward")
        ctx.save_for_backward(root_angles, other_angles, bone_lengths, K, scale, root)
        ctx.swap = int(swap)
        ctx.set_materialize_grads(False)
        return xyz, uv

    @staticmethod
    @_cabi.on_tensor_device
    def backward(ctx, g_xyz, g_uv):
        root_angles, other_angles, bone_lengths, K, scale, root = ctx.saved_tensors
        lib = _cabi.lib()
        B = root_angles.shape[0]
        dev = root_angles.device
        g_ra = torch.empty_like(root_angles)
        g_oa = torch.empty_like(other_angles)
        g_bl = torch.empty_like(bone_lengths)
        if g_xyz is not None:
            g_xyz = g_xyz.to(torch.float32).contiguous()
        if g_uv is not None:
            g_uv = g_uv.to(torch.float32).contiguous()
        _cabi.check(lib.mb_fk_backward(root_angles.data_ptr(), other_angles.data_ptr(), bone_lengths.data_ptr(),
                                       K.data_ptr(), scale.data_ptr(), root.data_ptr(), _cabi.ptr(g_xyz), _cabi.ptr(g_uv),
                                       B, ctx.swap, g_ra.data_ptr(), g_oa.data_ptr(), g_bl.data_ptr(),
                                       _cabi.stream_handle(dev)), "mb_fk_backward")
        # K, index_root_bone_length and kp_coord_xyz_root come from the dataset in the
        # reference (trainval.py:283-290) and never require grad.
        return g_ra, g_oa, g_bl, None, None, None, None


class ForwardKinematics(nn.Module):
    """``ForwardKinematics(device='cpu')`` as in forwardKinematicsLayer.py:143-145.  The
    ``device`` argument is kept for signature compatibility; the inputs' CUDA device is used.
    ``joint_order_switched=None`` follows the reference's ``config.joint_order_switched``
    global when that module is loaded, else True (config.py:68)."""

    def __init__(self, device="cpu", joint_order_switched=None):
        super().__init__()
        self.device = device
        self.joint_order_switched = joint_order_switched

    def forward(self, root_angles, other_angles, bone_lengths, camera_intrinsic_matrix, index_root_bone_length,
                kp_coord_xyz_root):
        assert isinstance(camera_intrinsic_matrix, torch.Tensor)          # forwardKinematicsLayer.py:206
        self.camera_intrinsic_matrix = camera_intrinsic_matrix
        if not isinstance(root_angles, torch.Tensor) or root_angles.device.type != "cuda":
            raise _cabi.ManoB200Error("ForwardKinematics only runs on CUDA tensors (sm_100a); there is no CPU fallback")
        dev = root_angles.device
        ra = _as_f32_cuda(root_angles, "root_angles", dev)
        oa = _as_f32_cuda(other_angles, "other_angles", dev)
        bl = _as_f32_cuda(bone_lengths, "bone_lengths", dev)
        K = _as_f32_cuda(camera_intrinsic_matrix, "camera_intrinsic_matrix", dev)
        sc = _as_f32_cuda(index_root_bone_length, "index_root_bone_length", dev)
        root = _as_f32_cuda(kp_coord_xyz_root, "kp_coord_xyz_root", dev)
        B = ra.shape[0]
        if (ra.shape != (B, 3) or oa.shape != (B, 23) or bl.shape != (B, 20) or K.shape != (B, 3, 3)
                or sc.numel() != B or root.shape != (B, 3)):
            raise RuntimeError("expected root_angles[B,3], other_angles[B,23], bone_lengths[B,20], K[B,3,3], "
                               "index_root_bone_length[B,1], kp_coord_xyz_root[B,3]")
        switched = self.joint_order_switched
        if switched is None:
            switched = _reference_joint_order_switched()
        xyz, uv = _FKFunction.apply(ra, oa, bl, K, sc.reshape(B), root, not switched)
        return [xyz, uv, None]                                            # forwardKinematicsLayer.py:330

    def convert_rel_normalized_to_absolute(self, kp_coord_xyz21_rel_normed, index_root_bone_length, kp_coord_xyz_root):
        """forwardKinematicsLayer.py:333-358 (plain elementwise helper kept for API parity;
        inside ``forward`` it is fused into the FK kernel)."""
        return kp_coord_xyz21_rel_normed * index_root_bone_length.unsqueeze(-1) + kp_coord_xyz_root.unsqueeze(1)


class _FKLossFunction(torch.autograd.Function):
    """loss_xyz, loss_uv, xyz, uv = f(root_angles, other_angles, bone_lengths | K, scale, root, gt_xyz, gt_uv, vis); xyz / uv
    carry no gradient (they are what the head returns; the two loss terms are differentiated here).  The terms leave as two
    0-dim outputs of the node (not as indexed views of one tensor: two SelectBackward nodes and their zero-filled gradient
    buffers cost more host time than the two kernels of this step)."""

    @staticmethod
    @_cabi.on_tensor_device
    def forward(ctx, root_angles, other_angles, bone_lengths, K, scale, root, gt_xyz, gt_uv, vis, swap, flags):
        lib = _cabi.lib()
        B = root_angles.shape[0]
        dev = root_angles.device
        xyz = torch.empty((B, 21, 3), dtype=torch.float32, device=dev)
        uv = torch.empty((B, 21, 2), dtype=torch.float32, device=dev)
        losses = torch.empty((2,), dtype=torch.float32, device=dev)
        ws = torch.empty((8,), dtype=torch.float64, device=dev)
        p = _cabi.ptr
        _cabi.check(lib.mb_fk_loss_forward(root_angles.data_ptr(), other_angles.data_ptr(), bone_lengths.data_ptr(), K.data_ptr(),
                                           scale.data_ptr(), root.data_ptr(), p(gt_xyz), p(gt_uv), p(vis), B, int(swap), flags,
                                           xyz.data_ptr(), uv.data_ptr(), losses.data_ptr(), ws.data_ptr(), 64,
                                           _cabi.stream_handle(dev)), "mb_fk_loss_forward")
        ctx.save_for_backward(root_angles, other_angles, bone_lengths, K, scale, root, gt_xyz, gt_uv, vis, ws)
        ctx.swap, ctx.flags = int(swap), flags
        ctx.mark_non_differentiable(xyz, uv)
        ctx.set_materialize_grads(False)
        return losses[0], losses[1], xyz, uv

    @staticmethod
    @_cabi.on_tensor_device
    def backward(ctx, g_lx, g_lu, _g_xyz, _g_uv):
        root_angles, other_angles, bone_lengths, K, scale, root, gt_xyz, gt_uv, vis, ws = ctx.saved_tensors
        lib = _cabi.lib()
        B = root_angles.shape[0]
        dev = root_angles.device
        zero = ws.new_zeros((), dtype=torch.float32) if g_lx is None or g_lu is None else None
        g_losses = torch.stack([(g if g is not None else zero).to(torch.float32) for g in (g_lx, g_lu)])
        g_ra, g_oa, g_bl = torch.empty_like(root_angles), torch.empty_like(other_angles), torch.empty_like(bone_lengths)
        p = _cabi.ptr
        _cabi.check(lib.mb_fk_loss_backward(root_angles.data_ptr(), other_angles.data_ptr(), bone_lengths.data_ptr(), K.data_ptr(),
                                            scale.data_ptr(), root.data_ptr(), p(gt_xyz), p(gt_uv), p(vis), B, ctx.swap, ctx.flags,
                                            g_losses.data_ptr(), g_ra.data_ptr(), g_oa.data_ptr(), g_bl.data_ptr(), ws.data_ptr(), 64,
                                            _cabi.stream_handle(dev)), "mb_fk_loss_backward")
        return (g_ra, g_oa, g_bl) + (None,) * 8


class ForwardKinematicsLoss(nn.Module):
    """The FK heads' training tail as one call per direction: ``ForwardKinematics.forward``
    (forwardKinematicsLayer.py:147-330) followed by ``LossCalculation``'s ``compute_3d_coord_loss`` /
    ``compute_uv_coord_loss`` (criterions/loss.py:83-87 -> ``L2Loss``, :10-25) on its outputs, the way
    network/TwoDimHandPoseWithFK.py + trainval.py:328-358 chain them.  ``comp_xyz_loss`` / ``comp_uv_loss`` are
    ``LossCalculation``'s flags (loss.py:63).

    ``forward(root_angles, other_angles, bone_lengths, camera_intrinsic_matrix, index_root_bone_length, kp_coord_xyz_root,
    gt_xyz, gt_uv, keypoint_vis)`` returns ``(loss_xyz, loss_uv, xyz, uv)`` — 0-dim loss tensors (``None`` for a term that is
    switched off, as ``LossCalculation.forward`` returns) and the layer's outputs, through which no gradient flows
    (differentiate the terms).  One kernel each way: the reductions run inside the FK kernel, the L2 gradients are formed
    inside the FK backward kernel.  At the reference's sizes the separate drop-ins (``ForwardKinematics`` + ``L2Loss`` x2)
    cost 3 autograd nodes and 7 launches per step and are bound by that overhead, not by the GPU."""

    def __init__(self, device="cpu", comp_xyz_loss=True, comp_uv_loss=True, joint_order_switched=None):
        super().__init__()
        self.device = device
        self.comp_xyz_loss, self.comp_uv_loss = bool(comp_xyz_loss), bool(comp_uv_loss)
        self.joint_order_switched = joint_order_switched

    def forward(self, root_angles, other_angles, bone_lengths, camera_intrinsic_matrix, index_root_bone_length,
                kp_coord_xyz_root, gt_xyz, gt_uv, keypoint_vis):
        if not isinstance(root_angles, torch.Tensor) or root_angles.device.type != "cuda":
            raise _cabi.ManoB200Error("ForwardKinematicsLoss only runs on CUDA tensors (sm_100a); there is no CPU fallback")
        dev = root_angles.device
        ra = _as_f32_cuda(root_angles, "root_angles", dev)
        oa = _as_f32_cuda(other_angles, "other_angles", dev)
        bl = _as_f32_cuda(bone_lengths, "bone_lengths", dev)
        K = _as_f32_cuda(camera_intrinsic_matrix, "camera_intrinsic_matrix", dev)
        sc = _as_f32_cuda(index_root_bone_length, "index_root_bone_length", dev)
        root = _as_f32_cuda(kp_coord_xyz_root, "kp_coord_xyz_root", dev)
        B = ra.shape[0]
        if (ra.shape != (B, 3) or oa.shape != (B, 23) or bl.shape != (B, 20) or K.shape != (B, 3, 3)
                or sc.numel() != B or root.shape != (B, 3)):
            raise RuntimeError("expected root_angles[B,3], other_angles[B,23], bone_lengths[B,20], K[B,3,3], "
                               "index_root_bone_length[B,1], kp_coord_xyz_root[B,3]")
        flags, vis = 0, None
        if self.comp_xyz_loss or self.comp_uv_loss:
            vis = _as_f32_cuda(keypoint_vis, "keypoint_vis", dev)
            if vis.numel() != B * 21:
                raise RuntimeError("expected keypoint_vis[B,21,1]")
            vis = vis.reshape(B, 21)
        if self.comp_xyz_loss:
            flags |= _cabi.HEAD_XYZ
            gt_xyz = _as_f32_cuda(gt_xyz, "gt_xyz", dev)
            if gt_xyz.shape != (B, 21, 3):
                raise RuntimeError("expected gt_xyz[B,21,3]")
        else:
            gt_xyz = None
        if self.comp_uv_loss:
            flags |= _cabi.HEAD_UV
            gt_uv = _as_f32_cuda(gt_uv, "gt_uv", dev)
            if gt_uv.shape != (B, 21, 2):
                raise RuntimeError("expected gt_uv[B,21,2]")
        else:
            gt_uv = None
        switched = self.joint_order_switched
        if switched is None:
            switched = _reference_joint_order_switched()
        lx, lu, xyz, uv = _FKLossFunction.apply(ra, oa, bl, K, sc.reshape(B), root, gt_xyz, gt_uv, vis, not switched, flags)
        return (lx if self.comp_xyz_loss else None, lu if self.comp_uv_loss else None, xyz, uv)


class _ProjectFunction(torch.autograd.Function):
    @staticmethod
    @_cabi.on_tensor_device
    def forward(ctx, xyz, K):
        lib = _cabi.lib()
        B, N = xyz.shape[0], xyz.shape[1]
        uv = torch.empty((B, N, 2), dtype=torch.float32, device=xyz.device)
        _cabi.check(lib.mb_project_uv_forward(xyz.data_ptr(), K.data_ptr(), B, N, uv.data_ptr(),
                                              _cabi.stream_handle(xyz.device)), "mb_project_uv_forward")
        ctx.save_for_backward(xyz, K)
        return uv

    @staticmethod
    @_cabi.on_tensor_device
    def backward(ctx, g_uv):
        xyz, K = ctx.saved_tensors
        lib = _cabi.lib()
        B, N = xyz.shape[0], xyz.shape[1]
        g_xyz = torch.empty_like(xyz)
        g_uv = g_uv.to(torch.float32).contiguous()
        _cabi.check(lib.mb_project_uv_backward(xyz.data_ptr(), K.data_ptr(), g_uv.data_ptr(), B, N, g_xyz.data_ptr(),
                                               _cabi.stream_handle(xyz.device)), "mb_project_uv_backward")
        return g_xyz, None


def batch_project_xyz_to_uv(positions_xyz, camera_intrinsic_matrix):
    """utils/coordinate_trans.py:29-73: uv[B,N,2] = (K xyz)_xy / (K xyz)_z with z==0 -> 1e-10."""
    if not isinstance(positions_xyz, torch.Tensor) or positions_xyz.device.type != "cuda":
        raise _cabi.ManoB200Error("batch_project_xyz_to_uv only runs on CUDA tensors (sm_100a); there is no CPU fallback")
    dev = positions_xyz.device
    xyz = _as_f32_cuda(positions_xyz, "positions_xyz", dev)
    K = _as_f32_cuda(camera_intrinsic_matrix, "camera_intrinsic_matrix", dev)
    if xyz.dim() != 3 or xyz.shape[2] != 3 or K.shape != (xyz.shape[0], 3, 3):
        raise RuntimeError("expected positions_xyz[B,N,3] and camera_intrinsic_matrix[B,3,3]")
    return _ProjectFunction.apply(xyz, K)


class _JointEpilogueFunction(torch.autograd.Function):
    """joints -> (rel_normalized, joint_xyz21[, uv21]) in one kernel (joint_epilogue.cu)."""

    @staticmethod
    @_cabi.on_tensor_device
    def forward(ctx, joints, scale, root, K, swap):
        lib = _cabi.lib()
        B = joints.shape[0]
        rel = torch.empty_like(joints)
        xyz = torch.empty_like(joints)
        uv = torch.empty((B, 21, 2), dtype=torch.float32, device=joints.device) if K is not None else None
        _cabi.check(lib.mb_joint_epilogue_forward(joints.data_ptr(), scale.data_ptr(), root.data_ptr(),
                                                  K.data_ptr() if K is not None else 0, B, int(swap), rel.data_ptr(),
                                                  xyz.data_ptr(), uv.data_ptr() if uv is not None else 0,
                                                  _cabi.stream_handle(joints.device)), "mb_joint_epilogue_forward")
        ctx.save_for_backward(joints, scale, root, K)
        ctx.swap = int(swap)
        return rel, xyz, uv                  # uv is None when no intrinsics were given

    @staticmethod
    @_cabi.on_tensor_device
    def backward(ctx, g_rel, g_xyz, g_uv):
        joints, scale, root, K = ctx.saved_tensors
        lib = _cabi.lib()
        B = joints.shape[0]
        prep = lambda g: None if g is None else g.to(torch.float32).contiguous()
        g_rel, g_xyz = prep(g_rel), prep(g_xyz)
        g_uv = prep(g_uv) if K is not None else None
        ptr = lambda t: t.data_ptr() if t is not None else 0
        g_joints = torch.empty_like(joints)
        g_scale = torch.empty_like(scale) if ctx.needs_input_grad[1] else None
        g_root = torch.empty_like(root) if ctx.needs_input_grad[2] else None
        _cabi.check(lib.mb_joint_epilogue_backward(joints.data_ptr(), scale.data_ptr(), root.data_ptr(), ptr(K), ptr(g_rel),
                                                   ptr(g_xyz), ptr(g_uv), B, ctx.swap, g_joints.data_ptr(), ptr(g_scale),
                                                   ptr(g_root), _cabi.stream_handle(joints.device)),
                    "mb_joint_epilogue_backward")
        return g_joints, g_scale, g_root, None, None


def _joint_epilogue(mano_joints, index_root_bone_length, kp_coord_xyz_root, K, joint_order_switched):
    if not isinstance(mano_joints, torch.Tensor) or mano_joints.device.type != "cuda":
        raise _cabi.ManoB200Error("match_mano_to_RHD only runs on CUDA tensors (sm_100a); there is no CPU fallback")
    dev = mano_joints.device
    joints = _as_f32_cuda(mano_joints, "mano_joints", dev)
    B = joints.shape[0]
    if joints.dim() != 3 or joints.shape[1:] != (21, 3):
        raise RuntimeError("expected mano_joints[B,21,3]")
    scale = _as_f32_cuda(index_root_bone_length, "index_root_bone_length", dev)
    root = _as_f32_cuda(kp_coord_xyz_root, "kp_coord_xyz_root", dev)
    if scale.numel() != B or root.shape != (B, 3):
        raise RuntimeError("expected index_root_bone_length[B,1] and kp_coord_xyz_root[B,3]")
    if K is not None:
        K = _as_f32_cuda(K, "camera_intrinsic_matrix", dev)
        if K.shape != (B, 3, 3):
            raise RuntimeError("expected camera_intrinsic_matrix[B,3,3]")
    if joint_order_switched is None:
        joint_order_switched = _reference_joint_order_switched()
    return _JointEpilogueFunction.apply(joints, scale, root, K, not joint_order_switched)


def match_mano_to_RHD(mano_joints, index_root_bone_length, kp_coord_xyz_root, joint_order_switched=None):
    """``match_mano_to_RHD`` of the MANO heads (network/Resnet50MANO3DHandPose.py:35-60,
    network/MANO3DHandPose.py:30-55) -> ``(mano_joints_rel_normalized, joint_xyz21)``.
    ``joint_order_switched`` defaults to the reference's global ``config.joint_order_switched``
    (True when the reference tree is not loaded).  Unlike the reference, ``mano_joints`` is not
    permuted in place."""
    rel, xyz, _ = _joint_epilogue(mano_joints, index_root_bone_length, kp_coord_xyz_root, None, joint_order_switched)
    return rel, xyz


def mano_joints_to_rhd_uv(mano_joints, index_root_bone_length, kp_coord_xyz_root, camera_intrinsic_matrix,
                          joint_order_switched=None):
    """``match_mano_to_RHD`` followed by ``batch_project_xyz_to_uv`` (Resnet50MANO3DHandPose.py:71-73)
    as one kernel -> ``(rel_normalized, joint_xyz21, uv21)``."""
    return _joint_epilogue(mano_joints, index_root_bone_length, kp_coord_xyz_root, camera_intrinsic_matrix,
                           joint_order_switched)
