"""The MANO heads' tail as one call: MANO parameters -> joints -> [scale / translation] -> [match_mano_to_RHD] ->
projection -> {L2 xyz, L2 uv, regulariser}.  Reference chain: ``network/sub_modules/resnet50MANO.py:76-87``,
``network/Resnet50MANO3DHandPose.py:35-60,71-73``, ``criterions/loss.py:10-25,83-87,113-117``, combined as in
``trainval.py:328-358``.  One ctypes call per direction (``mb_mano_head_loss_forward`` / ``_backward``) enqueues the
same kernels the separate drop-ins launch, without the ~10 autograd nodes, allocations and Python frames between
them — at the heads' batch size (``config.py:79``: 200) that overhead, not the GPU, is the step time.  No CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _cabi
from .fk_layer import _reference_joint_order_switched
from .mano_layer import ManoLayer, _as_f32_cuda


class _HeadLossFunction(torch.autograd.Function):
    """loss_xyz, loss_uv, loss_reg, joint_xyz21, uv21 = f(rot, coeffs, betas, transl, scale | constants); the last two outputs
    carry no gradient (they are what the head returns for logging / metrics; the loss terms are differentiated here).  The
    terms leave as three 0-dim outputs of the node (indexing one tensor afterwards would add three SelectBackward nodes)."""

    @staticmethod
    @_cabi.on_tensor_device
    def forward(ctx, rot, coeffs, betas, transl, scale, L, root, K, gt_xyz, gt_uv, vis, layer, flags, swap, alpha_beta):
        lib = _cabi.lib()
        B, nc, dev = rot.shape[0], coeffs.shape[1], rot.device
        xyz = torch.empty((B, 21, 3), dtype=torch.float32, device=dev)
        uv = torch.empty((B, 21, 2), dtype=torch.float32, device=dev)
        losses = torch.empty((3,), dtype=torch.float32, device=dev)
        ws = torch.empty((max(lib.mb_mano_head_loss_workspace_bytes(B), 16),), dtype=torch.uint8, device=dev)
        p = _cabi.ptr
        _cabi.check(lib.mb_mano_head_loss_forward(layer._blob.data_ptr(), nc, rot.data_ptr(), coeffs.data_ptr(), betas.data_ptr(),
                                                  p(transl), p(scale), p(L), p(root), K.data_ptr(), p(gt_xyz), p(gt_uv), p(vis),
                                                  B, layer._mode, flags, swap, alpha_beta, xyz.data_ptr(), uv.data_ptr(),
                                                  losses.data_ptr(), ws.data_ptr(), ws.numel(), _cabi.stream_handle(dev)),
                    "mb_mano_head_loss_forward")
        ctx.save_for_backward(rot, coeffs, betas, transl, scale, L, root, K, gt_xyz, gt_uv, vis, xyz, uv, ws)
        ctx.layer, ctx.flags, ctx.swap, ctx.alpha_beta = layer, flags, swap, alpha_beta
        ctx.mark_non_differentiable(xyz, uv)
        ctx.set_materialize_grads(False)
        return losses[0], losses[1], losses[2], xyz, uv

    @staticmethod
    @_cabi.on_tensor_device
    def backward(ctx, g_l0, g_l1, g_l2, _g_xyz, _g_uv):
        rot, coeffs, betas, transl, scale, L, root, K, gt_xyz, gt_uv, vis, xyz, uv, ws = ctx.saved_tensors
        lib = _cabi.lib()
        layer = ctx.layer
        B, nc, dev = rot.shape[0], coeffs.shape[1], rot.device
        gs = (g_l0, g_l1, g_l2)
        zero = ws.new_zeros((), dtype=torch.float32) if any(g is None for g in gs) else None
        g_losses = torch.stack([(g if g is not None else zero).to(torch.float32) for g in gs])
        g_rot, g_coeffs, g_betas = torch.empty_like(rot), torch.empty_like(coeffs), torch.empty_like(betas)
        g_transl = torch.empty_like(transl) if transl is not None and ctx.needs_input_grad[3] else None
        g_scale = torch.empty_like(scale) if scale is not None and ctx.needs_input_grad[4] else None
        p = _cabi.ptr
        _cabi.check(lib.mb_mano_head_loss_backward(layer._blob.data_ptr(), nc, rot.data_ptr(), coeffs.data_ptr(), betas.data_ptr(),
                                                   p(transl), p(scale), p(L), p(root), K.data_ptr(), p(gt_xyz), p(gt_uv), p(vis),
                                                   B, layer._mode, ctx.flags, ctx.swap, ctx.alpha_beta, xyz.data_ptr(), uv.data_ptr(),
                                                   g_losses.data_ptr(), g_rot.data_ptr(), g_coeffs.data_ptr(), g_betas.data_ptr(),
                                                   p(g_transl), p(g_scale), ws.data_ptr(), ws.numel(), _cabi.stream_handle(dev)),
                    "mb_mano_head_loss_backward")
        return (g_rot, g_coeffs, g_betas, g_transl, g_scale) + (None,) * 10


class ManoHeadLoss(nn.Module):
    """``ManoHeadLoss(mano_layer, comp_xyz_loss=True, comp_uv_loss=True, comp_regularization_loss=True,
    match_to_rhd=False)`` — the flags are ``LossCalculation``'s (criterions/loss.py:63); ``match_to_rhd`` inserts
    ``match_mano_to_RHD`` (Resnet50MANO3DHandPose.py:35-60) between the joints and the projection.

    ``forward(rot, pose, beta, camera_intrinsic_matrix, gt_xyz, gt_uv, keypoint_vis, *, index_root_bone_length=None,
    kp_coord_xyz_root=None, transl=None, scale=None)`` returns ``(loss_xyz, loss_uv, loss_regularization, joint_xyz21,
    uv21)``: the three terms as 0-dim tensors (``None`` for a term that is switched off, as ``LossCalculation.forward``
    does) and the head's outputs (no gradient flows through those two — differentiate the terms)."""

    def __init__(self, mano_layer: ManoLayer, comp_xyz_loss=True, comp_uv_loss=True, comp_regularization_loss=True,
                 match_to_rhd=False, joint_order_switched=None, alpha_beta=10.0):
        super().__init__()
        self.mano_layer = mano_layer
        self.comp_xyz_loss, self.comp_uv_loss = bool(comp_xyz_loss), bool(comp_uv_loss)
        self.comp_regularization_loss = bool(comp_regularization_loss)
        self.match_to_rhd = bool(match_to_rhd)
        self.joint_order_switched = joint_order_switched
        self.alpha_beta = float(alpha_beta)

    def forward(self, rot, pose, beta, camera_intrinsic_matrix, gt_xyz, gt_uv, keypoint_vis, *, index_root_bone_length=None,
                kp_coord_xyz_root=None, transl=None, scale=None):
        layer = self.mano_layer
        if not isinstance(rot, torch.Tensor) or rot.device.type != "cuda":
            raise _cabi.ManoB200Error("ManoHeadLoss only runs on CUDA tensors (sm_100a); there is no CPU fallback")
        dev = layer._require_device()
        rot = _as_f32_cuda(rot, "rot", dev)
        pose = _as_f32_cuda(pose, "pose", dev)
        beta = _as_f32_cuda(beta, "beta", dev)
        B = rot.shape[0]
        if rot.shape != (B, 3) or pose.shape != (B, layer.pose_num) or beta.shape != (B, 10):
            raise RuntimeError(f"expected rot[B,3], pose[B,{layer.pose_num}], beta[B,10]")
        K = _as_f32_cuda(camera_intrinsic_matrix, "camera_intrinsic_matrix", dev)
        if K.shape != (B, 3, 3):
            raise RuntimeError("expected camera_intrinsic_matrix[B,3,3]")
        flags = 0
        vis = None
        if self.comp_xyz_loss or self.comp_uv_loss:
            vis = _as_f32_cuda(keypoint_vis, "keypoint_vis", dev).reshape(B, 21)
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
        if self.comp_regularization_loss:
            flags |= _cabi.HEAD_REG
        L = root = None
        swap = 0
        if self.match_to_rhd:
            flags |= _cabi.HEAD_MATCH
            if index_root_bone_length is None or kp_coord_xyz_root is None:
                raise RuntimeError("match_to_rhd needs index_root_bone_length[B,1] and kp_coord_xyz_root[B,3]")
            L = _as_f32_cuda(index_root_bone_length, "index_root_bone_length", dev).reshape(B)
            root = _as_f32_cuda(kp_coord_xyz_root, "kp_coord_xyz_root", dev)
            if root.shape != (B, 3):
                raise RuntimeError("expected kp_coord_xyz_root[B,3]")
            switched = self.joint_order_switched
            if switched is None:
                switched = _reference_joint_order_switched()
            swap = int(not switched)
        if transl is not None:
            transl = _as_f32_cuda(transl, "transl", dev)
            if transl.shape != (B, 3):
                raise RuntimeError("expected transl[B,3]")
        if scale is not None:
            scale = _as_f32_cuda(scale, "scale", dev).reshape(B)
        lx, lu, lr, xyz, uv = _HeadLossFunction.apply(rot, pose, beta, transl, scale, L, root, K, gt_xyz, gt_uv, vis, layer, flags, swap,
                                                       self.alpha_beta)
        return (lx if self.comp_xyz_loss else None, lu if self.comp_uv_loss else None,
                lr if self.comp_regularization_loss else None, xyz, uv)
