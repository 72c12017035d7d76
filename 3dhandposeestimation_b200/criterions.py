"""MPJPE / L2Loss / MANO regulariser / hand-mask loss — drop-ins for ``criterions/metrics.py:MPJPE`` (:6-27),
``criterions/loss.py:L2Loss`` (:6-25), ``LossCalculation.compute_regularization_loss`` (:113-117) and
``LossCalculation.compute_hand_mask_loss`` (:92-111).  The two masked reductions run in one fused sm_100a kernel each (reduce.cu):
no masked_select, no host sync on numel(); the all-invisible case returns 0 from the device.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _cabi
from .mano_layer import _as_f32_cuda


def _prep(pre_xyz, gt_xyz, keypoint_vis):
    if not isinstance(pre_xyz, torch.Tensor) or pre_xyz.device.type != "cuda":
        raise _cabi.ManoB200Error("the masked joint reductions only run on CUDA tensors (sm_100a); no CPU fallback")
    dev = pre_xyz.device
    pre = _as_f32_cuda(pre_xyz, "pre_xyz", dev)
    gt = _as_f32_cuda(gt_xyz, "gt_xyz", dev)
    if pre.shape != gt.shape or pre.dim() != 3 or not 1 <= pre.shape[2] <= 4:
        raise RuntimeError("expected pre_xyz and gt_xyz of shape [B, J, D], D = 3 (xyz) or 2 (uv, loss.py:86-87)")
    n = pre.shape[0] * pre.shape[1]
    vis = keypoint_vis
    if vis.device != dev:
        raise RuntimeError("keypoint_vis is on a different device")
    if vis.numel() != n:
        raise RuntimeError("keypoint_vis must have B*J elements ([B, J, 1])")
    if vis.dtype in (torch.uint8, torch.bool):
        vis = vis.contiguous().view(torch.uint8)
        kind = _cabi.VIS_U8
    else:
        vis = vis.to(torch.float32).contiguous()
        kind = _cabi.VIS_F32
    return pre, gt, vis, kind, n


class _MaskedReduce(torch.autograd.Function):
    @staticmethod
    @_cabi.on_tensor_device
    def forward(ctx, pre, gt, vis, vis_kind, n, kind):
        lib = _cabi.lib()
        dev = pre.device
        accum = torch.empty((2,), dtype=torch.float64, device=dev)
        out = torch.empty((), dtype=torch.float32, device=dev)
        _cabi.check(lib.mb_masked_joint_reduce(pre.data_ptr(), gt.data_ptr(), vis.data_ptr(), vis_kind, n, pre.shape[2], kind,
                                               accum.data_ptr(), out.data_ptr(), _cabi.stream_handle(dev)),
                    "mb_masked_joint_reduce")
        ctx.save_for_backward(pre, gt, vis, accum)
        ctx.vis_kind, ctx.n, ctx.kind = vis_kind, n, kind
        return out

    @staticmethod
    @_cabi.on_tensor_device
    def backward(ctx, g_out):
        if ctx.kind != _cabi.REDUCE_L2:
            raise RuntimeError("MPJPE is a metric (used under no_grad, trainval.py:313-320); use L2Loss for training")
        pre, gt, vis, accum = ctx.saved_tensors
        lib = _cabi.lib()
        g_pre = torch.empty_like(pre)
        g_out = g_out.to(torch.float32).contiguous()
        _cabi.check(lib.mb_masked_l2_backward(pre.data_ptr(), gt.data_ptr(), vis.data_ptr(), ctx.vis_kind, ctx.n,
                                              pre.shape[2], accum.data_ptr(), g_out.data_ptr(), g_pre.data_ptr(),
                                              _cabi.stream_handle(pre.device)), "mb_masked_l2_backward")
        return g_pre, None, None, None, None, None


class MPJPE(nn.Module):
    """criterions/metrics.py:6-27: mean over visible joints of ||pre - gt|| in mm (x1000)."""

    def forward(self, pre_xyz, gt_xyz, keypoint_vis):
        pre, gt, vis, kind, n = _prep(pre_xyz, gt_xyz, keypoint_vis)
        with torch.no_grad():
            return _MaskedReduce.apply(pre, gt, vis, kind, n, _cabi.REDUCE_MPJPE_MM)


class L2Loss(nn.Module):
    """criterions/loss.py:6-25: mean over visible joints of ||pre - gt||^2 (differentiable in pre_xyz)."""

    def forward(self, pre_xyz, gt_xyz, keypoint_vis):
        pre, gt, vis, kind, n = _prep(pre_xyz, gt_xyz, keypoint_vis)
        return _MaskedReduce.apply(pre, gt, vis, kind, n, _cabi.REDUCE_L2)


class _Regulariser(torch.autograd.Function):
    @staticmethod
    @_cabi.on_tensor_device
    def forward(ctx, theta, beta, alpha_beta):
        lib = _cabi.lib()
        dev = theta.device
        accum = torch.empty((2,), dtype=torch.float64, device=dev)
        out = torch.empty((), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _cabi.check(lib.mb_regulariser_forward(theta.data_ptr(), theta.numel(), beta.data_ptr(), beta.numel(), alpha_beta,
                                                   accum.data_ptr(), out.data_ptr(), _cabi.stream_handle(dev)),
                        "mb_regulariser_forward")
        ctx.save_for_backward(theta, beta, accum)
        ctx.alpha_beta = alpha_beta
        return out

    @staticmethod
    @_cabi.on_tensor_device
    def backward(ctx, g_out):
        theta, beta, accum = ctx.saved_tensors
        dev = theta.device
        g_theta = torch.empty_like(theta)
        g_beta = torch.empty_like(beta)
        g_out = g_out.to(torch.float32).contiguous()
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().mb_regulariser_backward(theta.data_ptr(), theta.numel(), beta.data_ptr(), beta.numel(),
                                                            ctx.alpha_beta, accum.data_ptr(), g_out.data_ptr(),
                                                            g_theta.data_ptr(), g_beta.data_ptr(), _cabi.stream_handle(dev)),
                        "mb_regulariser_backward")
        return g_theta, g_beta, None


def compute_regularization_loss(theta, beta, alpha_beta=10.0):
    """criterions/loss.py:113-117 — the MANO regulariser over the whole batch, as two kernels (a fp64 sum-of-squares
    reduction + finalise; an elementwise backward), differentiable in both arguments.  ``alpha_beta`` is the
    reference's hard-coded weight of the shape term."""
    if not isinstance(theta, torch.Tensor) or theta.device.type != "cuda":
        raise _cabi.ManoB200Error("compute_regularization_loss only runs on CUDA tensors (sm_100a); no CPU fallback")
    dev = theta.device
    return _Regulariser.apply(_as_f32_cuda(theta, "theta", dev), _as_f32_cuda(beta, "beta", dev), float(alpha_beta))


@_cabi.on_tensor_device
def compute_hand_mask_loss(pred_uv, gt_uv, hand_mask):
    """criterions/loss.py:92-111: 1 - (mask samples at the predicted keypoints) / (mask samples at the ground-truth
    keypoints + 1e-8); uv[B,N,2] are truncated to integers and clamped to [0, W-1], ``hand_mask`` is [B,H,W].
    One gather-reduce kernel, no host sync; the result carries no gradient (neither does the reference's)."""
    if not isinstance(pred_uv, torch.Tensor) or pred_uv.device.type != "cuda":
        raise _cabi.ManoB200Error("compute_hand_mask_loss only runs on CUDA tensors (sm_100a); no CPU fallback")
    dev = pred_uv.device
    pred = _as_f32_cuda(pred_uv.detach(), "pred_uv", dev)
    gt = _as_f32_cuda(gt_uv.detach(), "gt_uv", dev)
    if pred.shape != gt.shape or pred.dim() != 3 or pred.shape[2] != 2:
        raise RuntimeError("expected pred_uv and gt_uv of shape [B, N, 2]")
    if hand_mask.device != dev or hand_mask.dim() != 3 or hand_mask.shape[0] != pred.shape[0]:
        raise RuntimeError("expected hand_mask[B, H, W] on the device of pred_uv")
    B, N = pred.shape[0], pred.shape[1]
    H, W = hand_mask.shape[1], hand_mask.shape[2]
    if W > H:
        raise IndexError("hand_mask rows are clamped with the last dimension (loss.py:94-95): needs W <= H")
    if hand_mask.dtype in (torch.uint8, torch.bool):
        mask, kind = hand_mask.contiguous().view(torch.uint8), _cabi.VIS_U8
    else:
        mask, kind = hand_mask.to(torch.float32).contiguous(), _cabi.VIS_F32
    accum = torch.empty((2,), dtype=torch.float64, device=dev)
    out = torch.empty((), dtype=torch.float32, device=dev)
    _cabi.check(_cabi.lib().mb_hand_mask_loss(pred.data_ptr(), gt.data_ptr(), mask.data_ptr(), kind, B, N, H, W,
                                              accum.data_ptr(), out.data_ptr(), _cabi.stream_handle(dev)), "mb_hand_mask_loss")
    return out
