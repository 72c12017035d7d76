"""MPJPE / L2Loss / MANO regulariser — drop-ins for ``criterions/metrics.py:MPJPE`` (:6-27),
``criterions/loss.py:L2Loss`` (:6-25) and ``LossCalculation.compute_regularization_loss``
(:113-117).  The two masked reductions run in one fused sm_100a kernel each (reduce.cu):
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
    if pre.shape != gt.shape or pre.dim() != 3 or pre.shape[2] != 3:
        raise RuntimeError("expected pre_xyz and gt_xyz of shape [B, J, 3]")
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
    def forward(ctx, pre, gt, vis, vis_kind, n, kind):
        lib = _cabi.lib()
        dev = pre.device
        accum = torch.empty((2,), dtype=torch.float64, device=dev)
        out = torch.empty((), dtype=torch.float32, device=dev)
        _cabi.check(lib.mb_masked_joint_reduce(pre.data_ptr(), gt.data_ptr(), vis.data_ptr(), vis_kind, n, kind,
                                               accum.data_ptr(), out.data_ptr(), _cabi.stream_handle(dev)),
                    "mb_masked_joint_reduce")
        ctx.save_for_backward(pre, gt, vis, accum)
        ctx.vis_kind, ctx.n, ctx.kind = vis_kind, n, kind
        return out

    @staticmethod
    def backward(ctx, g_out):
        if ctx.kind != _cabi.REDUCE_L2:
            raise RuntimeError("MPJPE is a metric (used under no_grad, trainval.py:313-320); use L2Loss for training")
        pre, gt, vis, accum = ctx.saved_tensors
        lib = _cabi.lib()
        g_pre = torch.empty_like(pre)
        g_out = g_out.to(torch.float32).contiguous()
        _cabi.check(lib.mb_masked_l2_backward(pre.data_ptr(), gt.data_ptr(), vis.data_ptr(), ctx.vis_kind, ctx.n,
                                              accum.data_ptr(), g_out.data_ptr(), g_pre.data_ptr(),
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


def compute_regularization_loss(theta, beta):
    """criterions/loss.py:113-117: (||theta||_F + 10 ||beta||_F) / 100 over the whole batch."""
    alpha_beta = 10
    return (torch.norm(theta) + alpha_beta * torch.norm(beta)) / 100
