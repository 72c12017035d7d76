"""Drop-ins for the viewpoint epilogue of the canonical-pose heads: ``_get_rot_mat``
(``utils/general.py:191-226``) and the ``can_xyz_kps21 @ rot_mat`` (+ inference-branch scale, root and
projection) that follows it in ``network/Hand3DPoseNet.py:41-50`` / ``network/Hand3DPosePriorNetwork.py:38-40``,
backed by viewpoint.cu.  CUDA tensors only, no CPU path.
"""
from __future__ import annotations

import torch

from . import _cabi
from .mano_layer import _as_f32_cuda


def _ptr(t):
    return t.data_ptr() if t is not None else 0


class _ViewpointFunction(torch.autograd.Function):
    """(can, ux, uy, uz) -> (rot_mat, rel_normed); ``can`` may be None (rot_mat only)."""

    @staticmethod
    @_cabi.on_tensor_device
    def forward(ctx, can, ux, uy, uz):
        B = ux.shape[0]
        dev = ux.device
        rot = torch.empty((B, 3, 3), dtype=torch.float32, device=dev)
        rel = torch.empty_like(can) if can is not None else None
        _cabi.check(_cabi.lib().mb_viewpoint_forward(_ptr(can), ux.data_ptr(), uy.data_ptr(), uz.data_ptr(), 0, 0, 0, B,
                                                     rot.data_ptr(), _ptr(rel), 0, 0, _cabi.stream_handle(dev)),
                    "mb_viewpoint_forward")
        ctx.save_for_backward(can, ux, uy, uz)
        return rot, rel

    @staticmethod
    @_cabi.on_tensor_device
    def backward(ctx, g_rot, g_rel):
        can, ux, uy, uz = ctx.saved_tensors
        B = ux.shape[0]
        prep = lambda g: None if g is None else g.to(torch.float32).contiguous()
        g_rot, g_rel = prep(g_rot), prep(g_rel) if can is not None else None
        g_can = torch.empty_like(can) if can is not None and ctx.needs_input_grad[0] else None
        want_u = any(ctx.needs_input_grad[1:4])
        gu = [torch.empty_like(ux) for _ in range(3)] if want_u else [None] * 3
        _cabi.check(_cabi.lib().mb_viewpoint_backward(_ptr(can), ux.data_ptr(), uy.data_ptr(), uz.data_ptr(), _ptr(g_rot),
                                                      _ptr(g_rel), B, _ptr(g_can), _ptr(gu[0]), _ptr(gu[1]), _ptr(gu[2]),
                                                      _cabi.stream_handle(ux.device)), "mb_viewpoint_backward")
        return g_can, gu[0], gu[1], gu[2]


def _angles(ux_b, uy_b, uz_b):
    if not isinstance(ux_b, torch.Tensor) or ux_b.device.type != "cuda":
        raise _cabi.ManoB200Error("_get_rot_mat only runs on CUDA tensors (sm_100a); there is no CPU fallback")
    dev = ux_b.device
    u = [_as_f32_cuda(t, n, dev).reshape(-1) for t, n in ((ux_b, "ux_b"), (uy_b, "uy_b"), (uz_b, "uz_b"))]
    if not (u[0].shape == u[1].shape == u[2].shape):
        raise RuntimeError("expected ux_b, uy_b, uz_b of one shape [B,1]")
    return u, dev


def _get_rot_mat(ux_b, uy_b, uz_b):
    """utils/general.py:191-226: axis-angle (ux, uy, uz)[B,1] with the angle encoded in the norm ->
    rot_matrix[B,3,3]; theta = sqrt(|u|^2 + 1e-8)."""
    u, _ = _angles(ux_b, uy_b, uz_b)
    rot, _ = _ViewpointFunction.apply(None, *u)
    return rot


@_cabi.on_tensor_device
def viewpoint_transform(can_xyz_kps21, ux, uy, uz, index_root_bone_length=None, kp_coord_xyz_root=None,
                        camera_intrinsic_matrix=None):
    """network/Hand3DPoseNet.py:41-50 in one kernel.  Training branch (no scale / root given) ->
    ``(coord_xyz_rel_normed[B,21,3], rot_mat[B,3,3])`` with gradients to all four inputs; inference branch
    (``config.is_inference``: scale, root and intrinsics given) -> ``(joint_xyz21, uv21)``, no gradient."""
    u, dev = _angles(ux, uy, uz)
    B = u[0].shape[0]
    can = _as_f32_cuda(can_xyz_kps21, "can_xyz_kps21", dev)
    if can.numel() != B * 63:
        raise RuntimeError("expected can_xyz_kps21[B,21,3] (or [B,63])")
    can = can.reshape(B, 21, 3)                                                   # Hand3DPoseNet.py:37-38
    can = can.contiguous()
    if index_root_bone_length is None:
        rot, rel = _ViewpointFunction.apply(can, *u)
        return rel, rot
    scale = _as_f32_cuda(index_root_bone_length, "index_root_bone_length", dev)
    root = _as_f32_cuda(kp_coord_xyz_root, "kp_coord_xyz_root", dev)
    K = _as_f32_cuda(camera_intrinsic_matrix, "camera_intrinsic_matrix", dev)
    if scale.numel() != B or root.shape != (B, 3) or K.shape != (B, 3, 3):
        raise RuntimeError("expected index_root_bone_length[B,1], kp_coord_xyz_root[B,3], camera_intrinsic_matrix[B,3,3]")
    xyz = torch.empty_like(can)
    uv = torch.empty((B, 21, 2), dtype=torch.float32, device=dev)
    with torch.no_grad():
        _cabi.check(_cabi.lib().mb_viewpoint_forward(can.data_ptr(), u[0].data_ptr(), u[1].data_ptr(), u[2].data_ptr(),
                                                     scale.data_ptr(), root.data_ptr(), K.data_ptr(), B, 0, 0, xyz.data_ptr(),
                                                     uv.data_ptr(), _cabi.stream_handle(dev)), "mb_viewpoint_forward")
    return xyz, uv
