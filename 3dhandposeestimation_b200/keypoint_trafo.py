"""Batched drop-ins for the keypoint re-parameterisations of the reference's dataloaders
(``utils/relative_trafo.py:167-270`` ``bone_rel_trafo`` / ``bone_rel_trafo_inv``,
``utils/canonical_trafo.py:93-184`` ``canonical_trafo`` / ``flip_right_hand``), backed by
hand_trafo.cu.  Forward only (the reference applies them to ground-truth keypoints, no gradient is
ever taken through them); CUDA tensors only, no CPU path.
"""
from __future__ import annotations

import torch

from . import _cabi
from .mano_layer import _as_f32_cuda


def _coords21(t, name):
    if not isinstance(t, torch.Tensor) or t.device.type != "cuda":
        raise _cabi.ManoB200Error(f"{name} must be a CUDA tensor (sm_100a); there is no CPU fallback")
    x = _as_f32_cuda(t.detach(), name, t.device)
    if x.dim() == 2:
        x = x.unsqueeze(0)
    x = x.reshape(-1, 21, 3)                                      # relative_trafo.py:176, canonical_trafo.py:121
    return x.contiguous()


@_cabi.on_tensor_device
def bone_rel_trafo(coords_xyz):
    """utils/relative_trafo.py:167-216: xyz[B,21,3] -> [B,21,3] = (bone length, angle_x, angle_y) in the
    frames of the kinematic chain."""
    xyz = _coords21(coords_xyz, "coords_xyz")
    out = torch.empty_like(xyz)
    _cabi.check(_cabi.lib().mb_bone_rel_trafo(xyz.data_ptr(), xyz.shape[0], out.data_ptr(), _cabi.stream_handle(xyz.device)),
                "mb_bone_rel_trafo")
    return out


@_cabi.on_tensor_device
def bone_rel_trafo_inv(coords_rel):
    """utils/relative_trafo.py:219-270: (length, angle_x, angle_y)[B,21,3] -> xyz[B,21,3]."""
    rel = _coords21(coords_rel, "coords_rel")
    out = torch.empty_like(rel)
    _cabi.check(_cabi.lib().mb_bone_rel_trafo_inv(rel.data_ptr(), rel.shape[0], out.data_ptr(), _cabi.stream_handle(rel.device)),
                "mb_bone_rel_trafo_inv")
    return out


def _cond_bytes(cond, shape, dev):
    c = torch.as_tensor(cond, device=dev)
    return c.to(torch.bool).expand(shape).contiguous().to(torch.uint8)


@_cabi.on_tensor_device
def canonical_trafo(coords_xyz, cond_right=None):
    """utils/canonical_trafo.py:93-159 -> ``(coords_xyz_normed[B,21,3], total_rot_mat[B,3,3])``.  With
    ``cond_right`` (bool, one per hand) the flagged hands are additionally mirrored as by
    ``flip_right_hand`` in the same pass (dataloader/thirdPartyTemplate/BinaryDbReaderRHD.py:250-252)."""
    xyz = _coords21(coords_xyz, "coords_xyz")
    B = xyz.shape[0]
    can = torch.empty_like(xyz)
    rot = torch.empty((B, 3, 3), dtype=torch.float32, device=xyz.device)
    cond = _cond_bytes(cond_right, (B,), xyz.device) if cond_right is not None else None
    _cabi.check(_cabi.lib().mb_canonical_trafo(xyz.data_ptr(), cond.data_ptr() if cond is not None else 0, B, can.data_ptr(),
                                               rot.data_ptr(), _cabi.stream_handle(xyz.device)), "mb_canonical_trafo")
    return can, rot


def flip_right_hand(coords_xyz_canonical, cond_right):
    """utils/canonical_trafo.py:163-184: z -> -z where ``cond_right``; [N,3] or [B,N,3] coordinates,
    ``cond_right`` broadcastable to [B,N] (the reference broadcasts ``cond_right.unsqueeze(-1)``)."""
    return _mirror(coords_xyz_canonical, cond_right, 2)


def mirror_left_hand(keypoint_xyz21, hand_side):
    """dataloader/RHD/dataloaderRHD.py:224-225, batched: x -> -x for the hands with ``hand_side == 0`` (left), so
    every sample reaches the right-hand MANO / FK layers as a right hand.  ``hand_side``: one integer per hand."""
    side = torch.as_tensor(hand_side, device=keypoint_xyz21.device if isinstance(keypoint_xyz21, torch.Tensor) else None)
    return _mirror(keypoint_xyz21, side == 0, 0)


@_cabi.on_tensor_device
def _mirror(coords_xyz_canonical, cond_right, axis):
    if not isinstance(coords_xyz_canonical, torch.Tensor) or coords_xyz_canonical.device.type != "cuda":
        raise _cabi.ManoB200Error("coords_xyz_canonical must be a CUDA tensor (sm_100a); there is no CPU fallback")
    dev = coords_xyz_canonical.device
    x = _as_f32_cuda(coords_xyz_canonical.detach(), "coords_xyz_canonical", dev)
    expanded = x.dim() == 2
    if expanded:
        x = x.unsqueeze(0)
    if x.dim() != 3 or x.shape[2] != 3:
        raise RuntimeError("expected coords_xyz_canonical[N,3] or [B,N,3]")
    B, N = x.shape[0], x.shape[1]
    c = torch.as_tensor(cond_right, device=dev)
    if expanded:
        c = c.unsqueeze(0)
    if c.dim() == 1:
        c = c.unsqueeze(-1) if c.shape[0] == B else c.unsqueeze(0)
    cond = _cond_bytes(c, (B, N), dev)
    out = torch.empty_like(x)
    _cabi.check(_cabi.lib().mb_mirror_hand(x.data_ptr(), cond.data_ptr(), B, N, 1, axis, out.data_ptr(), _cabi.stream_handle(dev)),
                "mb_mirror_hand")
    return out.squeeze(0) if expanded else out
