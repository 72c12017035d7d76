"""ORACLE (test infrastructure only — never imported by the product): the MANO heads' tail composed from the piecewise
restatements: ManoLayer joints (MANOLayer.py:122-208) -> scale * p + transl (resnet50MANO.py:77-81) ->
[match_mano_to_RHD, Resnet50MANO3DHandPose.py:35-60] -> batch_project_xyz_to_uv (:73) -> L2Loss on xyz / uv
(criterions/loss.py:10-25, :83-87) + compute_regularization_loss (:113-117); weighted sum of the three terms and its
gradient w.r.t. (rot, pose, beta, transl, scale).  Pinned by tests/golden/head_loss_*.npz (the unmodified reference)."""
import numpy as np

from . import fk_oracle as fo
from . import mano_oracle as mo


def head_loss(model, rot, pose, beta, transl, scale, L, root, K, gt_xyz, gt_uv, vis, match, switched=True, dtype=np.float64):
    """-> dict(xyz, uv, losses[3], and a closure `grads(weights)` -> (g_rot, g_pose, g_beta, g_transl, g_scale))."""
    f = lambda a: None if a is None else np.asarray(a, dtype)
    rot, pose, beta, transl, scale, L, root, K, gt_xyz, gt_uv = map(f, (rot, pose, beta, transl, scale, L, root, K, gt_xyz, gt_uv))
    B = rot.shape[0]
    _, j0 = mo.mano_forward(model, rot, pose, beta, dtype=dtype)
    s = np.ones(B, dtype) if scale is None else scale.reshape(B)
    t = np.zeros((B, 3), dtype) if transl is None else transl
    j = s[:, None, None] * j0 + t[:, None, :]
    xyz = fo.match_mano_to_rhd(j, L, root, joint_order_switched=switched, dtype=dtype)[1] if match else j
    uv = fo.project_uv(xyz, K)
    losses = np.array([fo.l2loss(xyz, gt_xyz, vis, dtype), fo.l2loss(uv, gt_uv, vis, dtype), fo.regularizer(pose, beta, dtype)])

    def grads(weights):
        w = np.asarray(weights, dtype)
        g_xyz = w[0] * fo.l2loss_backward(xyz, gt_xyz, vis, dtype)
        g_uv = w[1] * fo.l2loss_backward(uv, gt_uv, vis, dtype)
        g_xyz = g_xyz + fo.project_uv_backward(xyz, K, g_uv)
        if match:
            g_j = fo.match_mano_to_rhd_backward(j, L, root, np.zeros_like(g_xyz), g_xyz, joint_order_switched=switched, dtype=dtype)[0]
        else:
            g_j = g_xyz
        g_transl = g_j.sum(axis=1)
        g_scale = np.sum(g_j * j0, axis=(1, 2))
        g_rot, g_pose, g_beta = mo.mano_backward(model, rot, pose, beta, None, s[:, None, None] * g_j, dtype=dtype)
        gt_, gb_ = fo.regularizer_backward(pose, beta, dtype)
        return g_rot, g_pose + w[2] * gt_, g_beta + w[2] * gb_, g_transl, g_scale

    return dict(xyz=xyz, uv=uv, losses=losses, grads=grads)
