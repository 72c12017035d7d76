"""ORACLE — test infrastructure only (see mano_oracle.py header for the rules
and for how parity is pinned: against outputs of the reference itself, the
reference having no golden vectors).

numpy restatement of the reference's RHD 21-joint forward-kinematics layer,
pinhole projection, and the masked joint reductions, with analytic backward.

Restated reference code:
  network/sub_modules/forwardKinematicsLayer.py:59-96   Rx*Ry*Rz
  network/sub_modules/forwardKinematicsLayer.py:147-330 ForwardKinematics.forward
  network/sub_modules/forwardKinematicsLayer.py:333-358 convert_rel_normalized_to_absolute
  utils/coordinate_trans.py:29-73                       batch_project_xyz_to_uv
  criterions/metrics.py:10-27                           MPJPE
  criterions/loss.py:10-25, 113-117                     L2Loss, MANO regulariser
"""
from __future__ import annotations

import numpy as np


def dof_map():
    """Angle sources of the 20 nodes A1..E4 (forwardKinematicsLayer.py:239-274).
    Returns a list of 20 triples; entry a is an index into other_angles[23] or
    -1 for a fixed zero angle.  Tips (*4) have identity local rotation."""
    m = []
    m.append((0, 1, 2))      # A1  x y z
    m.append((3, 4, 5))      # A2  x y z
    m.append((-1, 6, -1))    # A3  y only
    m.append((-1, -1, -1))   # A4  tip
    idx = 7
    for _ in range(4):       # B, C, D, E
        m.append((idx, idx + 1, -1))   # *1  x y
        m.append((idx + 2, -1, -1))    # *2  x
        m.append((idx + 3, -1, -1))    # *3  x
        m.append((-1, -1, -1))         # *4  tip
        idx += 4
    assert idx == 23
    return m


def euler_xyz(ang):
    """forwardKinematicsLayer.py:59-96: R = Rx(x) @ Ry(y) @ Rz(z).  ang [...,3]."""
    x, y, z = ang[..., 0], ang[..., 1], ang[..., 2]
    cx, sx, cy, sy, cz, sz = np.cos(x), np.sin(x), np.cos(y), np.sin(y), np.cos(z), np.sin(z)
    o, l = np.zeros_like(x), np.ones_like(x)
    Rx = np.stack([np.stack([l, o, o], -1), np.stack([o, cx, -sx], -1), np.stack([o, sx, cx], -1)], -2)
    Ry = np.stack([np.stack([cy, o, sy], -1), np.stack([o, l, o], -1), np.stack([-sy, o, cy], -1)], -2)
    Rz = np.stack([np.stack([cz, -sz, o], -1), np.stack([sz, cz, o], -1), np.stack([o, o, l], -1)], -2)
    return Rx @ Ry @ Rz, (Rx, Ry, Rz)


def euler_xyz_backward(ang, dR):
    """d<dR, Rx Ry Rz>/d(x,y,z)."""
    x, y, z = ang[..., 0], ang[..., 1], ang[..., 2]
    cx, sx, cy, sy, cz, sz = np.cos(x), np.sin(x), np.cos(y), np.sin(y), np.cos(z), np.sin(z)
    o = np.zeros_like(x)
    _, (Rx, Ry, Rz) = euler_xyz(ang)
    dRx = np.stack([np.stack([o, o, o], -1), np.stack([o, -sx, -cx], -1), np.stack([o, cx, -sx], -1)], -2)
    dRy = np.stack([np.stack([-sy, o, cy], -1), np.stack([o, o, o], -1), np.stack([-cy, o, -sy], -1)], -2)
    dRz = np.stack([np.stack([-sz, -cz, o], -1), np.stack([cz, -sz, o], -1), np.stack([o, o, o], -1)], -2)
    gx = np.sum(dR * (dRx @ Ry @ Rz), axis=(-1, -2))
    gy = np.sum(dR * (Rx @ dRy @ Rz), axis=(-1, -2))
    gz = np.sum(dR * (Rx @ Ry @ dRz), axis=(-1, -2))
    return np.stack([gx, gy, gz], -1)


def project_uv(xyz, K):
    """utils/coordinate_trans.py:48-65.  p = K xyz^T ; p_z == 0 -> 1e-10 ; uv = p_xy / p_z."""
    p = np.einsum("bij,bnj->bni", K, xyz)
    pz = np.where(p[..., 2] == 0, np.asarray(1e-10, dtype=p.dtype), p[..., 2])
    return p[..., :2] / pz[..., None]


def project_uv_backward(xyz, K, g_uv):
    """Gradient of ``project_uv`` w.r.t. xyz.  On the p_z==0 branch the divisor
    is a constant, so only the numerator carries gradient (autograd of the
    in-place ``where`` at coordinate_trans.py:59 behaves the same way)."""
    p = np.einsum("bij,bnj->bni", K, xyz)
    zero = p[..., 2] == 0
    pz = np.where(zero, np.asarray(1e-10, dtype=p.dtype), p[..., 2])
    uv = p[..., :2] / pz[..., None]
    dp = np.zeros_like(p)
    dp[..., :2] = g_uv / pz[..., None]
    dp[..., 2] = np.where(zero, 0.0, -np.sum(g_uv * uv, axis=-1) / pz)
    return np.einsum("bij,bni->bnj", K, dp)      # K^T dp


def swap_joint_order(x):
    """forwardKinematicsLayer.py:324-327 (applied when config.joint_order_switched is False)."""
    x = x.copy()
    for i in range(1, 21, 4):
        x[:, [i, i + 3]] = x[:, [i + 3, i]]
        x[:, [i + 1, i + 2]] = x[:, [i + 2, i + 1]]
    return x


def fk_forward(root_angles, other_angles, bone_lengths, K, index_root_bone_length, kp_coord_xyz_root,
               joint_order_switched=True, dtype=np.float64, return_cache=False):
    """``ForwardKinematics.forward`` (forwardKinematicsLayer.py:147-330) -> (xyz[B,21,3], uv[B,21,2])."""
    ra = np.asarray(root_angles, dtype=dtype)
    oa = np.asarray(other_angles, dtype=dtype)
    bl = np.asarray(bone_lengths, dtype=dtype)
    K = np.asarray(K, dtype=dtype)
    s = np.asarray(index_root_bone_length, dtype=dtype).reshape(-1, 1, 1)
    root = np.asarray(kp_coord_xyz_root, dtype=dtype).reshape(-1, 1, 3)
    B = ra.shape[0]
    dmap = dof_map()
    Rg = np.zeros((B, 21, 3, 3), dtype=dtype)
    P = np.zeros((B, 21, 3), dtype=dtype)
    Rl = np.zeros((B, 20, 3, 3), dtype=dtype)
    ang = np.zeros((B, 20, 3), dtype=dtype)
    Rg[:, 0] = euler_xyz(ra)[0]                          # :214-215
    for i in range(20):
        par = 0 if i % 4 == 0 else i                     # :225-230
        for a in range(3):
            if dmap[i][a] >= 0:
                ang[:, i, a] = oa[:, dmap[i][a]]
        if i % 4 == 3:
            Rl[:, i] = np.eye(3, dtype=dtype)            # :254, :274
        else:
            Rl[:, i] = euler_xyz(ang[:, i])[0]
        Rg[:, i + 1] = Rg[:, par] @ Rl[:, i]             # :286
        P[:, i + 1] = P[:, par] + bl[:, i, None] * Rg[:, i + 1, :, 2]   # :290-308
    xyz = P * s + root                                   # :321, :333-358
    if not joint_order_switched:
        xyz = swap_joint_order(xyz)                      # :324-327
    uv = project_uv(xyz, K)                              # :329
    if return_cache:
        return xyz, uv, dict(Rg=Rg, P=P, Rl=Rl, ang=ang, dmap=dmap, s=s, bl=bl, ra=ra)
    return xyz, uv


def fk_backward(root_angles, other_angles, bone_lengths, K, index_root_bone_length, kp_coord_xyz_root,
                g_xyz, g_uv, joint_order_switched=True, dtype=np.float64):
    """Analytic gradient w.r.t. (root_angles, other_angles, bone_lengths) — the
    three tensors that require grad in the heads."""
    xyz, uv, k = fk_forward(root_angles, other_angles, bone_lengths, K, index_root_bone_length,
                            kp_coord_xyz_root, joint_order_switched, dtype, return_cache=True)
    K = np.asarray(K, dtype=dtype)
    B = xyz.shape[0]
    dx = np.zeros_like(xyz) if g_xyz is None else np.asarray(g_xyz, dtype=dtype).copy()
    if g_uv is not None:
        dx = dx + project_uv_backward(xyz, K, np.asarray(g_uv, dtype=dtype))
    if not joint_order_switched:
        dx = swap_joint_order(dx)          # the swap is an involution
    dP = dx * k["s"]
    Rg, Rl, bl, dmap = k["Rg"], k["Rl"], k["bl"], k["dmap"]
    dRg = np.zeros_like(Rg)
    g_bl = np.zeros((B, 20), dtype=dtype)
    g_oa = np.zeros((B, 23), dtype=dtype)
    for i in range(19, -1, -1):
        n = i + 1
        par = 0 if i % 4 == 0 else i
        g_bl[:, i] = np.sum(dP[:, n] * Rg[:, n, :, 2], axis=-1)
        dRg[:, n, :, 2] += bl[:, i, None] * dP[:, n]
        dP[:, par] += dP[:, n]
        dRl = np.swapaxes(Rg[:, par], -1, -2) @ dRg[:, n]
        dRg[:, par] += dRg[:, n] @ np.swapaxes(Rl[:, i], -1, -2)
        if i % 4 != 3:
            gang = euler_xyz_backward(k["ang"][:, i], dRl)
            for a in range(3):
                if dmap[i][a] >= 0:
                    g_oa[:, dmap[i][a]] += gang[:, a]
    g_ra = euler_xyz_backward(k["ra"], dRg[:, 0])
    return g_ra, g_oa, g_bl


def mpjpe(pre_xyz, gt_xyz, keypoint_vis, dtype=np.float64):
    """criterions/metrics.py:10-27: mean over visible joints of ||pre-gt|| * 1000; 0 if none visible."""
    d = np.sqrt(np.sum((np.asarray(pre_xyz, dtype) - np.asarray(gt_xyz, dtype)) ** 2, axis=2))
    m = np.asarray(keypoint_vis).reshape(d.shape).astype(bool)
    if m.sum() == 0:
        return dtype(0.0)
    return dtype(d[m].mean() * 1000.0)


def l2loss(pre_xyz, gt_xyz, keypoint_vis, dtype=np.float64):
    """criterions/loss.py:10-25: mean over visible joints of ||pre-gt||^2; 0 if none visible."""
    d = np.sum((np.asarray(pre_xyz, dtype) - np.asarray(gt_xyz, dtype)) ** 2, axis=2)
    m = np.asarray(keypoint_vis).reshape(d.shape).astype(bool)
    if m.sum() == 0:
        return dtype(0.0)
    return dtype(d[m].mean())


def l2loss_backward(pre_xyz, gt_xyz, keypoint_vis, dtype=np.float64):
    """d l2loss / d pre_xyz = 2 (pre-gt) * vis / N_vis."""
    diff = np.asarray(pre_xyz, dtype) - np.asarray(gt_xyz, dtype)
    m = np.asarray(keypoint_vis).reshape(diff.shape[:2]).astype(bool)
    n = m.sum()
    if n == 0:
        return np.zeros_like(diff)
    return 2.0 * diff * m[..., None] / n


def hand_mask_loss(pred_uv, gt_uv, hand_mask):
    """criterions/loss.py:92-111: uv truncated to int64 and clamped to [0, W-1] on both axes, mask sampled at
    [b, v, u]; 1 - sum(pred samples) / (sum(gt samples) + 1e-8), the last step in fp32 like the reference."""
    mask = np.asarray(hand_mask)
    hi = mask.shape[-1] - 1
    b = np.arange(mask.shape[0]).reshape(-1, 1)
    def samples(uv):
        q = np.clip(np.trunc(np.asarray(uv, np.float64)).astype(np.int64), 0, hi)
        return mask[b, q[..., 1], q[..., 0]].astype(np.float64).sum()
    return np.float32(1.0) - np.float32(samples(pred_uv)) / (np.float32(samples(gt_uv)) + np.float32(1e-8))


def regularizer(theta, beta, dtype=np.float64):
    """criterions/loss.py:113-117: (||theta||_F + 10 ||beta||_F) / 100 over the batch."""
    return dtype((np.linalg.norm(np.asarray(theta, dtype)) + 10.0 * np.linalg.norm(np.asarray(beta, dtype))) / 100.0)


def regularizer_backward(theta, beta, dtype=np.float64):
    """Gradient of ``regularizer`` (autograd of criterions/loss.py:113-117): theta / (100 ||theta||), 10 beta / (100 ||beta||)."""
    t, b = np.asarray(theta, dtype), np.asarray(beta, dtype)
    nt, nb = np.linalg.norm(t), np.linalg.norm(b)
    return (t / (100.0 * nt) if nt > 0 else np.zeros_like(t)), (10.0 * b / (100.0 * nb) if nb > 0 else np.zeros_like(b))


def match_mano_to_rhd(mano_joints, index_root_bone_length, kp_coord_xyz_root, joint_order_switched=True, dtype=np.float64):
    """``match_mano_to_RHD`` (network/Resnet50MANO3DHandPose.py:35-60, the same body at
    network/MANO3DHandPose.py:30-55): optional per-finger joint reversal, root-relative
    coordinates, division by ||joint 12 - root||, then * index_root_bone_length + root.
    Returns (rel_normalized[B,21,3], joint_xyz21[B,21,3]).  The reference permutes its argument
    in place; this restatement (and the product) leave the input alone."""
    j = np.asarray(mano_joints, dtype=dtype)
    if not joint_order_switched:
        j = swap_joint_order(j)
    rel = j - j[:, :1]
    s = np.sqrt(np.sum(rel[:, 12] ** 2, axis=-1))[:, None, None]
    reln = rel / s
    L = np.asarray(index_root_bone_length, dtype=dtype).reshape(-1, 1, 1)
    root = np.asarray(kp_coord_xyz_root, dtype=dtype).reshape(-1, 1, 3)
    return reln, reln * L + root


def match_mano_to_rhd_backward(mano_joints, index_root_bone_length, kp_coord_xyz_root, g_rel, g_xyz,
                               joint_order_switched=True, dtype=np.float64):
    """Gradients of ``match_mano_to_rhd`` w.r.t. (mano_joints, index_root_bone_length,
    kp_coord_xyz_root) for upstream gradients of both outputs."""
    j = np.asarray(mano_joints, dtype=dtype)
    if not joint_order_switched:
        j = swap_joint_order(j)
    L = np.asarray(index_root_bone_length, dtype=dtype).reshape(-1, 1, 1)
    g_rel = np.asarray(g_rel, dtype=dtype)
    g_xyz = np.asarray(g_xyz, dtype=dtype)
    rel = j - j[:, :1]
    s = np.sqrt(np.sum(rel[:, 12] ** 2, axis=-1))[:, None, None]
    reln = rel / s
    g_L = np.sum(reln * g_xyz, axis=(1, 2)).reshape(-1, 1)
    g_root = g_xyz.sum(axis=1)
    gn = g_rel + L * g_xyz
    g_s = -np.sum(gn * rel, axis=(1, 2), keepdims=True) / (s * s)
    gr = gn / s
    gr[:, 12] += (g_s * rel[:, 12:13] / s)[:, 0]
    gj = gr.copy()
    gj[:, 0] -= gr.sum(axis=1)
    if not joint_order_switched:
        gj = swap_joint_order(gj)          # the reversal is its own inverse
    return gj, g_L, g_root
