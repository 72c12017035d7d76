"""ORACLE — test infrastructure only (see mano_oracle.py header for the rules: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product never does).

numpy restatement of the reference's keypoint re-parameterisations that run per sample in its
dataloaders (dataloaderRHD.py:242-250):
  utils/relative_trafo.py:167-270   bone_rel_trafo / bone_rel_trafo_inv
  utils/canonical_trafo.py:93-184   canonical_trafo / flip_right_hand
Pinned against the live reference (tests/test_oracle_vs_reference.py) and the fixtures it produced
(tests/golden/trafo.npz, tests/golden/make_golden.py).
"""
import numpy as np

# relative_trafo.py:125-166: child -> parent (-1 = 'root': the origin, not joint 0) and evaluation order
PARENT = {0: -1, 4: -1, 3: 4, 2: 3, 1: 2, 8: -1, 7: 8, 6: 7, 5: 6, 12: -1, 11: 12, 10: 11, 9: 10,
          16: -1, 15: 16, 14: 15, 13: 14, 20: -1, 19: 20, 18: 19, 17: 18}
ORDER = [0, 4, 3, 2, 1, 8, 7, 6, 5, 12, 11, 10, 9, 16, 15, 14, 13, 20, 19, 18, 17]


def _rot_x(a):
    c, s, o, z = np.cos(a), np.sin(a), np.ones_like(a), np.zeros_like(a)
    return np.stack([np.stack([o, z, z], -1), np.stack([z, c, -s], -1), np.stack([z, s, c], -1)], -2)


def _rot_y(a):
    c, s, o, z = np.cos(a), np.sin(a), np.ones_like(a), np.zeros_like(a)
    return np.stack([np.stack([c, z, s], -1), np.stack([z, o, z], -1), np.stack([-s, z, c], -1)], -2)


def _rot_z(a):
    c, s, o, z = np.cos(a), np.sin(a), np.ones_like(a), np.zeros_like(a)
    return np.stack([np.stack([c, -s, z], -1), np.stack([s, c, z], -1), np.stack([z, z, o], -1)], -2)


def bone_rel_trafo(coords_xyz, dtype=np.float64):
    """relative_trafo.py:167-216 with :103-123 ``_backward``.  The 4x4 transforms of the reference
    are affine [R | t]; a bone vector is a difference of two points in the parent frame, so only R
    enters: delta = R (child - parent) (children of 'root': delta = the point itself, R = I).
    -> [B,21,3] = (length, angle_x, angle_y)."""
    c = np.asarray(coords_xyz, dtype=dtype).reshape(-1, 21, 3)
    B = c.shape[0]
    R = {}
    out = np.zeros((B, 21, 3), dtype=dtype)
    eps = dtype(1e-8)
    for b in ORDER:
        p = PARENT[b]
        if p < 0:
            Rp = np.tile(np.eye(3, dtype=dtype), (B, 1, 1))
            d = c[:, b]
        else:
            Rp = R[p]
            d = np.einsum("bij,bj->bi", Rp, c[:, b] - c[:, p])
        length = np.sqrt(d[:, 0] ** 2 + d[:, 1] ** 2 + d[:, 2] ** 2)
        ay = np.arctan2(d[:, 0], d[:, 2] + eps)
        tmp = np.einsum("bij,bj->bi", _rot_y(-ay), d)
        ax = np.arctan2(-tmp[:, 1], tmp[:, 2] + eps)
        R[b] = _rot_x(-ax) @ _rot_y(-ay) @ Rp
        out[:, b] = np.stack([length, ax, ay], -1)
    return out


def bone_rel_trafo_inv(coords_rel, dtype=np.float64):
    """relative_trafo.py:219-270 with :87-100 ``_forward``: T <- Trans_z(-len) Rx(-ax) Ry(-ay) T and the
    joint is inverse(T) applied to the origin, i.e. parent position + len * (third row of the new R)."""
    r = np.asarray(coords_rel, dtype=dtype)
    if r.ndim == 2:
        r = r[None]
    B = r.shape[0]
    R, P = {}, {}
    out = np.zeros((B, 21, 3), dtype=dtype)
    for b in ORDER:
        p = PARENT[b]
        Rp = np.tile(np.eye(3, dtype=dtype), (B, 1, 1)) if p < 0 else R[p]
        Pp = np.zeros((B, 3), dtype=dtype) if p < 0 else P[p]
        R[b] = _rot_x(-r[:, b, 1]) @ _rot_y(-r[:, b, 2]) @ Rp
        P[b] = Pp + r[:, b, 0:1] * R[b][:, 2, :]
        out[:, b] = P[b]
    return out


def atan2_reference(y, x):
    """canonical_trafo.py:23-41 ``atan2_pytorch``: atan(y / (x + 1e-8)) moved to (-pi, pi] by quadrant."""
    x = np.asarray(x)
    y = np.asarray(y)
    xe = x + x.dtype.type(1e-8)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.arctan(y / xe)
    pi = x.dtype.type(3.141592653589793)
    t = t + np.where(xe < 0, pi, 0)
    t = t + np.where(t < 0, 2 * pi, 0)
    t = t + np.where(t > pi, -2 * pi, 0)
    return t


def canonical_trafo(coords_xyz, dtype=np.float64):
    """canonical_trafo.py:93-159 -> (coords_xyz_normed[B,21,3], total_rot_mat[B,3,3])."""
    c = np.asarray(coords_xyz, dtype=dtype).reshape(-1, 21, 3)
    t = c - c[:, :1]
    p = t[:, 12]
    Rz = _rot_z(atan2_reference(p[:, 0], p[:, 1]))
    r1 = t @ np.swapaxes(Rz, 1, 2)
    total = Rz
    p = r1[:, 12]
    Rx = _rot_x(-atan2_reference(p[:, 2], p[:, 1]) + dtype(3.141592653589793))
    r2 = r1 @ np.swapaxes(Rx, 1, 2)
    total = total @ Rx
    p = r2[:, 20]
    Ry = _rot_y(atan2_reference(p[:, 2], p[:, 0]))
    out = r2 @ np.swapaxes(Ry, 1, 2)
    total = total @ Ry
    return out, total


def flip_right_hand(coords_xyz_canonical, cond_right):
    """canonical_trafo.py:163-184: z -> -z for the hands where cond_right is true."""
    c = np.asarray(coords_xyz_canonical)
    cond = np.asarray(cond_right).astype(bool)
    m = c * np.array([1, 1, -1], dtype=c.dtype)
    return np.where(cond.reshape(cond.shape + (1,) * (c.ndim - cond.ndim)), m, c)


def get_rot_mat(ux, uy, uz, dtype=np.float64):
    """utils/general.py:191-226 ``_get_rot_mat``: theta = sqrt(|u|^2 + 1e-8), axis = u * (1 / theta),
    Rodrigues -> [B,3,3].  Inputs [B,1] or [B]."""
    ux, uy, uz = (np.asarray(a, dtype=dtype).reshape(-1) for a in (ux, uy, uz))
    th = np.sqrt(ux ** 2 + uy ** 2 + uz ** 2 + dtype(1e-8))
    st, ct, oc = np.sin(th), np.cos(th), 1.0 - np.cos(th)
    nx, ny, nz = ux / th, uy / th, uz / th
    top = np.stack([ct + nx * nx * oc, nx * ny * oc - nz * st, nx * nz * oc + ny * st], 1)
    mid = np.stack([ny * nx * oc + nz * st, ct + ny * ny * oc, ny * nz * oc - nx * st], 1)
    bot = np.stack([nz * nx * oc - ny * st, nz * ny * oc + nx * st, ct + nz * nz * oc], 1)
    return np.stack([top, mid, bot], 1)


def viewpoint_forward(can_xyz, ux, uy, uz, dtype=np.float64):
    """network/Hand3DPoseNet.py:41-43: (rot_mat, coord_xyz_rel_normed = can_xyz_kps21 @ rot_mat)."""
    R = get_rot_mat(ux, uy, uz, dtype)
    return R, np.asarray(can_xyz, dtype=dtype) @ R


def viewpoint_backward(can_xyz, ux, uy, uz, g_rot, g_rel, dtype=np.float64, h=1e-6):
    """Gradients of ``viewpoint_forward`` w.r.t. (can_xyz, ux, uy, uz): g_can analytically, the three
    axis-angle gradients by central differences in float64 of <g_rot, R> + <g_rel, can @ R> (a check that is
    independent of the kernel's closed form)."""
    can = np.asarray(can_xyz, dtype=dtype)
    g_rot = np.asarray(g_rot, dtype=dtype)
    g_rel = np.asarray(g_rel, dtype=dtype)
    u = [np.asarray(a, dtype=dtype).reshape(-1) for a in (ux, uy, uz)]
    R = get_rot_mat(*u, dtype=dtype)
    g_can = g_rel @ np.swapaxes(R, 1, 2)
    G = g_rot + np.swapaxes(can, 1, 2) @ g_rel                  # dL/dR
    gu = []
    for k in range(3):
        up = [a.copy() for a in u]
        um = [a.copy() for a in u]
        up[k] += h
        um[k] -= h
        dR = (get_rot_mat(*up, dtype=dtype) - get_rot_mat(*um, dtype=dtype)) / (2 * h)
        gu.append(np.sum(G * dR, axis=(1, 2)))
    return g_can, gu[0], gu[1], gu[2]
