"""ORACLE — test infrastructure only.  CPU (numpy) restatement of the reference's
batched MANO layer, forward and analytic backward.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  The product
(``3dhandposeestimation_b200``) never does: it calls the sm_100a CUDA library
and fails loudly without it.

Parity pin: the reference has no tests or golden vectors of its own
(SURVEY.md §4), so this restatement is pinned against OUTPUTS OF THE REFERENCE
ITSELF — ``tests/golden/*.npz`` were produced by importing the unmodified
reference classes in the authoring container (``tests/golden/make_golden.py``),
and ``tests/test_oracle_vs_reference.py`` re-runs the live reference when
``/root/reference`` is present.

Every function cites the reference lines it restates
(/root/reference/network/sub_modules/MANOLayer.py unless another file is named).
The model is a dict of numpy arrays with the MANO pkl's keys
(v_template, shapedirs, posedirs, J_regressor, weights, hands_components,
hands_mean, kintree_table).
"""
from __future__ import annotations

import numpy as np

TIP_VERTS = (333, 444, 672, 555, 745)   # :196-200
TIP_SLOTS = (4, 8, 12, 16, 20)
CHAIN_SLOTS = (0, 1, 2, 3, 5, 6, 7, 9, 10, 11, 13, 14, 15, 17, 18, 19)
ROOT_ROT = (np.pi, 0.0, 0.0)            # :76


def parents_of(model) -> list:
    """:65-67 id_to_col / parent."""
    kt = np.asarray(model["kintree_table"])
    id_to_col = {int(kt[1, i]): i for i in range(kt.shape[1])}
    return [-1] + [id_to_col[int(kt[0, i])] for i in range(1, kt.shape[1])]


def constants(model, nc, dtype):
    """:69-76 — each constant is first rounded to fp32 (the reference stores
    fp32 tensors) and then, for the fp64 arbiter, widened exactly."""
    f32 = lambda a: np.asarray(a, dtype=np.float64).astype(np.float32).astype(dtype)
    return {
        "mesh_mu": f32(model["v_template"]),                    # [778,3]
        "mesh_pca": f32(model["shapedirs"]),                    # [778,3,10]
        "posedirs": f32(model["posedirs"]),                     # [778,3,135]
        "J_regressor": f32(model["J_regressor"]),               # [16,778]
        "weights": f32(model["weights"]),                       # [778,16]
        "hands_components": f32(model["hands_components"][:nc]),  # [nc,45]
        "hands_mean": f32(model["hands_mean"]),                 # [45]
        "root_rot": np.asarray(ROOT_ROT, dtype=np.float32).astype(dtype),
    }


def skew(r):
    """S(n) of :85-89.  r [...,3] -> [...,3,3]."""
    z = np.zeros_like(r[..., 0])
    return np.stack([
        np.stack([z, -r[..., 2], r[..., 1]], -1),
        np.stack([r[..., 2], z, -r[..., 0]], -1),
        np.stack([-r[..., 1], r[..., 0], z], -1),
    ], -2)


def rodrigues(r):
    """:82-112.  R = I + sin(t) S(n) + (1-cos t) S(n)^2 with n = r/t; rows with
    t < 1e-30 use the Taylor form I + (1-t^2/6) S(r) + (1/2 - t^2/24) S(r)^2."""
    theta = np.sqrt(np.sum(r * r, axis=-1))
    small = theta < 1e-30
    safe = np.where(small, 1.0, theta).astype(r.dtype)
    n = r / safe[..., None]
    Sn = skew(n)
    eye = np.eye(3, dtype=r.dtype)
    R = eye + np.sin(theta)[..., None, None] * Sn + (1.0 - np.cos(theta))[..., None, None] * (Sn @ Sn)
    if np.any(small):
        Sr = skew(r)
        t2 = (theta * theta)[..., None, None]
        R2 = eye + (1.0 - t2 / 6.0) * Sr + (0.5 - t2 / 24.0) * (Sr @ Sr)
        R = np.where(small[..., None, None], R2, R)
    return R.astype(r.dtype)


def rodrigues_backward(r, dR):
    """Gradient of ``rodrigues``.  At theta -> 0
    the analytic limits are used; the reference's autograd returns NaN there
    (r/theta at :91) — a documented, deliberate deviation."""
    dt = r.dtype
    t2 = np.sum(r * r, axis=-1)
    theta = np.sqrt(t2)
    small = theta < (1e-4 if dt == np.float64 else 1e-3)
    th = np.where(small, 1.0, theta)
    a = np.where(small, 1.0 - t2 / 6.0, np.sin(th) / th)
    b = np.where(small, 0.5 - t2 / 24.0, (1.0 - np.cos(th)) / (th * th))
    a2 = np.where(small, -1.0 / 3.0 + t2 / 30.0, (th * np.cos(th) - np.sin(th)) / th ** 3)
    b2 = np.where(small, -1.0 / 12.0 + t2 / 180.0, (th * np.sin(th) - 2.0 * (1.0 - np.cos(th))) / th ** 4)
    S = skew(r)
    S2 = S @ S
    gS = np.sum(dR * S, axis=(-1, -2))
    gS2 = np.sum(dR * S2, axis=(-1, -2))
    out = np.zeros_like(r)
    eye = np.eye(3, dtype=dt)
    for i in range(3):
        E = skew(np.broadcast_to(eye[i], r.shape).astype(dt))
        gE = np.sum(dR * E, axis=(-1, -2))
        gES = np.sum(dR * (E @ S + S @ E), axis=(-1, -2))
        out[..., i] = (a2 * gS + b2 * gS2) * r[..., i] + a * gE + b * gES
    return out.astype(dt)


def mano_forward(model, rots, poses, betas, dtype=np.float64, return_cache=False):
    """``ManoLayer.rot_pose_beta_to_mesh`` (:122-208) == ``forward`` (:238-240).

    rots [B,3] global axis-angle, poses [B,nc] PCA coefficients, betas [B,10]
    -> vertices [B,778,3], joint [B,21,3] (contiguous)."""
    rots = np.asarray(rots, dtype=dtype)
    poses = np.asarray(poses, dtype=dtype)
    betas = np.asarray(betas, dtype=dtype)
    B, nc = poses.shape
    c = constants(model, nc, dtype)
    parent = parents_of(model)
    nj = len(parent)

    # :126 PCA -> axis-angle, :128 constant root [pi,0,0] prepended
    theta = c["hands_mean"] + poses @ c["hands_components"]
    pose = np.concatenate([np.broadcast_to(c["root_rot"], (B, 1, 3)), theta.reshape(B, nj - 1, 3)], axis=1)

    # :130-132 shape blend
    v_shaped = c["mesh_mu"][None] + (betas @ c["mesh_pca"].reshape(-1, betas.shape[1]).T).reshape(B, -1, 3)
    # :116-119 pose feature, :134-137 pose blend
    R = rodrigues(pose.reshape(-1, 3)).reshape(B, nj, 3, 3)
    pf = (R[:, 1:] - np.eye(3, dtype=dtype)).reshape(B, (nj - 1) * 9)
    v_posed = v_shaped + (pf @ c["posedirs"].reshape(-1, pf.shape[1]).T).reshape(B, -1, 3)
    # :139-141 joint regression from the SHAPED mesh
    J = np.matmul(c["J_regressor"], v_shaped)

    # :159-165 chain
    Rg = np.zeros((B, nj, 3, 3), dtype=dtype)
    tg = np.zeros((B, nj, 3), dtype=dtype)
    Rg[:, 0] = R[:, 0]
    tg[:, 0] = J[:, 0]
    for i in range(1, nj):
        p = parent[i]
        Rg[:, i] = Rg[:, p] @ R[:, i]
        tg[:, i] = tg[:, p] + np.einsum("bij,bj->bi", Rg[:, p], J[:, i] - J[:, p])
    # :169-175 remove rest pose: A_i = [Rg_i | tg_i - Rg_i J_i]
    tA = tg - np.einsum("bkij,bkj->bki", Rg, J)
    # :177-185 LBS
    Tr = np.matmul(c["weights"], Rg.reshape(B, nj, 9)).reshape(B, -1, 3, 3)
    Tt = np.matmul(c["weights"], tA)
    v = np.matmul(Tr, v_posed[..., None])[..., 0] + Tt
    # :190-202 joints: 16 chain translations with the 5 tips inserted
    Jtr = np.zeros((B, 21, 3), dtype=dtype)
    Jtr[:, list(CHAIN_SLOTS)] = tg
    Jtr[:, list(TIP_SLOTS)] = v[:, list(TIP_VERTS)]
    # :188, :204-205 global rotation about the origin
    Rq = rodrigues(rots)
    vertices = np.matmul(v, np.swapaxes(Rq, -1, -2))
    joint = np.matmul(Jtr, np.swapaxes(Rq, -1, -2))
    if return_cache:
        cache = dict(c=c, parent=parent, pose=pose, R=R, pf=pf, v_posed=v_posed, J=J,
                     Rg=Rg, tg=tg, tA=tA, v=v, Jtr=Jtr, Rq=Rq)
        return vertices.astype(dtype), joint.astype(dtype), cache
    return vertices.astype(dtype), joint.astype(dtype)


def mano_backward(model, rots, poses, betas, g_verts, g_joints, dtype=np.float64):
    """Analytic gradient of ``mano_forward`` w.r.t. (rots, poses, betas) for
    upstream ``g_verts`` [B,778,3] (or None) and ``g_joints`` [B,21,3].  The
    reference obtains this from the autograd tape of :122-208; the closed form
    follows SURVEY Appendix A.2 and is validated against the reference's
    autograd in tests/test_oracle_vs_reference.py."""
    _, _, k = mano_forward(model, rots, poses, betas, dtype=dtype, return_cache=True)
    c, parent = k["c"], k["parent"]
    B = k["pose"].shape[0]
    nj = len(parent)
    nc = np.asarray(poses).shape[1]
    g_joints = np.asarray(g_joints, dtype=dtype)
    gv_out = np.zeros((B, 778, 3), dtype=dtype) if g_verts is None else np.asarray(g_verts, dtype=dtype)

    Rq, Rg, tg, J, R = k["Rq"], k["Rg"], k["tg"], k["J"], k["R"]
    # global rotation: vertices = Rq v ; joint = Rq Jtr
    dRq = np.matmul(np.swapaxes(gv_out, -1, -2), k["v"]) + np.matmul(np.swapaxes(g_joints, -1, -2), k["Jtr"])
    gv = np.matmul(gv_out, Rq)                          # Rq^T g
    gJtr = np.matmul(g_joints, Rq)
    gv[:, list(TIP_VERTS)] += gJtr[:, list(TIP_SLOTS)]
    dtg = gJtr[:, list(CHAIN_SLOTS)].copy()

    # LBS: v = (sum_k w Rg_k) v_posed + sum_k w tA_k
    W = c["weights"]
    outer = (gv[..., :, None] * k["v_posed"][..., None, :]).reshape(B, -1, 9)
    dRg = np.matmul(W.T, outer).reshape(B, nj, 3, 3)
    dtA = np.matmul(W.T, gv)
    Tr = np.matmul(W, Rg.reshape(B, nj, 9)).reshape(B, -1, 3, 3)
    dv_posed = np.matmul(np.swapaxes(Tr, -1, -2), gv[..., None])[..., 0]
    # tA = tg - Rg J
    dtg += dtA
    dRg -= np.einsum("bki,bkj->bkij", dtA, J)
    dJ = -np.einsum("bkij,bki->bkj", Rg, dtA)
    # reverse chain
    dR = np.zeros_like(R)
    for i in range(nj - 1, 0, -1):
        p = parent[i]
        dR[:, i] = np.einsum("bji,bjk->bik", Rg[:, p], dRg[:, i])
        dRg[:, p] += dRg[:, i] @ np.swapaxes(R[:, i], -1, -2)
        dRg[:, p] += np.einsum("bi,bj->bij", dtg[:, i], J[:, i] - J[:, p])
        dtg[:, p] += dtg[:, i]
        dd = np.einsum("bji,bj->bi", Rg[:, p], dtg[:, i])
        dJ[:, i] += dd
        dJ[:, p] -= dd
    dJ[:, 0] += dtg[:, 0]          # tg_0 = J_0 ; R_0 is constant
    # v_posed = v_shaped + posedirs pf ; J = Jreg v_shaped
    dpf = dv_posed.reshape(B, -1) @ c["posedirs"].reshape(-1, (nj - 1) * 9)
    dR[:, 1:] += dpf.reshape(B, nj - 1, 3, 3)
    dv_shaped = dv_posed + np.matmul(c["J_regressor"].T, dJ)
    g_betas = dv_shaped.reshape(B, -1) @ c["mesh_pca"].reshape(-1, c["mesh_pca"].shape[-1])
    dtheta = rodrigues_backward(k["pose"][:, 1:].reshape(-1, 3), dR[:, 1:].reshape(-1, 3, 3)).reshape(B, 45)
    g_poses = dtheta @ c["hands_components"].T
    g_rots = rodrigues_backward(np.asarray(rots, dtype=dtype), dRq)
    assert g_poses.shape == (B, nc)
    return g_rots.astype(dtype), g_poses.astype(dtype), g_betas.astype(dtype)
