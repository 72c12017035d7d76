"""Generate tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_import.py) on CPU, fp32.

Run in the authoring container only:  python tests/golden/make_golden.py
The MANO model used is the seeded synthetic MANO-shaped model
(assets.synthetic_mano) written to a temporary pkl that the reference's own
ManoLayer constructor opens — so no MANO-licensed data enters the fixtures.
Known-answer scalars from the real MANO_RIGHT.pkl go to kat_real_mano.json
(a handful of sums/coordinates only).
"""
import importlib
import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
assets = importlib.import_module("3dhandposeestimation_b200.assets")
from oracle import ref_import  # noqa: E402

ref = ref_import.load()
torch.set_num_threads(1)


def mano_case(pkl, nc, B, seed, scale_pose=np.pi):
    layer = ref.ManoLayer("cpu", pkl, pose_num=nc)
    g = torch.Generator().manual_seed(seed)
    rot = ((torch.rand(B, 3, generator=g) - .5) * 2 * np.pi).requires_grad_()
    pose = ((torch.rand(B, 45, generator=g) - .5) * scale_pose)[:, :nc].clone().requires_grad_()
    beta = (torch.rand(B, 10, generator=g) - .5).requires_grad_()
    v, j = layer(rot, pose, beta)
    gv = torch.randn(v.shape, generator=g)
    gj = torch.randn(j.shape, generator=g)
    ((v * gv).sum() + (j * gj).sum()).backward()
    full = [t.grad.clone() for t in (rot, pose, beta)]
    for t in (rot, pose, beta):
        t.grad = None
    v2, j2 = layer(rot, pose, beta)
    (j2 * gj).sum().backward()          # the heads' joints-only case
    jo = [t.grad.clone() for t in (rot, pose, beta)]
    n = lambda t: t.detach().contiguous().numpy()
    return dict(rot=n(rot), pose=n(pose), beta=n(beta), verts=n(v), joints=n(j), g_verts=n(gv), g_joints=n(gj),
                g_rot=n(full[0]), g_pose=n(full[1]), g_beta=n(full[2]),
                gj_rot=n(jo[0]), gj_pose=n(jo[1]), gj_beta=n(jo[2]))


def mano_affine_case(pkl, nc, B, seed):
    """The reference layer followed by the callers' post-ops (resnet50MANO.py:77-81: scale * x3d, + trans), with a
    3-D translation: p' = scale * p + transl on vertices and joints, and the reference's autograd for all five inputs."""
    layer = ref.ManoLayer("cpu", pkl, pose_num=nc)
    g = torch.Generator().manual_seed(seed)
    rot = ((torch.rand(B, 3, generator=g) - .5) * 2 * np.pi).requires_grad_()
    pose = ((torch.rand(B, nc, generator=g) - .5) * 2).requires_grad_()
    beta = (torch.rand(B, 10, generator=g) - .5).requires_grad_()
    transl = (torch.randn(B, 3, generator=g) * .1 + torch.tensor([0, 0, .6])).requires_grad_()
    scale = (torch.rand(B, generator=g) + .5).requires_grad_()
    v, j = layer(rot, pose, beta)
    v = scale.unsqueeze(1).unsqueeze(2) * v + transl.unsqueeze(1)
    j = scale.unsqueeze(1).unsqueeze(2) * j + transl.unsqueeze(1)
    gv = torch.randn(v.shape, generator=g)
    gj = torch.randn(j.shape, generator=g)
    ((v * gv).sum() + (j * gj).sum()).backward()
    n = lambda t: t.detach().contiguous().numpy()
    return dict(rot=n(rot), pose=n(pose), beta=n(beta), transl=n(transl), scale=n(scale), verts=n(v), joints=n(j),
                g_verts=n(gv), g_joints=n(gj), g_rot=n(rot.grad), g_pose=n(pose.grad), g_beta=n(beta.grad),
                g_transl=n(transl.grad), g_scale=n(scale.grad))


def fk_case(B, seed, switched):
    ref.config.joint_order_switched = switched
    fk = ref.ForwardKinematics("cpu")
    g = torch.Generator().manual_seed(seed)
    ra = ((torch.rand(B, 3, generator=g) - .5) * 2 * np.pi).requires_grad_()
    oa = ((torch.rand(B, 23, generator=g) - .5) * np.pi).requires_grad_()
    bl = (torch.rand(B, 20, generator=g) + .1).requires_grad_()
    K = torch.tensor([[282.9, 0, 160], [0, 282.9, 160], [0, 0, 1.]]).repeat(B, 1, 1)
    K = K + torch.rand(B, 3, 3, generator=g) * torch.tensor([[1., 0, 1], [0, 1, 1], [0, 0, 0]])
    sc = torch.rand(B, 1, generator=g) * .05 + .02
    root = torch.randn(B, 3, generator=g) * .05 + torch.tensor([0, 0, .6])
    xyz, uv, _ = fk(ra, oa, bl, K, sc, root)
    gx = torch.randn(xyz.shape, generator=g)
    gu = torch.randn(uv.shape, generator=g) * 1e-3
    ((xyz * gx).sum() + (uv * gu).sum()).backward()
    ref.config.joint_order_switched = True
    n = lambda t: t.detach().contiguous().numpy()
    return dict(root_angles=n(ra), other_angles=n(oa), bone_lengths=n(bl), K=n(K), scale=n(sc), root=n(root),
                xyz=n(xyz), uv=n(uv), g_xyz=n(gx), g_uv=n(gu), g_root_angles=n(ra.grad),
                g_other_angles=n(oa.grad), g_bone_lengths=n(bl.grad), switched=np.array(switched))


def reduce_case(B, seed, p_vis):
    g = torch.Generator().manual_seed(seed)
    pre = (torch.randn(B, 21, 3, generator=g) * .05).requires_grad_()
    gt = torch.randn(B, 21, 3, generator=g) * .05
    vis = (torch.rand(B, 21, 1, generator=g) < p_vis).float()
    m = ref.MPJPE()(pre.detach(), gt, vis)
    l2 = ref.L2Loss()(pre, gt, vis)
    if l2.requires_grad:
        l2.backward()
        gpre = pre.grad
    else:
        gpre = torch.zeros_like(pre)
    n = lambda t: t.detach().contiguous().numpy()
    return dict(pre=n(pre), gt=n(gt), vis=n(vis), mpjpe=np.float32(m.item()), l2=np.float32(l2.item()), g_pre=n(gpre))


def reduce_uv_case(B, seed, p_vis):
    """L2Loss on [B,21,2] — LossCalculation.compute_uv_coord_loss (criterions/loss.py:86-87) sends uv through the same class."""
    g = torch.Generator().manual_seed(seed)
    pre = (torch.rand(B, 21, 2, generator=g) * 320).requires_grad_()
    gt = torch.rand(B, 21, 2, generator=g) * 320
    vis = (torch.rand(B, 21, 1, generator=g) < p_vis).float()
    l2 = ref.L2Loss()(pre, gt, vis)
    l2.backward()
    n = lambda t: t.detach().contiguous().numpy()
    return dict(pre=n(pre), gt=n(gt), vis=n(vis), l2=np.float32(l2.item()), g_pre=n(pre.grad))


def regulariser_case(B, nc, seed):
    """LossCalculation.compute_regularization_loss (criterions/loss.py:113-117) and its autograd."""
    from criterions.loss import LossCalculation
    g = torch.Generator().manual_seed(seed)
    theta = ((torch.rand(B, nc, generator=g) - .5) * 4).requires_grad_()
    beta = ((torch.rand(B, 10, generator=g) - .5) * .1).requires_grad_()
    loss = LossCalculation(comp_regularization_loss=True).compute_regularization_loss(theta, beta)
    (loss * 3.0).backward()
    n = lambda t: t.detach().contiguous().numpy()
    return dict(theta=n(theta), beta=n(beta), loss=np.float32(loss.item()), g_theta=n(theta.grad), g_beta=n(beta.grad),
                g_out=np.float32(3.0))


def proj_case():
    g = torch.Generator().manual_seed(7)
    B = 3
    xyz = torch.randn(B, 21, 3, generator=g) * .1 + torch.tensor([0, 0, .5])
    xyz[0, 0] = 0.0           # exercises the p_z == 0 -> 1e-10 branch (coordinate_trans.py:59)
    xyz[1, 5, 2] = 0.0
    xyz.requires_grad_()
    K = torch.tensor([[600., 0, 300], [0, 600., 300], [0, 0, 1.]]).repeat(B, 1, 1)
    uv = ref.batch_project_xyz_to_uv(xyz, K)
    gu = torch.randn(uv.shape, generator=g)
    (uv * gu).sum().backward()
    n = lambda t: t.detach().contiguous().numpy()
    return dict(xyz=n(xyz), K=n(K), uv=n(uv), g_uv=n(gu), g_xyz=n(xyz.grad))


def match_case(B, seed, switched):
    """match_mano_to_RHD -> batch_project_xyz_to_uv, the heads' joint epilogue
    (network/Resnet50MANO3DHandPose.py:35-60,73)."""
    from network.Resnet50MANO3DHandPose import Resnet50MANO3DHandPose
    ref.config.joint_order_switched = switched
    g = torch.Generator().manual_seed(seed)
    j = (torch.randn(B, 21, 3, generator=g) * .04).requires_grad_()
    L = (torch.rand(B, 1, generator=g) * .05 + .02).requires_grad_()
    root = (torch.randn(B, 3, generator=g) * .05 + torch.tensor([0, 0, .6])).requires_grad_()
    K = torch.tensor([[282.9, 0, 160], [0, 282.9, 160], [0, 0, 1.]]).repeat(B, 1, 1)
    rel, xyz = Resnet50MANO3DHandPose.match_mano_to_RHD(None, j.clone(), L, root)   # the reference permutes its argument in place
    uv = ref.batch_project_xyz_to_uv(xyz, K)
    gr = torch.randn(rel.shape, generator=g)
    gx = torch.randn(xyz.shape, generator=g)
    gu = torch.randn(uv.shape, generator=g) * 1e-3
    ((rel * gr).sum() + (xyz * gx).sum() + (uv * gu).sum()).backward()
    ref.config.joint_order_switched = True
    n = lambda t: t.detach().contiguous().numpy()
    return dict(joints=n(j), scale=n(L), root=n(root), K=n(K), rel=n(rel), xyz=n(xyz), uv=n(uv), g_rel=n(gr), g_xyz=n(gx),
                g_uv=n(gu), g_joints=n(j.grad), g_scale=n(L.grad), g_root=n(root.grad), switched=np.array(switched))


def head_loss_case(pkl, nc, B, seed, match, switched=True):
    """The MANO heads' tail end to end through the unmodified reference: ManoLayer -> scale * joints + transl
    (resnet50MANO.py:77-81) -> [match_mano_to_RHD, Resnet50MANO3DHandPose.py:35-60] -> batch_project_xyz_to_uv (:73) ->
    LossCalculation(xyz, uv, regularisation) (criterions/loss.py:62-153), weighted sum, autograd to every input."""
    from criterions.loss import LossCalculation
    from network.Resnet50MANO3DHandPose import Resnet50MANO3DHandPose
    ref.config.joint_order_switched = switched
    layer = ref.ManoLayer("cpu", pkl, pose_num=nc)
    g = torch.Generator().manual_seed(seed)
    rot = ((torch.rand(B, 3, generator=g) - .5) * 2).requires_grad_()
    pose = ((torch.rand(B, nc, generator=g) - .5) * 2).requires_grad_()
    beta = (torch.rand(B, 10, generator=g) - .5).requires_grad_()
    transl = (torch.randn(B, 3, generator=g) * .05 + torch.tensor([0, 0, .6])).requires_grad_()
    scale = (torch.rand(B, generator=g) * .4 + .8).requires_grad_()
    L = torch.rand(B, 1, generator=g) * .05 + .02
    root = torch.randn(B, 3, generator=g) * .05 + torch.tensor([0, 0, .6])
    K = torch.tensor([[282.9, 0, 160], [0, 282.9, 160], [0, 0, 1.]]).repeat(B, 1, 1)
    vis = (torch.rand(B, 21, 1, generator=g) < .8).float()
    _, j = layer(rot, pose, beta)
    j = scale.unsqueeze(1).unsqueeze(2) * j + transl.unsqueeze(1)
    if match:
        _, xyz = Resnet50MANO3DHandPose.match_mano_to_RHD(None, j.clone(), L, root)
    else:
        xyz = j
    uv = ref.batch_project_xyz_to_uv(xyz, K)
    gt_xyz = xyz.detach() + torch.randn(B, 21, 3, generator=g) * .05     # offsets well above the fp32 noise of the chain
    gt_uv = uv.detach() + torch.randn(B, 21, 2, generator=g) * 20.0
    crit = LossCalculation("cpu", comp_xyz_loss=True, comp_uv_loss=True, comp_regularization_loss=True)
    loss_xyz, loss_uv, _, _, loss_reg = crit(xyz, gt_xyz, uv, gt_uv, vis, theta=pose, beta=beta)
    w = torch.tensor([1.0, 1e-4, 0.5])
    (w[0] * loss_xyz + w[1] * loss_uv + w[2] * loss_reg).backward()
    ref.config.joint_order_switched = True
    n = lambda t: t.detach().contiguous().numpy()
    return dict(rot=n(rot), pose=n(pose), beta=n(beta), transl=n(transl), scale=n(scale), L=n(L), root=n(root), K=n(K), vis=n(vis),
                gt_xyz=n(gt_xyz), gt_uv=n(gt_uv), xyz=n(xyz), uv=n(uv), losses=np.array([loss_xyz.item(), loss_uv.item(), loss_reg.item()]),
                weights=n(w), g_rot=n(rot.grad), g_pose=n(pose.grad), g_beta=n(beta.grad), g_transl=n(transl.grad),
                g_scale=n(scale.grad), match=np.array(match), switched=np.array(switched))


def trafo_case(B, seed):
    """bone_rel_trafo / bone_rel_trafo_inv / canonical_trafo / flip_right_hand on hand-like keypoints
    (root-relative, normalised: joint 0 at the origin, bones ~0.3-1 long)."""
    from utils.canonical_trafo import canonical_trafo, flip_right_hand
    from utils.relative_trafo import bone_rel_trafo, bone_rel_trafo_inv
    g = torch.Generator().manual_seed(seed)
    spread = torch.arange(21)[:, None] * torch.tensor([.09, .06, .03])
    xyz = torch.randn(B, 21, 3, generator=g) * .5 + spread
    xyz[:, 0] = 0
    xyz[1] += torch.tensor([.3, -.2, .1])                    # one hand whose root is not at the origin
    rel = bone_rel_trafo(xyz)
    inv = bone_rel_trafo_inv(rel)
    can, rot = canonical_trafo(xyz)
    cond = torch.rand(B, generator=g) < .5
    flipped = flip_right_hand(can, cond[:, None].expand(B, 21))
    n = lambda t: t.detach().contiguous().numpy()
    return dict(xyz=n(xyz), rel=n(rel), inv=n(inv), can=n(can), rot=n(rot), cond_right=n(cond), flipped=n(flipped))


def viewpoint_case(B, seed):
    """_get_rot_mat -> can @ R -> * scale + root -> projection (network/Hand3DPoseNet.py:41-50) with autograd."""
    from utils.general import _get_rot_mat
    g = torch.Generator().manual_seed(seed)
    can = (torch.randn(B, 21, 3, generator=g) * .5).requires_grad_()
    u = [((torch.rand(B, 1, generator=g) - .5) * 4).requires_grad_() for _ in range(3)]
    with torch.no_grad():
        u[0][0] = u[1][0] = u[2][0] = 0.0                     # theta = 1e-4: the epsilon branch
    L = torch.rand(B, 1, generator=g) * .05 + .02
    root = torch.randn(B, 3, generator=g) * .05 + torch.tensor([0, 0, .6])
    K = torch.tensor([[282.9, 0, 160], [0, 282.9, 160], [0, 0, 1.]]).repeat(B, 1, 1)
    R = _get_rot_mat(*u)
    rel = torch.matmul(can, R)
    xyz = rel * L.unsqueeze(-1) + root.unsqueeze(1)
    uv = ref.batch_project_xyz_to_uv(xyz, K)
    gR = torch.randn(R.shape, generator=g)
    gr = torch.randn(rel.shape, generator=g)
    ((R * gR).sum() + (rel * gr).sum()).backward()
    n = lambda t: t.detach().contiguous().numpy()
    return dict(can=n(can), ux=n(u[0]), uy=n(u[1]), uz=n(u[2]), scale=n(L), root=n(root), K=n(K), rot=n(R), rel=n(rel),
                xyz=n(xyz), uv=n(uv), g_rot=n(gR), g_rel=n(gr), g_can=n(can.grad), g_ux=n(u[0].grad), g_uy=n(u[1].grad),
                g_uz=n(u[2].grad))


def hand_mask_case(B, seed, empty_gt=False):
    """LossCalculation.compute_hand_mask_loss (criterions/loss.py:92-111) on a blob mask; some uv outside the image."""
    from criterions.loss import LossCalculation
    g = torch.Generator().manual_seed(seed)
    H = W = 64
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    c = torch.rand(B, 2, generator=g) * 24 + 20
    mask = (((xx[None] - c[:, 0, None, None]) ** 2 + (yy[None] - c[:, 1, None, None]) ** 2) < 15 ** 2).float()
    gt_uv = c[:, None, :] + torch.randn(B, 21, 2, generator=g) * 9
    pred_uv = gt_uv + torch.randn(B, 21, 2, generator=g) * 8
    pred_uv[0, 0] = torch.tensor([-7.5, 300.25])               # clamped on both sides
    pred_uv[1, 3] = torch.tensor([-0.9, 63.999])               # truncation toward zero
    if empty_gt:
        mask[:] = 0
    loss = LossCalculation("cpu", comp_hand_mask_loss=True).compute_hand_mask_loss(pred_uv, gt_uv, mask)
    n = lambda t: t.detach().contiguous().numpy()
    return dict(pred_uv=n(pred_uv), gt_uv=n(gt_uv), hand_mask=n(mask), loss=np.float32(loss.item()))


def main():
    model = assets.synthetic_mano()
    with tempfile.TemporaryDirectory() as td:
        pkl = os.path.join(td, "synthetic_mano.pkl")
        assets.write_reference_style_pkl(model, pkl)
        np.savez_compressed(os.path.join(HERE, "head_loss_match.npz"), **head_loss_case(pkl, 10, 7, 91, True, switched=False))
        np.savez_compressed(os.path.join(HERE, "head_loss_plain.npz"), **head_loss_case(pkl, 45, 6, 92, False))
        if len(sys.argv) > 1 and sys.argv[1] == "head_loss":       # only the fixtures added in round 2
            return
        np.savez_compressed(os.path.join(HERE, "mano_synth_nc45.npz"), **mano_case(pkl, 45, 4, 1234))
        np.savez_compressed(os.path.join(HERE, "mano_synth_nc10.npz"), **mano_case(pkl, 10, 4, 1234))
        np.savez_compressed(os.path.join(HERE, "mano_synth_nc6.npz"), **mano_case(pkl, 6, 3, 99, scale_pose=4.0))
        np.savez_compressed(os.path.join(HERE, "mano_affine_nc10.npz"), **mano_affine_case(pkl, 10, 5, 77))
    np.savez_compressed(os.path.join(HERE, "fk_switched.npz"), **fk_case(8, 1234, True))
    np.savez_compressed(os.path.join(HERE, "fk_unswitched.npz"), **fk_case(8, 4321, False))
    np.savez_compressed(os.path.join(HERE, "reduce_vis80.npz"), **reduce_case(16, 5, .8))
    np.savez_compressed(os.path.join(HERE, "reduce_none_visible.npz"), **reduce_case(4, 6, -1.0))
    np.savez_compressed(os.path.join(HERE, "reduce_uv.npz"), **reduce_uv_case(9, 15, .7))
    np.savez_compressed(os.path.join(HERE, "regulariser.npz"), **regulariser_case(13, 10, 21))
    np.savez_compressed(os.path.join(HERE, "project_uv.npz"), **proj_case())
    np.savez_compressed(os.path.join(HERE, "trafo.npz"), **trafo_case(12, 31))
    np.savez_compressed(os.path.join(HERE, "viewpoint.npz"), **viewpoint_case(7, 41))
    np.savez_compressed(os.path.join(HERE, "hand_mask.npz"), **hand_mask_case(5, 51))
    np.savez_compressed(os.path.join(HERE, "hand_mask_empty.npz"), **hand_mask_case(3, 52, True))
    np.savez_compressed(os.path.join(HERE, "match_switched.npz"), **match_case(6, 77, True))
    np.savez_compressed(os.path.join(HERE, "match_unswitched.npz"), **match_case(6, 78, False))

    # model checksum pin + real-pkl known answers (scalars only)
    chk = {k: float(np.asarray(v, dtype=np.float64).sum()) for k, v in model.items()}
    kat = {"synthetic_model_field_sums": chk}
    if os.path.isfile(ref_import.REAL_PKL):
        layer = ref.ManoLayer("cpu", ref_import.REAL_PKL, pose_num=45)
        z = lambda *s: torch.zeros(*s)
        v, j = layer(z(1, 3), z(1, 45), z(1, 10))
        kat["KAT-MANO-0"] = dict(verts_sum=v.sum().item(), joints_sum=j.sum().item(), verts_absmax=v.abs().max().item(),
                                 joint0=j[0, 0].tolist(), joint4=j[0, 4].tolist(), joint20=j[0, 20].tolist())
        g = torch.Generator().manual_seed(1234)
        rot = (torch.rand(4, 3, generator=g) - .5) * 2 * np.pi
        pose = (torch.rand(4, 45, generator=g) - .5) * np.pi
        beta = torch.rand(4, 10, generator=g) - .5
        v, j = layer(rot, pose, beta)
        kat["KAT-MANO-1"] = dict(verts_sum=v.sum().item(), joints_sum=j.sum().item(), joint3_0=j[3, 0].tolist(),
                                 joint3_8=j[3, 8].tolist(), joint3_17=j[3, 17].tolist())
        layer10 = ref.ManoLayer("cpu", ref_import.REAL_PKL, pose_num=10)
        v, j = layer10(rot, pose[:, :10], beta)
        kat["KAT-MANO-2"] = dict(verts_sum=v.sum().item(), joints_sum=j.sum().item())
    with open(os.path.join(HERE, "kat.json"), "w") as fh:
        json.dump(kat, fh, indent=1)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
