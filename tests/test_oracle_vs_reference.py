"""CPU, authoring container only: the oracle against the LIVE unmodified reference
(imported from /root/reference).  Skipped where the reference is absent (GPU box)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import fk_oracle as fo
from oracle import mano_oracle as mo
from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    import torch

    torch.set_num_threads(2)
    return ref_import.load()


@pytest.mark.parametrize("nc", [45, 10])
def test_real_pkl_forward_backward(pkg, ref, nc):
    import torch

    model = pkg.assets.read_mano_pkl(ref_import.REAL_PKL)
    layer = ref.ManoLayer("cpu", ref_import.REAL_PKL, pose_num=nc)
    g = torch.Generator().manual_seed(1234)
    B = 4
    rot = ((torch.rand(B, 3, generator=g) - .5) * 2 * np.pi).requires_grad_()
    pose = ((torch.rand(B, 45, generator=g) - .5) * np.pi)[:, :nc].clone().requires_grad_()
    beta = (torch.rand(B, 10, generator=g) - .5).requires_grad_()
    v, j = layer(rot, pose, beta)
    kat = json.load(open(os.path.join(GOLDEN, "kat.json")))
    key = "KAT-MANO-1" if nc == 45 else "KAT-MANO-2"
    assert v.sum().item() == pytest.approx(kat[key]["verts_sum"], abs=1e-4)
    gv, gj = torch.randn(v.shape, generator=g), torch.randn(j.shape, generator=g)
    ((v * gv).sum() + (j * gj).sum()).backward()
    n = lambda t: t.detach().numpy()
    ov, oj = mo.mano_forward(model, n(rot), n(pose), n(beta))
    assert np.abs(ov - n(v)).max() < 2e-7 and np.abs(oj - n(j)).max() < 2e-7
    gr, gp, gb = mo.mano_backward(model, n(rot), n(pose), n(beta), n(gv), n(gj))
    for got, want in ((gr, rot.grad), (gp, pose.grad), (gb, beta.grad)):
        assert np.abs(got - n(want)).max() / np.abs(n(want)).max() < 1e-4


def test_real_pkl_zero_pose_kat(pkg):
    model = pkg.assets.read_mano_pkl(ref_import.REAL_PKL)
    v, j = mo.mano_forward(model, np.zeros((1, 3)), np.zeros((1, 45)), np.zeros((1, 10)))
    kat = json.load(open(os.path.join(GOLDEN, "kat.json")))["KAT-MANO-0"]
    assert v.sum() == pytest.approx(kat["verts_sum"], abs=2e-5)
    assert np.abs(j[0, 0] - kat["joint0"]).max() < 2e-7
    assert np.abs(j[0, 4] - kat["joint4"]).max() < 2e-7 and np.abs(v[0, 333] - kat["joint4"]).max() < 2e-7
    assert np.abs(j[0, 20] - kat["joint20"]).max() < 2e-7


def test_pkl_reader_equals_reference_constants(pkg, ref):
    model = pkg.assets.read_mano_pkl(ref_import.REAL_PKL)
    layer = ref.ManoLayer("cpu", ref_import.REAL_PKL, pose_num=45)
    assert np.array_equal(model["shapedirs"].astype(np.float32), layer.mesh_pca[0].numpy())
    assert np.array_equal(model["posedirs"].astype(np.float32), layer.posedirs[0].numpy())
    assert np.array_equal(model["weights"].astype(np.float32), layer.weights[0].numpy())
    assert np.array_equal(model["J_regressor"].astype(np.float32), layer.J_regressor[0].numpy())
    assert layer.parent == {i: int(p) for i, p in enumerate(pkg.assets.parents_from_kintree(model["kintree_table"])) if i}


def test_fk_live(ref):
    import torch

    fk = ref.ForwardKinematics("cpu")
    g = torch.Generator().manual_seed(77)
    B = 6
    ra = (torch.rand(B, 3, generator=g) - .5) * 2 * np.pi
    oa = (torch.rand(B, 23, generator=g) - .5) * np.pi
    bl = torch.rand(B, 20, generator=g) + .1
    K = torch.tensor([[282.9, 0, 160], [0, 282.9, 160], [0, 0, 1.]]).repeat(B, 1, 1)
    sc = torch.rand(B, 1, generator=g) * .05 + .02
    root = torch.randn(B, 3, generator=g) * .05 + torch.tensor([0, 0, .6])
    for sw in (True, False):
        ref.config.joint_order_switched = sw
        xyz, uv, _ = fk(ra, oa, bl, K, sc, root)
        oxyz, ouv = fo.fk_forward(*[t.numpy() for t in (ra, oa, bl, K, sc, root)], joint_order_switched=sw)
        assert np.abs(oxyz - xyz.numpy()).max() < 2e-7
        assert np.abs(ouv - uv.numpy()).max() < 1e-3
    ref.config.joint_order_switched = True


def test_match_mano_to_rhd_live(ref):
    """Both copies of match_mano_to_RHD in the reference (the two MANO heads) against the oracle,
    including the in-place permutation of the argument that the product does not reproduce."""
    import torch
    from network.MANO3DHandPose import MANO3DHandPose
    from network.Resnet50MANO3DHandPose import Resnet50MANO3DHandPose

    g = torch.Generator().manual_seed(5)
    B = 9
    j = torch.randn(B, 21, 3, generator=g) * .04
    L = torch.rand(B, 1, generator=g) * .05 + .02
    root = torch.randn(B, 3, generator=g) * .05 + torch.tensor([0, 0, .6])
    try:
        for head in (MANO3DHandPose, Resnet50MANO3DHandPose):
            for sw in (True, False):
                ref.config.joint_order_switched = sw
                arg = j.clone()
                reln, xyz = head.match_mano_to_RHD(None, arg, L, root)
                orel, oxyz = fo.match_mano_to_rhd(j.numpy(), L.numpy(), root.numpy(), sw)
                assert np.abs(orel - reln.numpy()).max() < 2e-6 * np.abs(orel).max()
                assert np.abs(oxyz - xyz.numpy()).max() < 2e-7
                assert torch.equal(arg, j) == sw                   # permuted in place when not switched
    finally:
        ref.config.joint_order_switched = True


def test_keypoint_trafos_live(ref):
    """utils/relative_trafo.py and utils/canonical_trafo.py run live against the oracle restatement."""
    import torch
    from oracle import trafo_oracle as tro
    from utils.canonical_trafo import canonical_trafo, flip_right_hand
    from utils.relative_trafo import bone_rel_trafo, bone_rel_trafo_inv

    g = torch.Generator().manual_seed(8)
    B = 10
    xyz = torch.randn(B, 21, 3, generator=g) * .5 + torch.arange(21)[:, None] * torch.tensor([.09, .06, .03])
    xyz[:, 0] = 0
    rel = bone_rel_trafo(xyz)
    assert np.abs(tro.bone_rel_trafo(xyz.numpy()) - rel.numpy()).max() < 5e-6
    assert np.abs(tro.bone_rel_trafo_inv(rel.numpy()) - bone_rel_trafo_inv(rel).numpy()).max() < 5e-6
    can, rot = canonical_trafo(xyz)
    ocan, orot = tro.canonical_trafo(xyz.numpy())
    assert np.abs(ocan - can.numpy()).max() < 1e-5 and np.abs(orot - rot.numpy()).max() < 1e-5
    cond = torch.rand(B, generator=g) < .5
    assert np.array_equal(tro.flip_right_hand(can.numpy(), cond.numpy()),
                          flip_right_hand(can, cond[:, None].expand(B, 21)).numpy())
    # single-sample form of the dataloaders ([21,3] in, leading batch dimension of 1 out)
    assert np.abs(tro.bone_rel_trafo(xyz[0].numpy())[0] - bone_rel_trafo(xyz[0]).numpy()[0]).max() < 5e-6


def test_viewpoint_live(ref):
    """utils/general.py _get_rot_mat and the matmul of network/Hand3DPoseNet.py:41-43 live against the oracle."""
    import torch
    from oracle import trafo_oracle as tro
    from utils.general import _get_rot_mat

    g = torch.Generator().manual_seed(9)
    B = 16
    can = torch.randn(B, 21, 3, generator=g)
    u = [(torch.rand(B, 1, generator=g) - .5) * 6 for _ in range(3)]
    R = _get_rot_mat(*u)
    oR, orel = tro.viewpoint_forward(can.numpy(), *[t.numpy() for t in u])
    assert np.abs(oR - R.numpy()).max() < 1e-6
    assert np.abs(orel - torch.matmul(can, R).numpy()).max() < 2e-6


def test_dropin_rebinds_the_reference_in_place(ref):
    """dropin.install_into_reference(): the reference's own modules — and the heads that imported the symbols by name —
    end up holding the B200 classes / functions; uninstall() restores the originals."""
    import importlib
    import sys

    pkg = importlib.import_module("3dhandposeestimation_b200")
    import network.sub_modules.MANOLayer as ML
    import network.sub_modules.forwardKinematicsLayer as FKL
    import network.sub_modules.resnetMANO as RM            # holds `ManoLayer` by `from ... import` (resnetMANO.py)
    import utils.general as UG
    import utils.relative_trafo as RT
    import criterions.loss as CL
    import criterions.metrics as CM
    orig = (ML.ManoLayer, FKL.ForwardKinematics, UG._get_rot_mat, RT.bone_rel_trafo, CL.L2Loss, CM.MPJPE)
    assert RM.ManoLayer is ML.ManoLayer
    try:
        reg0 = CL.LossCalculation.__dict__["compute_regularization_loss"]
        done = pkg.install_into_reference()
        assert ML.ManoLayer is pkg.ManoLayer and RM.ManoLayer is pkg.ManoLayer
        assert FKL.ForwardKinematics is pkg.ForwardKinematics and UG._get_rot_mat is pkg._get_rot_mat
        assert CL.L2Loss is pkg.L2Loss and CM.MPJPE is pkg.MPJPE
        # the per-sample CPU helpers of the reference's Dataset.__getitem__ are NOT swapped by default (the B200
        # versions are batched GPU-side functions and would break the DataLoader workers) ...
        assert RT.bone_rel_trafo is orig[3]
        # ... the two LossCalculation methods are (the rest of the class is untouched)
        assert CL.LossCalculation.__dict__["compute_regularization_loss"].__mb_replacement__ is pkg.compute_regularization_loss
        assert CL.LossCalculation.__dict__["compute_hand_mask_loss"].__mb_replacement__ is pkg.compute_hand_mask_loss
        assert "network.sub_modules.MANOLayer.ManoLayer" in done
        assert "criterions.loss.LossCalculation.compute_regularization_loss" in done
        assert pkg.install_into_reference() == []             # idempotent
        more = pkg.install_into_reference(dataloader=True)    # opt-in
        assert RT.bone_rel_trafo is pkg.bone_rel_trafo and "utils.relative_trafo.bone_rel_trafo" in more
    finally:
        assert pkg.dropin.uninstall() > 0
    assert CL.LossCalculation.__dict__["compute_regularization_loss"] is reg0
    assert (ML.ManoLayer, FKL.ForwardKinematics, UG._get_rot_mat, RT.bone_rel_trafo, CL.L2Loss, CM.MPJPE) == orig
    assert RM.ManoLayer is orig[0]
    # the oracle helper's handle on the live reference was never touched
    assert ref.ManoLayer is orig[0]
