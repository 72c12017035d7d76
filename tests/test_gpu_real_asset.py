"""GPU (B200): the REAL MANO_RIGHT.pkl and the LIVE, unmodified reference on the CUDA path.

``oracle/_ref/`` is the byte-for-byte travelling copy of the reference's files of the path (+ the pkl) that
``oracle/make_ref.py`` stages in the authoring container (git-ignored, shipped by gpurun).  These tests skip only
when it is absent.  They check what synthetic-model tests cannot: the real asset's skin program (386 entries, its own
slot schedule and sparsity), its PCA basis and hands_mean, and elementwise agreement with the reference's own
PyTorch forward and autograd at pose_num = 45 and 10 in both pose-kernel regimes (one warp per hand below 8 192 hands,
one thread per hand from there on).
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, assert_positions
from oracle import mano_oracle as mo
from oracle import ref_import

pytestmark = pytest.mark.gpu

POS_TOL_REF = 2e-7       # vs the reference's fp32 outputs (its own fp32-vs-fp64 noise is ~1e-7 m)
GRAD_TOL = 1e-4


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.fixture(scope="module")
def ref():
    if not (ref_import.available() and ref_import.real_pkl_available()):
        pytest.skip("oracle/_ref (staged reference + MANO_RIGHT.pkl) is not present")
    return ref_import.load()


@pytest.fixture(scope="module")
def real_model(pkg, ref):
    return pkg.assets.read_mano_pkl(ref_import.REAL_PKL)


def kat_inputs():
    import torch

    g = torch.Generator().manual_seed(1234)
    rot = (torch.rand(4, 3, generator=g) - .5) * 2 * np.pi
    pose = (torch.rand(4, 45, generator=g) - .5) * np.pi
    beta = torch.rand(4, 10, generator=g) - .5
    return rot, pose, beta


def test_real_pkl_known_answers_on_gpu(pkg, ref, cuda_device):
    """KAT-MANO-0/1/2 (tests/golden/kat.json, recorded from the reference with the real pkl) through the CUDA path."""
    import torch

    kat = json.load(open(os.path.join(GOLDEN, "kat.json")))
    layer = pkg.ManoLayer(cuda_device, ref_import.REAL_PKL, pose_num=45)
    z = lambda *s: torch.zeros(*s, device=cuda_device)
    v, j = layer(z(1, 3), z(1, 45), z(1, 10))
    k0 = kat["KAT-MANO-0"]
    assert v.sum().item() == pytest.approx(k0["verts_sum"], abs=3e-5)
    assert j.sum().item() == pytest.approx(k0["joints_sum"], abs=2e-6)
    assert v.abs().max().item() == pytest.approx(k0["verts_absmax"], abs=POS_TOL_REF)
    for slot, key in ((0, "joint0"), (4, "joint4"), (20, "joint20")):
        assert np.abs(j[0, slot].cpu().numpy() - np.array(k0[key])).max() < POS_TOL_REF
    assert torch.equal(j[0, 4], v[0, 333]) and torch.equal(j[0, 20], v[0, 745])     # tips are vertices

    rot, pose, beta = kat_inputs()
    v, j = layer(rot.to(cuda_device), pose.to(cuda_device), beta.to(cuda_device))
    k1 = kat["KAT-MANO-1"]
    assert v.sum().item() == pytest.approx(k1["verts_sum"], abs=6e-5)
    assert j.sum().item() == pytest.approx(k1["joints_sum"], abs=3e-6)
    for slot, key in ((0, "joint3_0"), (8, "joint3_8"), (17, "joint3_17")):
        assert np.abs(j[3, slot].cpu().numpy() - np.array(k1[key])).max() < POS_TOL_REF

    layer10 = pkg.ManoLayer(cuda_device, ref_import.REAL_PKL, pose_num=10)
    v, j = layer10(rot.to(cuda_device), pose[:, :10].contiguous().to(cuda_device), beta.to(cuda_device))
    k2 = kat["KAT-MANO-2"]
    assert v.sum().item() == pytest.approx(k2["verts_sum"], abs=6e-5)
    assert j.sum().item() == pytest.approx(k2["joints_sum"], abs=3e-6)


def _reference_fwd_bwd(ref, nc, rot, pose, beta, gv, gj):
    """The reference's own ManoLayer forward + autograd on CPU, in chunks of 512 hands (it materialises ~1.8 MB per hand)."""
    import torch

    layer = ref.ManoLayer("cpu", ref_import.REAL_PKL, pose_num=nc)
    outs = []
    for s in range(0, rot.shape[0], 512):
        sl = slice(s, s + 512)
        t = [torch.from_numpy(a[sl]).clone().requires_grad_() for a in (rot, pose, beta)]
        v, j = layer(*t)
        ((v * torch.from_numpy(gv[sl])).sum() + (j * torch.from_numpy(gj[sl])).sum()).backward()
        outs.append((v.detach().numpy().copy(), j.detach().numpy().copy(), *[x.grad.numpy().copy() for x in t]))
    return [np.concatenate(c, 0) for c in zip(*outs)]


@pytest.mark.parametrize("B,nc", [(64, 45), (64, 10), (700, 45), (8200, 45), (8200, 10)])
def test_real_pkl_matches_live_reference_forward_and_autograd(pkg, ref, real_model, cuda_device, B, nc):
    """Elementwise against the unmodified reference (PyTorch CPU fp32, its own autograd) and the fp64 oracle, real pkl.
    B = 64 is BASELINE config 1's batch; 8 200 hands run the one-thread-per-hand kernels and the hand-tile-resident
    blend forward (the reference is evaluated on a 320-hand subset of that batch)."""
    import torch

    rs = np.random.RandomState(B + nc)
    rot = ((rs.rand(B, 3) - .5) * 2 * np.pi).astype(np.float32)
    pose = ((rs.rand(B, nc) - .5) * np.pi).astype(np.float32)
    beta = (rs.rand(B, 10) - .5).astype(np.float32)
    gv = rs.randn(B, 778, 3).astype(np.float32)
    gj = rs.randn(B, 21, 3).astype(np.float32)
    layer = pkg.ManoLayer(cuda_device, ref_import.REAL_PKL, pose_num=nc)
    t = [torch.from_numpy(a).to(cuda_device).requires_grad_() for a in (rot, pose, beta)]
    v, j = layer(*t)
    ((v * torch.from_numpy(gv).to(cuda_device)).sum() + (j * torch.from_numpy(gj).to(cuda_device)).sum()).backward()
    idx = np.arange(B) if B <= 700 else np.unique(np.r_[np.arange(128), np.arange(B - 64, B),
                                                        np.random.RandomState(1).choice(B, 128, replace=False)])
    rv, rj, rg_rot, rg_pose, rg_beta = _reference_fwd_bwd(ref, nc, rot[idx], pose[idx], beta[idx], gv[idx], gj[idx])
    got_v, got_j = v.detach().cpu().numpy()[idx], j.detach().cpu().numpy()[idx]
    assert np.abs(got_v - rv).max() < POS_TOL_REF
    assert np.abs(got_j - rj).max() < POS_TOL_REF
    for x, want in zip(t, (rg_rot, rg_pose, rg_beta)):
        assert rel(x.grad.cpu().numpy()[idx], want) < GRAD_TOL
    # and the fp64 arbiter on the same real model
    sub = idx[:96]
    ov, oj = mo.mano_forward(real_model, rot[sub], pose[sub], beta[sub])
    assert_positions(v.detach().cpu().numpy()[sub], ov)
    assert_positions(j.detach().cpu().numpy()[sub], oj)


def test_real_pkl_joints_only_head_path(pkg, ref, cuda_device):
    """The heads' path (resnet50MANO.py:76,87: vertices discarded) at the reference's batch size (config.py:79,
    batch_size = 200) and value ranges (resnet50MANO.py:73-75), pose_num = config.mano_pose_num = 10."""
    import torch

    B, nc = 200, 10
    rs = np.random.RandomState(5)
    sig = lambda *s: rs.rand(*s).astype(np.float32)
    rot = (sig(B, 3) - .5) * 2 * np.pi
    pose = (sig(B, nc) - .5) * 4
    beta = (sig(B, 10) - .5) * .1
    gj = rs.randn(B, 21, 3).astype(np.float32)
    layer = pkg.ManoLayer(cuda_device, ref_import.REAL_PKL, pose_num=nc)
    t = [torch.from_numpy(a).to(cuda_device).requires_grad_() for a in (rot, pose, beta)]
    _, j = layer.rot_pose_beta_to_mesh(*t, joints_only=True)
    (j * torch.from_numpy(gj).to(cuda_device)).sum().backward()
    rl = ref.ManoLayer("cpu", ref_import.REAL_PKL, pose_num=nc)
    rt = [torch.from_numpy(a).clone().requires_grad_() for a in (rot, pose, beta)]
    _, rj = rl(*rt)
    (rj * torch.from_numpy(gj)).sum().backward()
    assert np.abs(j.detach().cpu().numpy() - rj.detach().numpy()).max() < POS_TOL_REF
    for x, want in zip(t, rt):
        assert rel(x.grad.cpu().numpy(), want.grad.numpy()) < GRAD_TOL


def test_reference_loss_calculation_runs_on_swapped_classes(pkg, ref, cuda_device):
    """After ``install_into_reference()`` the reference's own ``LossCalculation`` (criterions/loss.py:62-153) — which
    constructs ``L2Loss`` and sends BOTH [B,21,3] xyz and [B,21,2] uv through it (:83-87), plus the hand-mask and
    regularisation terms — runs on CUDA tensors through the kernels and matches the untouched reference on CPU."""
    import torch
    import criterions.loss as CL

    B = 37
    g = torch.Generator().manual_seed(3)
    pre_xyz = torch.randn(B, 21, 3, generator=g)
    gt_xyz = torch.randn(B, 21, 3, generator=g)
    pre_uv = torch.rand(B, 21, 2, generator=g) * 140 - 6
    gt_uv = torch.rand(B, 21, 2, generator=g) * 128
    vis = (torch.rand(B, 21, 1, generator=g) < .8).float()
    mask = (torch.rand(B, 128, 128, generator=g) < .4).float()
    theta = torch.randn(B, 10, generator=g)
    beta = torch.randn(B, 10, generator=g) * .1
    kw = dict(comp_xyz_loss=True, comp_uv_loss=True, comp_hand_mask_loss=True, comp_regularization_loss=True)

    def run(dev):
        calc = CL.LossCalculation(device=dev, **kw)
        a = [x.to(dev).clone().requires_grad_() for x in (pre_xyz, pre_uv, theta, beta)]
        out = calc(a[0], gt_xyz.to(dev), a[1], gt_uv.to(dev), vis.to(dev), hand_mask=mask.to(dev), theta=a[2], beta=a[3])
        (out[0] + out[1] / 1e3 + out[4]).backward()
        return [float(o) for o in (out[0], out[1], out[3], out[4])], [x.grad.cpu().numpy() for x in a]

    want, want_g = run("cpu")
    try:
        done = pkg.install_into_reference()
        assert any(d.endswith("LossCalculation.compute_regularization_loss") for d in done)
        assert CL.L2Loss is pkg.L2Loss
        got, got_g = run(cuda_device)
    finally:
        pkg.dropin.uninstall()
    for a, b in zip(got, want):
        assert a == pytest.approx(b, rel=2e-5, abs=1e-7)
    for a, b in zip(got_g, want_g):
        assert rel(a, b) < 1e-5
