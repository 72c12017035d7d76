"""GPU (B200): the fused lane = vertex forward kernel (csrc/vskin.cu) — blend shapes and skinning as tcgen05 products
with M = vertices — against the fp64 oracle: its intermediates (blended 3x4 transforms, rest positions) through the
diagnostics entry point, its outputs at ragged batch sizes, the rest-pose scratch it leaves for the backward, and
agreement with the separate blend + skinning kernels."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import ROOT, assert_positions
from oracle import mano_oracle as mo

pytestmark = pytest.mark.gpu



def mano_inputs(B, nc, seed):
    rs = np.random.RandomState(seed)
    rot = ((rs.rand(B, 3) - .5) * 2 * np.pi).astype(np.float32)
    pose = ((rs.rand(B, nc) - .5) * np.pi).astype(np.float32)
    beta = (rs.rand(B, 10) - .5).astype(np.float32)
    return rot, pose, beta


def run_debug(pkg, layer, dev, rot, pose, beta, variant):
    import torch

    lib = pkg.load_library()
    B, nc = pose.shape
    t = [torch.from_numpy(a).to(dev) for a in (rot, pose, beta)]
    verts = torch.zeros(B, 778, 3, device=dev)
    joints = torch.zeros(B, 21, 3, device=dev)
    dbg = torch.zeros(4, 128, 16, device=dev)
    ws = torch.zeros(lib.mb_mano_workspace_bytes(B, layer._mode), dtype=torch.uint8, device=dev)
    rc = lib.mb_mano_forward_debug(layer._blob.data_ptr(), nc, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), B, layer._mode,
                                   verts.data_ptr(), joints.data_ptr(), ws.data_ptr(), ws.numel(), dbg.data_ptr(), variant,
                                   torch.cuda.current_stream(dev).cuda_stream)
    assert rc == 0, lib.mb_error_string(rc)
    torch.cuda.synchronize(dev)
    return verts.cpu().numpy(), joints.cpu().numpy(), dbg.cpu().numpy()


@pytest.mark.parametrize("B", [100, 64, 3])
def test_fused_forward_intermediates_and_outputs(pkg, synth_model, cuda_device, B):
    """Blended transforms T_v = sum_k w_vk A'_k and rest positions as the epilogue sees them (hands 0-3, vertices 0-127),
    then every vertex and joint of a batch that is not a multiple of the 64-hand tile."""
    nc = 45
    rot, pose, beta = mano_inputs(B, nc, seed=B)
    layer = pkg.ManoLayer(cuda_device, model=synth_model, pose_num=nc)
    verts, joints, dbg = run_debug(pkg, layer, cuda_device, rot, pose, beta, 0)
    ov, oj, cache = mo.mano_forward(synth_model, rot, pose, beta, return_cache=True)
    n = min(B, 4)
    Rp = np.einsum("bij,bkjl->bkil", cache["Rq"], cache["Rg"])                    # Rq Rg_k
    tp = np.einsum("bij,bkj->bki", cache["Rq"], cache["tA"])                      # Rq (tg_k - Rg_k J_k)
    A = np.concatenate([Rp, tp[..., None]], axis=-1)                              # [B,16,3,4]
    W = np.asarray(synth_model["weights"], np.float64)                            # [778,16]
    T_ref = np.einsum("vk,bkij->bvij", W[:128], A[:n]).reshape(n, 128, 12)
    err_T = np.abs(dbg[:n, :, :12] - T_ref).max()
    err_vp = np.abs(dbg[:n, :, 12:15] - cache["v_posed"][:n, :128]).max()
    err_v = np.abs(verts - ov).max()
    err_j = np.abs(joints - oj).max()
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, f"vskin_debug_B{B}.json"), "w") as fh:
        json.dump({"err_T": float(err_T), "err_vp": float(err_vp), "err_verts": float(err_v), "err_joints": float(err_j)}, fh)
    assert err_vp < 8e-8, err_vp                    # stated bound of the f16x3 contraction (measured 3.5e-8 .. 5.8e-8)
    assert err_T < 1e-6, err_T                      # |R| <= 1, |t| < 0.3 m: up to five truncating fp32 accumulations at magnitude 1 (measured 4.6e-7 .. 6.1e-7)
    assert_positions(verts, ov)
    assert_positions(joints, oj)


@pytest.mark.parametrize("products,bound", [(3, 2e-7), (4, 1e-7), (6, 1e-7)])
def test_fused_forward_split_products(pkg, synth_model, cuda_device, products, bound):
    """Number of fp16 split products of the transform contraction: 4 (default) and 6 are fp32-accurate, 3 is bounded."""
    B, nc = 130, 45
    rot, pose, beta = mano_inputs(B, nc, seed=7)
    layer = pkg.ManoLayer(cuda_device, model=synth_model, pose_num=nc)
    verts, joints, _ = run_debug(pkg, layer, cuda_device, rot, pose, beta, products << 4)
    ov, oj = mo.mano_forward(synth_model, rot, pose, beta)
    assert np.abs(verts - ov).max() < bound


@pytest.mark.parametrize("B,nc", [(8192, 45), (8195, 10), (9473, 45), (20001, 45)])
def test_fused_forward_through_the_layer(pkg, synth_model, cuda_device, B, nc):
    """``fused_forward=True`` runs the fused kernel from 8 192 hands on: outputs against the fp64 oracle (prefix, suffix, random
    sample), against the separate kernels, and — with gradients enabled — the backward on the scratch it leaves."""
    import torch

    rot, pose, beta = mano_inputs(B, nc, seed=B + nc)
    layer = pkg.ManoLayer(cuda_device, model=synth_model, pose_num=nc, fused_forward=True)
    unfused = pkg.ManoLayer(cuda_device, model=synth_model, pose_num=nc)
    t = [torch.from_numpy(a).to(cuda_device) for a in (rot, pose, beta)]
    with torch.no_grad():
        v, j = layer(*t)                                     # inference: no scratch kept
        v0, j0 = unfused(*t)
    idx = np.unique(np.r_[np.arange(160), np.arange(B - 70, B), np.random.RandomState(B).choice(B, 160, replace=False)])
    ov, oj = mo.mano_forward(synth_model, rot[idx], pose[idx], beta[idx])
    assert_positions(v.cpu().numpy()[idx], ov)
    assert_positions(j.cpu().numpy()[idx], oj)
    assert float((v - v0).abs().max()) < 1.5e-7 and float((j - j0).abs().max()) < 1.5e-7
    # training: the same values, and gradients through the saved rest-pose scratch
    tg = [x.clone().requires_grad_() for x in t]
    v2, j2 = layer(*tg)
    assert torch.equal(v2.detach(), v) and torch.equal(j2.detach(), j)
    rs = np.random.RandomState(1)
    gv = rs.randn(B, 778, 3).astype(np.float32)
    gj = rs.randn(B, 21, 3).astype(np.float32)
    ((v2 * torch.from_numpy(gv).to(cuda_device)).sum() + (j2 * torch.from_numpy(gj).to(cuda_device)).sum()).backward()
    sub = idx[:200]
    og = mo.mano_backward(synth_model, rot[sub], pose[sub], beta[sub], gv[sub], gj[sub])
    for x, want in zip(tg, og):
        got = x.grad.cpu().numpy()[sub]
        assert float(np.abs(got - want).max() / np.abs(want).max()) < 1e-4
    # the default layer (separate kernels; its backward starts from the scratch ITS forward wrote) agrees
    tu = [x.clone().requires_grad_() for x in t]
    v3, j3 = unfused(*tu)
    ((v3 * torch.from_numpy(gv).to(cuda_device)).sum() + (j3 * torch.from_numpy(gj).to(cuda_device)).sum()).backward()
    for a, b in zip(tg, tu):
        assert float((a.grad - b.grad).abs().max() / b.grad.abs().max()) < 1e-4


def test_fused_forward_is_batch_position_independent(pkg, synth_model, cuda_device):
    """A hand's result does not depend on where it sits in a 64-hand tile / 4-hand chunk (bit-exact)."""
    import torch

    B, nc = 8192 + 64, 45
    rot, pose, beta = mano_inputs(B, nc, seed=3)
    layer = pkg.ManoLayer(cuda_device, model=synth_model, pose_num=nc, fused_forward=True)
    t = [torch.from_numpy(a).to(cuda_device) for a in (rot, pose, beta)]
    with torch.no_grad():
        v, j = layer(*t)
        perm = torch.roll(torch.arange(B, device=cuda_device), 37)
        v2, j2 = layer(*[x[perm].contiguous() for x in t])
    assert torch.equal(v[perm], v2) and torch.equal(j[perm], j2)
