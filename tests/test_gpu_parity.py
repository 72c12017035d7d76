"""GPU (B200): the CUDA path, called through the nn.Module wrappers and the C ABI, against the
golden vectors of the unmodified reference and against the numpy oracle on seeded inputs.

Tolerances (SURVEY A.4 / BASELINE north_star):
  verts / joints       : 2e-7 m absolute vs the fp32 reference goldens (the reference's own fp32 noise is ~1e-7 m),
                         vs the fp64 oracle (conftest.assert_positions): 1e-7 m (1e-4 mm, north_star) at the 99.999th
                         percentile of the coordinates and 2e-7 m on the worst one (measured worst: 1.0e-7 .. 1.55e-7 m;
                         the reference's own fp32 path: 6e-8 .. 9e-8 m)
  FK xyz               : 2e-7 m vs the fp64 oracle (coordinates ~0.6 m from the camera: 1 ulp = 6e-8 m)
  uv                   : 1e-3 px for |z| >= 0.1 m
  gradients            : 1e-4 relative to the tensor's max-abs
  MPJPE / L2           : 1e-5 relative
"""
import importlib

import numpy as np
import pytest

from conftest import assert_positions, load_golden
from oracle import fk_oracle as fo
from oracle import mano_oracle as mo

pytestmark = pytest.mark.gpu

POS_TOL_REF = 2e-7
# against the reference's fp32 goldens both sides carry fp32 noise: 2e-7 m; against the fp64 arbiter: conftest.assert_positions
POS_TOL_FK = 2e-7
GRAD_TOL = 1e-4
ACCURATE_MODES = ["fp32", "f16x3"]
FAST_TOL = 5e-5        # MB_MODE_F16: one fp16 product, stated bound for the blend contraction


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def to_dev(dev, *arrs, grad=False):
    import torch

    out = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in arrs]
    if grad:
        out = [t.requires_grad_() for t in out]
    return out


def mano_inputs(B, nc, seed, pose_scale=np.pi):
    rs = np.random.RandomState(seed)
    rot = ((rs.rand(B, 3) - .5) * 2 * np.pi).astype(np.float32)
    pose = ((rs.rand(B, nc) - .5) * pose_scale).astype(np.float32)
    beta = (rs.rand(B, 10) - .5).astype(np.float32)
    return rot, pose, beta


@pytest.mark.parametrize("mode", ACCURATE_MODES)
@pytest.mark.parametrize("name,nc", [("mano_synth_nc45.npz", 45), ("mano_synth_nc10.npz", 10), ("mano_synth_nc6.npz", 6)])
def test_mano_matches_reference_golden(pkg, synth_model, cuda_device, name, nc, mode):
    import torch

    g = load_golden(name)
    layer = pkg.ManoLayer(cuda_device, model=synth_model, pose_num=nc, mode=mode)
    rot, pose, beta = to_dev(cuda_device, g["rot"], g["pose"], g["beta"], grad=True)
    gv, gj = to_dev(cuda_device, g["g_verts"], g["g_joints"])
    verts, joints = layer(rot, pose, beta)
    assert verts.shape == (rot.shape[0], 778, 3) and joints.shape == (rot.shape[0], 21, 3)
    assert verts.is_contiguous() and joints.is_contiguous()
    assert np.abs(verts.detach().cpu().numpy() - g["verts"]).max() < POS_TOL_REF
    assert np.abs(joints.detach().cpu().numpy() - g["joints"]).max() < POS_TOL_REF
    ((verts * gv).sum() + (joints * gj).sum()).backward()
    assert rel(rot.grad.cpu().numpy(), g["g_rot"]) < GRAD_TOL
    assert rel(pose.grad.cpu().numpy(), g["g_pose"]) < GRAD_TOL
    assert rel(beta.grad.cpu().numpy(), g["g_beta"]) < GRAD_TOL
    # the heads' case: only the joints are consumed -> joints-only backward kernel
    for t in (rot, pose, beta):
        t.grad = None
    _, joints2 = layer(rot, pose, beta)
    (joints2 * gj).sum().backward()
    assert rel(rot.grad.cpu().numpy(), g["gj_rot"]) < GRAD_TOL
    assert rel(pose.grad.cpu().numpy(), g["gj_pose"]) < GRAD_TOL
    assert rel(beta.grad.cpu().numpy(), g["gj_beta"]) < GRAD_TOL
    # joints-only forward (extension) gives the same joints without the vertex contraction
    none, joints3 = layer.rot_pose_beta_to_mesh(rot.detach(), pose.detach(), beta.detach(), joints_only=True)
    assert none is None
    assert np.abs(joints3.cpu().numpy() - g["joints"]).max() < POS_TOL_REF
    torch.cuda.synchronize()


def test_mano_transl_scale_keywords_match_reference_golden(pkg, synth_model, cuda_device):
    """``ManoLayer.forward(..., transl=, scale=)`` (keyword-only extension; north_star "global rotation and translation
    in") against the reference layer followed by its callers' post-ops (resnet50MANO.py:77-81) with the reference's
    autograd for all five inputs — full layer and the joints-only path; the default call is unchanged."""
    import torch

    g = load_golden("mano_affine_nc10.npz")
    layer = pkg.ManoLayer(cuda_device, model=synth_model, pose_num=10)
    rot, pose, beta, transl, scale = to_dev(cuda_device, g["rot"], g["pose"], g["beta"], g["transl"], g["scale"], grad=True)
    gv, gj = to_dev(cuda_device, g["g_verts"], g["g_joints"])
    v, j = layer(rot, pose, beta, transl=transl, scale=scale)
    assert np.abs(v.detach().cpu().numpy() - g["verts"]).max() < 3e-7        # coordinates ~0.7 m: 1 ulp = 6e-8
    assert np.abs(j.detach().cpu().numpy() - g["joints"]).max() < 3e-7
    ((v * gv).sum() + (j * gj).sum()).backward()
    for t, key in ((rot, "g_rot"), (pose, "g_pose"), (beta, "g_beta"), (transl, "g_transl"), (scale, "g_scale")):
        assert rel(t.grad.cpu().numpy(), g[key]) < GRAD_TOL, key
    # joints only, translation only
    for t in (rot, pose, beta, transl, scale):
        t.grad = None
    _, j2 = layer.rot_pose_beta_to_mesh(rot, pose, beta, joints_only=True, transl=transl)
    want = (g["joints"] - g["transl"][:, None]) / g["scale"][:, None, None] + g["transl"][:, None]
    assert np.abs(j2.detach().cpu().numpy() - want).max() < 3e-7
    (j2 * gj).sum().backward()
    assert rel(transl.grad.cpu().numpy(), g["g_joints"].sum(1)) < 1e-5 and scale.grad is None
    # positional signature untouched
    v0, j0 = layer(rot.detach(), pose.detach(), beta.detach())
    assert np.abs(v0.cpu().numpy() * g["scale"][:, None, None] + g["transl"][:, None] - g["verts"]).max() < 3e-7


@pytest.mark.parametrize("mode", ACCURATE_MODES)
@pytest.mark.parametrize("B,nc", [(1, 45), (2, 45), (7, 10), (129, 45), (1000, 45), (1344, 45), (2720, 10), (4096, 10), (4133, 45), (5000, 6), (8192, 10), (9001, 45),
                                  (20001, 45), (40001, 45), (77777, 10)])
def test_mano_matches_fp64_oracle(pkg, synth_model, cuda_device, B, nc, mode):
    """Ragged sizes (partial hand groups, partial tcgen05 tiles) up to BASELINE config 2 (B=4096,
    nc=10, the Resnet50MANO3DHandPose head workload).  B >= 8192 in the tensor-core modes runs the
    one-thread-per-hand pose kernels, below that the one-warp-per-hand kernels; from 8 192 hands (64 hand
    tiles) the blend forward is the hand-tile-resident kernel (several hand tiles per CTA from 18 945)."""
    rot, pose, beta = mano_inputs(B, nc, seed=B + nc)
    layer = pkg.ManoLayer(cuda_device, model=synth_model, pose_num=nc, mode=mode)
    trot, tpose, tbeta = to_dev(cuda_device, rot, pose, beta, grad=True)
    verts, joints = layer(trot, tpose, tbeta)
    nchk = min(B, 256)                                         # oracle on a bounded prefix + suffix
    idx = np.unique(np.r_[np.arange(nchk), np.arange(B - min(B, 64), B),
                          np.random.RandomState(B).choice(B, min(B, 192), replace=False)])     # + a sample of every region
    ov, oj = mo.mano_forward(synth_model, rot[idx], pose[idx], beta[idx])
    assert_positions(verts.detach().cpu().numpy()[idx], ov)
    assert_positions(joints.detach().cpu().numpy()[idx], oj)
    rs = np.random.RandomState(1)
    gv = rs.randn(B, 778, 3).astype(np.float32)
    gj = rs.randn(B, 21, 3).astype(np.float32)
    tgv, tgj = to_dev(cuda_device, gv, gj)
    ((verts * tgv).sum() + (joints * tgj).sum()).backward()
    og = mo.mano_backward(synth_model, rot[idx], pose[idx], beta[idx], gv[idx], gj[idx])
    for t, want in zip((trot, tpose, tbeta), og):
        assert rel(t.grad.cpu().numpy()[idx], want) < GRAD_TOL


@pytest.mark.parametrize("B,nc", [(4099, 45), (8195, 45), (12000, 10)])
def test_mano_joints_only_large_batch_matches_fp64_oracle(pkg, synth_model, cuda_device, B, nc):
    """The heads' / fitting loop's case at a batch that runs the one-thread-per-hand joints-only kernels
    (>= 8192 hands): 21 joints and their gradients without the 778-vertex contraction."""
    import torch

    rot, pose, beta = mano_inputs(B, nc, seed=B + nc)
    layer = pkg.ManoLayer(cuda_device, model=synth_model, pose_num=nc)
    trot, tpose, tbeta = to_dev(cuda_device, rot, pose, beta, grad=True)
    none, joints = layer.rot_pose_beta_to_mesh(trot, tpose, tbeta, joints_only=True)
    assert none is None
    idx = np.unique(np.r_[np.arange(96), np.arange(B - 40, B)])
    _, oj = mo.mano_forward(synth_model, rot[idx], pose[idx], beta[idx])
    assert_positions(joints.detach().cpu().numpy()[idx], oj)
    # the full layer gives the same joints
    _, joints_full = layer(trot.detach(), tpose.detach(), tbeta.detach())
    assert float((joints_full - joints.detach()).abs().max()) < 2e-7
    gj = np.random.RandomState(2).randn(B, 21, 3).astype(np.float32)
    (joints * torch.from_numpy(gj).to(cuda_device)).sum().backward()
    og = mo.mano_backward(synth_model, rot[idx], pose[idx], beta[idx], np.zeros((idx.size, 778, 3), np.float32), gj[idx])
    for t, want in zip((trot, tpose, tbeta), og):
        assert rel(t.grad.cpu().numpy()[idx], want) < GRAD_TOL


@pytest.mark.parametrize("B,nc", [(8200, 45), (9000, 20)])
def test_mano_identity_pca_model_matches_fp64_oracle(pkg, synth_model, cuda_device, B, nc):
    """BASELINE config 4 ("no PCA": hands_components = I, the coefficients are the axis-angles): the one-thread-per-hand
    kernels detect the identity and skip the 45 x 45 products — full layer, joints-only path and gradients; with a
    non-zero mean and nc < 45 the remaining angles stay at the mean."""
    import torch

    model = dict(synth_model)
    model["hands_components"] = np.eye(45)
    rot, pose, beta = mano_inputs(B, nc, seed=B, pose_scale=np.pi / 2)
    layer = pkg.ManoLayer(cuda_device, model=model, pose_num=nc)
    t = to_dev(cuda_device, rot, pose, beta, grad=True)
    verts, joints = layer(*t)
    idx = np.unique(np.r_[np.arange(64), np.arange(B - 40, B)])
    ov, oj = mo.mano_forward(model, rot[idx], pose[idx], beta[idx])
    assert_positions(verts.detach().cpu().numpy()[idx], ov)
    assert_positions(joints.detach().cpu().numpy()[idx], oj)
    rs = np.random.RandomState(4)
    gv = rs.randn(B, 778, 3).astype(np.float32)
    gj = rs.randn(B, 21, 3).astype(np.float32)
    tgv, tgj = to_dev(cuda_device, gv, gj)
    ((verts * tgv).sum() + (joints * tgj).sum()).backward()
    og = mo.mano_backward(model, rot[idx], pose[idx], beta[idx], gv[idx], gj[idx])
    for x, want in zip(t, og):
        assert rel(x.grad.cpu().numpy()[idx], want) < GRAD_TOL
    t2 = to_dev(cuda_device, rot, pose, beta, grad=True)
    _, j2 = layer.rot_pose_beta_to_mesh(*t2, joints_only=True)
    assert float((j2.detach() - joints.detach()).abs().max()) < 2e-7
    (j2 * tgj).sum().backward()
    og2 = mo.mano_backward(model, rot[idx], pose[idx], beta[idx], np.zeros((idx.size, 778, 3), np.float32), gj[idx])
    for x, want in zip(t2, og2):
        assert rel(x.grad.cpu().numpy()[idx], want) < GRAD_TOL


def test_mano_fast_f16_mode_error_is_bounded(pkg, synth_model, cuda_device):
    """MB_MODE_F16 (single fp16 product on the tensor cores): bounded, stated error."""
    rot, pose, beta = mano_inputs(512, 45, seed=3)
    layer = pkg.ManoLayer(cuda_device, model=synth_model, pose_num=45, mode="f16")
    v, j = layer(*to_dev(cuda_device, rot, pose, beta))
    ov, oj = mo.mano_forward(synth_model, rot, pose, beta)
    err = np.abs(v.cpu().numpy() - ov).max()
    assert err < FAST_TOL, err
    assert np.abs(j.cpu().numpy() - oj).max() < FAST_TOL


def test_mano_backward_recompute_equals_saved_workspace(pkg, synth_model, cuda_device):
    rot, pose, beta = mano_inputs(37, 45, seed=5)
    rs = np.random.RandomState(2)
    gv, gj = to_dev(cuda_device, rs.randn(37, 778, 3).astype(np.float32), rs.randn(37, 21, 3).astype(np.float32))
    grads = []
    for keep in (True, False):
        layer = pkg.ManoLayer(cuda_device, model=synth_model, pose_num=45, keep_workspace=keep)
        t = to_dev(cuda_device, rot, pose, beta, grad=True)
        v, j = layer(*t)
        loss = (v * gv).sum() + (j * gj).sum()
        loss.backward(retain_graph=keep)
        first = [x.grad.clone() for x in t]
        if keep:                                                 # second backward must recompute, not reuse
            for x in t:
                x.grad = None
            loss.backward()
            for a, b in zip(first, t):
                assert np.array_equal(a.cpu().numpy(), b.grad.cpu().numpy())
        grads.append([a.cpu().numpy() for a in first])
    for a, b in zip(*grads):
        assert np.array_equal(a, b)


def test_mano_zero_angle_is_finite_where_the_reference_is_nan(pkg, synth_model, cuda_device):
    """rots == 0 gives NaN gradients in the reference (r/theta); the analytic limit here."""
    import torch

    layer = pkg.ManoLayer(cuda_device, model=synth_model, pose_num=45)
    rot = torch.zeros(3, 3, device=cuda_device, requires_grad=True)
    pose = torch.zeros(3, 45, device=cuda_device, requires_grad=True)
    beta = torch.zeros(3, 10, device=cuda_device, requires_grad=True)
    v, j = layer(rot, pose, beta)
    ov, oj = mo.mano_forward(synth_model, np.zeros((3, 3)), np.zeros((3, 45)), np.zeros((3, 10)))
    assert_positions(v.detach().cpu().numpy(), ov)
    (v.sum() + j.sum()).backward()
    og = mo.mano_backward(synth_model, np.zeros((3, 3)), np.zeros((3, 45)), np.zeros((3, 10)),
                          np.ones((3, 778, 3)), np.ones((3, 21, 3)))
    for t, want in zip((rot, pose, beta), og):
        assert torch.isfinite(t.grad).all()
        assert rel(t.grad.cpu().numpy(), want) < GRAD_TOL


def test_mano_empty_batch_and_input_handling(pkg, synth_model, cuda_device):
    import torch

    layer = pkg.ManoLayer(cuda_device, model=synth_model, pose_num=10)
    v, j = layer(torch.zeros(0, 3, device=cuda_device), torch.zeros(0, 10, device=cuda_device),
                 torch.zeros(0, 10, device=cuda_device))
    assert v.shape == (0, 778, 3) and j.shape == (0, 21, 3)
    # non-contiguous / fp64 inputs are accepted like the reference tolerates views
    rot, pose, beta = mano_inputs(6, 10, seed=9)
    big = torch.from_numpy(np.concatenate([rot, rot], 1)).to(cuda_device)
    v1, j1 = layer(big[:, :3], torch.from_numpy(pose).to(cuda_device).double(), torch.from_numpy(beta).to(cuda_device))
    ov, oj = mo.mano_forward(synth_model, rot, pose, beta)
    assert_positions(v1.cpu().numpy(), ov)
    with pytest.raises(RuntimeError):
        layer(torch.zeros(2, 3, device=cuda_device), torch.zeros(2, 11, device=cuda_device),
              torch.zeros(2, 10, device=cuda_device))


def test_mano_linearity_property_full_size(pkg, synth_model, cuda_device):
    """Size-independent property at a large batch (65536 hands): the layer is affine in the
    shape coefficients for fixed pose up to the joint-regression path — check instead the exact
    invariances: (a) verts of hand i do not depend on the batch it sits in, (b) a global
    rotation about the origin preserves every vertex norm."""
    import torch

    B = 65536
    rot, pose, beta = mano_inputs(B, 45, seed=11)
    layer = pkg.ManoLayer(cuda_device, model=synth_model, pose_num=45)
    t = to_dev(cuda_device, rot, pose, beta)
    v, j = layer(*t)
    sel = np.array([0, 1, 2, 777, 4095, 4096, 32767, 65534, 65535])
    # (a) bit-exact in another batch served by the same kernels (>= 8192 hands: one thread per hand),
    #     at other positions of a hand group / tcgen05 tile ...
    order = torch.from_numpy(np.r_[np.arange(5000, 5003), sel, np.arange(7000, 7000 + 8192 - 12)]).to(cuda_device)
    v_mid, j_mid = layer(*[x[order] for x in t])
    assert torch.equal(v[sel], v_mid[3:12]) and torch.equal(j[sel], j_mid[3:12])
    # ... and to fp32 rounding in a 9-hand batch (one warp per hand pose kernels)
    v_small, j_small = layer(*[x[torch.from_numpy(sel).to(cuda_device)] for x in t])
    assert float((v[sel] - v_small).abs().max()) < 2e-7 and float((j[sel] - j_small).abs().max()) < 2e-7
    zero_rot = torch.zeros_like(t[0])
    v0, _ = layer(zero_rot, t[1], t[2])
    n1 = v.norm(dim=2)
    n0 = v0.norm(dim=2)
    assert float((n1 - n0).abs().max()) < 3e-7
    ov, oj = mo.mano_forward(synth_model, rot[sel], pose[sel], beta[sel])
    assert_positions(v_small.cpu().numpy(), ov)


def test_lbs_stage_alone(pkg, synth_model, cuda_device):
    """mb_lbs_forward through the C ABI: identity bones reproduce the input; random bones match numpy."""
    import torch

    lib = pkg.load_library()
    layer = pkg.ManoLayer(cuda_device, model=synth_model, pose_num=45)
    B = 5
    rs = np.random.RandomState(3)
    vp = np.zeros((B, 2336), np.float32)
    vp[:, :2334] = rs.randn(B, 2334).astype(np.float32) * .1
    bone = rs.randn(B, 16, 3, 4).astype(np.float32)
    tvp, tbone = to_dev(cuda_device, vp, bone)
    verts = torch.empty(B, 778, 3, device=cuda_device)
    joints = torch.zeros(B, 21, 3, device=cuda_device)
    ws = torch.empty(lib.mb_lbs_workspace_bytes(B), dtype=torch.uint8, device=cuda_device)
    pkg._cabi.check(lib.mb_lbs_forward(layer._blob.data_ptr(), tvp.data_ptr(), 2336, tbone.data_ptr(), B,
                                       verts.data_ptr(), joints.data_ptr(), ws.data_ptr(), ws.numel(),
                                       pkg._cabi.stream_handle(cuda_device)), "lbs")
    W = synth_model["weights"].astype(np.float32).astype(np.float64)
    T = np.einsum("vk,bkij->bvij", W, bone.astype(np.float64))
    x = vp[:, :2334].reshape(B, 778, 3).astype(np.float64)
    want = np.einsum("bvij,bvj->bvi", T[..., :3], x) + T[..., 3]
    assert np.abs(verts.cpu().numpy() - want).max() < 2e-6
    assert np.abs(joints.cpu().numpy()[:, [4, 8, 12, 16, 20]] - want[:, [333, 444, 672, 555, 745]]).max() < 2e-6
    assert lib.mb_lbs_forward(layer._blob.data_ptr(), tvp.data_ptr(), 2334, tbone.data_ptr(), B,
                              verts.data_ptr(), None, ws.data_ptr(), ws.numel(), None) == -2
    assert lib.mb_lbs_forward(layer._blob.data_ptr(), tvp.data_ptr(), 2336, tbone.data_ptr(), B,
                              verts.data_ptr(), None, ws.data_ptr(), 16, None) == -3


# ------------------------------------------------------------------------------- FK
@pytest.mark.parametrize("name", ["fk_switched.npz", "fk_unswitched.npz"])
def test_fk_matches_reference_golden(pkg, cuda_device, name):
    g = load_golden(name)
    sw = bool(g["switched"])
    fk = pkg.ForwardKinematics(cuda_device, joint_order_switched=sw)
    ra, oa, bl = to_dev(cuda_device, g["root_angles"], g["other_angles"], g["bone_lengths"], grad=True)
    K, sc, root, gx, gu = to_dev(cuda_device, g["K"], g["scale"], g["root"], g["g_xyz"], g["g_uv"])
    out = fk(ra, oa, bl, K, sc, root)
    assert isinstance(out, list) and len(out) == 3 and out[2] is None
    xyz, uv = out[0], out[1]
    assert np.abs(xyz.detach().cpu().numpy() - g["xyz"]).max() < POS_TOL_REF
    assert np.abs(uv.detach().cpu().numpy() - g["uv"]).max() < 1e-3
    ((xyz * gx).sum() + (uv * gu).sum()).backward()
    assert rel(ra.grad.cpu().numpy(), g["g_root_angles"]) < GRAD_TOL
    assert rel(oa.grad.cpu().numpy(), g["g_other_angles"]) < GRAD_TOL
    assert rel(bl.grad.cpu().numpy(), g["g_bone_lengths"]) < GRAD_TOL


def fk_inputs(B, seed):
    rs = np.random.RandomState(seed)
    ra = ((rs.rand(B, 3) - .5) * 2 * np.pi).astype(np.float32)
    oa = ((rs.rand(B, 23) - .5) * np.pi).astype(np.float32)
    bl = (rs.rand(B, 20) + .1).astype(np.float32)
    K = np.tile(np.array([[282.9, 0, 160], [0, 282.9, 160], [0, 0, 1]], np.float32), (B, 1, 1))
    sc = (rs.rand(B, 1) * .05 + .02).astype(np.float32)
    root = (rs.randn(B, 3) * .05 + np.array([0, 0, .6])).astype(np.float32)
    return ra, oa, bl, K, sc, root


@pytest.mark.parametrize("B", [1, 63, 64, 65, 1000, 65536])
def test_fk_matches_oracle_config3(pkg, cuda_device, B):
    """Up to BASELINE config 3 (B=65536) with the visible-joint MPJPE reduction."""
    import torch

    args = fk_inputs(B, seed=B)
    fk = pkg.ForwardKinematics(cuda_device, joint_order_switched=True)
    t = to_dev(cuda_device, *args[:3], grad=True) + to_dev(cuda_device, *args[3:])
    xyz, uv, _ = fk(*t)
    oxyz, ouv = fo.fk_forward(*args)
    assert np.abs(xyz.detach().cpu().numpy() - oxyz).max() < POS_TOL_FK
    assert np.abs(uv.detach().cpu().numpy() - ouv).max() < 1e-3
    rs = np.random.RandomState(7)
    gx = rs.randn(B, 21, 3).astype(np.float32)
    gu = (rs.randn(B, 21, 2) * 1e-3).astype(np.float32)
    tgx, tgu = to_dev(cuda_device, gx, gu)
    ((xyz * tgx).sum() + (uv * tgu).sum()).backward()
    ora, ooa, obl = fo.fk_backward(*args, gx, gu)
    for got, want in zip(t[:3], (ora, ooa, obl)):
        assert rel(got.grad.cpu().numpy(), want) < GRAD_TOL
    vis = (rs.rand(B, 21, 1) < .8).astype(np.float32)
    gt = (oxyz + rs.randn(B, 21, 3) * .05).astype(np.float32)
    tvis, tgt = to_dev(cuda_device, vis, gt)
    m = pkg.MPJPE()(xyz.detach(), tgt, tvis)
    assert m.dim() == 0
    want = fo.mpjpe(xyz.detach().cpu().numpy(), gt, vis)
    assert abs(float(m) - want) <= 1e-5 * want
    # xyz-only upstream gradient (uv unused) and uv-only
    for x in t[:3]:
        x.grad = None
    xyz2, uv2, _ = fk(*t)
    (xyz2 * tgx).sum().backward()
    ora, ooa, obl = fo.fk_backward(*args, gx, None)
    for got, want in zip(t[:3], (ora, ooa, obl)):
        assert rel(got.grad.cpu().numpy(), want) < GRAD_TOL
    torch.cuda.synchronize()


@pytest.mark.parametrize("B,skip", [(97, 1), (1 << 18, 0), (100003, 3)])
def test_fk_bulk_copy_and_per_element_paths_agree(pkg, cuda_device, B, skip):
    """fk.cu moves whole 32-sample tiles with bulk copies when every pointer is 16-byte aligned and falls back to per-element
    loops otherwise (row-offset views: rows of 3 / 23 / 9 / 1 floats are not 16-byte multiples) and for the last partial tile.
    Both paths run the same per-sample code: bit-identical results, several tiles per persistent warp at the large sizes."""
    import torch

    args = fk_inputs(B + skip, seed=11 + B)
    fk = pkg.ForwardKinematics(cuda_device, joint_order_switched=False)
    full = to_dev(cuda_device, *args)
    rs = np.random.RandomState(5)
    gx, gu = to_dev(cuda_device, rs.randn(B + skip, 21, 3).astype(np.float32), (rs.randn(B + skip, 21, 2) * 1e-3).astype(np.float32))

    def run(ts, gxs, gus):
        leaves = [x.detach().requires_grad_() for x in ts[:3]]
        xyz, uv, _ = fk(*leaves, *ts[3:])
        ((xyz * gxs).sum() + (uv * gus).sum()).backward()
        return [xyz.detach(), uv.detach()] + [x.grad for x in leaves]

    # (a) views that start `skip` rows in: unaligned when skip is odd -> per-element path for every tile
    view = run([x[skip:] for x in full], gx[skip:], gu[skip:])
    # (b) fresh, aligned copies of the same rows -> bulk path for the full tiles
    copy = run([x[skip:].clone() for x in full], gx[skip:].clone(), gu[skip:].clone())
    for a, b in zip(view, copy):
        assert torch.equal(a, b)
    n = min(B, 4096)
    sub = [a[skip:skip + n] for a in args]
    oxyz, ouv = fo.fk_forward(*sub, joint_order_switched=False)
    assert np.abs(copy[0][:n].cpu().numpy() - oxyz).max() < POS_TOL_FK
    # the one-kernel loss call on the same views / copies (ground truths and visibility offset too)
    vis = (rs.rand(B + skip, 21) < .8).astype(np.float32)
    tv, = to_dev(cuda_device, vis)
    crit = pkg.ForwardKinematicsLoss(cuda_device, joint_order_switched=False)

    def run_loss(ts, gxs, gus, vs):
        leaves = [x.detach().requires_grad_() for x in ts[:3]]
        lx, lu, xyz, uv = crit(*leaves, *ts[3:], gxs, gus * 1e3, vs)
        (lx + 1e-3 * lu).backward()
        return [lx.detach(), lu.detach(), xyz, uv] + [x.grad for x in leaves]

    lv = run_loss([x[skip:] for x in full], gx[skip:], gu[skip:], tv[skip:])
    lc = run_loss([x[skip:].clone() for x in full], gx[skip:].clone(), gu[skip:].clone(), tv[skip:].clone())
    for a, b in zip(lv[2:], lc[2:]):
        assert torch.equal(a, b)
    for a, b in zip(lv[:2], lc[:2]):                        # the same fp64 partial sums, added in a different order
        assert abs(float(a) - float(b)) <= 1e-6 * abs(float(b))
    assert torch.equal(lv[2], copy[0]) and torch.equal(lv[3], copy[1])
    torch.cuda.synchronize()


@pytest.mark.parametrize("B,switched,terms", [(1, True, (1, 1)), (200, True, (1, 1)), (333, False, (1, 1)), (65536, True, (1, 1)),
                                              (77, True, (1, 0)), (77, False, (0, 1)), (64, True, (0, 0))])
def test_fk_loss_one_call_matches_oracle_and_the_separate_dropins(pkg, cuda_device, B, switched, terms):
    """ForwardKinematicsLoss (mb_fk_loss_forward / _backward: FK + both L2Loss terms, one kernel per direction) against the fp64
    oracle composition fk_forward -> l2loss (+ backward) and against ForwardKinematics + L2Loss x2 through autograd."""
    import torch

    args = fk_inputs(B, seed=3 * B + 1)
    rs = np.random.RandomState(B)
    oxyz, ouv = fo.fk_forward(*args, joint_order_switched=switched)
    gt_xyz = (oxyz + rs.randn(B, 21, 3) * .02).astype(np.float32)
    gt_uv = (ouv + rs.randn(B, 21, 2) * 3.).astype(np.float32)
    vis = (rs.rand(B, 21, 1) < .8).astype(np.float32)
    if B == 1:
        vis[:] = 1.
    wx, wu = 0.7, 1e-3                                          # weights of the two terms in the total (upstream gradients)
    use_xyz, use_uv = terms
    crit = pkg.ForwardKinematicsLoss(cuda_device, comp_xyz_loss=bool(use_xyz), comp_uv_loss=bool(use_uv), joint_order_switched=switched)
    t = to_dev(cuda_device, *args[:3], grad=True) + to_dev(cuda_device, *args[3:])
    tg = to_dev(cuda_device, gt_xyz, gt_uv, vis)
    lx, lu, xyz, uv = crit(*t, *tg)
    assert (lx is None) == (not use_xyz) and (lu is None) == (not use_uv)
    assert not xyz.requires_grad and not uv.requires_grad
    assert np.abs(xyz.cpu().numpy() - oxyz).max() < POS_TOL_FK
    assert np.abs(uv.cpu().numpy() - ouv).max() < 1e-3
    # oracle losses on the kernel's own fp32 outputs (the reduction), then the full fp64 chain for the gradients
    if use_xyz:
        want = fo.l2loss(xyz.cpu().numpy(), gt_xyz, vis)
        assert abs(float(lx.detach()) - want) <= 1e-5 * abs(want) + 1e-12
    if use_uv:
        want = fo.l2loss(uv.cpu().numpy(), gt_uv, vis)
        assert abs(float(lu.detach()) - want) <= 1e-5 * abs(want) + 1e-12
    if not (use_xyz or use_uv):
        return
    total = (wx * lx if use_xyz else 0.) + (wu * lu if use_uv else 0.)
    total.backward()
    g_xyz = wx * fo.l2loss_backward(oxyz, gt_xyz, vis) if use_xyz else None
    g_uv = wu * fo.l2loss_backward(ouv, gt_uv, vis) if use_uv else None
    want = fo.fk_backward(*args, g_xyz, g_uv, joint_order_switched=switched)
    got = [x.grad.clone() for x in t[:3]]
    for g, w in zip(got, want):
        assert rel(g.cpu().numpy(), w) < GRAD_TOL
    # the separate drop-ins through autograd: same kernels' arithmetic, so agreement to fp32 rounding of the loss scalars
    for x in t[:3]:
        x.grad = None
    fk, l2 = pkg.ForwardKinematics(cuda_device, joint_order_switched=switched), pkg.L2Loss()
    xyz2, uv2, _ = fk(*t)
    assert torch.equal(xyz2.detach(), xyz) and torch.equal(uv2.detach(), uv)
    total2 = (wx * l2(xyz2, tg[0], tg[2]) if use_xyz else 0.) + (wu * l2(uv2, tg[1], tg[2]) if use_uv else 0.)
    total2.backward()
    assert abs(float(total2.detach()) - float(total.detach())) <= 2e-6 * abs(float(total.detach()))
    for g, x in zip(got, t[:3]):
        assert rel(g.cpu().numpy(), x.grad.cpu().numpy()) < 1e-5
    if use_xyz and use_uv:
        # both terms computed, only ONE differentiated: the other term's upstream gradient is None inside the node
        for x in t[:3]:
            x.grad = None
        lx3, lu3, _, _ = crit(*t, *tg)
        lu3.backward()
        want = fo.fk_backward(*args, None, fo.l2loss_backward(ouv, gt_uv, vis), joint_order_switched=switched)
        for x, w in zip(t[:3], want):
            assert rel(x.grad.cpu().numpy(), w) < GRAD_TOL
    torch.cuda.synchronize()


def test_fk_hand_typed_kat(pkg, cuda_device):
    """KAT-FK-0."""
    import torch

    oa = torch.zeros(1, 23, device=cuda_device)
    oa[0, 1] = np.pi / 2
    K = torch.tensor([[[600., 0, 300], [0, 600., 300], [0, 0, 1]]], device=cuda_device)
    xyz, uv, _ = pkg.ForwardKinematics(cuda_device, joint_order_switched=True)(
        torch.tensor([[0, 0, np.pi / 2]], device=cuda_device, dtype=torch.float32), oa,
        torch.ones(1, 20, device=cuda_device), K, torch.ones(1, 1, device=cuda_device), torch.zeros(1, 3, device=cuda_device))
    xyz, uv = xyz.cpu().numpy(), uv.cpu().numpy()
    for k in range(1, 5):
        assert np.allclose(xyz[0, k], [0, k, 0], atol=1e-6)
    for f in range(1, 5):
        for k in range(1, 5):
            assert np.allclose(xyz[0, 4 * f + k], [0, 0, k], atol=1e-6)
            assert np.allclose(uv[0, 4 * f + k], [300, 300], atol=1e-3)
    assert np.allclose(uv[0, 0], [0, 0])


def test_projection_and_z0_branch(pkg, cuda_device):
    g = load_golden("project_uv.npz")
    xyz, = to_dev(cuda_device, g["xyz"], grad=True)
    K, gu = to_dev(cuda_device, g["K"], g["g_uv"])
    uv = pkg.batch_project_xyz_to_uv(xyz, K)
    got = uv.detach().cpu().numpy()
    ok = np.abs(g["uv"]) < 1e6
    assert np.abs(got[ok] - g["uv"][ok]).max() < 2e-3
    assert np.array_equal(got[0, 0], g["uv"][0, 0])            # (0,0,0) -> 0 / 1e-10 = 0 exactly
    (uv * gu).sum().backward()
    fin = np.isfinite(g["g_xyz"]) & (np.abs(g["g_xyz"]) < 1e6)
    assert rel(xyz.grad.cpu().numpy()[fin], g["g_xyz"][fin]) < GRAD_TOL


# ---------------------------------------------------------------- joint epilogue
@pytest.mark.parametrize("name", ["match_switched.npz", "match_unswitched.npz"])
def test_match_mano_to_rhd_matches_reference_golden(pkg, cuda_device, name):
    """match_mano_to_RHD -> batch_project_xyz_to_uv of the reference heads, forward and the three input
    gradients, against fixtures produced by the unmodified reference (tests/golden/make_golden.py)."""
    g = load_golden(name)
    switched = bool(g["switched"])
    j, L, root = to_dev(cuda_device, g["joints"], g["scale"], g["root"], grad=True)
    K, gr, gx, gu = to_dev(cuda_device, g["K"], g["g_rel"], g["g_xyz"], g["g_uv"])
    keep = j.detach().clone()
    rel_n, xyz, uv = pkg.mano_joints_to_rhd_uv(j, L, root, K, joint_order_switched=switched)
    assert bool((j.detach() == keep).all())                    # the input is not permuted in place
    assert np.abs(rel_n.detach().cpu().numpy() - g["rel"]).max() < 2e-6 * np.abs(g["rel"]).max()
    assert np.abs(xyz.detach().cpu().numpy() - g["xyz"]).max() < 2e-7
    assert np.abs(uv.detach().cpu().numpy() - g["uv"]).max() < 2e-4          # pixels, |uv| ~ 160
    ((rel_n * gr).sum() + (xyz * gx).sum() + (uv * gu).sum()).backward()
    for t, key in ((j, "g_joints"), (L, "g_scale"), (root, "g_root")):
        assert rel(t.grad.cpu().numpy(), g[key]) < GRAD_TOL, key
    # the two-output form (no projection) gives the same tensors
    rel2, xyz2 = pkg.match_mano_to_RHD(j.detach(), L.detach(), root.detach(), joint_order_switched=switched)
    import torch
    assert torch.equal(rel2, rel_n.detach()) and torch.equal(xyz2, xyz.detach())


@pytest.mark.parametrize("B", [1, 31, 33, 4096, 100003])
def test_match_mano_to_rhd_matches_fp64_oracle(pkg, cuda_device, B):
    """Ragged batches up to beyond config 2, both joint orders, optional outputs / gradients."""
    rs = np.random.RandomState(B)
    # hand-like joints: every joint a few centimetres from the wrist, so ||r_12|| (either joint order) stays away from 0
    spread = np.arange(21)[:, None] * np.array([.006, .004, .002]) + np.array([0, 0, 0.])
    joints = (spread[None] + rs.randn(B, 21, 3) * .01).astype(np.float32)
    L = (rs.rand(B, 1) * .05 + .02).astype(np.float32)
    root = (rs.randn(B, 3) * .05 + np.array([0, 0, .6])).astype(np.float32)
    K = np.tile(np.array([[282.9, 0, 160], [0, 282.9, 160], [0, 0, 1]], np.float32), (B, 1, 1))
    gr, gx = rs.randn(B, 21, 3).astype(np.float32), rs.randn(B, 21, 3).astype(np.float32)
    gu = (rs.randn(B, 21, 2) * 1e-3).astype(np.float32)
    for switched in (True, False):
        tj, tL, troot = to_dev(cuda_device, joints, L, root, grad=True)
        tK, tgr, tgx, tgu = to_dev(cuda_device, K, gr, gx, gu)
        rel_n, xyz, uv = pkg.mano_joints_to_rhd_uv(tj, tL, troot, tK, joint_order_switched=switched)
        orel, oxyz = fo.match_mano_to_rhd(joints, L, root, switched)
        ouv = fo.project_uv(oxyz, K.astype(np.float64))
        # a hand whose ||r_12|| happens to be tiny has huge normalised coordinates: tolerances are relative,
        # and the projection is checked where it is well conditioned (z away from 0)
        assert (np.abs(rel_n.detach().cpu().numpy() - orel) <= 1e-6 * (1 + np.abs(orel))).all()
        assert (np.abs(xyz.detach().cpu().numpy() - oxyz) <= 2e-7 + 1e-6 * np.abs(oxyz)).all()
        okz = oxyz[..., 2] > 0.1
        assert okz.mean() > 0.9
        assert (np.abs(uv.detach().cpu().numpy() - ouv)[okz] <= 2e-4 + 1e-5 * np.abs(ouv)[okz]).all()
        ((rel_n * tgr).sum() + (xyz * tgx).sum() + (uv * tgu).sum()).backward()
        gx_tot = gx.astype(np.float64) + fo.project_uv_backward(oxyz, K.astype(np.float64), gu.astype(np.float64))
        want = fo.match_mano_to_rhd_backward(joints, L, root, gr, gx_tot, switched)
        for t, w in zip((tj, tL, troot), want):
            assert rel(t.grad.cpu().numpy(), w) < GRAD_TOL
        # xyz-only upstream gradient, no scale / root gradient wanted
        tj2, = to_dev(cuda_device, joints, grad=True)
        _, xyz2 = pkg.match_mano_to_RHD(tj2, tL.detach(), troot.detach(), joint_order_switched=switched)
        (xyz2 * tgx).sum().backward()
        want2 = fo.match_mano_to_rhd_backward(joints, L, root, np.zeros_like(gr), gx, switched)[0]
        assert rel(tj2.grad.cpu().numpy(), want2) < GRAD_TOL


# ------------------------------------------------------- keypoint re-parameterisations
def test_keypoint_trafos_match_reference_golden(pkg, cuda_device):
    import torch

    g = load_golden("trafo.npz")
    xyz, rel_in = to_dev(cuda_device, g["xyz"], g["rel"])
    assert np.abs(pkg.bone_rel_trafo(xyz).cpu().numpy() - g["rel"]).max() < 5e-6
    assert np.abs(pkg.bone_rel_trafo_inv(rel_in).cpu().numpy() - g["inv"]).max() < 5e-6
    can, rot = pkg.canonical_trafo(xyz)
    assert np.abs(can.cpu().numpy() - g["can"]).max() < 1e-5 and np.abs(rot.cpu().numpy() - g["rot"]).max() < 1e-5
    cond = torch.from_numpy(g["cond_right"]).to(cuda_device)
    can_in, = to_dev(cuda_device, g["can"])
    assert np.array_equal(pkg.flip_right_hand(can_in, cond).cpu().numpy(), g["flipped"])
    assert np.array_equal(pkg.flip_right_hand(can_in, cond[:, None].expand(-1, 21)).cpu().numpy(), g["flipped"])
    assert np.array_equal(pkg.flip_right_hand(can_in[3], cond[3]).cpu().numpy(), g["flipped"][3])       # [21,3] form
    # left -> right mirroring of the RHD dataloader (dataloaderRHD.py:224-225): x -> -x where hand_side == 0
    side = torch.tensor([0, 1] * 6, device=cuda_device)
    want = torch.where((side == 0)[:, None, None], torch.cat([-xyz[..., :1], xyz[..., 1:]], dim=-1), xyz)
    assert torch.equal(pkg.mirror_left_hand(xyz, side), want)
    assert torch.equal(pkg.mirror_left_hand(xyz[0], side[0]), want[0])
    # canonical_trafo + flip in one pass
    can_f, _ = pkg.canonical_trafo(xyz, cond_right=cond)
    assert torch.equal(can_f, pkg.flip_right_hand(can, cond))
    # single-sample form of the dataloaders
    assert np.abs(pkg.bone_rel_trafo(xyz[2]).cpu().numpy()[0] - g["rel"][2]).max() < 5e-6


@pytest.mark.parametrize("B", [1, 33, 1000, 262147])
def test_keypoint_trafos_match_fp64_oracle_and_round_trip(pkg, cuda_device, B):
    """Ragged batches; size-independent properties at the large size: bone_rel_trafo_inv o bone_rel_trafo = id,
    the canonical frame puts joint 12 on the y axis and joint 20 in the xy-plane, and keeps every distance."""
    from oracle import trafo_oracle as tro

    rs = np.random.RandomState(B % 1000)
    xyz = (rs.randn(B, 21, 3) * .5 + np.arange(21)[:, None] * np.array([.09, .06, .03])).astype(np.float32)
    xyz[:, 0] = 0
    t, = to_dev(cuda_device, xyz)
    rel_n = pkg.bone_rel_trafo(t)
    back = pkg.bone_rel_trafo_inv(rel_n)
    assert float((back - t).abs().max()) < 2e-5
    can, rot = pkg.canonical_trafo(t)
    c = can.cpu().numpy()
    assert np.abs(c[:, 0]).max() == 0 and np.abs(c[:, 12, [0, 2]]).max() < 2e-5 and np.abs(c[:, 20, 2]).max() < 2e-5
    assert np.abs(np.linalg.norm(c, axis=2) - np.linalg.norm(xyz, axis=2)).max() < 2e-5
    r = rot.cpu().numpy().astype(np.float64)
    assert np.abs(r @ np.swapaxes(r, 1, 2) - np.eye(3)).max() < 1e-5
    idx = np.unique(np.r_[np.arange(min(B, 512)), np.arange(max(B - 64, 0), B)])
    orel = tro.bone_rel_trafo(xyz[idx])
    got = rel_n.cpu().numpy()[idx]
    # angles of a bone nearly parallel to the local y axis are ill-conditioned: compare where the projected length is sane
    d = np.abs(got - orel)
    d[..., 1:] = np.minimum(d[..., 1:], 2 * np.pi - d[..., 1:])
    well = orel[..., 0] * np.abs(np.cos(orel[..., 1])) > 0.05
    assert np.abs(d[..., 0]).max() < 2e-6 and d[..., 1:][well].max() < 2e-5
    assert np.abs(pkg.bone_rel_trafo_inv(to_dev(cuda_device, orel.astype(np.float32))[0]).cpu().numpy()
                  - tro.bone_rel_trafo_inv(orel.astype(np.float32))).max() < 5e-6
    ocan, orot = tro.canonical_trafo(xyz[idx])
    assert np.abs(c[idx] - ocan).max() < 2e-5 and np.abs(r[idx] - orot).max() < 2e-5


# --------------------------------------------------------------------- viewpoint epilogue
def test_viewpoint_matches_reference_golden(pkg, cuda_device):
    """_get_rot_mat -> can @ R (-> * scale + root -> projection) and its gradients against the reference's autograd."""
    g = load_golden("viewpoint.npz")
    can, ux, uy, uz = to_dev(cuda_device, g["can"], g["ux"], g["uy"], g["uz"], grad=True)
    gR, gr = to_dev(cuda_device, g["g_rot"], g["g_rel"])
    rel_n, R = pkg.viewpoint_transform(can, ux, uy, uz)
    assert np.abs(R.detach().cpu().numpy() - g["rot"]).max() < 5e-7
    assert np.abs(rel_n.detach().cpu().numpy() - g["rel"]).max() < 1e-6
    ((R * gR).sum() + (rel_n * gr).sum()).backward()
    assert rel(can.grad.cpu().numpy(), g["g_can"]) < GRAD_TOL
    for t, name in ((ux, "g_ux"), (uy, "g_uy"), (uz, "g_uz")):
        assert t.grad.shape == g[name].shape and rel(t.grad.cpu().numpy(), g[name]) < GRAD_TOL
    # stand-alone _get_rot_mat with its own gradient
    u2 = to_dev(cuda_device, g["ux"], g["uy"], g["uz"], grad=True)
    R2 = pkg._get_rot_mat(*u2)
    assert np.array_equal(R2.detach().cpu().numpy(), R.detach().cpu().numpy())
    (R2 * gR).sum().backward()
    from oracle import trafo_oracle as tro
    want = tro.viewpoint_backward(np.zeros_like(g["can"]), g["ux"], g["uy"], g["uz"], g["g_rot"], np.zeros_like(g["g_rel"]))
    for t, w in zip(u2, want[1:]):
        assert rel(t.grad.cpu().numpy()[:, 0], w) < GRAD_TOL
    # inference branch
    L, root, K = to_dev(cuda_device, g["scale"], g["root"], g["K"])
    xyz, uv = pkg.viewpoint_transform(can.detach().reshape(-1, 63), ux.detach(), uy.detach(), uz.detach(), L, root, K)
    assert np.abs(xyz.cpu().numpy() - g["xyz"]).max() < 1e-6
    assert np.abs(uv.cpu().numpy() - g["uv"]).max() < 5e-3


@pytest.mark.parametrize("B", [1, 31, 129, 4096, 300007])
def test_viewpoint_matches_fp64_oracle_and_properties(pkg, cuda_device, B):
    """Ragged batches against the fp64 restatement (subsample at the large size) plus size-independent properties:
    R is a rotation (up to the reference's 1e-8 epsilon), it keeps every norm, and R(-u) = R(u)^T."""
    import torch
    from oracle import trafo_oracle as tro

    rs = np.random.RandomState(B % 977)
    can = (rs.randn(B, 21, 3) * .5).astype(np.float32)
    u = [((rs.rand(B, 1) - .5) * 6).astype(np.float32) for _ in range(3)]
    gR = rs.randn(B, 3, 3).astype(np.float32)
    gr = rs.randn(B, 21, 3).astype(np.float32)
    tc, tx, ty, tz = to_dev(cuda_device, can, *u, grad=True)
    tgR, tgr = to_dev(cuda_device, gR, gr)
    rel_n, R = pkg.viewpoint_transform(tc, tx, ty, tz)
    ((R * tgR).sum() + (rel_n * tgr).sum()).backward()
    r = R.detach().cpu().numpy().astype(np.float64)
    assert np.abs(r @ np.swapaxes(r, 1, 2) - np.eye(3)).max() < 2e-6
    assert np.abs(np.linalg.norm(rel_n.detach().cpu().numpy(), axis=2) - np.linalg.norm(can, axis=2)).max() < 2e-6
    Rm = pkg._get_rot_mat(-tx.detach(), -ty.detach(), -tz.detach())
    assert float((Rm - R.detach().transpose(1, 2)).abs().max()) < 1e-6
    idx = np.unique(np.r_[np.arange(min(B, 512)), np.arange(max(B - 64, 0), B)])
    oR, orel = tro.viewpoint_forward(can[idx], u[0][idx], u[1][idx], u[2][idx])
    assert np.abs(r[idx] - oR).max() < 5e-7 and np.abs(rel_n.detach().cpu().numpy()[idx] - orel).max() < 2e-6
    ogc, ogx, ogy, ogz = tro.viewpoint_backward(can[idx], u[0][idx], u[1][idx], u[2][idx], gR[idx], gr[idx])
    assert rel(tc.grad.cpu().numpy()[idx], ogc) < GRAD_TOL
    for t, w in ((tx, ogx), (ty, ogy), (tz, ogz)):
        assert rel(t.grad.cpu().numpy()[idx, 0], w) < GRAD_TOL
    assert pkg.viewpoint_transform(tc.detach()[:0], tx.detach()[:0], ty.detach()[:0], tz.detach()[:0])[0].shape == (0, 21, 3)


# ------------------------------------------------------------------------ reductions
@pytest.mark.parametrize("name", ["reduce_vis80.npz", "reduce_none_visible.npz"])
def test_reductions_match_reference_golden(pkg, cuda_device, name):
    import torch

    g = load_golden(name)
    pre, = to_dev(cuda_device, g["pre"], grad=True)
    gt, vis = to_dev(cuda_device, g["gt"], g["vis"])
    m = pkg.MPJPE()(pre, gt, vis)
    l2 = pkg.L2Loss()(pre, gt, vis)
    assert float(m) == pytest.approx(float(g["mpjpe"]), rel=1e-5, abs=1e-12)
    assert float(l2) == pytest.approx(float(g["l2"]), rel=1e-5, abs=1e-12)
    l2.backward()
    assert np.abs(pre.grad.cpu().numpy() - g["g_pre"]).max() <= 1e-5 * max(np.abs(g["g_pre"]).max(), 1e-12) + 1e-12
    # uint8 / bool masks behave like the float ones
    m2 = pkg.MPJPE()(pre.detach(), gt, vis.bool())
    assert float(m2) == pytest.approx(float(m), rel=1e-6, abs=1e-12)
    torch.cuda.synchronize()


def test_l2loss_on_uv_and_regulariser_kernels_match_reference_golden(pkg, cuda_device):
    """L2Loss with [B,21,2] inputs (LossCalculation.compute_uv_coord_loss, loss.py:86-87) and the regulariser kernels
    (loss.py:113-117), forward and backward, against goldens recorded from the unmodified reference."""
    import torch

    g = load_golden("reduce_uv.npz")
    pre, = to_dev(cuda_device, g["pre"], grad=True)
    gt, vis = to_dev(cuda_device, g["gt"], g["vis"])
    l2 = pkg.L2Loss()(pre, gt, vis)
    assert float(l2) == pytest.approx(float(g["l2"]), rel=1e-5)
    l2.backward()
    assert np.abs(pre.grad.cpu().numpy() - g["g_pre"]).max() <= 1e-5 * np.abs(g["g_pre"]).max()
    r = load_golden("regulariser.npz")
    theta, beta = to_dev(cuda_device, r["theta"], r["beta"], grad=True)
    loss = pkg.compute_regularization_loss(theta, beta)
    assert float(loss) == pytest.approx(float(r["loss"]), rel=1e-6)
    (loss * float(r["g_out"])).backward()
    assert np.abs(theta.grad.cpu().numpy() - r["g_theta"]).max() <= 1e-5 * np.abs(r["g_theta"]).max()
    assert np.abs(beta.grad.cpu().numpy() - r["g_beta"]).max() <= 1e-5 * np.abs(r["g_beta"]).max()
    # zero norms: 0 loss, finite (zero) gradients — torch.norm's subgradient
    z1 = torch.zeros(4, 10, device=cuda_device, requires_grad=True)
    z2 = torch.zeros(4, 10, device=cuda_device, requires_grad=True)
    lz = pkg.compute_regularization_loss(z1, z2)
    lz.backward()
    assert float(lz) == 0.0 and float(z1.grad.abs().max()) == 0.0 and float(z2.grad.abs().max()) == 0.0


@pytest.mark.parametrize("name", ["hand_mask.npz", "hand_mask_empty.npz"])
def test_hand_mask_loss_matches_reference_golden(pkg, cuda_device, name):
    import torch

    g = load_golden(name)
    p, q, m = to_dev(cuda_device, g["pred_uv"], g["gt_uv"], g["hand_mask"])
    assert float(pkg.compute_hand_mask_loss(p, q, m)) == float(g["loss"])
    assert float(pkg.compute_hand_mask_loss(p, q, m.to(torch.uint8))) == float(g["loss"])       # byte masks too
    assert float(pkg.compute_hand_mask_loss(p, q, m > 0)) == float(g["loss"])


@pytest.mark.parametrize("B,H,W", [(1, 8, 8), (37, 40, 32), (4096, 64, 64)])
def test_hand_mask_loss_matches_oracle(pkg, cuda_device, B, H, W):
    """Bit-exact (integer gather + exactly representable sums) against the oracle, uv far outside the image,
    non-finite uv (NaN -> 0 like the reference's float -> int64 cast of this platform is not relied on: excluded)."""
    rs = np.random.RandomState(B)
    mask = (rs.rand(B, H, W) < .4).astype(np.float32)
    gt = (rs.rand(B, 21, 2) * (W + 20) - 10).astype(np.float32)
    pred = (gt + rs.randn(B, 21, 2) * 6).astype(np.float32)
    pred[0, 0] = [-1e9, 1e9]
    p, q, m = to_dev(cuda_device, pred, gt, mask)
    assert float(pkg.compute_hand_mask_loss(p, q, m)) == float(fo.hand_mask_loss(pred, gt, mask))
    with pytest.raises(IndexError):
        pkg.compute_hand_mask_loss(p, q, m[:, : W // 2, :])


def test_adam_step_matches_torch(pkg, cuda_device):
    import torch

    lib = pkg.load_library()
    torch.manual_seed(0)
    p = torch.randn(10001, device=cuda_device)
    ref_p = p.clone().requires_grad_()
    opt = torch.optim.Adam([ref_p], lr=1e-2)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    for step in range(1, 6):
        gr = torch.randn_like(p)
        ref_p.grad = gr.clone()
        opt.step()
        pkg._cabi.check(lib.mb_adam_step(p.data_ptr(), gr.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(),
                                         1e-2, 0.9, 0.999, 1e-8, step, pkg._cabi.stream_handle(cuda_device)), "adam")
    assert float((p - ref_p.detach()).abs().max()) < 1e-6


# ------------------------------------------------------------------------ fitting loop
@pytest.mark.parametrize("fused,B,nc", [(True, 64, 45), (False, 64, 45), (True, 8231, 45), (True, 100, 10)])
def test_fitting_loop_matches_oracle_adam(pkg, synth_model, cuda_device, fused, B, nc):
    """BASELINE config 5 parity: 64 hands x 10 Adam iterations against the oracle's
    objective (L2Loss + regulariser) and gradients with a numpy Adam (torch.optim.Adam semantics) — as one
    kernel per iteration (mb_mano_fit_step) and as the separate forward / reduce / backward / Adam kernels;
    a ragged 8 231-hand batch for the fused kernel's multi-group path."""
    import torch

    fitting = importlib.import_module("3dhandposeestimation_b200.fitting")
    iters = 10
    rs = np.random.RandomState(5)
    hidden = mano_inputs(B, nc, seed=77, pose_scale=1.0)
    _, tj = mo.mano_forward(synth_model, *hidden)
    target = (tj + rs.randn(B, 21, 3) * 1e-3).astype(np.float32)
    vis = (rs.rand(B, 21, 1) < .8).astype(np.float32)
    start = [a * 0.5 + 0.01 for a in hidden]
    layer = pkg.ManoLayer(cuda_device, model=synth_model, pose_num=nc)
    fit = fitting.ManoFitter(layer, B, lr=1e-2, fused=fused)
    assert fit.fused == fused
    for dst, src in zip((fit.rot, fit.pose, fit.beta), start):
        dst.copy_(torch.from_numpy(src.astype(np.float32)))
    ttarget, tvis = to_dev(cuda_device, target, vis)
    # numpy reference
    p = [a.astype(np.float64) for a in start]
    m = [np.zeros_like(a) for a in p]
    v = [np.zeros_like(a) for a in p]
    losses_ref, losses = [], []
    for it in range(1, iters + 1):
        losses.append(float(fit.step(ttarget, tvis)))
        _, j = mo.mano_forward(synth_model, *p)
        losses_ref.append(float(fo.l2loss(j, target, vis) + fo.regularizer(p[1], p[2])))
        gj = fo.l2loss_backward(j, target, vis)
        g = list(mo.mano_backward(synth_model, *p, None, gj))
        g[1] = g[1] + p[1] / (100.0 * np.linalg.norm(p[1]))
        g[2] = g[2] + 10.0 * p[2] / (100.0 * np.linalg.norm(p[2]))
        for k in range(3):
            m[k] = 0.9 * m[k] + 0.1 * g[k]
            v[k] = 0.999 * v[k] + 0.001 * g[k] ** 2
            p[k] = p[k] - 1e-2 / (1 - 0.9 ** it) * m[k] / (np.sqrt(v[k]) / np.sqrt(1 - 0.999 ** it) + 1e-8)
    assert np.allclose(losses, losses_ref, rtol=2e-4, atol=1e-7), (losses, losses_ref)
    assert losses[-1] < losses[0]
    for got, want in zip((fit.rot, fit.pose, fit.beta), p):
        assert np.abs(got.cpu().numpy() - want).max() < 2e-3      # Adam's sign-like first steps amplify fp32 noise


def test_module_api_does_not_leak_device_memory(pkg, synth_model, cuda_device):
    import torch

    layer = pkg.ManoLayer(cuda_device, model=synth_model, pose_num=45, mode="f16x3")
    rot, pose, beta = mano_inputs(2048, 45, seed=1)
    gv = torch.randn(2048, 778, 3, device=cuda_device)
    gj = torch.randn(2048, 21, 3, device=cuda_device)

    def one():
        t = to_dev(cuda_device, rot, pose, beta, grad=True)
        v, j = layer(*t)
        torch.autograd.backward([v, j], [gv, gj])

    for _ in range(3):
        one()
    torch.cuda.synchronize()
    base = torch.cuda.memory_allocated(cuda_device)
    for _ in range(10):
        one()
    torch.cuda.synchronize()
    assert torch.cuda.memory_allocated(cuda_device) - base < (1 << 20)


def test_layers_on_a_non_current_device(pkg, synth_model):
    """The C ABI launches on the CURRENT device; the wrappers switch to the tensors' device (a layer on
    cuda:1 while cuda:0 is current failed every launch).  Needs two GPUs: skipped on a single-GPU box."""
    import torch

    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    assert torch.cuda.current_device() == 0
    dev = torch.device("cuda", 1)
    B, nc = 97, 45
    rot, pose, beta = mano_inputs(B, nc, 5)
    layer = pkg.ManoLayer(dev, model=synth_model, pose_num=nc)
    t = to_dev(dev, rot, pose, beta, grad=True)
    verts, joints = layer(*t)
    (verts.sum() + joints.sum()).backward()
    ov, oj = mo.mano_forward(synth_model, rot, pose, beta)
    assert_positions(verts.detach().cpu().numpy(), ov)
    og = mo.mano_backward(synth_model, rot, pose, beta, np.ones((B, 778, 3), np.float32), np.ones((B, 21, 3), np.float32))
    for got, want in zip(t, og):
        assert rel(got.grad.cpu().numpy(), want) < GRAD_TOL
    args = fk_inputs(B, seed=9)
    xyz, uv, _ = pkg.ForwardKinematics(dev)(*to_dev(dev, *args))
    oxyz, _ = fo.fk_forward(*args)
    assert np.abs(xyz.cpu().numpy() - oxyz).max() < POS_TOL_FK
    assert torch.cuda.current_device() == 0
    torch.cuda.synchronize(dev)


@pytest.mark.parametrize("B,shift", [(32, 0), (33, 0), (4099, 0), (4099, 1), (70000, 0)])
def test_fk_entry_points_write_only_their_outputs(pkg, cuda_device, B, shift):
    """Guard bands around every output of the FK entry points (bulk-copy path: shift 0 keeps the 16-byte alignment; per-element
    path: shift 1 float): nothing outside [0, B x width) changes.  compute-sanitizer is not available on the GPU pool."""
    import torch

    cabi = pkg._cabi
    lib = pkg.load_library()
    args = fk_inputs(B, seed=B + shift)
    ins = to_dev(cuda_device, *args)
    PAD, SENT = 64, -777.25

    def guarded(width):
        buf = torch.full((PAD + shift + B * width + PAD,), SENT, device=cuda_device)
        return buf, buf[PAD + shift:PAD + shift + B * width]

    def intact(buf, width):
        return bool((buf[:PAD + shift] == SENT).all()) and bool((buf[PAD + shift + B * width:] == SENT).all())

    P = lambda t: t.data_ptr()
    st = cabi.stream_handle(cuda_device)
    bx, xyz = guarded(63)
    bu, uv = guarded(42)
    cabi.check(lib.mb_fk_forward(*[P(t) for t in ins], B, 0, P(xyz), P(uv), st), "fk_forward")
    assert intact(bx, 63) and intact(bu, 42)
    assert not bool((xyz == SENT).any()) and not bool((uv == SENT).any())
    rs = np.random.RandomState(1)
    gx, gu, vis = to_dev(cuda_device, rs.randn(B, 21, 3).astype(np.float32), rs.randn(B, 21, 2).astype(np.float32),
                         (rs.rand(B, 21) < .8).astype(np.float32))
    outs = [guarded(w) for w in (3, 23, 20)]
    cabi.check(lib.mb_fk_backward(*[P(t) for t in ins], P(gx), P(gu), B, 0, *[P(o[1]) for o in outs], st), "fk_backward")
    for (buf, view), w in zip(outs, (3, 23, 20)):
        assert intact(buf, w) and not bool((view == SENT).any())
    # the one-kernel loss pair
    bx2, xyz2 = guarded(63)
    bu2, uv2 = guarded(42)
    losses = torch.full((2 + 2 * PAD,), SENT, device=cuda_device)
    ws = torch.zeros(8, dtype=torch.float64, device=cuda_device)
    both = cabi.HEAD_XYZ | cabi.HEAD_UV
    cabi.check(lib.mb_fk_loss_forward(*[P(t) for t in ins], P(gx), P(gu), P(vis), B, 0, both, P(xyz2), P(uv2), P(losses[PAD:]),
                                      P(ws), 64, st), "fk_loss_forward")
    assert intact(bx2, 63) and intact(bu2, 42) and torch.equal(xyz2, xyz) and torch.equal(uv2, uv)
    assert bool((losses[:PAD] == SENT).all()) and bool((losses[PAD + 2:] == SENT).all()) and not bool((losses[PAD:PAD + 2] == SENT).any())
    g_l = torch.tensor([1.0, 1e-3], device=cuda_device)
    outs2 = [guarded(w) for w in (3, 23, 20)]
    cabi.check(lib.mb_fk_loss_backward(*[P(t) for t in ins], P(gx), P(gu), P(vis), B, 0, both, P(g_l), *[P(o[1]) for o in outs2],
                                       P(ws), 64, st), "fk_loss_backward")
    for (buf, view), w in zip(outs2, (3, 23, 20)):
        assert intact(buf, w) and not bool((view == SENT).any())
    torch.cuda.synchronize()


@pytest.mark.parametrize("B,inference", [(8195, False), (8195, True), (37921, False), (100, False)])
def test_mano_entry_points_write_only_their_outputs(pkg, synth_model, cuda_device, B, inference):
    """Guard bands around verts / joints / the three gradients and BEHIND the declared workspace of mb_mano_forward /
    mb_mano_backward (ragged batches: the fused forward's 64-hand tiles, its 256-bit scratch stores into 32-hand groups, the
    lane = hand backward).  compute-sanitizer is not available on the GPU pool."""
    import torch

    cabi = pkg._cabi
    lib = pkg.load_library()
    nc = 45
    layer = pkg.ManoLayer(cuda_device, model=synth_model, pose_num=nc)
    layer._require_device()
    rot, pose, beta = to_dev(cuda_device, *mano_inputs(B, nc, 77))
    PAD, SENT = 4096, -777.25
    mode = layer._mode | (cabi.FWD_INFERENCE if inference else 0)

    def guarded(n):
        buf = torch.full((PAD + n + PAD,), SENT, device=cuda_device)
        return buf, buf[PAD:PAD + n]

    def intact(buf, n):
        return bool((buf[:PAD] == SENT).all()) and bool((buf[PAD + n:] == SENT).all())

    P = lambda t: t.data_ptr()
    st = cabi.stream_handle(cuda_device)
    nws = lib.mb_mano_workspace_bytes(B, layer._mode)
    wsbuf = torch.full((nws + 65536,), 0x5a, dtype=torch.uint8, device=cuda_device)
    bv, verts = guarded(B * 778 * 3)
    bj, joints = guarded(B * 21 * 3)
    cabi.check(lib.mb_mano_forward(P(layer._blob), nc, P(rot), P(pose), P(beta), B, mode, P(verts), P(joints), P(wsbuf), nws, st), "fwd")
    assert intact(bv, B * 778 * 3) and intact(bj, B * 63)
    assert not bool((verts == SENT).any()) and not bool((joints == SENT).any())
    assert bool((wsbuf[nws:] == 0x5a).all())
    ov, oj = mo.mano_forward(synth_model, *[t.cpu().numpy() for t in (rot[-3:], pose[-3:], beta[-3:])])
    assert np.abs(verts.view(B, 778, 3)[-3:].cpu().numpy() - ov).max() < POS_TOL_REF
    if inference:
        return
    gv = torch.randn(B, 778, 3, device=cuda_device)
    gj = torch.randn(B, 21, 3, device=cuda_device)
    outs = [guarded(B * w) for w in (3, nc, 10)]
    cabi.check(lib.mb_mano_backward(P(layer._blob), nc, P(rot), P(pose), P(beta), P(gv), P(gj), B, layer._mode, cabi.BWD_WORKSPACE_VALID,
                                    *[P(o[1]) for o in outs], P(wsbuf), nws, st), "bwd")
    for (buf, view), w in zip(outs, (3, nc, 10)):
        assert intact(buf, B * w) and not bool((view == SENT).any())
    assert bool((wsbuf[nws:] == 0x5a).all())
    torch.cuda.synchronize()

