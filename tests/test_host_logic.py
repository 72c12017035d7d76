"""CPU: host-side logic — the C ABI loads and exports every declared symbol, constants pack,
argument checking, module interface parity with the reference, no-CPU-fallback behaviour, and
the __host__ __device__ kernel math (compiled for the host) against the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import fk_oracle as fo
from oracle import mano_oracle as mo


def test_library_exports_every_header_symbol(pkg):
    lib = pkg.load_library()
    declared = pkg._cabi.header_symbols()
    assert len(declared) >= 16
    assert set(declared) == set(pkg._cabi.SIGNATURES), "ctypes table and header disagree"
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mb_abi_version() == 4
    assert b"NULL" in lib.mb_error_string(-1)


def test_no_torch_types_in_the_abi():
    import re

    text = open(os.path.join(ROOT, "include", "mano_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)          # strip comments
    assert "torch" not in code.lower() and "at::" not in code and "#include <cuda" not in code


def test_product_never_imports_the_oracle():
    pkgdir = os.path.join(ROOT, "3dhandposeestimation_b200")
    for dirpath, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, f"{f} mentions the oracle"


def test_argument_errors_without_a_gpu(pkg):
    lib = pkg.load_library()
    assert lib.mb_mano_forward(None, 45, None, None, None, -1, 0, None, None, None, 0, None) == -2     # B < 0
    assert lib.mb_mano_forward(None, 45, None, None, None, 0, 0, None, None, None, 0, None) == 0      # empty batch
    assert lib.mb_mano_forward(None, 46, None, None, None, 4, 0, None, None, None, 0, None) == -2     # nc range
    assert lib.mb_mano_forward(None, 45, None, None, None, 4, 0, None, None, None, 0, None) == -1     # NULL
    assert lib.mb_fk_forward(None, None, None, None, None, None, 0, 0, None, None, None) == 0
    assert lib.mb_fk_forward(None, None, None, None, None, None, 3, 0, None, None, None) == -1
    assert lib.mb_adam_step(None, None, None, None, 10, 0.1, 0.9, 0.999, 1e-8, 0, None) == -2          # step < 1
    assert lib.mb_mano_workspace_bytes(1024, 0) >= 1024 * (148 + 192 + 2336 + 192 + 148) * 4


def test_pack_constants_layout_and_tree(pkg, synth_model):
    lib = pkg.load_library()
    packed = pkg.assets.pack_mano(synth_model, 45)
    assert packed.basis.shape == (148, 2334) and packed.basis[145].reshape(778, 3)[5, 1] == np.float32(synth_model["v_template"][5, 1])
    assert np.all(packed.basis[146:] == 0)
    assert list(packed.parents) == [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14]
    assert packed.depth.max() == 3
    # folded joint regressor equals the reference's dense product on the shaped mesh
    beta = np.random.RandomState(1).rand(10) - .5
    v_shaped = synth_model["v_template"] + synth_model["shapedirs"] @ beta
    J = synth_model["J_regressor"] @ v_shaped
    assert np.abs(packed.j0 + packed.jb @ beta - J).max() < 1e-7
    host = np.zeros(lib.mb_mano_blob_bytes(), np.uint8)
    P = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)
    keep = [np.ascontiguousarray(a) for a in (packed.basis, packed.j0, packed.jb, packed.pca, packed.pose_mean,
                                              packed.skin_w, packed.skin_b.astype(np.int32), packed.parents.astype(np.int32))]
    rc = lib.mb_mano_pack_constants(*[a.ctypes.data_as(C.c_void_p) for a in keep[:4]], 45,
                                    *[a.ctypes.data_as(C.c_void_p) for a in keep[4:]], host.ctypes.data_as(C.c_void_p))
    assert rc == 0
    hdr = host[:16 * 4].view(np.int32)
    assert hdr[0] == 0x4d423230 and hdr[2] == 45 and hdr[3] == 3
    bad = keep[7].copy(); bad[3] = 9      # parent after child
    rc = lib.mb_mano_pack_constants(*[a.ctypes.data_as(C.c_void_p) for a in keep[:4]], 45,
                                    *[a.ctypes.data_as(C.c_void_p) for a in keep[4:7]], bad.ctypes.data_as(C.c_void_p),
                                    host.ctypes.data_as(C.c_void_p))
    assert rc == -5


def test_mano_layer_interface_matches_reference(pkg, synth_model):
    import torch

    layer = pkg.ManoLayer("cuda", model=synth_model, pose_num=10)
    # no parameters / buffers / state-dict keys
    assert list(layer.parameters()) == [] and list(layer.buffers()) == [] and len(layer.state_dict()) == 0
    assert (layer.pose_num, layer.bases_num, layer.mesh_num, layer.keypoints_num) == (10, 10, 778, 16)
    assert layer.faces.shape == (1538, 3) and layer.kintree_table.shape == (2, 16)
    assert layer.parent[4] == 0 and layer.parent[15] == 14
    with pytest.raises(FileNotFoundError):
        pkg.ManoLayer("cuda", "/nonexistent/MANO_RIGHT.pkl")
    with pytest.raises(TypeError):
        pkg.ManoLayer("cuda")
    # no CPU fallback: CPU tensors are refused loudly
    with pytest.raises(pkg.ManoB200Error):
        layer(torch.zeros(2, 3), torch.zeros(2, 10), torch.zeros(2, 10))
    with pytest.raises(pkg.ManoB200Error):
        pkg.ForwardKinematics()(torch.zeros(1, 3), torch.zeros(1, 23), torch.zeros(1, 20), torch.eye(3)[None],
                                torch.ones(1, 1), torch.zeros(1, 3))
    with pytest.raises(pkg.ManoB200Error):
        pkg.MPJPE()(torch.zeros(1, 21, 3), torch.zeros(1, 21, 3), torch.ones(1, 21, 1))
    with pytest.raises(pkg.ManoB200Error):
        pkg.batch_project_xyz_to_uv(torch.zeros(1, 21, 3), torch.eye(3)[None])
    # ... and so do the drop-ins of the rows either side of the path
    z21 = torch.zeros(1, 21, 3)
    for call in (lambda: pkg.match_mano_to_RHD(z21, torch.ones(1, 1), torch.zeros(1, 3)),
                 lambda: pkg.mano_joints_to_rhd_uv(z21, torch.ones(1, 1), torch.zeros(1, 3), torch.eye(3)[None]),
                 lambda: pkg.bone_rel_trafo(z21), lambda: pkg.bone_rel_trafo_inv(z21), lambda: pkg.canonical_trafo(z21),
                 lambda: pkg.flip_right_hand(z21, torch.ones(1, dtype=torch.bool)),
                 lambda: pkg._get_rot_mat(torch.zeros(1, 1), torch.zeros(1, 1), torch.zeros(1, 1)),
                 lambda: pkg.viewpoint_transform(z21, torch.zeros(1, 1), torch.zeros(1, 1), torch.zeros(1, 1)),
                 lambda: pkg.compute_hand_mask_loss(torch.zeros(1, 21, 2), torch.zeros(1, 21, 2), torch.zeros(1, 8, 8)),
                 lambda: pkg.L2Loss()(z21, z21, torch.ones(1, 21, 1))):
        with pytest.raises(pkg.ManoB200Error):
            call()


def test_missing_library_fails_loudly(pkg, monkeypatch):
    monkeypatch.setattr(pkg._cabi, "_lib", None)
    monkeypatch.setattr(pkg._cabi, "LIB_PATH", "/nonexistent/libmano_b200.so")
    with pytest.raises(pkg.ManoB200Error, match="no CPU/PyTorch fallback"):
        pkg._cabi.lib()


def test_real_pkl_reader_if_present(pkg):
    path = "/root/reference/config/mano/models/MANO_RIGHT.pkl"
    if not os.path.isfile(path):
        pytest.skip("real MANO pkl not on this machine")
    m = pkg.assets.read_mano_pkl(path)
    assert m["shapedirs"].shape == (778, 3, 10) and (m["weights"] != 0).sum() == 2028
    pkg.assets.pack_mano(m, 45)


# ---- kernel math compiled for the host -------------------------------------------------
@pytest.fixture(scope="module")
def hostlib():
    out = os.path.join(ROOT, "tests", "host", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libhostcheck.so")
    src = os.path.join(ROOT, "tests", "host", "host_check.cu")
    subprocess.run(["nvcc", "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-gencode",
                    "arch=compute_100a,code=sm_100a", "-o", so, src], check=True)
    return C.CDLL(so)


P = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("scale", [3.0, 1.0, 0.05, 1e-3, 0.0])
def test_device_rodrigues_math_on_host(hostlib, scale):
    rs = np.random.RandomState(3)
    for _ in range(100):
        r = ((rs.rand(3) - .5) * 2 * scale).astype(np.float32)
        dR = rs.randn(9).astype(np.float32)
        R, g = np.zeros(9, np.float32), np.zeros(3, np.float32)
        hostlib.hc_rodrigues(P(r), P(R))
        hostlib.hc_rodrigues_bwd(P(r), P(dR), P(g))
        Ro = mo.rodrigues(r.astype(np.float64)[None])[0].ravel()
        go = mo.rodrigues_backward(r.astype(np.float64)[None], dR.astype(np.float64).reshape(1, 3, 3))[0]
        assert np.abs(R - Ro).max() < 1e-6
        assert np.isfinite(g).all()                        # analytic limit at theta -> 0
        assert np.abs(g - go).max() <= 5e-6 * max(1.0, np.abs(go).max())


@pytest.mark.parametrize("swap", [0, 1])
def test_device_fk_math_on_host(hostlib, swap):
    rs = np.random.RandomState(4)
    B = 48
    ra = ((rs.rand(B, 3) - .5) * 2 * np.pi).astype(np.float32)
    oa = ((rs.rand(B, 23) - .5) * np.pi).astype(np.float32)
    bl = (rs.rand(B, 20) + .1).astype(np.float32)
    K = np.tile(np.array([[282.9, 0, 160], [0, 282.9, 160], [0, 0, 1]], np.float32), (B, 1, 1))
    s = (rs.rand(B) * .05 + .02).astype(np.float32)
    root = (rs.randn(B, 3) * .05 + np.array([0, 0, .6])).astype(np.float32)
    gx = rs.randn(B, 21, 3).astype(np.float32)
    gu = (rs.randn(B, 21, 2) * 1e-3).astype(np.float32)
    xyz, uv = np.zeros((B, 21, 3), np.float32), np.zeros((B, 21, 2), np.float32)
    hostlib.hc_fk_forward(B, P(ra), P(oa), P(bl), P(K), P(s), P(root), swap, P(xyz), P(uv))
    oxyz, ouv = fo.fk_forward(ra, oa, bl, K, s, root, joint_order_switched=not swap)
    assert np.abs(xyz - oxyz).max() < 2e-7 and np.abs(uv - ouv).max() < 1e-3
    for a, b in ((gx, gu), (gx, None), (None, gu)):
        gra, goa, gbl = np.zeros((B, 3), np.float32), np.zeros((B, 23), np.float32), np.zeros((B, 20), np.float32)
        hostlib.hc_fk_backward(B, P(ra), P(oa), P(bl), P(K), P(s), P(root), swap, P(a), P(b), P(gra), P(goa), P(gbl))
        ora, ooa, obl = fo.fk_backward(ra, oa, bl, K, s, root, a, b, joint_order_switched=not swap)
        for got, want in ((gra, ora), (goa, ooa), (gbl, obl)):
            assert np.abs(got - want).max() / np.abs(want).max() < 1e-4


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_fk_specialised_rotations_equal_the_general_form(hostlib, kind):
    """fk_math.cuh chains local rotations by their zero structure (xyz / xy / x / y angles): the result must be the
    SAME floats as euler_xyz + m3_mul (sin 0, cos 0 exact; zero products add nothing in the fmaf chains), and the
    specialised angle gradients must agree with euler_xyz_bwd on the angles that exist."""
    rs = np.random.RandomState(10 + kind)
    live = {0: [0, 1, 2], 1: [0, 1], 2: [0], 3: [1]}[kind]
    for _ in range(200):
        A = rs.randn(9).astype(np.float32)
        ang = ((rs.rand(3) - .5) * 8).astype(np.float32)
        sp, ge = np.zeros(9, np.float32), np.zeros(9, np.float32)
        hostlib.hc_fk_chain(kind, P(A), P(ang), P(sp), P(ge))
        assert np.array_equal(sp, ge), (kind, sp, ge)
        dR = rs.randn(9).astype(np.float32)
        gs, gg = np.zeros(3, np.float32), np.zeros(3, np.float32)
        hostlib.hc_fk_angle_grad(kind, P(ang), P(dR), P(gs), P(gg))
        assert np.abs(gs[live] - gg[live]).max() <= 4e-6 * max(1.0, np.abs(gg[live]).max())


def _pack_host_blob(pkg, model, nc=45):
    lib = pkg.load_library()
    packed = pkg.assets.pack_mano(model, nc)
    host = np.zeros(lib.mb_mano_blob_bytes(), dtype=np.uint8)
    keep = [np.ascontiguousarray(a) for a in (packed.basis, packed.j0, packed.jb, packed.pca, packed.pose_mean,
                                              packed.skin_w, packed.skin_b.astype(np.int32), packed.parents.astype(np.int32))]
    args = [a.ctypes.data_as(C.c_void_p) for a in keep]
    rc = lib.mb_mano_pack_constants(args[0], args[1], args[2], args[3], nc, args[4], args[5], args[6], args[7],
                                    host.ctypes.data_as(C.c_void_p))
    return rc, host, keep


def test_skin_program_schedule_is_valid(pkg, synth_model):
    """The skinning kernels trust a host-built program: blocks of 8 vertices inside 16-vertex segments
    and a static (Belady) schedule of which shared-memory slot holds which bone when.  The library
    replays the schedule on the host over two consecutive hand groups; here the replay must accept the
    synthetic model (and the real MANO pickle when the reference checkout is present) and the bone
    loads per sweep must stay far below one per (block, bone) entry."""
    lib = pkg.load_library()
    models = [("synthetic", synth_model)]
    real = "/root/reference/config/mano/models/MANO_RIGHT.pkl"
    if os.path.isfile(real):
        models.append(("real", pkg.assets.read_mano_pkl(real)))
    for name, model in models:
        rc, host, _ = _pack_host_blob(pkg, model)
        assert rc == 0, name
        stats = (C.c_int32 * 4)()
        assert lib.mb_mano_skin_program_stats(host.ctypes.data_as(C.c_void_p), stats) == 0, name
        entries, loads, blocks, max_bones = list(stats)
        assert blocks == 98 and 98 <= entries <= 512 and max_bones <= 12, (name, list(stats))
        assert loads < entries // 2, (name, list(stats))
    # an empty (all-zero) blob is rejected
    bad = np.zeros_like(host)
    assert lib.mb_mano_skin_program_stats(bad.ctypes.data_as(C.c_void_p), stats) == -5


def test_model_flags(pkg, synth_model):
    lib = pkg.load_library()
    parents = pkg.assets.pack_mano(synth_model, 45).parents.astype(np.int32)
    assert lib.mb_mano_model_flags(parents.ctypes.data_as(C.c_void_p)) == 0x100
    other = parents.copy()
    other[5] = 1
    assert lib.mb_mano_model_flags(other.ctypes.data_as(C.c_void_p)) == 0


def test_blob_file_round_trip(pkg, synth_model, tmp_path):
    """Packed constants on disk: save_blob() / ManoLayer(device, '*.mb20.npz') reproduce the blob byte for byte
    and refuse files of another ABI, pose_num or size (no pickle is involved in loading)."""
    layer = pkg.ManoLayer("cpu", model=synth_model, pose_num=10)
    path = str(tmp_path / "synthetic.mb20.npz")
    layer.save_blob(path)
    again = pkg.ManoLayer("cpu", path, pose_num=10)
    assert np.array_equal(again._blob_host.numpy(), layer._blob_host.numpy())
    assert again._mode == layer._mode and again.parent == layer.parent
    assert np.array_equal(np.asarray(again.faces), np.asarray(layer.faces))
    with pytest.raises(pkg._cabi.ManoB200Error):
        pkg.ManoLayer("cpu", path, pose_num=45)
    with pytest.raises(ValueError):
        layer.save_blob(str(tmp_path / "x.bin"))
    z = dict(np.load(path))
    z["abi"] = np.int32(1)
    bad = str(tmp_path / "old.mb20.npz")
    np.savez(bad, **z)
    with pytest.raises(pkg._cabi.ManoB200Error):
        pkg.ManoLayer("cpu", bad, pose_num=10)
