"""ManoLayer — drop-in for the reference's ``network/sub_modules/MANOLayer.py:ManoLayer``.

Same constructor and ``forward`` / ``rot_pose_beta_to_mesh`` signatures
(MANOLayer.py:52, :122, :238): ``(root_angles[B,3], other_angles[B,pose_num],
betas[B,10]) -> (vertices[B,778,3], joint[B,21,3])``; same public attributes;
no parameters, buffers or state-dict keys.  The arithmetic runs in
hand-written sm_100a CUDA behind the C ABI of include/mano_b200.h through a
``torch.autograd.Function``; PyTorch only provides device memory, the current
stream and the autograd plumbing.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _cabi, assets


def _as_f32_cuda(t: torch.Tensor, name: str, device: torch.device) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if t.device.type != "cuda":
        raise _cabi.ManoB200Error(
            f"{name} is on {t.device}; this layer only runs on CUDA (sm_100a) — there is no CPU fallback")
    if t.device != device:
        raise RuntimeError(f"{name} is on {t.device} but the layer's constants are on {device}")
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    return t.contiguous()


class _ManoFunction(torch.autograd.Function):
    """verts, joints = f(rot, coeffs, betas).  Saves only the three inputs and the
    forward's scratch (feature rows, bone transforms, rest-pose vertices); the
    backward re-uses that scratch once and recomputes it if called again."""

    @staticmethod
    @_cabi.on_tensor_device
    def forward(ctx, rot, coeffs, betas, transl, scale, layer, want_verts):
        lib = _cabi.lib()
        B = rot.shape[0]
        dev = rot.device
        nc = layer.pose_num
        joints = torch.empty((B, 21, 3), dtype=torch.float32, device=dev)
        stream = _cabi.stream_handle(dev)
        ws = None
        # no gradient will be asked for (torch.no_grad(), or no input requires grad): the forward keeps no scratch
        need_bwd = any(ctx.needs_input_grad[:3])
        if want_verts:
            verts = torch.empty((B, 778, 3), dtype=torch.float32, device=dev)
            nbytes = lib.mb_mano_workspace_bytes(B, layer._mode)
            ws = torch.empty((max(nbytes, 16),), dtype=torch.uint8, device=dev)
            fwd_mode = layer._mode | layer._fwd_flags
            if not (need_bwd and layer.keep_workspace):
                fwd_mode |= _cabi.FWD_INFERENCE
            _cabi.check(lib.mb_mano_forward(layer._blob.data_ptr(), nc, rot.data_ptr(), coeffs.data_ptr(),
                                            betas.data_ptr(), B, fwd_mode, verts.data_ptr(), joints.data_ptr(),
                                            ws.data_ptr(), ws.numel(), stream), "mb_mano_forward")
        else:
            verts = None
            _cabi.check(lib.mb_mano_forward(layer._blob.data_ptr(), nc, rot.data_ptr(), coeffs.data_ptr(),
                                            betas.data_ptr(), B, layer._mode, None, joints.data_ptr(),
                                            None, 0, stream), "mb_mano_forward(joints only)")
        ctx.layer = layer
        ctx.ws = ws if (layer.keep_workspace and want_verts and need_bwd) else None
        ctx.ws_valid = ctx.ws is not None
        ctx.affine = transl is not None or scale is not None
        if ctx.affine:
            # p' = scale * p + transl per hand, in place on the outputs (csrc/affine.cu)
            _cabi.check(lib.mb_affine_forward(_cabi.ptr(verts), joints.data_ptr(), _cabi.ptr(scale), _cabi.ptr(transl), B, stream),
                        "mb_affine_forward")
        ctx.set_materialize_grads(False)
        if want_verts:
            if ctx.affine and any(ctx.needs_input_grad[:5]):
                ctx.save_for_backward(rot, coeffs, betas, transl, scale, verts, joints)
            else:
                ctx.save_for_backward(rot, coeffs, betas, transl, scale)
            return verts, joints
        if ctx.affine and any(ctx.needs_input_grad[:5]):
            ctx.save_for_backward(rot, coeffs, betas, transl, scale, None, joints)
        else:
            ctx.save_for_backward(rot, coeffs, betas, transl, scale)
        placeholder = joints.new_empty((0,))
        ctx.mark_non_differentiable(placeholder)
        return placeholder, joints

    @staticmethod
    @_cabi.on_tensor_device
    def backward(ctx, g_verts, g_joints):
        saved = ctx.saved_tensors
        rot, coeffs, betas, transl, scale = saved[:5]
        verts_out, joints_out = (saved[5], saved[6]) if len(saved) > 5 else (None, None)
        layer = ctx.layer
        lib = _cabi.lib()
        B = rot.shape[0]
        dev = rot.device
        nc = layer.pose_num
        stream = _cabi.stream_handle(dev)
        if g_joints is None:
            g_joints = torch.zeros((B, 21, 3), dtype=torch.float32, device=dev)
        else:
            g_joints = g_joints.to(torch.float32).contiguous()
        g_rot = torch.empty_like(rot)
        g_coeffs = torch.empty_like(coeffs)
        g_betas = torch.empty_like(betas)
        if g_verts is not None and g_verts.numel() == 0:
            g_verts = None
        if g_verts is None:
            _cabi.check(lib.mb_mano_backward(layer._blob.data_ptr(), nc, rot.data_ptr(), coeffs.data_ptr(),
                                             betas.data_ptr(), None, g_joints.data_ptr(), B, layer._mode, 0,
                                             g_rot.data_ptr(), g_coeffs.data_ptr(), g_betas.data_ptr(),
                                             None, 0, stream), "mb_mano_backward(joints only)")
        else:
            g_verts = g_verts.to(torch.float32).contiguous()
            flags = 0
            ws = ctx.ws
            if ws is not None and ctx.ws_valid:
                flags = _cabi.BWD_WORKSPACE_VALID
                ctx.ws_valid = False          # the backward consumes v_posed (dv_posed aliases it)
            elif ws is None:
                nbytes = lib.mb_mano_workspace_bytes(B, layer._mode)
                ws = torch.empty((max(nbytes, 16),), dtype=torch.uint8, device=dev)
            _cabi.check(lib.mb_mano_backward(layer._blob.data_ptr(), nc, rot.data_ptr(), coeffs.data_ptr(),
                                             betas.data_ptr(), g_verts.data_ptr(), g_joints.data_ptr(), B,
                                             layer._mode, flags, g_rot.data_ptr(), g_coeffs.data_ptr(),
                                             g_betas.data_ptr(), ws.data_ptr(), ws.numel(), stream),
                        "mb_mano_backward")
            if not ctx.ws_valid:
                ctx.ws = None                 # consumed: let the caching allocator reuse it (stream-ordered)
        g_transl = g_scale = None
        if ctx.affine:
            if transl is not None and ctx.needs_input_grad[3]:
                g_transl = torch.empty_like(transl)
            if scale is not None and ctx.needs_input_grad[4]:
                g_scale = torch.empty_like(scale)
            _cabi.check(lib.mb_affine_backward(_cabi.ptr(g_verts), g_joints.data_ptr(), _cabi.ptr(verts_out), _cabi.ptr(joints_out),
                                               _cabi.ptr(scale), _cabi.ptr(transl), B, nc, _cabi.ptr(g_scale), _cabi.ptr(g_transl),
                                               g_rot.data_ptr(), g_coeffs.data_ptr(), g_betas.data_ptr(), stream),
                        "mb_affine_backward")
        return g_rot, g_coeffs, g_betas, g_transl, g_scale, None, None


BLOB_SUFFIX = ".mb20.npz"
BLOB_MAGIC = b"MB20BLOB"


class ManoLayer(nn.Module):
    """B200-native MANO layer with the reference's interface.

    Parameters mirror MANOLayer.py:52: ``ManoLayer(device, MANO_RIGHT_pkl=None,
    bases_num=10, pose_num=6)``.  Keyword-only extensions: ``model`` (a dict of
    numpy arrays with the pkl's keys, e.g. ``assets.synthetic_mano()``) instead
    of a pkl path; ``mode`` in {"fp32", "f16x3", "f16"} selects the blend-shape
    contraction precision; ``keep_workspace`` trades 12 KB/hand of retained
    memory for not recomputing the forward in the backward; ``fused_forward=False``
    selects the two separate forward kernels (blend contraction + lane = hand skinning)
    instead of the fused blend + skinning kernel with lane = vertex (csrc/vskin.cu) that
    runs from 8 192 hands on by default (the measured faster one, with and without a
    backward to follow).
    """

    def __init__(self, device, MANO_RIGHT_pkl=None, bases_num=10, pose_num=6, *, model=None, mode="f16x3",
                 keep_workspace=True, fused_forward=None):
        super().__init__()
        self.device = device
        self.bases_num = bases_num
        self.pose_num = int(pose_num)
        self.mesh_num = 778
        self.keypoints_num = 16
        if mode not in _cabi.MODES:
            raise ValueError(f"mode must be one of {sorted(_cabi.MODES)}, got {mode!r}")
        self._mode = _cabi.MODES[mode]      # model property bits (mb_mano_model_flags) are OR-ed in below
        self.mode = mode
        self.keep_workspace = bool(keep_workspace)
        self._fwd_flags = 0 if fused_forward is None else (_cabi.FWD_FUSED if fused_forward else _cabi.FWD_UNFUSED)

        if model is None:
            if MANO_RIGHT_pkl is None:
                raise TypeError("MANO_RIGHT_pkl (path to MANO_RIGHT.pkl) or model= is required")
            if str(MANO_RIGHT_pkl).endswith(BLOB_SUFFIX):     # packed constants written by save_blob()
                self._init_from_blob_file(MANO_RIGHT_pkl, device)
                return
            model = assets.read_mano_pkl(MANO_RIGHT_pkl)      # FileNotFoundError like MANOLayer.py:63
        self.kintree_table = model["kintree_table"]
        self.id_to_col = {int(self.kintree_table[1, i]): i for i in range(self.kintree_table.shape[1])}
        self.parent = {i: self.id_to_col[int(self.kintree_table[0, i])] for i in range(1, self.kintree_table.shape[1])}
        self.faces = model["f"]

        packed = assets.pack_mano(model, self.pose_num)
        lib = _cabi.lib()
        nbytes = lib.mb_mano_blob_bytes()
        host = np.zeros(nbytes, dtype=np.uint8)
        keep = [np.ascontiguousarray(a) for a in (packed.basis, packed.j0, packed.jb, packed.pca, packed.pose_mean,
                                                  packed.skin_w, packed.skin_b.astype(np.int32),
                                                  packed.parents.astype(np.int32))]
        args = [a.ctypes.data_as(C.c_void_p) for a in keep]
        _cabi.check(lib.mb_mano_pack_constants(args[0], args[1], args[2], args[3], self.pose_num, args[4], args[5],
                                               args[6], args[7], host.ctypes.data_as(C.c_void_p)),
                    "mb_mano_pack_constants")
        self._mode |= int(lib.mb_mano_model_flags(args[7]))
        self._blob_host = torch.from_numpy(host)
        self._blob = None
        dev = torch.device(device)
        if dev.type == "cuda" and torch.cuda.is_available():
            self._upload(dev)

    # ---- packed constants on disk: no pickle / chumpy at run time ----------------
    def save_blob(self, path) -> None:
        """Write the packed constant blob plus the few host attributes of the layer to ``path``
        (must end in ``.mb20.npz``); ``ManoLayer(device, path)`` loads it back without unpickling."""
        if not str(path).endswith(BLOB_SUFFIX):
            raise ValueError(f"blob files end in {BLOB_SUFFIX}")
        with open(path, "wb") as fh:
            np.savez(fh, magic=np.frombuffer(BLOB_MAGIC, dtype=np.uint8), abi=np.int32(_cabi.lib().mb_abi_version()),
                     pose_num=np.int32(self.pose_num), model_flags=np.int32(self._mode & ~0xff),
                     blob=self._blob_host.numpy(), kintree_table=np.asarray(self.kintree_table), faces=np.asarray(self.faces))

    def _init_from_blob_file(self, path, device) -> None:
        lib = _cabi.lib()
        with np.load(path, allow_pickle=False) as z:
            if bytes(z["magic"].tobytes()) != BLOB_MAGIC:
                raise _cabi.ManoB200Error(f"{path}: not a packed MANO blob")
            if int(z["abi"]) != lib.mb_abi_version():
                raise _cabi.ManoB200Error(f"{path}: packed for ABI {int(z['abi'])}, library is ABI {lib.mb_abi_version()}; re-pack it")
            if int(z["pose_num"]) != self.pose_num:
                raise _cabi.ManoB200Error(f"{path}: packed for pose_num={int(z['pose_num'])}, layer asks for {self.pose_num}")
            host = np.ascontiguousarray(z["blob"], dtype=np.uint8)
            self.kintree_table = z["kintree_table"]
            self.faces = z["faces"]
            flags = int(z["model_flags"])
        if host.size != lib.mb_mano_blob_bytes():
            raise _cabi.ManoB200Error(f"{path}: blob has {host.size} bytes, library expects {lib.mb_mano_blob_bytes()}")
        stats = (C.c_int32 * 4)()
        _cabi.check(lib.mb_mano_skin_program_stats(host.ctypes.data_as(C.c_void_p), stats), f"{path}: skin program")
        self.id_to_col = {int(self.kintree_table[1, i]): i for i in range(self.kintree_table.shape[1])}
        self.parent = {i: self.id_to_col[int(self.kintree_table[0, i])] for i in range(1, self.kintree_table.shape[1])}
        self._mode |= flags
        self._blob_host = torch.from_numpy(host.copy())
        self._blob = None
        dev = torch.device(device)
        if dev.type == "cuda" and torch.cuda.is_available():
            self._upload(dev)

    # constants are plain attributes, like the reference's (no state-dict keys)
    def _upload(self, dev: torch.device) -> None:
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().mb_check_device(), "mb_check_device")
        self._blob = self._blob_host.to(dev)
        self._dev = dev

    def _require_device(self) -> torch.device:
        if self._blob is None:
            dev = torch.device(self.device)
            if dev.type != "cuda":
                raise _cabi.ManoB200Error(
                    f"ManoLayer was constructed for device {self.device!r}; this implementation only runs on "
                    "CUDA (B200, sm_100a) and has no CPU fallback")
            if not torch.cuda.is_available():
                raise _cabi.ManoB200Error("CUDA is not available; ManoLayer has no CPU fallback")
            self._upload(dev)
        return self._dev

    def rot_pose_beta_to_mesh(self, rots, poses, betas, joints_only=False, transl=None, scale=None):
        """MANOLayer.py:122-208.  ``joints_only=True`` (extension) skips the 778-vertex
        contraction and returns ``(None, joint)`` — the only output the heads use.
        ``transl[B,3]`` / ``scale[B]`` or ``[B,1]`` (extensions, default None = the reference's
        behaviour): every vertex and joint of hand b becomes ``scale[b] * p + transl[b]``,
        differentiable in both (the post-ops of resnet50MANO.py:77-81 folded into the layer)."""
        dev = self._require_device()
        if self.bases_num != 10:
            raise RuntimeError("bases_num must be 10 (the reference's view at MANOLayer.py:131 fails otherwise)")
        rots = _as_f32_cuda(rots, "rots", dev)
        poses = _as_f32_cuda(poses, "poses", dev)
        betas = _as_f32_cuda(betas, "betas", dev)
        B = rots.shape[0]
        if rots.shape != (B, 3) or poses.shape != (B, self.pose_num) or betas.shape != (B, 10):
            raise RuntimeError(f"expected rots[B,3], poses[B,{self.pose_num}], betas[B,10]; got "
                               f"{tuple(rots.shape)}, {tuple(poses.shape)}, {tuple(betas.shape)}")
        if transl is not None:
            transl = _as_f32_cuda(transl, "transl", dev)
            if transl.shape != (B, 3):
                raise RuntimeError(f"expected transl[B,3], got {tuple(transl.shape)}")
        if scale is not None:
            scale = _as_f32_cuda(scale, "scale", dev)
            if scale.numel() != B:
                raise RuntimeError(f"expected scale[B] or [B,1], got {tuple(scale.shape)}")
        verts, joints = _ManoFunction.apply(rots, poses, betas, transl, scale, self, not joints_only)
        return (None if joints_only else verts), joints

    def forward(self, root_angles, other_angles, betas, *, transl=None, scale=None):
        """MANOLayer.py:238-240; ``transl`` / ``scale`` are keyword-only extensions (see rot_pose_beta_to_mesh)."""
        return self.rot_pose_beta_to_mesh(root_angles, other_angles, betas, transl=transl, scale=scale)
