"""ctypes binding of libmano_b200.so (the C ABI declared in include/mano_b200.h).

There is no fallback: if the shared library is missing, cannot be loaded or the
device is not a B200, every entry point raises.  The library is built in-tree by
``build.py`` (``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
# MANO_B200_LIB selects another build of the same library (A/B timing of kernel variants: profiles/tools)
LIB_PATH = os.environ.get("MANO_B200_LIB") or os.path.join(_HERE, "libmano_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "mano_b200.h")

MODE_FP32, MODE_F16X3, MODE_F16 = 0, 1, 2
MODES = {"fp32": MODE_FP32, "f16x3": MODE_F16X3, "f16": MODE_F16}
BWD_WORKSPACE_VALID = 1
MODEL_CHAINS_5X3 = 0x100
FWD_INFERENCE = 0x200
FWD_FUSED = 0x400
FWD_UNFUSED = 0x800
REDUCE_MPJPE_MM, REDUCE_L2 = 0, 1
HEAD_XYZ, HEAD_UV, HEAD_REG, HEAD_MATCH = 1, 2, 4, 8
VIS_F32, VIS_U8 = 0, 1

_p = C.c_void_p
_i = C.c_int
_ll = C.c_longlong
_f = C.c_float
_sz = C.c_size_t

# name -> (restype, argtypes); must list every symbol include/mano_b200.h declares
SIGNATURES = {
    "mb_abi_version": (_i, []),
    "mb_error_string": (C.c_char_p, [_i]),
    "mb_check_device": (_i, []),
    "mb_mano_blob_bytes": (_sz, []),
    "mb_mano_pack_constants": (_i, [_p, _p, _p, _p, _i, _p, _p, _p, _p, _p]),
    "mb_mano_model_flags": (_i, [_p]),
    "mb_mano_skin_program_stats": (_i, [_p, _p]),
    "mb_mano_workspace_bytes": (_sz, [_i, _i]),
    "mb_mano_forward": (_i, [_p, _i, _p, _p, _p, _i, _i, _p, _p, _p, _sz, _p]),
    "mb_mano_forward_debug": (_i, [_p, _i, _p, _p, _p, _i, _i, _p, _p, _p, _sz, _p, _i, _p]),
    "mb_mano_backward": (_i, [_p, _i, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _sz, _p]),
    "mb_affine_forward": (_i, [_p, _p, _p, _p, _i, _p]),
    "mb_affine_backward": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _p, _p, _p, _p, _p, _p]),
    "mb_lbs_workspace_bytes": (_sz, [_i]),
    "mb_lbs_forward": (_i, [_p, _p, _i, _p, _i, _p, _p, _p, _sz, _p]),
    "mb_fk_forward": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _p, _p, _p]),
    "mb_fk_backward": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p, _p, _p, _p]),
    "mb_fk_loss_workspace_bytes": (_sz, [_i]),
    "mb_fk_loss_forward": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _sz, _p]),
    "mb_fk_loss_backward": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _sz, _p]),
    "mb_project_uv_forward": (_i, [_p, _p, _i, _i, _p, _p]),
    "mb_project_uv_backward": (_i, [_p, _p, _p, _i, _i, _p, _p]),
    "mb_joint_epilogue_forward": (_i, [_p, _p, _p, _p, _i, _i, _p, _p, _p, _p]),
    "mb_joint_epilogue_backward": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _p, _p, _p, _p]),
    "mb_bone_rel_trafo": (_i, [_p, _i, _p, _p]),
    "mb_bone_rel_trafo_inv": (_i, [_p, _i, _p, _p]),
    "mb_canonical_trafo": (_i, [_p, _p, _i, _p, _p, _p]),
    "mb_flip_right_hand": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "mb_mirror_hand": (_i, [_p, _p, _i, _i, _i, _i, _p, _p]),
    "mb_viewpoint_forward": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p]),
    "mb_viewpoint_backward": (_i, [_p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p]),
    "mb_masked_joint_reduce": (_i, [_p, _p, _p, _i, _ll, _i, _i, _p, _p, _p]),
    "mb_masked_l2_backward": (_i, [_p, _p, _p, _i, _ll, _i, _p, _p, _p, _p]),
    "mb_regulariser_forward": (_i, [_p, _ll, _p, _ll, _f, _p, _p, _p]),
    "mb_regulariser_backward": (_i, [_p, _ll, _p, _ll, _f, _p, _p, _p, _p, _p]),
    "mb_mano_head_loss_workspace_bytes": (_sz, [_i]),
    "mb_mano_head_loss_forward": (_i, [_p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _p, _p, _p, _p, _sz, _p]),
    "mb_mano_head_loss_backward": (_i, [_p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _p, _p, _p, _p, _p, _p,
                                        _p, _p, _p, _sz, _p]),
    "mb_hand_mask_loss": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p]),
    "mb_mano_fit_step": (_i, [_p, _i, _p, _p, _p, _p, _p, _i, _i, _p, _p, _f, _f, _f, _f, _i, _i, _p]),
    "mb_fit_finalize": (_i, [_p, _p, _i, _p, _p, _p]),
    "mb_adam_step": (_i, [_p, _p, _p, _p, _ll, _f, _f, _f, _f, _i, _p]),
    "mb_launch_count": (_ll, []),
    "mb_profile_enable": (None, [_i]),
    "mb_profile_collect": (_i, [_p, _p]),
}
STAGES = ("pose_fwd", "blend_fwd", "lbs_fwd", "lbs_bwd", "blend_bwd", "pose_bwd",
          "joints_only_fwd", "joints_only_bwd", "fk_fwd", "fk_bwd", "fused_fwd")


class ManoB200Error(RuntimeError):
    pass


_lib = None


def header_symbols() -> list:
    """Function names declared (MB_API ...) in include/mano_b200.h."""
    with open(HEADER_PATH) as fh:
        text = fh.read()
    return re.findall(r"MB_API\s+[\w\s\*]+?\b(mb_\w+)\s*\(", text)


def profile_collect() -> dict:
    """{stage: (total_ms, launches)} since profiling was enabled / last collected."""
    ms = (C.c_double * len(STAGES))()
    cnt = (C.c_longlong * len(STAGES))()
    check(lib().mb_profile_collect(ms, cnt), "mb_profile_collect")
    return {name: (ms[i], cnt[i]) for i, name in enumerate(STAGES) if cnt[i]}


def lib() -> C.CDLL:
    """Load the shared library (once).  Raises ManoB200Error when it is absent —
    the product has no CPU or PyTorch fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ManoB200Error(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU/PyTorch fallback for this path.")
    try:
        handle = C.CDLL(LIB_PATH)
    except OSError as exc:
        raise ManoB200Error(f"cannot load {LIB_PATH}: {exc}") from exc
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(handle, name)
        except AttributeError as exc:
            raise ManoB200Error(f"{LIB_PATH} does not export {name}") from exc
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().mb_error_string(rc).decode()
        raise ManoB200Error(f"{what} failed ({rc}): {msg}")


def on_tensor_device(fn):
    """Run ``fn`` with the CUDA device of its first CUDA tensor argument (or of the autograd context's saved
    tensors) current.  The C ABI launches on the CURRENT device and only receives a stream, so a layer built on
    ``cuda:1`` while ``cuda:0`` is current would otherwise fail every launch (cudaErrorInvalidResourceHandle)."""
    import functools

    import torch

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = None
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.Tensor) and a.is_cuda:
                dev = a.device
                break
        if dev is None and args and hasattr(args[0], "saved_tensors"):
            for a in args[0].saved_tensors:
                if isinstance(a, torch.Tensor) and a.is_cuda:
                    dev = a.device
                    break
        if dev is None or dev.index == torch.cuda.current_device():          # the common case: no context switch to pay for
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)

    return wrapper


def ptr(t):
    """Raw device/host pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_handle(device=None):
    """Raw cudaStream_t of torch's current stream on ``device`` (the private fast accessor where this torch has it: the
    public ``torch.cuda.current_stream`` builds a Stream object, ~12 us per call — measured 1 200 calls in 15 ms)."""
    import torch

    try:
        idx = device.index if isinstance(device, torch.device) else device
        if idx is None:
            idx = torch.cuda.current_device()
        return torch._C._cuda_getCurrentRawStream(idx)
    except (AttributeError, TypeError):
        return torch.cuda.current_stream(device).cuda_stream
