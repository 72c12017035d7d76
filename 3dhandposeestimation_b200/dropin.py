"""Swap the reference's classes and functions of the hot path for the B200 ones IN PLACE, without editing the
reference tree: ``install_into_reference()`` imports the reference's own modules (its checkout must be on
``sys.path``, as its scripts arrange with ``sys.path.append``) and rebinds

    network.sub_modules.MANOLayer.ManoLayer                      -> ManoLayer
    network.sub_modules.forwardKinematicsLayer.ForwardKinematics -> ForwardKinematics
    criterions.metrics.MPJPE, criterions.loss.L2Loss             -> MPJPE, L2Loss
    criterions.loss.LossCalculation.compute_regularization_loss  -> compute_regularization_loss   (method)
    criterions.loss.LossCalculation.compute_hand_mask_loss       -> compute_hand_mask_loss        (method)
    utils.coordinate_trans.batch_project_xyz_to_uv               -> batch_project_xyz_to_uv
    utils.general._get_rot_mat                                   -> _get_rot_mat

in their defining modules AND in every already-imported module that holds the same object under any name (the
heads do ``from network.sub_modules.MANOLayer import ManoLayer`` at import time, e.g. resnet50MANO.py:16,
TwoDimHandPoseWithFK.py:11, Hand3DPoseNet.py:14-15), so it works before or after the heads are imported.

NOT swapped by default — ``DATALOADER_TARGETS``, opt in with ``dataloader=True`` or by name through ``only=``:

    utils.relative_trafo.bone_rel_trafo / bone_rel_trafo_inv     -> bone_rel_trafo / bone_rel_trafo_inv
    utils.canonical_trafo.canonical_trafo / flip_right_hand      -> canonical_trafo / flip_right_hand

The reference calls these four per sample on CPU tensors inside ``Dataset.__getitem__`` (dataloaderRHD.py:243,248,
BinaryDbReaderSTB.py:201,206, dataloaderInterHand2M6.py:413,418), usually in forked DataLoader workers, where CUDA
cannot be used at all.  The B200 versions are BATCHED, GPU-side replacements (they refuse CPU tensors): use them
after the batch is on the device, not inside ``__getitem__``.

``uninstall()`` restores everything.  Host-side glue only: no arithmetic, and the replacements still refuse CPU
tensors (there is no fallback).
"""
from __future__ import annotations

import importlib
import sys

# (reference module, attribute or Class.method) -> name in this package
TARGETS = {
    ("network.sub_modules.MANOLayer", "ManoLayer"): "ManoLayer",
    ("network.sub_modules.forwardKinematicsLayer", "ForwardKinematics"): "ForwardKinematics",
    ("criterions.metrics", "MPJPE"): "MPJPE",
    ("criterions.loss", "L2Loss"): "L2Loss",
    ("criterions.loss", "LossCalculation.compute_regularization_loss"): "compute_regularization_loss",
    ("criterions.loss", "LossCalculation.compute_hand_mask_loss"): "compute_hand_mask_loss",
    ("utils.coordinate_trans", "batch_project_xyz_to_uv"): "batch_project_xyz_to_uv",
    ("utils.general", "_get_rot_mat"): "_get_rot_mat",
}
# per-sample CPU calls inside the reference's Dataset.__getitem__: opt-in only (see the module docstring)
DATALOADER_TARGETS = {
    ("utils.relative_trafo", "bone_rel_trafo"): "bone_rel_trafo",
    ("utils.relative_trafo", "bone_rel_trafo_inv"): "bone_rel_trafo_inv",
    ("utils.canonical_trafo", "canonical_trafo"): "canonical_trafo",
    ("utils.canonical_trafo", "flip_right_hand"): "flip_right_hand",
}

_undo = []          # (holder, attribute, original object)


def _as_method(fn):
    """A module-level function of this package bound as a method of a reference class: drops ``self``."""
    def method(self, *args, **kwargs):
        return fn(*args, **kwargs)
    method.__name__ = fn.__name__
    method.__doc__ = fn.__doc__
    method.__mb_replacement__ = fn
    return method


def install_into_reference(only=None, strict=False, dataloader=False) -> list:
    """Rebind the reference's hot-path symbols to this package's.  ``only``: iterable of attribute names to restrict
    the swap (e.g. ``["ManoLayer"]``; names from ``DATALOADER_TARGETS`` are accepted here).  ``dataloader=True`` also
    swaps the four keypoint re-parameterisations the reference calls per sample in its datasets.  Reference modules
    that cannot be imported are skipped (``strict=True`` raises instead).  Returns the list of
    ``"module.attribute"`` names that were rebound."""
    pkg = sys.modules[__package__]
    done = []
    targets = dict(TARGETS)
    if dataloader:
        targets.update(DATALOADER_TARGETS)
    elif only is not None:
        targets.update({k: v for k, v in DATALOADER_TARGETS.items() if k[1] in only})
    for (modname, attr), ours in targets.items():
        leaf = attr.split(".")[-1]
        if only is not None and attr not in only and leaf not in only:
            continue
        try:
            mod = importlib.import_module(modname)
            if "." in attr:                                   # Class.method
                cls = getattr(mod, attr.split(".")[0])
                original = cls.__dict__[leaf]
            else:
                original = getattr(mod, attr)
        except Exception:
            if strict:
                raise
            continue
        replacement = getattr(pkg, ours)
        if "." in attr:
            if getattr(original, "__mb_replacement__", None) is replacement:
                continue
            _undo.append((cls, leaf, original))
            setattr(cls, leaf, _as_method(replacement))
            done.append(f"{modname}.{attr}")
            continue
        if original is replacement:
            continue
        # the defining module and every loaded module that imported the object by name
        for m in list(sys.modules.values()):
            d = getattr(m, "__dict__", None)
            if not d or m is pkg or getattr(m, "__name__", "").startswith(pkg.__name__ + "."):
                continue
            for name, val in list(d.items()):
                if val is original:
                    _undo.append((m, name, original))
                    setattr(m, name, replacement)
                    done.append(f"{m.__name__}.{name}")
    return done


def uninstall() -> int:
    """Restore every binding ``install_into_reference`` changed; returns how many."""
    n = len(_undo)
    while _undo:
        m, name, original = _undo.pop()
        setattr(m, name, original)
    return n
