"""Swap the reference's classes and functions of the hot path for the B200 ones IN PLACE, without editing the
reference tree: ``install_into_reference()`` imports the reference's own modules (its checkout must be on
``sys.path``, as its scripts arrange with ``sys.path.append``) and rebinds

    network.sub_modules.MANOLayer.ManoLayer                      -> ManoLayer
    network.sub_modules.forwardKinematicsLayer.ForwardKinematics -> ForwardKinematics
    criterions.metrics.MPJPE, criterions.loss.L2Loss             -> MPJPE, L2Loss
    utils.coordinate_trans.batch_project_xyz_to_uv               -> batch_project_xyz_to_uv
    utils.general._get_rot_mat                                   -> _get_rot_mat
    utils.relative_trafo.bone_rel_trafo / bone_rel_trafo_inv     -> bone_rel_trafo / bone_rel_trafo_inv
    utils.canonical_trafo.canonical_trafo / flip_right_hand      -> canonical_trafo / flip_right_hand

in their defining modules AND in every already-imported module that holds the same object under any name (the
heads do ``from network.sub_modules.MANOLayer import ManoLayer`` at import time, e.g. resnet50MANO.py:16,
TwoDimHandPoseWithFK.py:11, Hand3DPoseNet.py:14-15), so it works before or after the heads are imported.
``uninstall()`` restores everything.  Host-side glue only: no arithmetic, and the replacements still refuse CPU
tensors (there is no fallback).
"""
from __future__ import annotations

import importlib
import sys

# (reference module, attribute) -> name in this package
TARGETS = {
    ("network.sub_modules.MANOLayer", "ManoLayer"): "ManoLayer",
    ("network.sub_modules.forwardKinematicsLayer", "ForwardKinematics"): "ForwardKinematics",
    ("criterions.metrics", "MPJPE"): "MPJPE",
    ("criterions.loss", "L2Loss"): "L2Loss",
    ("utils.coordinate_trans", "batch_project_xyz_to_uv"): "batch_project_xyz_to_uv",
    ("utils.general", "_get_rot_mat"): "_get_rot_mat",
    ("utils.relative_trafo", "bone_rel_trafo"): "bone_rel_trafo",
    ("utils.relative_trafo", "bone_rel_trafo_inv"): "bone_rel_trafo_inv",
    ("utils.canonical_trafo", "canonical_trafo"): "canonical_trafo",
    ("utils.canonical_trafo", "flip_right_hand"): "flip_right_hand",
}

_undo = []          # (module, attribute, original object)


def install_into_reference(only=None, strict=False) -> list:
    """Rebind the reference's hot-path symbols to this package's.  ``only``: iterable of attribute names to restrict
    the swap (e.g. ``["ManoLayer"]``).  Reference modules that cannot be imported are skipped (``strict=True``
    raises instead).  Returns the list of ``"module.attribute"`` names that were rebound."""
    pkg = sys.modules[__package__]
    done = []
    for (modname, attr), ours in TARGETS.items():
        if only is not None and attr not in only:
            continue
        try:
            mod = importlib.import_module(modname)
            original = getattr(mod, attr)
        except Exception:
            if strict:
                raise
            continue
        replacement = getattr(pkg, ours)
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
